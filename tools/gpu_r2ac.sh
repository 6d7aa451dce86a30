#!/bin/bash
# round 2ac: batch host path -- neighbouring copies that are contiguous on both sides merged into one (H2D and D2H) vs one copy per image
mkdir -p gpurun_out/r2ac
for lib in libastc_b200.so libastc_b200_nomerge.so libastc_b200.so libastc_b200_nomerge.so; do
  echo "== $lib"
  ASTC_B200_LIB=astc_encoder_b200/$lib python bench.py --no-cpu --e2e-steps 1 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['config5']['batch']['e2e_host']
print('all levels uploaded', e['ms'], e['value'], e['matches_device_batch'], '| from bases', e['from_bases']['ms'], e['from_bases']['value'], e['from_bases']['matches_all_levels_uploaded'], '| pageable', e['pageable']['ms'], e['pageable']['value'])"
done | tee gpurun_out/r2ac/copy_merge.txt
