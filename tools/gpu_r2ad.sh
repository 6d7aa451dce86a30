#!/bin/bash
# round 2ad: group size again, with merged uploads
mkdir -p gpurun_out/r2ad
for lib in libastc_b200_g16.so libastc_b200_g32.so libastc_b200_g48.so libastc_b200.so libastc_b200_g32.so libastc_b200.so; do
  echo "== $lib"
  ASTC_B200_LIB=astc_encoder_b200/$lib python bench.py --no-cpu --e2e-steps 1 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['config5']['batch']['e2e_host']
print('all levels uploaded', e['ms'], e['value'], e['launches_per_call'], e['matches_device_batch'], '| from bases', e['from_bases']['ms'], e['from_bases']['value'], e['from_bases']['matches_all_levels_uploaded'], '| pageable', e['pageable']['ms'], e['pageable']['value'])"
done | tee gpurun_out/r2ad/group_size_merged.txt
