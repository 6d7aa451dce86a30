#!/bin/bash
# round 2v: block walk -- one subtract when a thread crosses the end of a block row (division only for rows shorter than
# the stride) vs a division at every row end ("divalways" = the kernel as it was); batch of 512 mip chains + single textures
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
for lib in libastc_b200.so libastc_b200_divalways.so libastc_b200.so libastc_b200_divalways.so; do
  echo "== $lib"
  ASTC_B200_LIB=astc_encoder_b200/$lib python bench.py --no-cpu --e2e-steps 1 --no-host-batch --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config5']
print('16384^2', d['ms_per_step'], 'others', [o['kernel_ms'] for o in d['others']], 'batch', c['batch']['ms'], c['batch']['value'], c['batch']['bytes_identical_to_per_texture_encode'], 'both', c['both']['ms'])"
done | tee $O/ab_walk.txt
