#!/bin/bash
# round 2x: batch start-up -- warp-cooperative 32-ary search of the descriptor table (3 dependent loads for 6144 images)
# vs the per-thread binary search (13); batch of 512 mip chains; full GPU suite on the new default
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
for lib in libastc_b200.so libastc_b200_binsearch.so libastc_b200.so libastc_b200_binsearch.so; do
  echo "== $lib"
  ASTC_B200_LIB=astc_encoder_b200/$lib python bench.py --no-cpu --e2e-steps 1 --no-host-batch --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config5']
print('16384^2', d['ms_per_step'], 'batch', c['batch']['ms'], c['batch']['value'], c['batch']['bytes_identical_to_per_texture_encode'], c['batch']['checksum_of_all_blocks'], 'both', c['both']['ms'])"
done | tee $O/ab_search.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 600 python tools/soak.py 1500 11 2>&1 | tail -1
