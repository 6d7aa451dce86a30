#!/bin/bash
# round 2, call C: same-box A/B -- round-1 library vs the tile-queue library (product defaults) vs hooks variants
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
P=$PWD/astc_encoder_b200
for rep in 1 2; do
for v in r1 base; do
  lib=$P/libastc_b200.so; [ $v = r1 ] && lib=$P/libastc_b200_r1.so
  echo "== $v (rep $rep)"; ASTC_B200_LIB=$lib timeout 300 python tools/quick_bench.py short
done
for cfg in "1 1" "1 2" "1 4" "0 2" "0 8"; do
  set -- $cfg
  echo "== hooks dynamic=$1 passes=$2 (rep $rep)"; ASTC_B200_LIB=$P/libastc_b200_hooks.so ASTC_B200_DYNAMIC=$1 ASTC_B200_PASSES=$2 timeout 300 python tools/quick_bench.py short
done
done > $O/ab.txt 2>&1
cat $O/ab.txt
