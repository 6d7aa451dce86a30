#!/bin/bash
# round 2j: first-block fetch issued before the table prologue ("early") vs the committed kernels, same box
mkdir -p gpurun_out/r2j
python tools/variants.py run base early base early > gpurun_out/r2j/ab_early.txt 2>&1
cat gpurun_out/r2j/ab_early.txt
for lib in libastc_b200.so libastc_b200_early.so libastc_b200.so libastc_b200_early.so; do echo "== $lib"; ASTC_B200_LIB=astc_encoder_b200/$lib python tools/small_sizes.py 2>&1 | tail -10; done | tee gpurun_out/r2j/small_sizes_early.txt
