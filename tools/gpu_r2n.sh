#!/bin/bash
# round 2n: tapered CTA schedule (long CTAs first, one wave each of passes/2 ... 1 last) vs the uniform schedule, same box
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
python tools/variants.py run base taper base taper > $O/ab_taper.txt 2>&1
cat $O/ab_taper.txt
for lib in libastc_b200.so libastc_b200_taper.so libastc_b200.so libastc_b200_taper.so; do echo "== $lib"; ASTC_B200_LIB=astc_encoder_b200/$lib python tools/small_sizes.py 2>&1 | tail -10; done | tee $O/small_sizes_taper.txt
ASTC_B200_LIB=astc_encoder_b200/libastc_b200_taper.so timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_build.py > $O/pytest_taper.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_taper.txt
tail -4 $O/pytest_taper.txt
