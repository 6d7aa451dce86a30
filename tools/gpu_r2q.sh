#!/bin/bash
# round 2q: programmatic dependent launch (next launch resident + tables fetched while the previous one drains) vs plain launches
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
for lib in libastc_b200.so libastc_b200_pdl.so libastc_b200.so libastc_b200_pdl.so; do ASTC_B200_LIB=astc_encoder_b200/$lib python tools/back_to_back.py 2>&1 | tail -9; done | tee $O/back_to_back.txt
for lib in libastc_b200.so libastc_b200_pdl.so; do echo "== $lib"; ASTC_B200_LIB=astc_encoder_b200/$lib python tools/small_sizes.py 2>&1 | tail -10; done | tee $O/small_sizes.txt
ASTC_B200_LIB=astc_encoder_b200/libastc_b200_pdl.so timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_build.py > $O/pytest_pdl.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_pdl.txt
tail -4 $O/pytest_pdl.txt
