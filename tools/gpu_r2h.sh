#!/bin/bash
# round 2h: wide-table 4x4 kernel (ASTC_WIDE_LUT) against the committed kernel, same box
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
python tools/variants.py run base wide base wide > $O/ab_wide.txt 2>&1
ASTC_B200_LIB=astc_encoder_b200/libastc_b200_wide.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q > $O/pytest_wide.txt 2>&1; echo "rc=$?" >> $O/pytest_wide.txt
cat $O/ab_wide.txt; tail -5 $O/pytest_wide.txt
