#!/bin/bash
# round 2k: sRGB conversion variants (packed rounded-product sums; alpha through its own UNORM table in the 6x6 kernel)
# vs the committed arithmetic ("old"), same box; parity of the candidate against the oracle on the -srgb tests
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
python tools/variants.py run old psum alut both old both > $O/ab_srgb.txt 2>&1
cat $O/ab_srgb.txt | grep -E "==|srgb"
ASTC_B200_LIB=astc_encoder_b200/libastc_b200_both.so timeout 900 python -m pytest tests -m gpu -x -q -k "srgb or parity or edges" > $O/pytest_both.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_both.txt
tail -4 $O/pytest_both.txt
