#!/bin/bash
# round 2t: 6x6 -- next block's rows loaded into registers under the pack tail (2 / 4 / 8 passes) vs the committed kernel
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
python tools/variants.py run base regpf regpf4 regpf8 base regpf regpf4 regpf8 > $O/ab_regpf.txt 2>&1
grep -E "==|6x6" $O/ab_regpf.txt
