#!/bin/bash
# Short GPU session: parity tests + kernel timings (+ optional ncu of the 4x4 kernel with NCU=1).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/quick_bench.py 2>&1 | tail -12
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:encode4x4 -s 2 -c 1 -f -o gpurun_out/prof_4x4rgb16k \
    python tools/profile_target.py 4x4rgb16k 3 > gpurun_out/ncu_4x4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode6x6 -s 2 -c 1 -f -o gpurun_out/prof_6x6rgba8k \
    python tools/profile_target.py 6x6rgba8k 3 > gpurun_out/ncu_6x6.log 2>&1
fi
