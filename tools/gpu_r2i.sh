#!/bin/bash
# round 2i: parity + timings after the quad-friendly 6x6 parking store; fresh full captures of both 6x6 kernels
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
python tools/quick_bench.py short > $O/quick_bench.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode6x6 -s 2 -c 1 -f -o $O/prof_6x6rgb8k python tools/profile_target.py 6x6rgb8k 3 > $O/ncu_6x6rgb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode6x6 -s 2 -c 1 -f -o $O/prof_6x6rgba8k python tools/profile_target.py 6x6rgba8k 3 > $O/ncu_6x6rgba.log 2>&1
tail -3 $O/pytest_gpu.txt; cat $O/quick_bench.txt
