#!/bin/bash
# round 2, call A: CLC microbenchmark, small-launch sweep, ncu of the 4096^2 launch, compute-sanitizer evidence
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
( cd tools/microbench && timeout 120 ./clc ) > $O/clc.txt 2>&1; echo "clc rc=$?" >> $O/clc.txt
timeout 300 python tools/small_launch.py > $O/small_launch.txt 2>&1; echo "rc=$?" >> $O/small_launch.txt
timeout 300 python tools/quick_bench.py short > $O/quick_bench.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode4x4 -s 3 -c 1 -f -o $O/prof_4x4rgb4k \
    python tools/profile_target.py 4x4rgb4k 3 > $O/ncu_4x4rgb4k.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?" >> $O/sanitizer_synccheck.txt
ls -la $O
tail -3 $O/*.txt
