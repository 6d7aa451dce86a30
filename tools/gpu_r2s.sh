#!/bin/bash
# round 2s: final evidence with the committed kernels (tapered schedule + programmatic dependent launch):
# GPU suite, bench (both arms), ncu launch list of the bench command, full ncu captures, sanitizer runs
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode4x4 -s 2 -c 1 -f -o $O/prof_4x4rgb16k python tools/profile_target.py 4x4rgb16k 3 > $O/ncu_4x4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode6x6 -s 2 -c 1 -f -o $O/prof_6x6rgba8k python tools/profile_target.py 6x6rgba8k 3 > $O/ncu_6x6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode4x4 -s 2 -c 1 -f -o $O/prof_4x4rgb4k python tools/profile_target.py 4x4rgb4k 3 > $O/ncu_4x4_4k.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?" >> $O/sanitizer_synccheck.txt
tail -n 4 $O/pytest_gpu.txt; tail -n 3 $O/bench.err; tail -n 3 $O/sanitizer_*.txt; cat $O/bench_ref.json; ls -la $O
