#!/bin/bash
# round 2, call B: tile-queue (CLC) kernels -- parity, A/B against one-tile-per-CTA, racecheck of the new handshake
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 200 python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
ASTC_B200_LIB=$PWD/astc_encoder_b200/libastc_b200_hooks.so timeout 600 python tools/small_launch.py > $O/small_launch.txt 2>&1; echo "rc=$?" >> $O/small_launch.txt
timeout 300 python tools/quick_bench.py short > $O/quick_bench.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > $O/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.txt
tail -n 4 $O/pytest_gpu.txt $O/smoke.txt $O/sanitizer_racecheck.txt $O/sanitizer_memcheck.txt
cat $O/small_launch.txt $O/quick_bench.txt
