#!/bin/bash
# One GPU-box session: tests, smoke, bench, launch list, full ncu capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python tools/quick_bench.py > gpurun_out/quick_bench.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
kill $SMI
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode4x4 -s 2 -c 1 -f -o gpurun_out/prof_4x4rgb16k \
    python tools/profile_target.py 4x4rgb16k 3 > gpurun_out/ncu_4x4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode6x6 -s 2 -c 1 -f -o gpurun_out/prof_6x6rgba8k \
    python tools/profile_target.py 6x6rgba8k 3 > gpurun_out/ncu_6x6.log 2>&1
ls -la gpurun_out
