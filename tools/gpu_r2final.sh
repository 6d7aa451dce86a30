#!/bin/bash
# round 2 final: the driver's own sequence (GPU suite, smoke, both bench arms) on the committed tree, the ncu launch list
# of the bench command and full captures of the three kernels the summaries under profiles/ quote
mkdir -p gpurun_out/r2final
O=gpurun_out/r2final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
if [ "$1" != "quick" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode4x4 -s 2 -c 1 -f -o $O/prof_4x4rgb16k python tools/profile_target.py 4x4rgb16k 3 > $O/ncu_4x4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode6x6 -s 2 -c 1 -f -o $O/prof_6x6rgba8k python tools/profile_target.py 6x6rgba8k 3 > $O/ncu_6x6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mip_chain_fused -c 1 -f -o $O/prof_mipchain16k python tools/mip_time.py > $O/ncu_mip.log 2>&1
fi
tail -n 4 $O/pytest_gpu.txt; cat $O/smoke.txt; tail -n 2 $O/bench.err; cat $O/bench_ref.json | cut -c1-300
