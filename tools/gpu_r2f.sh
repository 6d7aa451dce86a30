#!/bin/bash
# round 2, call F: full GPU suite, full bench line (config5 + host batch + small-texture e2e), axis heuristic comparison
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
free -g > $O/host.txt; nproc >> $O/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 300 python tools/axis_compare.py > $O/axis_compare.txt 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
tail -n 6 $O/pytest_gpu.txt; cat $O/axis_compare.txt; tail -n 5 $O/bench.err; cat $O/host.txt
