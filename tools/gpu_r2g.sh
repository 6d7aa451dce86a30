#!/bin/bash
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 600 python -m pytest tests/test_gpu_host.py tests/test_gpu_edges.py -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 300 python tools/host_path_probe.py > $O/host_path_probe.txt 2>&1
timeout 900 python bench.py --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
tail -n 4 $O/pytest_gpu.txt; cat $O/host_path_probe.txt; tail -n 3 $O/bench.err
