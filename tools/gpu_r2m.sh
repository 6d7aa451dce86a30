#!/bin/bash
# round 2m: batch host path with per-group pinned slots (pageable levels through the copy workers); full GPU suite; bench
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
python tools/pageable_probe.py > $O/pageable_probe.txt 2>&1; cat $O/pageable_probe.txt
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 3000 $O/bench.json; tail -5 $O/bench.err
