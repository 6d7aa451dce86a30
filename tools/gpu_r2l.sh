#!/bin/bash
# round 2l: staged pipeline for pageable host memory (worker threads -> pinned slots), full GPU suite incl. the
# from-source build test, pageable-vs-pinned probe
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
nproc > $O/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> $O/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
python tools/pageable_probe.py > $O/pageable_probe.txt 2>&1
cat $O/pageable_probe.txt
