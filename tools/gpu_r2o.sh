#!/bin/bash
# round 2o (2 GPUs): the full GPU suite incl. the multi-GPU parity tests, and config 5 on 2 GPUs with the tapered schedule
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
tail -n 5 $O/pytest_gpu.txt; tail -n 5 $O/bench_2gpu.err; cat $O/bench_2gpu.json
