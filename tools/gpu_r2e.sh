#!/bin/bash
# round 2, call E (8 GPUs): config 5 strong scaling + the 8-rank bare-copy ceiling next to the e2e number
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m >> $O/gpus.txt 2>&1
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err; echo "rc=$?" >> $O/bench_${n}gpu.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 tools/pcie_all_ranks.py > $O/pcie_8ranks.txt 2>&1
tail -n 3 $O/*.err; cat $O/bench_8gpu.json; cat $O/pcie_8ranks.txt | tail -20
