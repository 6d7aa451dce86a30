#!/bin/bash
# round 2p (8 GPUs): config 5 strong scaling on 8 and 4 GPUs with the tapered schedule
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
nvidia-smi -L > $O/gpus.txt
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err; echo "rc=$?" >> $O/bench_${n}gpu.err
done
tail -n 2 $O/*.err; cat $O/bench_8gpu.json
