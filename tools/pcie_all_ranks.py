"""What the HOST side of a box can feed N GPUs at once: every rank copies the traffic of one
astc_b200_encode_host call of the 16384^2 config (1 GiB H2D, 256 MiB D2H, pinned buffers, both
directions concurrently) with no kernel at all, first one rank at a time and then all ranks together.
The end-to-end (`e2e`) line of bench.py at N GPUs cannot beat the aggregate figure printed here.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_all_ranks.py
"""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)                               # NCCL prints its banner on stdout
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
    n_in, n_out = 1 << 30, 1 << 28
    h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n_out, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()

    def copies():
        d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(side):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)

    def timed(iters=5):
        copies()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            copies()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    def gather(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(v.item()) for v in out]

    # one rank at a time
    solo = 0.0
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            solo = timed()
    solos = gather(solo)
    # all ranks together
    if world > 1:
        dist.barrier()
    together = timed()
    alls = gather(together)
    if rank == 0:
        gb = (n_in + n_out) / 1e6
        print(f"traffic per rank per call: {n_in >> 20} MiB H2D + {n_out >> 20} MiB D2H, concurrent, pinned host memory, no kernel")
        print("one rank at a time   ms: " + " ".join(f"{v:.2f}" for v in solos) + f"   -> {gb / max(solos):.1f} .. {gb / min(solos):.1f} GB/s per rank")
        print(f"all {world} ranks together ms: " + " ".join(f"{v:.2f}" for v in alls)
              + f"   -> slowest {max(alls):.2f} ms, aggregate {world * gb / max(alls):.1f} GB/s"
              + f"  = at most {world * 16384 * 16384 / max(alls) / 1e6:.1f} Gtexel/s end to end (16384^2 4x4 per rank)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
