"""Per-call cost of the host-buffer path (astc_b200_encode_host on its persistent context) by texture size,
pinned host memory; and the raw link for the same bytes.   python tools/host_path_probe.py   (under gpurun)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth


def main():
    opt = A.encode_option()
    for size in (4, 64, 256, 512, 1024, 2048, 4096, 8192):
        src = synth.synth_rgba(size, size, 11)
        pin = torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True)
        pin.copy_(src)
        out = torch.empty((A.output_size(size, size, opt) // 16, 16), dtype=torch.uint8, pin_memory=True)
        d = torch.empty_like(pin, device="cuda")
        a, o = pin.numpy(), out.numpy()
        for _ in range(5):
            A.encode_astc_host(a, opt, out=o)
        iters = 300 if size <= 1024 else 30
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            A.encode_astc_host(a, opt, out=o)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        med = ts[len(ts) // 2]
        # the raw link: one flat H2D of the same bytes + sync
        cs = []
        for _ in range(iters):
            t0 = time.perf_counter()
            d.copy_(pin, non_blocking=True)
            torch.cuda.synchronize()
            cs.append(time.perf_counter() - t0)
        cs.sort()
        print(f"{size:5d}^2: encode_host {med * 1e6:9.1f} us/call ({size * size / med / 1e9:6.2f} Gtexel/s, {size * size * 4 / med / 1e9:5.1f} GB/s in)"
              f"   bare H2D of the same bytes + sync {cs[len(cs) // 2] * 1e6:9.1f} us", flush=True)


if __name__ == "__main__":
    main()
