"""astc_b200_encode_host from PAGEABLE host memory (numpy / malloc -- what the reference's caller holds) against
pinned memory, by texture size.  ASTC_B200_LIB selects the build.   python tools/pageable_probe.py   (under gpurun)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth


def rate(fn, iters):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    opt = A.encode_option()
    print(A.lib()._name)
    sizes = (512, 1024, 2048, 4096, 8192, 16384)
    if len(sys.argv) > 1 and sys.argv[1] == "pinned-first":
        # pinned memory only, before any pageable call has started the copy workers
        for size in sizes:
            src = synth.synth_rgba(size, size, 11)
            pin = torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True)
            pin.copy_(src)
            out_pin = torch.empty((A.output_size(size, size, opt) // 16, 16), dtype=torch.uint8, pin_memory=True)
            t = rate(lambda: A.encode_astc_host(pin.numpy(), opt, out=out_pin.numpy()), 100 if size <= 1024 else 20 if size <= 4096 else 5)
            print(f"{size:6d}^2: pinned (no workers yet) {t * 1e6:10.1f} us", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "threads":
        # pageable memory by number of copy workers (besides the caller)
        ctx = A.Context()
        for size in (1024, 4096, 16384):
            page = synth.synth_rgba(size, size, 11).numpy().copy()
            out_page = np.empty((A.output_size(size, size, opt) // 16, 16), np.uint8)
            for threads in (0, 1, 2, 3, 4, 6, 8, 12):
                ctx.set_copy_threads(threads)
                t = rate(lambda: ctx.encode_host(page, opt, out=out_page), 50 if size <= 1024 else 10 if size <= 4096 else 3)
                print(f"{size:6d}^2 pageable, {threads:2d} workers + caller: {t * 1e6:10.1f} us ({size * size / t / 1e9:6.2f} Gtexel/s, {size * size * 4 / t / 1e9:5.1f} GB/s in)", flush=True)
        return
    for size in sizes:
        src = synth.synth_rgba(size, size, 11)
        pin = torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True)
        pin.copy_(src)
        n = A.output_size(size, size, opt) // 16
        out_pin = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
        page = src.numpy().copy()
        out_page = np.empty((n, 16), np.uint8)
        iters = 100 if size <= 1024 else 20 if size <= 4096 else 5
        t_pin = rate(lambda: A.encode_astc_host(pin.numpy(), opt, out=out_pin.numpy()), iters)
        t_page = rate(lambda: A.encode_astc_host(page, opt, out=out_page), iters)
        ok = np.array_equal(out_page, out_pin.numpy())
        print(f"{size:6d}^2: pinned {t_pin * 1e6:10.1f} us ({size * size / t_pin / 1e9:6.2f} Gtexel/s)   pageable {t_page * 1e6:10.1f} us "
              f"({size * size / t_page / 1e9:6.2f} Gtexel/s, {size * size * 4 / t_page / 1e9:5.1f} GB/s in)   same bytes: {ok}", flush=True)


if __name__ == "__main__":
    main()
