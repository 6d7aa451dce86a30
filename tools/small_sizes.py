"""Per-launch time of small jobs (L2 flushed between launches, median / best of 30): the sizes one GPU sees
when a large texture is cut eight ways or a mip chain is encoded level by level.
    [ASTC_B200_LIB=...] python tools/small_sizes.py        (under gpurun)"""
import sys

import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth


def time_one(img, opt, flush, iters=30):
    out = A.encode_astc(img, opt)
    for _ in range(3):
        A.encode_astc(img, opt, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        A.encode_astc(img, opt, out=out)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rgb, rgba6 = A.encode_option(), A.encode_option(is6x6=True, has_alpha=True, srgb=True)
    cases = [("256^2 4x4 rgb", 256, 256, rgb), ("512^2 4x4 rgb", 512, 512, rgb), ("1024^2 4x4 rgb", 1024, 1024, rgb),
             ("2048^2 4x4 rgb", 2048, 2048, rgb), ("4096^2 4x4 rgb", 4096, 4096, rgb), ("16384x2048 4x4 rgb (1/8 band)", 16384, 2048, rgb),
             ("16384x4096 4x4 rgb (1/4 band)", 16384, 4096, rgb), ("1024^2 6x6 rgba srgb", 1024, 1024, rgba6),
             ("4096^2 6x6 rgba srgb", 4096, 4096, rgba6), ("8192x1368 6x6 rgba srgb (1/6 band)", 8192, 1368, rgba6)]
    for name, w, h, opt in cases:
        img = synth.synth_rgba(w, h, synth.SEED_CFG2, device="cuda")
        med, best = time_one(img, opt, flush)
        print(f"{name:38s} median {med:7.1f} us  best {best:7.1f} us  -> {w * h / med / 1e3:7.1f} Gtexel/s", flush=True)


if __name__ == "__main__":
    main()
