"""Per-launch time of encode launches issued BACK TO BACK on one stream (what a pipeline of bands, mip levels or
benchmark steps looks like), over rotating inputs that together exceed the L2; two events around the whole run, so the
timer's resolution is amortised.   [ASTC_B200_LIB=...] python tools/back_to_back.py      (under gpurun)"""
import sys

import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth


def per_launch_us(imgs, opt, rounds=7, reps=4):
    outs = [A.encode_astc(i, opt) for i in imgs]
    torch.cuda.synchronize()
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            for i, o in zip(imgs, outs):
                A.encode_astc(i, opt, out=o)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / (len(imgs) * reps))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    rgb, rgba6 = A.encode_option(), A.encode_option(is6x6=True, has_alpha=True, srgb=True)
    cases = [("1024^2 4x4 rgb", 1024, 1024, rgb), ("2048^2 4x4 rgb", 2048, 2048, rgb), ("4096^2 4x4 rgb", 4096, 4096, rgb),
             ("16384x2048 4x4 rgb (1/8 band)", 16384, 2048, rgb), ("16384x4096 4x4 rgb (1/4 band)", 16384, 4096, rgb),
             ("16384^2 4x4 rgb", 16384, 16384, rgb), ("4096^2 6x6 rgba srgb", 4096, 4096, rgba6), ("8192^2 6x6 rgba srgb", 8192, 8192, rgba6)]
    print(A.lib()._name)
    for name, w, h, opt in cases:
        img = synth.synth_rgba(w, h, synth.SEED_CFG2, device="cuda")
        copies = max(2, min(16, (700 << 20) // (w * h * 4)))
        imgs = [img] + [img.clone() for _ in range(copies - 1)]
        med, best = per_launch_us(imgs, opt)
        print(f"{name:34s} {copies:2d} inputs  median {med:8.2f} us  best {best:8.2f} us per launch -> {w * h / med / 1e3:7.1f} Gtexel/s", flush=True)


if __name__ == "__main__":
    main()
