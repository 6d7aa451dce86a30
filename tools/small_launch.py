"""Small-launch loss of the 4x4 kernel: time per launch vs. blocks per thread (passes) for the two
job sizes that matter for strong scaling -- 4096^2 (BASELINE config 2) and the 16384x2048 band one
GPU gets when a 16384^2 texture is cut eight ways.  Experiment builds read ASTC_B200_PASSES per launch.
    python tools/small_launch.py            (under gpurun)
"""
import os
import sys

import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth


def time_one(img, opt, iters=20, flush=None):
    out = A.encode_astc(img, opt)
    for _ in range(3):
        A.encode_astc(img, opt, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        A.encode_astc(img, opt, out=out)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = [("4096x4096 4x4 rgb", 4096, 4096, A.encode_option()),
             ("16384x2048 4x4 rgb (1/8 band)", 16384, 2048, A.encode_option()),
             ("16384x4096 4x4 rgb (1/4 band)", 16384, 4096, A.encode_option()),
             ("8192x1368 6x6 rgba srgb (1/6 band)", 8192, 1368, A.encode_option(is6x6=True, has_alpha=True, srgb=True)),
             ("2048x2048 4x4 rgb", 2048, 2048, A.encode_option()),
             ("1024x1024 4x4 rgb", 1024, 1024, A.encode_option())]
    big = synth.synth_rgba(16384, 16384, synth.SEED_CFG5, device="cuda")
    ref_ms, _ = time_one(big, A.encode_option(), iters=10)
    rate = 16384 * 16384 / ref_ms                      # texels per ms at full size
    print(f"16384^2 4x4 rgb: {ref_ms:.4f} ms -> {rate / 1e6:.1f} Gtexel/s", flush=True)
    for name, w, h, opt in cases:
        img = synth.synth_rgba(w, h, synth.SEED_CFG2, device="cuda")
        row = []
        for passes in ("", "1", "2", "3", "4", "6", "8"):
            if passes:
                os.environ["ASTC_B200_PASSES"] = passes
            else:
                os.environ.pop("ASTC_B200_PASSES", None)
            med, best = time_one(img, opt, flush=flush)
            row.append(f"{passes or 'auto'}:{med * 1e3:.1f}/{best * 1e3:.1f}")
        os.environ.pop("ASTC_B200_PASSES", None)
        warm, _ = time_one(img, opt, flush=None)
        ideal = w * h / rate * 1e3 if opt.is4x4 and not opt.is6x6 else float("nan")
        print(f"{name}: us median/best by passes  {'  '.join(row)}  | no-flush auto {warm * 1e3:.1f}  | ideal at the 16384^2 rate {ideal:.1f}", flush=True)


if __name__ == "__main__":
    main()
