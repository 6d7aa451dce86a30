"""Small-launch loss of the 4x4 kernel: time per launch vs. blocks per thread (passes) for the two
job sizes that matter for strong scaling -- 4096^2 (BASELINE config 2) and the 16384x2048 band one
GPU gets when a 16384^2 texture is cut eight ways.  Needs an experiment build (-DASTC_TUNING_HOOKS: reads ASTC_B200_PASSES / ASTC_B200_DYNAMIC per launch):
    python tools/variants.py build hooks=ASTC_TUNING_HOOKS ; ASTC_B200_LIB=astc_encoder_b200/libastc_b200_hooks.so python tools/small_launch.py
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


def time_rotating(imgs, opt, rounds=6):
    """Back-to-back launches over inputs that together exceed the L2: per-launch time with event
    resolution amortised over len(imgs) launches (includes the gap between launches)."""
    outs = [A.encode_astc(i, opt) for i in imgs]
    torch.cuda.synchronize()
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i, o in zip(imgs, outs):
            A.encode_astc(i, opt, out=o)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / len(imgs))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = [("4096x4096 4x4 rgb", 4096, 4096, A.encode_option()),
             ("16384x2048 4x4 rgb (1/8 band)", 16384, 2048, A.encode_option()),
             ("16384x4096 4x4 rgb (1/4 band)", 16384, 4096, A.encode_option()),
             ("8192x1368 6x6 rgba srgb (1/6 band)", 8192, 1368, A.encode_option(is6x6=True, has_alpha=True, srgb=True)),
             ("2048x2048 4x4 rgb", 2048, 2048, A.encode_option()),
             ("1024x1024 4x4 rgb", 1024, 1024, A.encode_option())]
    big = synth.synth_rgba(16384, 16384, synth.SEED_CFG5, device="cuda")
    ref_ms = 1e9
    for dyn in ("1", "0"):
        os.environ["ASTC_B200_DYNAMIC"] = dyn
        for passes in ("", "1", "2", "4", "8"):
            if passes:
                os.environ["ASTC_B200_PASSES"] = passes
            else:
                os.environ.pop("ASTC_B200_PASSES", None)
            ms, _ = time_one(big, A.encode_option(), iters=10)
            ref_ms = min(ref_ms, ms)
            print(f"16384^2 4x4 rgb dynamic={dyn} passes={passes or 'auto'}: {ms:.4f} ms -> {16384 * 16384 / ms / 1e6:.1f} Gtexel/s", flush=True)
    os.environ.pop("ASTC_B200_PASSES", None)
    os.environ.pop("ASTC_B200_DYNAMIC", None)
    rate = 16384 * 16384 / ref_ms                      # texels per ms at full size
    for name, w, h, opt in cases:
        img = synth.synth_rgba(w, h, synth.SEED_CFG2, device="cuda")
        copies = max(2, min(16, (600 << 20) // (w * h * 4)))
        imgs = [img] + [img.clone() for _ in range(copies - 1)]
        ideal = w * h / rate * 1e3 if opt.is4x4 and not opt.is6x6 else float("nan")
        print(f"{name}:  (ideal at the 16384^2 rate {ideal:.1f} us)", flush=True)
        for dyn in ("1", "0"):
            os.environ["ASTC_B200_DYNAMIC"] = dyn
            row = []
            for passes in ("", "1", "2", "3", "4", "8"):
                if passes:
                    os.environ["ASTC_B200_PASSES"] = passes
                else:
                    os.environ.pop("ASTC_B200_PASSES", None)
                med, best = time_one(img, opt, flush=flush)
                row.append(f"{passes or 'auto'}:{med * 1e3:.1f}/{best * 1e3:.1f}")
            os.environ.pop("ASTC_B200_PASSES", None)
            warm, _ = time_one(img, opt, flush=None)
            rot = []
            for passes in ("", "1", "2", "4"):
                if passes:
                    os.environ["ASTC_B200_PASSES"] = passes
                else:
                    os.environ.pop("ASTC_B200_PASSES", None)
                rot.append(f"{passes or 'auto'}:{time_rotating(imgs, opt) * 1e3:.2f}")
            os.environ.pop("ASTC_B200_PASSES", None)
            print(f"    {'tile queue (CLC)' if dyn == '1' else 'one tile per CTA '}: us median/best by passes  {'  '.join(row)}  | no-flush auto {warm * 1e3:.1f}"
                  f"  | {copies} rotating inputs back to back: {'  '.join(rot)}", flush=True)
        os.environ.pop("ASTC_B200_DYNAMIC", None)


if __name__ == "__main__":
    main()
