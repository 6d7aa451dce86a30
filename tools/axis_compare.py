"""PCA (what the reference ships) vs max_accumulation_pixel_direction (ASTC_Encode.hlsl:170-227, opt-in
axis_method = 1) on the BASELINE configs: kernel time, throughput and decoded PSNR against the image as
the encoder sees it (texel*255 after the UNORM / sRGB conversion).
    python tools/axis_compare.py      (under gpurun)
"""
import sys

import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth


def timed(img, opt, iters=10):
    out = A.encode_astc(img, opt)
    for _ in range(3):
        A.encode_astc(img, opt, out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        A.encode_astc(img, opt, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], out


def psnr(dec, seen, nch):
    d = dec[..., :nch].double() - seen[..., :nch]
    mse = (d * d).reshape(-1, nch).mean(dim=0)
    return (10.0 * torch.log10(255.0 * 255.0 / mse)).tolist()


def main():
    cases = [
        ("cfg2 4096^2 4x4 RGB linear", synth.synth_rgba(4096, 4096, synth.SEED_CFG2, device="cuda"), dict(), 4, 3),
        ("cfg3 8192^2 6x6 -alpha -srgb", synth.synth_rgba(8192, 8192, synth.SEED_CFG3, device="cuda"), dict(is6x6=True, has_alpha=True, srgb=True), 6, 4),
        ("cfg4 4096^2 4x4 -norm", synth.synth_normal(4096, 4096, synth.SEED_CFG4, device="cuda"), dict(is_normal_map=True), 4, 2),
        ("cfg5 16384^2 4x4 RGB linear", synth.synth_rgba(16384, 16384, synth.SEED_CFG5, device="cuda"), dict(), 4, 3),
        ("leaf.png 4x4 -alpha", torch.from_numpy(A.load_image("tests/golden/leaf.png", True)).cuda(), dict(has_alpha=True), 4, 4),
        ("leaf.png 6x6 -alpha", torch.from_numpy(A.load_image("tests/golden/leaf.png", True)).cuda(), dict(is6x6=True, has_alpha=True), 6, 4),
    ]
    for name, img, kw, dim, nch in cases:
        h, w = int(img.shape[0]), int(img.shape[1])
        seen = img.double()
        if kw.get("srgb"):
            lut = torch.from_numpy(A.unorm_lut(True)).double().cuda() * 255.0
            seen[..., :3] = lut[img[..., :3].long()]
        row = []
        for axis in (0, 1):
            opt = A.encode_option(axis_method=axis, **kw)
            ms, out = timed(img, opt)
            dec = A.decode_astc(out, w, h, dim)
            p = psnr(dec, seen, nch)
            row.append((ms, p))
            print(f"{name:32s} {'PCA power iteration    ' if axis == 0 else 'max accumulation (opt-in)'}: {ms:.4f} ms  {w * h / ms / 1e6:7.1f} Gtexel/s"
                  f"   PSNR {' / '.join(f'{v:.3f}' for v in p)} dB", flush=True)
        print(f"{'':32s} -> {row[0][0] / row[1][0]:.2f}x the speed, PSNR change {' / '.join(f'{b - a:+.3f}' for a, b in zip(row[0][1], row[1][1]))} dB\n", flush=True)


if __name__ == "__main__":
    main()
