#!/usr/bin/env python3
"""How much rides on the modelling choices no reference fixture pins?

The canonical arithmetic is pinned by the reference's one golden vector for RGBA 4x4
linear only.  For the other variants several choices are free: x RN(1/n) vs true division for the
mean and the covariance scale (ASTC_Encode.hlsl:147,162), the FMA shape of sample_texel (:307-314),
how the sRGB decode of the texture unit is rounded (main.cpp:38), MUFU vs correctly rounded rcp / rsq
(:103,332,366), fused vs unfused texel*255 - base.  This script re-encodes the BASELINE configs with
the oracle under each alternative (oracle switches, astc_oracle_set_variant) and reports the share of
blocks that change and the decoded-PSNR difference, so every "parity unpinned" statement carries a number.

    python tools/pin_sensitivity.py [--quick] [--json profiles/r2_pin_sensitivity.json]
CPU only (oracle + numpy); ~2 minutes on 8 cores at full size.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O                                    # noqa: E402


def leaf():
    """tests/golden/leaf.png as the reference loads it (RGBA8, flipped); read with PIL so that this tool
    needs neither the CUDA library nor a GPU."""
    from PIL import Image
    img = np.asarray(Image.open(ROOT / "tests" / "golden" / "leaf.png").convert("RGBA"))
    return np.ascontiguousarray(img[::-1])


def as_encoded(img, srgb, normal):
    """The image as the encoder sees it (texel * 255 after the texture-unit conversion): the reference
    PSNR is measured against.  For -srgb that is the LINEARISED colour, not the sRGB byte."""
    lut = O.unorm_lut(bool(srgb)).astype(np.float64) * 255.0
    out = img.astype(np.float64)
    if srgb and not normal:
        out[..., :3] = lut[img[..., :3]]
    return out


def psnr(dec, seen, channels):
    d = dec[..., :channels].astype(np.float64) - seen[..., :channels]
    mse = (d * d).reshape(-1, channels).mean(axis=0)
    with np.errstate(divide="ignore"):
        return 10.0 * np.log10(255.0 * 255.0 / mse)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="1024^2 crops instead of the full BASELINE sizes")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import torch  # noqa: F401  (synth uses torch on the CPU)
    from astc_encoder_b200 import synth
    O.lib()
    s = 1024 if args.quick else None
    cfgs = [
        ("cfg2 4096^2 4x4 RGB linear", synth.synth_rgba(s or 4096, s or 4096, synth.SEED_CFG2).numpy(), dict(block_dim=4), 3),
        ("cfg3 8192^2 6x6 -alpha -srgb", synth.synth_rgba(s or 8192, s or 8192, synth.SEED_CFG3).numpy(),
         dict(block_dim=6, has_alpha=True, srgb=True), 4),
        ("cfg4 4096^2 4x4 -norm", synth.synth_normal(s or 4096, s or 4096, synth.SEED_CFG4).numpy(), dict(block_dim=4, is_normal_map=True), 2),
        ("leaf 1024^2 4x4 -alpha -srgb (README example)", leaf(), dict(block_dim=4, has_alpha=True, srgb=True), 4),
        ("leaf 1024^2 6x6 -alpha", leaf(), dict(block_dim=6, has_alpha=True), 4),
    ]
    variants = [
        ("mean, cov: true division instead of x RN(1/n)", O.VAR_TRUE_DIVISION),
        ("6x6 sample_texel: unfused", O.VAR_UNFUSED_SAMPLE),
        ("sRGB decode evaluated in float (powf)", O.VAR_SRGB_POWF),
        ("rcp / rsq correctly rounded instead of MUFU", O.VAR_EXACT_RCP_RSQ),
        ("texel*255 - base: unfused", O.VAR_UNFUSED_DEV),
        ("all of the above together", O.VAR_TRUE_DIVISION | O.VAR_UNFUSED_SAMPLE | O.VAR_SRGB_POWF | O.VAR_EXACT_RCP_RSQ | O.VAR_UNFUSED_DEV),
    ]
    rows = []
    for name, img, kw, nch in cfgs:
        dim = kw["block_dim"]
        h, w = img.shape[:2]
        O.set_variant(0)
        base = O.encode_image(img, **kw)
        dec0, bad = O.decode_image(base, w, h, dim)
        assert bad == 0
        seen = as_encoded(img, kw.get("srgb"), kw.get("is_normal_map"))
        nch = 2 if kw.get("is_normal_map") else nch              # only r, g are kept by a normal-map encode
        p0 = psnr(dec0, seen, nch)
        print(f"\n{name}: {len(base)} blocks, decoded PSNR {' / '.join(f'{v:.3f}' for v in p0)} dB", flush=True)
        for vname, flags in variants:
            applies = True
            if flags == O.VAR_UNFUSED_SAMPLE and dim != 6:
                applies = False
            if flags == O.VAR_SRGB_POWF and not kw.get("srgb"):
                applies = False
            if flags == O.VAR_TRUE_DIVISION and dim == 4:
                pass                                             # 1/16 is exact, 1/15 is not: still applies
            if not applies:
                continue
            O.set_variant(flags)
            alt = O.encode_image(img, **kw)
            O.set_variant(0)
            changed = int((alt != base).any(axis=1).sum())
            dec1, _ = O.decode_image(alt, w, h, dim)
            p1 = psnr(dec1, seen, nch)
            dpsnr = float(np.max(np.abs(p1 - p0)))
            rows.append({"config": name, "alternative": vname, "blocks": int(len(base)), "blocks_changed": changed,
                         "pct_changed": round(100.0 * changed / len(base), 4), "max_abs_psnr_delta_db": round(dpsnr, 5)})
            print(f"  {vname:52s} {changed:9d} blocks change = {100.0 * changed / len(base):7.4f} %   max |dPSNR| {dpsnr:.5f} dB", flush=True)
    O.set_variant(0)
    if args.json:
        Path(args.json).write_text(json.dumps({"tool": "tools/pin_sensitivity.py", "quick": args.quick, "rows": rows}, indent=1) + "\n")
    worst = max(r["max_abs_psnr_delta_db"] for r in rows)
    print(f"\nlargest PSNR movement under any alternative: {worst:.5f} dB (north_star's bar: 0.05 dB)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
