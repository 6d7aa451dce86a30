#!/usr/bin/env python3
"""Capture MUFU.RCP / MUFU.RSQ from the GPU as delta tables for the CPU oracle (run on a GPU box).

For every mantissa m (and, for rsq, both parities of the exponent) the table holds
    bits(device result) - bits(base),   base = float32(1 / x) resp. float32(1 / sqrt(x)) evaluated in IEEE double
as an int8.  Both functions scale exactly with the exponent (checked here on random inputs over the
whole normal range), so the tables cover every positive normal float.  Output (xz-compressed int8):
    tests/golden/mufu_rcp.i8.xz   2^23 entries      tests/golden/mufu_rsq.i8.xz   2^24 entries
"""
import lzma
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import astc_encoder_b200 as A  # noqa: E402

OUT = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "tests" / "golden"


def base_rcp(x):
    return (1.0 / x.astype(np.float64)).astype(np.float32)


def base_rsq(x):
    return (1.0 / np.sqrt(x.astype(np.float64))).astype(np.float32)


def device(op, x):
    return A.mufu(op, torch.from_numpy(x).cuda()).cpu().numpy()


def main():
    m = np.arange(1 << 23, dtype=np.uint32)
    x = (np.uint32(0x3F800000) + m).view(np.float32)                        # [1, 2)
    d_rcp = device("rcp", x).view(np.int32).astype(np.int64) - base_rcp(x).view(np.int32).astype(np.int64)
    x2 = (np.uint32(0x3F800000) + np.arange(1 << 24, dtype=np.uint32)).view(np.float32)   # [1, 4): parity 0 then 1
    d_rsq = device("rsq", x2).view(np.int32).astype(np.int64) - base_rsq(x2).view(np.int32).astype(np.int64)
    for name, d in (("rcp", d_rcp), ("rsq", d_rsq)):
        print(name, "delta range", int(d.min()), int(d.max()), "histogram", {int(v): int(c) for v, c in zip(*np.unique(d, return_counts=True))})
        assert -127 <= d.min() and d.max() <= 127
        blob = lzma.compress(d.astype(np.int8).tobytes(), preset=9 | lzma.PRESET_EXTREME)
        (OUT / f"mufu_{name}.i8.xz").write_bytes(blob)
        print(name, "compressed", len(blob), "bytes")
    # scale invariance over the normal range, on 2^26 random positive normal floats
    rng = np.random.default_rng(7)
    bits = rng.integers(0x00800000, 0x7F800000, 1 << 26, dtype=np.uint32)
    xr = bits.view(np.float32)
    mant, expo = bits & 0x7FFFFF, bits >> 23
    got = device("rcp", xr).view(np.int32)
    want = base_rcp(xr).view(np.int32) + d_rcp[mant].astype(np.int32)
    ok = (xr < 2.0 ** 126) & (xr > 2.0 ** -126)                             # results that stay normal (ftz outside)
    print("rcp model mismatches:", int(((got != want) & ok).sum()), "of", int(ok.sum()))
    got = device("rsq", xr).view(np.int32)
    par = (expo + 1) & 1                                                    # exponent 127 (x in [1,2)) -> parity 0
    want = base_rsq(xr).view(np.int32) + d_rsq[(par.astype(np.int64) << 23) | mant].astype(np.int32)
    print("rsq model mismatches:", int((got != want).sum()), "of", len(xr))


if __name__ == "__main__":
    main()
