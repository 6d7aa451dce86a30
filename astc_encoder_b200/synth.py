"""Deterministic synthetic textures for the BASELINE.json configs.

Generated with torch ops on whichever device is asked for (CPU here, the GPU in
bench.py); the same bytes are then fed to the CUDA path and to the oracle, so
cross-device libm differences never matter for parity.

synth_rgba:   per channel a low-frequency sinusoidal gradient (so blocks have a
              real principal axis, alpha included: RGB mode still runs a 4-D
              PCA, ASTC_Encode.hlsl:141-166) plus +-16 of hashed noise.
synth_normal: tangent-space normal map of a hashed value-noise height field,
              (n.xy*0.5+0.5) in R,G, n.z in B, 255 in A.
"""
from __future__ import annotations

import math

import torch

SEED_CFG2 = 0xA57C0001   # 4096^2, 4x4, RGB linear
SEED_CFG3 = 0xA57C0002   # 8192^2, 6x6, -alpha -srgb
SEED_CFG4 = 0xA57C0003   # 4096^2, -norm -4x4
SEED_CFG5 = 0xA57C0004   # 16384^2, 4x4
SEED_BATCH = 0xA57C1000  # + texture index


def _i64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def _splitmix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic)."""
    x = x + _i64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def synth_rgba(width: int, height: int, seed: int, device="cpu", rows_per_chunk: int = 1024,
               row0: int = 0, rows: int | None = None) -> torch.Tensor:
    """(H, W, 4) uint8.  `row0` / `rows` produce only that slab of texel rows of the same
    width x height image (a rank's band of a sharded texture): texels depend on their absolute position."""
    if rows is None:
        row0, rows = 0, height
    out = torch.empty((rows, width, 4), dtype=torch.uint8, device=device)
    for y0 in range(row0, row0 + rows, rows_per_chunk):
        n = min(rows_per_chunk, row0 + rows - y0)
        out[y0 - row0:y0 - row0 + n] = _synth_rows(width, height, seed, device, y0, y0 + n)
    return out


def _synth_rows(width: int, height: int, seed: int, device, y0: int, y1: int) -> torch.Tensor:
    """Texel rows [y0, y1) of synth_rgba(width, height, seed)."""
    out = torch.empty((y1 - y0, width, 4), dtype=torch.uint8, device=device)
    xs = torch.arange(width, device=device, dtype=torch.float32)
    xi = torch.arange(width, device=device, dtype=torch.int64)
    rng = (seed * 2654435761) & 0xFFFFFFFF
    ys = torch.arange(y0, y1, device=device, dtype=torch.float32)[:, None]
    yi = torch.arange(y0, y1, device=device, dtype=torch.int64)[:, None]
    lin = (yi * width + xi[None, :]) * 4
    for c in range(4):
        fx = float(1 + ((rng >> (4 * c)) & 3))
        fy = float(1 + ((rng >> (4 * c + 2)) & 3))
        phase = 2.0 * math.pi * (((rng >> (16 + 3 * c)) & 7) / 8.0)
        base = 128.0 + 96.0 * torch.sin((2.0 * math.pi / max(width, 1)) * (xs[None, :] * fx + ys * fy) + phase)
        h = _splitmix64((lin + c) ^ _i64(seed))
        noise = (_lsr(h, 40) & 0xFFFF).to(torch.float32) * (32.0 / 65535.0) - 16.0
        out[:, :, c] = torch.clamp(torch.round(base + noise), 0, 255).to(torch.uint8)
    return out


def _value_noise(width: int, height: int, cell: int, seed: int, device) -> torch.Tensor:
    gx, gy = width // cell + 2, height // cell + 2
    idx = torch.arange(gx * gy, device=device, dtype=torch.int64)
    lattice = (_lsr(_splitmix64(idx ^ _i64(seed)), 40) & 0xFFFF).to(torch.float32).reshape(gy, gx) / 65535.0
    x = torch.arange(width, device=device, dtype=torch.float32) / cell
    y = torch.arange(height, device=device, dtype=torch.float32) / cell
    x0, y0 = x.floor().long(), y.floor().long()
    fx, fy = (x - x0)[None, :], (y - y0)[:, None]
    fx, fy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a = lattice[y0][:, x0]
    b = lattice[y0][:, x0 + 1]
    c = lattice[y0 + 1][:, x0]
    d = lattice[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def synth_normal(width: int, height: int, seed: int, device="cpu") -> torch.Tensor:
    """(H, W, 4) uint8 tangent-space normal map."""
    hgt = torch.zeros((height, width), dtype=torch.float32, device=device)
    for octave, cell in enumerate((64, 16, 4)):
        hgt += _value_noise(width, height, cell, seed + octave, device) * (cell / 8.0)
    dx = torch.zeros_like(hgt)
    dy = torch.zeros_like(hgt)
    dx[:, 1:-1] = (hgt[:, 2:] - hgt[:, :-2]) * 0.5
    dy[1:-1, :] = (hgt[2:, :] - hgt[:-2, :]) * 0.5
    inv = torch.rsqrt(dx * dx + dy * dy + 1.0)
    out = torch.empty((height, width, 4), dtype=torch.uint8, device=device)
    out[..., 0] = torch.clamp(torch.round((-dx * inv * 0.5 + 0.5) * 255.0), 0, 255).to(torch.uint8)
    out[..., 1] = torch.clamp(torch.round((-dy * inv * 0.5 + 0.5) * 255.0), 0, 255).to(torch.uint8)
    out[..., 2] = torch.clamp(torch.round(inv * 255.0), 0, 255).to(torch.uint8)
    out[..., 3] = 255
    return out


def mip_chain(base: torch.Tensor) -> list[torch.Tensor]:
    """Full chain down to 1x1 by 2x2 box filter with round-half-up (uint8)."""
    chain = [base]
    cur = base
    while cur.shape[0] > 1 or cur.shape[1] > 1:
        h, w = cur.shape[0], cur.shape[1]
        nh, nw = max(1, h // 2), max(1, w // 2)
        c = cur.to(torch.int32)
        ys = [0, 1] if h > 1 else [0, 0]
        xs = [0, 1] if w > 1 else [0, 0]
        acc = sum(c[ys[j]::2 if h > 1 else 1, xs[i]::2 if w > 1 else 1][:nh, :nw] for j in range(2) for i in range(2))
        cur = ((acc + 2) >> 2).to(torch.uint8).contiguous()
        chain.append(cur)
    return chain
