"""ctypes binding of the CPU oracle (oracle/astc_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of astc_oracle.h.  Importable only
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libastc_oracle.so"


class OracleOpt(C.Structure):
    _fields_ = [("block_dim", C.c_int), ("has_alpha", C.c_int),
                ("is_normal_map", C.c_int), ("srgb", C.c_int), ("axis_method", C.c_int)]


class OracleTrace(C.Structure):
    _fields_ = [("mean", C.c_float * 4), ("cov", C.c_float * 16), ("axis", C.c_float * 4),
                ("e0", C.c_float * 4), ("e1", C.c_float * 4), ("ep", C.c_uint8 * 8),
                ("projw", C.c_float * 16), ("q", C.c_uint8 * 16), ("qs", C.c_uint8 * 16)]


class OracleSymbolic(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("partitions", C.c_uint32), ("cem", C.c_uint32),
                ("weight_quant", C.c_uint32), ("grid_w", C.c_uint32), ("grid_h", C.c_uint32),
                ("ep", C.c_uint8 * 8), ("weights", C.c_uint8 * 64), ("weights_unq", C.c_uint8 * 64)]


def build(force: bool = False) -> Path:
    """Compile the oracle with the frozen float flags (no contraction)."""
    src = _HERE / "astc_oracle.c"
    hdr = _HERE / "astc_oracle.h"
    if (not force and _SO.exists()
            and _SO.stat().st_mtime >= max(src.stat().st_mtime, hdr.stat().st_mtime)):
        return _SO
    _SO.parent.mkdir(exist_ok=True)
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    base = [cc, "-O2", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
            "-shared", "-o", str(_SO), str(src), "-lm"]
    try:
        subprocess.run(base[:6] + ["-fopenmp"] + base[6:], check=True, capture_output=True)
    except subprocess.CalledProcessError:
        subprocess.run(base, check=True)      # single-threaded oracle
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        u8p = C.POINTER(C.c_uint8)
        L.astc_oracle_unorm_lut.argtypes = [C.c_int, C.POINTER(C.c_float)]
        L.astc_oracle_encode_block.argtypes = [C.c_void_p, C.POINTER(OracleOpt), u8p,
                                               C.POINTER(OracleTrace)]
        L.astc_oracle_encode_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                               C.POINTER(OracleOpt), C.c_void_p, C.c_int]
        L.astc_oracle_encode_image.restype = C.c_int
        L.astc_oracle_encode_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                              C.POINTER(OracleOpt), C.c_int, C.c_int,
                                              C.c_void_p, C.c_int]
        L.astc_oracle_encode_rows.restype = C.c_int
        L.astc_oracle_quant_layout.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 3
        L.astc_oracle_ise_bitcount.argtypes = [C.c_uint32, C.c_int]
        L.astc_oracle_ise_bitcount.restype = C.c_uint32
        L.astc_oracle_bise_encode.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.astc_oracle_bise_encode.restype = C.c_uint32
        L.astc_oracle_integer_from_trits.argtypes = [C.c_int] * 5
        L.astc_oracle_integer_from_trits.restype = C.c_uint8
        L.astc_oracle_integer_from_quints.argtypes = [C.c_int] * 3
        L.astc_oracle_integer_from_quints.restype = C.c_uint8
        L.astc_oracle_scramble.argtypes = [C.c_int, C.c_int]
        L.astc_oracle_scramble.restype = C.c_uint8
        L.astc_oracle_blockmode.argtypes = [C.c_int]
        L.astc_oracle_blockmode.restype = C.c_uint32
        L.astc_oracle_decode_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p, C.c_size_t]
        L.astc_oracle_decode_image.restype = C.c_int
        L.astc_oracle_unpack_block.argtypes = [C.c_void_p, C.POINTER(OracleSymbolic)]
        L.astc_oracle_unpack_block.restype = C.c_int
        L.astc_oracle_set_mufu_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.astc_oracle_mufu_rcp.argtypes = [C.c_float]
        L.astc_oracle_mufu_rcp.restype = C.c_float
        L.astc_oracle_mufu_rsq.argtypes = [C.c_float]
        L.astc_oracle_mufu_rsq.restype = C.c_float
        rcp, rsq = mufu_tables()
        L.astc_oracle_set_mufu_tables(rcp.ctypes.data, rsq.ctypes.data)
        _lib = L
    return _lib


_mufu = None


def mufu_tables() -> tuple[np.ndarray, np.ndarray]:
    """int8 delta tables of MUFU.RCP (2^23) and MUFU.RSQ (2^24), captured on a B200 by
    tools/gen_mufu_tables.py; kept alive for the lifetime of the process (the C side holds pointers)."""
    global _mufu
    if _mufu is None:
        import lzma
        out = []
        for name, n in (("rcp", 1 << 23), ("rsq", 1 << 24)):
            raw = lzma.decompress((_HERE / "tables" / f"mufu_{name}.i8.xz").read_bytes())
            if len(raw) != n:
                raise RuntimeError(f"oracle/tables/mufu_{name}.i8.xz: {len(raw)} entries, expected {n}")
            out.append(np.frombuffer(raw, dtype=np.int8).copy())
        _mufu = tuple(out)
    return _mufu


def mufu_rcp(x: np.ndarray) -> np.ndarray:
    """Vectorised emulation of rcp.approx.ftz.f32 for positive normal float32 inputs."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32)
    base = (1.0 / x.astype(np.float64)).astype(np.float32)
    return (base.view(np.int32) + mufu_tables()[0][b & 0x7FFFFF].astype(np.int32)).view(np.float32)


def mufu_rsq(x: np.ndarray) -> np.ndarray:
    """Vectorised emulation of rsqrt.approx.ftz.f32 for positive normal float32 inputs."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32)
    base = (1.0 / np.sqrt(x.astype(np.float64))).astype(np.float32)
    idx = ((((b >> 23) + 1) & 1).astype(np.int64) << 23) | (b & 0x7FFFFF)
    return (base.view(np.int32) + mufu_tables()[1][idx].astype(np.int32)).view(np.float32)


def make_opt(block_dim=4, has_alpha=False, is_normal_map=False, srgb=False, axis_method=0) -> OracleOpt:
    return OracleOpt(int(block_dim), int(bool(has_alpha)), int(bool(is_normal_map)), int(bool(srgb)), int(axis_method))


VAR_TRUE_DIVISION, VAR_UNFUSED_SAMPLE, VAR_SRGB_POWF, VAR_EXACT_RCP_RSQ, VAR_UNFUSED_DEV = 1, 2, 4, 8, 16


def set_mufu_bias(rcp_ulps: int, rsq_ulps: int) -> None:
    """tools/golden_residual.py only: ulps added to every emulated rcp / rsq result (0, 0 = as captured).  Process-wide."""
    lib().astc_oracle_set_mufu_bias(int(rcp_ulps), int(rsq_ulps))


def set_variant(flags: int) -> None:
    """Sensitivity switches of tools/pin_sensitivity.py (0 = canonical arithmetic).  Process-wide."""
    lib().astc_oracle_set_variant(int(flags))


def num_blocks(width: int, height: int, dim: int) -> tuple[int, int]:
    return (width + dim - 1) // dim, (height + dim - 1) // dim


def encode_image(rgba: np.ndarray, *, block_dim=4, has_alpha=False, is_normal_map=False,
                 srgb=False, threads=0, axis_method=0) -> np.ndarray:
    """rgba: (H, W, 4) uint8, row 0 first.  Returns (nblocks, 16) uint8."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, c = rgba.shape
    assert c == 4
    bw, bh = num_blocks(w, h, block_dim)
    out = np.zeros((bw * bh, 16), dtype=np.uint8)
    if bw * bh == 0:
        return out
    opt = make_opt(block_dim, has_alpha, is_normal_map, srgb, axis_method)
    lib().astc_oracle_encode_image(rgba.ctypes.data, w, h, w * 4, C.byref(opt),
                                   out.ctypes.data, threads)
    return out


def encode_rows(rgba: np.ndarray, row0: int, row1: int, *, block_dim=4, has_alpha=False,
                is_normal_map=False, srgb=False, threads=0, axis_method=0) -> tuple[np.ndarray, int]:
    """Encode block rows [row0,row1) only; returns (blocks, threads_used)."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, _ = rgba.shape
    bw, _bh = num_blocks(w, h, block_dim)
    out = np.zeros((bw * (row1 - row0), 16), dtype=np.uint8)
    opt = make_opt(block_dim, has_alpha, is_normal_map, srgb, axis_method)
    used = lib().astc_oracle_encode_rows(rgba.ctypes.data, w, h, w * 4, C.byref(opt),
                                         row0, row1, out.ctypes.data, threads)
    return out, used


def encode_block(raw: np.ndarray, *, block_dim=4, has_alpha=False, is_normal_map=False,
                 srgb=False, axis_method=0):
    """raw: (dim*dim, 4) float32 UNORM values.  Returns (16 bytes, trace)."""
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    assert raw.shape == (block_dim * block_dim, 4)
    out = (C.c_uint8 * 16)()
    tr = OracleTrace()
    opt = make_opt(block_dim, has_alpha, is_normal_map, srgb, axis_method)
    lib().astc_oracle_encode_block(raw.ctypes.data, C.byref(opt), out, C.byref(tr))
    return np.frombuffer(bytes(out), dtype=np.uint8).copy(), tr


def decode_image(blocks: np.ndarray, width: int, height: int, block_dim: int):
    """Returns ((H, W, 4) uint8, number_of_undecodable_blocks)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    out = np.zeros((height, width, 4), dtype=np.uint8)
    bad = lib().astc_oracle_decode_image(blocks.ctypes.data, width, height, block_dim,
                                         out.ctypes.data, width * 4)
    return out, bad


def unpack_blocks(blocks: np.ndarray):
    """Symbolic fields of each block: dict of arrays (mode, cem, ep, weights...)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 16)
    n = blocks.shape[0]
    res = {"ok": np.zeros(n, bool), "mode": np.zeros(n, np.uint32), "cem": np.zeros(n, np.uint32),
           "partitions": np.zeros(n, np.uint32), "ep": np.zeros((n, 8), np.uint8),
           "weights": np.zeros((n, 16), np.uint8)}
    s = OracleSymbolic()
    L = lib()
    for i in range(n):
        rc = L.astc_oracle_unpack_block(blocks[i].ctypes.data, C.byref(s))
        res["ok"][i] = rc == 0
        res["mode"][i] = s.mode
        res["cem"][i] = s.cem
        res["partitions"][i] = s.partitions
        res["ep"][i] = np.frombuffer(bytes(s.ep), np.uint8)
        res["weights"][i] = np.frombuffer(bytes(s.weights), np.uint8)[:16]
    return res


def unorm_lut(srgb: bool) -> np.ndarray:
    out = (C.c_float * 256)()
    lib().astc_oracle_unorm_lut(int(bool(srgb)), out)
    return np.frombuffer(bytes(out), dtype=np.float32).copy()


def psnr_per_channel(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """PSNR (dB, peak 255) per channel of two (H, W, 4) uint8 images."""
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = (d * d).reshape(-1, a.shape[-1]).mean(axis=0)
    with np.errstate(divide="ignore"):
        return 10.0 * np.log10(255.0 * 255.0 / mse)


def downsample2x2(img: np.ndarray) -> np.ndarray:
    """Checker for astc_b200_downsample2x2_device (not part of the reference: SURVEY.md 8f N3).
    2x2 box filter on (H, W, 4) uint8, (sum + 2) >> 2; odd trailing row / column dropped; a
    dimension of 1 stays 1 (its texel counted twice)."""
    h, w = img.shape[0], img.shape[1]
    oh, ow = max(1, h // 2), max(1, w // 2)
    c = img.astype(np.uint32)
    ys = (np.arange(oh) * 2, np.arange(oh) * 2 + 1) if h > 1 else (np.zeros(1, int), np.zeros(1, int))
    xs = (np.arange(ow) * 2, np.arange(ow) * 2 + 1) if w > 1 else (np.zeros(1, int), np.zeros(1, int))
    acc = sum(c[np.ix_(ys[j], xs[i])] for j in range(2) for i in range(2))
    return ((acc + 2) >> 2).astype(np.uint8)
