"""The exact-arithmetic identities the CUDA kernel relies on to match the oracle bit for
bit, checked on the CPU with numpy float32 / exact rationals (no GPU needed)."""
import re
from fractions import Fraction as Fr

import numpy as np

from conftest import ROOT

f32 = np.float32
KERNELS = (ROOT / "astc_encoder_b200" / "csrc" / "astc_kernels.cu").read_text()
BLOCK = (ROOT / "astc_encoder_b200" / "csrc" / "astc_block.cuh").read_text()


def _const(text, name):
    m = re.search(name + r"\s*=\s*(-?0x[0-9a-fA-F.]+p[-+]?\d+)f", text)
    assert m, name
    return float.fromhex(m.group(1))


def _rn(fr: Fr) -> np.float32:
    """Exact Fraction -> nearest float32, ties to even."""
    y = f32(float(fr))
    cands = [y, np.nextafter(y, f32(np.inf)), np.nextafter(y, f32(-np.inf))]
    key = lambda c: (abs(Fr(float(c)) - fr), int(np.frombuffer(f32(c).tobytes(), np.uint32)[0]) & 1)
    return f32(min(cands, key=key))


def test_unorm_two_constant_formula_equals_division():
    """raw = fma(c, kRcpHi, c*kRcpLo) == c/255.0f for every byte (astc_kernels.cu unorm2)."""
    hi, lo = _const(KERNELS, "kRcpHi"), _const(KERNELS, "kRcpLo")
    assert f32(hi) == f32(1.0) / f32(255.0)
    for c in range(256):
        q = Fr(c) * Fr(lo)
        assert Fr(float(f32(float(q)))) == q                    # c*kRcpLo is exact in float32
        got = _rn(Fr(c) * Fr(hi) + q)                           # the FMA rounds once
        assert got == f32(c) / f32(255.0), c


def test_texel_times_255_is_the_byte():
    """RN((c/255) * 255) == c: why the linear-mode mean may sum byte values (convert_texel)."""
    c = np.arange(256, dtype=np.float32)
    assert np.array_equal((c / f32(255.0)) * f32(255.0), c)


def test_magic_add_is_round_half_even():
    magic = f32(12582912.0)
    v = np.concatenate([np.arange(0, 256, 0.5, dtype=np.float32),
                        np.random.default_rng(0).uniform(0, 255, 100000).astype(np.float32),
                        np.nextafter(np.arange(0.5, 256, 1, dtype=np.float32), f32(0)),
                        np.nextafter(np.arange(0.5, 256, 1, dtype=np.float32), f32(1e9))])
    bits = (v + magic).view(np.uint32)
    assert np.array_equal(bits - np.uint32(0x4B400000), np.rint(v).astype(np.uint32))
    assert np.array_equal((bits & 0xFF), np.rint(v).astype(np.uint32) & 0xFF)


def test_small_length_threshold():
    """sqrtf(x) < 1e-5f  <=>  x < kSmallSq (astc_block.cuh), for correctly rounded sqrt."""
    k = f32(_const(BLOCK, "kSmallSq"))
    small = f32(1e-5)
    below = np.nextafter(k, f32(0))
    assert np.sqrt(k) >= small and np.sqrt(below) < small
    xs = np.abs(np.random.default_rng(1).normal(0, 1e-10, 200000)).astype(np.float32)
    assert np.array_equal(np.sqrt(xs) < small, xs < k)


def test_weight_clamp_never_acts(oracle):
    """x * rcp(x) <= 1 + 3 * 2^-23 with the MUFU reciprocal (good to 1 ulp), so round(w * range) <= range:
    the reference's clamp is dead code and the kernel's table index stays in its row."""
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(1e-5, 1500, 500000), [1e-5, 255.0, 441.67294]]).astype(np.float32)
    n = (x * oracle.mufu_rcp(x)).astype(np.float32)
    assert n.max() <= f32(1.0) + f32(3 * 2.0 ** -23) and n.min() >= f32(1.0) - f32(3 * 2.0 ** -23)
    for r in (f32(5.0), f32(11.0)):
        assert np.rint((n * r).astype(np.float32)).max() == r


def test_field_table_reconstructs_trit_packing(oracle):
    """The per-position field tables + scattered trit table (astc_block.cuh WeightPack) reproduce
    the oracle's bise_weights bit stream for random weight vectors."""
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(4)
    for method, n, levels in ((4, 1, 6), (7, 2, 12)):
        mpos = [0, n + 2, 2 * n + 4, 3 * n + 5, 4 * n + 7]
        field = [[0] * 16 for _ in range(5)]
        for j in range(5):
            for q in range(levels):
                v = L.astc_oracle_scramble(method, q)
                field[j][q] = (4 * (v >> n) * 3 ** j) | ((v & ((1 << n) - 1)) << (10 + mpos[j]))
        def scatter(T):
            return ((T & 3) << n) | (((T >> 2) & 3) << (2 * n + 2)) | (((T >> 4) & 1) << (3 * n + 4)) | \
                   (((T >> 5) & 3) << (4 * n + 5)) | (((T >> 7) & 1) << (5 * n + 7))
        trit = [scatter(L.astc_oracle_integer_from_trits(*[(i // 3 ** k) % 3 for k in range(5)])) for i in range(243)]
        gbits = 5 * n + 8
        for _ in range(300):
            q = rng.integers(0, levels, 16)
            stream = 0
            for g in range(4):
                s = sum(field[j][q[5 * g + j]] for j in range(5) if 5 * g + j < 16)
                assert (s & 0x3FF) % 4 == 0 and (s & 0x3FF) // 4 < 243
                stream |= (trit[(s & 0x3FC) >> 2] | (s >> 10)) << (g * gbits)
            qs = (C.c_uint8 * 16)(*[L.astc_oracle_scramble(method, int(x)) for x in q])
            buf = (C.c_uint8 * 16)()
            L.astc_oracle_bise_encode(qs, 16, method, buf)
            assert int.from_bytes(bytes(buf), "little") == stream


def test_fma_sum_of_unorm_texels_is_the_byte_sum():
    """fma(c/255, 255, sum) == sum + c exactly while sum is an integer below 2^14 (convert_texel,
    4x4 and 6x6: at most 36 * 255 = 9180): the linear-mode mean may accumulate by FMA."""
    raw = np.arange(256, dtype=np.float32) / f32(255.0)
    for c in range(256):
        prod = Fr(float(raw[c])) * 255                          # the FMA's unrounded product
        for s in (0, 1, 255, 1023, 2048, 4079, 4096, 8191, 8192, 9180 - c, 16383 - c):
            assert _rn(prod + s) == f32(s + c), (c, s)


def test_trace_bound_decides_the_early_exit():
    """power_iteration skips the exact |M v|^2 test when fl|M M v|^2 >= 1.02e-10 * trace(M)^2 (and
    >= 1e-30).  For random Gram matrices in float32 the implication |M v|^2 >= kSmallSq must
    hold whenever the bound fires -- including matrices scaled down to the threshold."""
    k = f32(_const(BLOCK, "kSmallSq"))
    rng = np.random.default_rng(7)
    fired = 0
    for trial in range(4000):
        n = int(rng.integers(2, 37))
        scale = f32(10.0 ** rng.uniform(-7, 2.5))
        d = (rng.normal(0, 1, (n, 4)) * np.array([1, rng.uniform(0, 1), rng.uniform(0, 1) ** 4, rng.uniform(0, 1) ** 8])).astype(np.float32) * scale
        m = np.zeros((4, 4), np.float32)
        for t in d:                                             # sequential float32 accumulation, like the kernel
            m = (m + np.outer(t, t).astype(np.float32)).astype(np.float32)
        m = (m * f32(1.0 / (n - 1))).astype(np.float32)
        v = rng.normal(0, 1, 4).astype(np.float32)
        v = (v / np.sqrt(np.sum(v * v, dtype=np.float32))).astype(np.float32)
        mv = lambda x: np.array([np.float32(sum(f32(m[i, j] * x[j]) for j in range(4))) for i in range(4)], np.float32)
        u = mv(v)
        w = mv(u)
        uu, ww = f32(np.sum(u * u, dtype=np.float32)), f32(np.sum(w * w, dtype=np.float32))
        tr = f32(m[0, 0] + m[1, 1] + m[2, 2] + m[3, 3])
        decided = max(f32(f32(tr * tr) * f32(1.02e-10)), f32(1e-30))
        if ww >= decided:
            fired += 1
            assert uu >= k, (trial, uu, ww, tr)
    assert fired > 1000                                         # the cheap test does decide the common case
