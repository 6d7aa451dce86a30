"""Integer sequence encoding known-answer and round-trip tests
(ASTC_IntegerSequenceEncoding.hlsl:142-276), CPU oracle here, device packer under -m gpu."""
import ctypes as C

import numpy as np
import pytest

LEVELS = [2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64, 80, 96, 128, 160, 192, 256]


def _bits(stream: bytes, pos: int, n: int) -> int:
    v = int.from_bytes(stream, "little")
    return (v >> pos) & ((1 << n) - 1)


def _oracle_encode(oracle, values, quant):
    buf = (C.c_uint8 * 16)()
    arr = (C.c_uint8 * len(values))(*values)
    n = oracle.lib().astc_oracle_bise_encode(arr, len(values), quant, buf)
    return bytes(buf), n


def _spec_decode(oracle, stream: bytes, count: int, quant: int):
    """Independent decoder written from the ASTC spec's trit/quint block layout."""
    b, t, q = C.c_int(), C.c_int(), C.c_int()
    oracle.lib().astc_oracle_quant_layout(quant, C.byref(b), C.byref(t), C.byref(q))
    n, out, pos = b.value, [], 0
    if t.value:
        inv = {}
        for i in range(243):
            tr = [(i // 3 ** k) % 3 for k in range(5)]
            inv.setdefault(oracle.lib().astc_oracle_integer_from_trits(*tr), tr)
        while len(out) < count:
            m, T = [], 0
            for width, shift in ((2, 0), (2, 2), (1, 4), (2, 5), (1, 7)):
                m.append(_bits(stream, pos, n)); pos += n
                T |= _bits(stream, pos, width) << shift; pos += width
            out += [(inv[T][j] << n) | m[j] for j in range(5)]
    elif q.value:
        inv = {}
        for i in range(125):
            qu = [(i // 5 ** k) % 5 for k in range(3)]
            inv.setdefault(oracle.lib().astc_oracle_integer_from_quints(*qu), qu)
        while len(out) < count:
            m, Q = [], 0
            for width, shift in ((3, 0), (2, 3), (2, 5)):
                m.append(_bits(stream, pos, n)); pos += n
                Q |= _bits(stream, pos, width) << shift; pos += width
            out += [(inv[Q][j] << n) | m[j] for j in range(3)]
    else:
        for _ in range(count):
            out.append(_bits(stream, pos, n)); pos += n
    return out[:count]


def _cases(quant, rng):
    lv = LEVELS[quant]
    b = {0: 0}
    counts = [1, 3, 5, 6, 8, 15, 16]
    for cnt in counts:
        bits = None
        yield cnt, [lv - 1] * cnt
        yield cnt, [0] * cnt
        yield cnt, list(rng.integers(0, lv, cnt))


@pytest.mark.parametrize("quant", range(21), ids=lambda q: f"QUANT_{LEVELS[q]}")
def test_oracle_roundtrip(oracle, quant):
    rng = np.random.default_rng(1000 + quant)
    for cnt, vals in _cases(quant, rng):
        if oracle.lib().astc_oracle_ise_bitcount(cnt, quant) > 128:
            continue
        stream, nbits = _oracle_encode(oracle, [int(v) for v in vals], quant)
        assert _spec_decode(oracle, stream, cnt, quant) == [int(v) for v in vals]


def test_known_answers(oracle):
    # 5 trits (2,1,0,2,1), no plain bits: T = integer_from_trits -> 8 bits, LSB first
    s, n = _oracle_encode(oracle, [2, 1, 0, 2, 1], 1)
    T = oracle.lib().astc_oracle_integer_from_trits(2, 1, 0, 2, 1)
    assert n == 8 and s[0] == T and not any(s[1:])
    # QUANT_6 weights (trit + 1 bit): m0 T[1:0] m1 T[3:2] m2 T[4] m3 T[6:5] m4 T[7]
    vals = [5, 0, 3, 4, 1]
    s, n = _oracle_encode(oracle, vals, 4)
    T = oracle.lib().astc_oracle_integer_from_trits(*[v >> 1 for v in vals])
    m = [v & 1 for v in vals]
    want = m[0] | (T & 3) << 1 | m[1] << 3 | ((T >> 2) & 3) << 4 | m[2] << 6 | ((T >> 4) & 1) << 7 | \
        m[3] << 8 | ((T >> 5) & 3) << 9 | m[4] << 11 | ((T >> 7) & 1) << 12
    assert n == 13 and int.from_bytes(s, "little") == want
    # 3 quints (4,4,4): the all-ones special case of the quint block
    s, n = _oracle_encode(oracle, [4, 4, 4], 3)
    assert n == 7 and s[0] == oracle.lib().astc_oracle_integer_from_quints(4, 4, 4)
    # QUANT_256 endpoints are plain bytes (bise_endpoints, :233-239)
    s, n = _oracle_encode(oracle, [1, 2, 3, 250, 251, 252, 253, 254], 20)
    assert n == 64 and s[:8] == bytes([1, 2, 3, 250, 251, 252, 253, 254])
    # 16 weights pad the fourth group with zeros: 42 payload bits, bits past them are 0
    s, n = _oracle_encode(oracle, [5] * 16, 4)
    assert n == 4 * 13 and int.from_bytes(s, "little") >> 42 == 0


@pytest.mark.gpu
@pytest.mark.parametrize("quant", range(21), ids=lambda q: f"QUANT_{LEVELS[q]}")
def test_device_packer_matches_oracle(native, oracle, quant):
    import torch
    rng = np.random.default_rng(7 + quant)
    for cnt in (1, 5, 6, 8, 16):
        if oracle.lib().astc_oracle_ise_bitcount(cnt, quant) > 128:
            continue
        vals = rng.integers(0, LEVELS[quant], size=(257, cnt)).astype(np.uint8)
        vals[0] = LEVELS[quant] - 1
        vals[1] = 0
        got = native.bise_encode(torch.from_numpy(vals).cuda(), quant).cpu().numpy()
        for i in range(vals.shape[0]):
            want, _ = _oracle_encode(oracle, [int(v) for v in vals[i]], quant)
            assert bytes(got[i]) == want, (quant, cnt, i)
