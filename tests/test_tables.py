"""Every lookup table (oracle's and the product library's, both derived from the ASTC
spec rules) against the literal tables of the reference, extracted into
tests/golden/ref_tables.json by tools/extract_ref_tables.py."""
import json

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def ref():
    return json.loads((GOLDEN / "ref_tables.json").read_text())


def _layouts(fn):
    import ctypes as C
    rows = []
    for q in range(21):
        b, t, qu = C.c_int(), C.c_int(), C.c_int()
        fn(q, C.byref(b), C.byref(t), C.byref(qu))
        rows += [b.value, t.value, qu.value]
    return rows


def test_bits_trits_quints(ref, oracle, native):
    assert _layouts(oracle.lib().astc_oracle_quant_layout) == ref["bits_trits_quints_table"]
    assert _layouts(native.lib().astc_b200_quant_layout) == ref["bits_trits_quints_table"]


def test_integer_from_trits(ref, oracle, native):
    want = ref["integer_from_trits"]
    for i in range(243):
        t = [(i // 3 ** k) % 3 for k in range(5)]
        assert oracle.lib().astc_oracle_integer_from_trits(*t) == want[i], i
        assert native.lib().astc_b200_integer_from_trits(*t) == want[i], i


def test_integer_from_quints(ref, oracle, native):
    want = ref["integer_from_quints"]
    for i in range(125):
        q = [(i // 5 ** k) % 5 for k in range(3)]
        assert oracle.lib().astc_oracle_integer_from_quints(*q) == want[i], i
        assert native.lib().astc_b200_integer_from_quints(*q) == want[i], i


def test_scramble_table(ref, oracle, native):
    want = np.array(ref["scramble_table"]).reshape(12, 32)
    levels = [2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32]
    for m in range(12):
        for q in range(levels[m]):
            assert oracle.lib().astc_oracle_scramble(m, q) == want[m, q], (m, q)
            assert native.lib().astc_b200_scramble(m, q) == want[m, q], (m, q)
    # the two rows the encoder really indexes (ASTC_Table.hlsl:27,42)
    assert list(want[4, :6]) == [0, 2, 4, 5, 3, 1]
    assert list(want[7, :12]) == [0, 4, 8, 2, 6, 10, 11, 7, 3, 9, 5, 1]


def test_blockmode_words(oracle, native):
    # assemble_blockmode (ASTC_Encode.hlsl:446-473): QUANT_6 -> 0x43, QUANT_12 -> 0x251
    for fn in (oracle.lib().astc_oracle_blockmode, native.lib().astc_b200_blockmode):
        assert fn(4) == 0x43
        assert fn(7) == 0x251


def test_ise_bitcount(oracle, native):
    for q in range(21):
        for n in (0, 1, 5, 6, 8, 16, 36, 64):
            assert oracle.lib().astc_oracle_ise_bitcount(n, q) == native.lib().astc_b200_ise_bitcount(n, q)
    assert native.lib().astc_b200_ise_bitcount(16, 4) == 42      # 16 x (trit + 1 bit)
    assert native.lib().astc_b200_ise_bitcount(16, 7) == 58      # 16 x (trit + 2 bits)
    assert native.lib().astc_b200_ise_bitcount(8, 20) == 64


def test_unorm_luts_match(oracle, native):
    """Kernel-side UNORM8->float tables == oracle's own evaluation, bit for bit."""
    for srgb in (False, True):
        a, b = native.unorm_lut(srgb), oracle.unorm_lut(srgb)
        assert a.tobytes() == b.tobytes()
    lin = native.unorm_lut(False)
    assert lin[0] == 0.0 and lin[255] == 1.0
    s = native.unorm_lut(True)
    assert s[0] == 0.0 and s[255] == 1.0 and np.all(np.diff(s) > 0)
    # (c/255)*255 == c exactly: what makes un-contracted texels exact integers
    assert np.array_equal(lin * np.float32(255.0), np.arange(256, dtype=np.float32))
