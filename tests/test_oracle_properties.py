"""Properties of the oracle's encodings: every block is a valid block of
the encoder's subset, endpoints are ordered so decoders never apply blue contraction,
flat / alpha-only / padded blocks behave as the reference's quirks dictate, and the
decoder round trip has sane quality."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

MODES = [dict(has_alpha=False), dict(has_alpha=True)]


def _unpack(oracle, blocks):
    return oracle.unpack_blocks(blocks)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), dim=st.sampled_from([4, 6]), alpha=st.booleans(), srgb=st.booleans(),
       flavour=st.sampled_from(["noise", "gradient", "two-tone", "flat", "alpha-only"]))
def test_random_blocks_are_valid_and_ordered(oracle, seed, dim, alpha, srgb, flavour):
    rng = np.random.default_rng(seed)
    n = 8
    img = np.zeros((dim, dim * n, 4), np.uint8)
    if flavour == "noise":
        img[:] = rng.integers(0, 256, img.shape)
    elif flavour == "gradient":
        ramp = np.linspace(rng.integers(0, 128), rng.integers(128, 256), dim * n)
        img[:] = ramp[None, :, None].astype(np.uint8)
        img[..., 3] = rng.integers(0, 256)
    elif flavour == "two-tone":
        a, b = rng.integers(0, 256, 4), rng.integers(0, 256, 4)
        img[:] = np.where(rng.random((dim, dim * n, 1)) < 0.5, a, b)
    elif flavour == "flat":
        img[:] = rng.integers(0, 256, 4)
    else:
        img[..., :3] = rng.integers(0, 256, 3)
        img[..., 3] = rng.integers(0, 256, (dim, dim * n))
    enc = oracle.encode_image(img, block_dim=dim, has_alpha=alpha, srgb=srgb)
    sym = _unpack(oracle, enc)
    assert sym["ok"].all()
    assert (sym["mode"] == (0x43 if alpha else 0x251)).all() and (sym["partitions"] == 1).all()
    assert (sym["cem"] == (12 if alpha else 8)).all()
    ep = sym["ep"].astype(int)
    assert (ep[:, 0] + ep[:, 2] + ep[:, 4] <= ep[:, 1] + ep[:, 3] + ep[:, 5]).all()     # no blue contraction
    assert (sym["weights"] <= (5 if alpha else 11)).all()
    if not alpha:
        assert (ep[:, 6:] == 255).all()
    if flavour == "flat":
        assert (sym["weights"] == 0).all() and (ep[:, 0::2] == ep[:, 1::2]).all()
    if flavour == "alpha-only":
        # the seed vector has w = 0 (ASTC_Encode.hlsl:96): a block varying only in alpha is encoded flat
        assert (sym["weights"] == 0).all()
    dec, bad = oracle.decode_image(enc, dim * n, dim, dim)
    assert bad == 0


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("mode", MODES, ids=["rgb", "rgba"])
def test_roundtrip_quality(oracle, dim, mode):
    from astc_encoder_b200 import synth
    img = synth.synth_rgba(192, 144, synth.SEED_CFG2).numpy()
    enc = oracle.encode_image(img, block_dim=dim, **mode)
    dec, bad = oracle.decode_image(enc, 192, 144, dim)
    assert bad == 0
    p = oracle.psnr_per_channel(dec, img)
    # noisy synthetic data, one partition, fixed modes: mid-20s dB on the coded channels
    assert (p[:3] > 22.0).all(), p
    if mode["has_alpha"]:
        assert p[3] > 20.0
    else:
        assert (dec[..., 3] == 255).all()


def test_padding_reads_zero(oracle):
    """W, H not multiples of the block: out-of-range texels are (0,0,0,0) (Texture2D.Load), and a
    normal map still forces b = a = 1 on them (ASTC_Encode.hlsl:574-578)."""
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    padded = np.zeros((8, 8, 4), np.uint8)
    padded[:5, :7] = img
    for kw in (dict(), dict(has_alpha=True), dict(is_normal_map=True)):
        a = oracle.encode_image(img, block_dim=4, **kw)
        b = oracle.encode_image(padded, block_dim=4, **kw)
        assert np.array_equal(a, b), kw
    a = oracle.encode_image(img, block_dim=6, has_alpha=True)
    padded6 = np.zeros((6, 12, 4), np.uint8)
    padded6[:5, :7] = img
    assert np.array_equal(a, oracle.encode_image(padded6, block_dim=6, has_alpha=True))


def test_tiny_mips(oracle):
    for w, h in ((1, 1), (2, 2), (3, 1), (1, 5)):
        img = np.full((h, w, 4), 200, np.uint8)
        for dim in (4, 6):
            enc = oracle.encode_image(img, block_dim=dim, has_alpha=True)
            assert enc.shape == (-(-w // dim) * -(-h // dim), 16)
            dec, bad = oracle.decode_image(enc, w, h, dim)
            assert bad == 0 and dec.shape == (h, w, 4)


def test_normal_map_ignores_srgb_and_blue_alpha(oracle):
    from astc_encoder_b200 import synth
    img = synth.synth_normal(64, 64, synth.SEED_CFG4).numpy()
    a = oracle.encode_image(img, block_dim=4, is_normal_map=True)
    b = oracle.encode_image(img, block_dim=4, is_normal_map=True, srgb=True)       # main.cpp:214
    assert np.array_equal(a, b)
    img2 = img.copy()
    img2[..., 2:] = np.random.default_rng(1).integers(0, 256, img2[..., 2:].shape)  # b, a are overwritten by the kernel
    assert np.array_equal(a, oracle.encode_image(img2, block_dim=4, is_normal_map=True))
    sym = oracle.unpack_blocks(a)
    assert (sym["ep"][:, 4:6] == 255).all()                                         # blue endpoints = 1.0


def test_empty_image(oracle):
    assert oracle.encode_image(np.zeros((0, 16, 4), np.uint8)).shape == (0, 16)


def test_max_accumulation_axis_oracle_properties(oracle):
    """max_accumulation_pixel_direction (ASTC_Encode.hlsl:170-227) in the oracle: valid blocks, ordered endpoints,
    flat blocks identical to the PCA path, and on a block varying along ONE channel the same endpoints as the PCA."""
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    for dim in (4, 6):
        for kw in (dict(), dict(has_alpha=True), dict(is_normal_map=True)):
            enc = oracle.encode_image(img, block_dim=dim, axis_method=1, **kw)
            sym = oracle.unpack_blocks(enc)
            assert sym["ok"].all()
            ep = sym["ep"].astype(int)
            assert np.all(ep[:, 0] + ep[:, 2] + ep[:, 4] <= ep[:, 1] + ep[:, 3] + ep[:, 5])     # no blue contraction on decode
            dec, bad = oracle.decode_image(enc, 64, 64, dim)
            assert bad == 0
    flat = np.full((8, 8, 4), 77, np.uint8)
    assert np.array_equal(oracle.encode_image(flat, block_dim=4, has_alpha=True, axis_method=1),
                          oracle.encode_image(flat, block_dim=4, has_alpha=True))
    ramp = np.zeros((4, 4, 4), np.uint8); ramp[..., 3] = 255; ramp[..., 1] = (np.arange(16).reshape(4, 4) * 9 + 20)
    a = oracle.unpack_blocks(oracle.encode_image(ramp, block_dim=4, axis_method=1))
    b = oracle.unpack_blocks(oracle.encode_image(ramp, block_dim=4))
    assert np.array_equal(a["ep"], b["ep"])
