"""BASELINE.json configs at FULL size on the GPU.  The oracle checks a band sample bit for
bit (a full-size CPU encode would take minutes); the rest is covered by size-independent
properties: block-row bands re-encoded separately equal the one-shot output, every block
header is the mode's constant, endpoints are ordered, and the decoded image is close to
the source."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [
    ("cfg2", 4096, 4096, 4, dict(), "rgba"),
    ("cfg3", 8192, 8192, 6, dict(has_alpha=True, srgb=True), "rgba"),
    ("cfg4", 4096, 4096, 4, dict(is_normal_map=True), "normal"),
    ("cfg5", 16384, 16384, 4, dict(), "rgba"),
]


@pytest.mark.parametrize("name,w,h,dim,kw,kind", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_full_size_config(native, oracle, name, w, h, dim, kw, kind):
    import torch
    from astc_encoder_b200 import synth
    seed = {"cfg2": synth.SEED_CFG2, "cfg3": synth.SEED_CFG3, "cfg4": synth.SEED_CFG4, "cfg5": synth.SEED_CFG5}[name]
    img = (synth.synth_normal if kind == "normal" else synth.synth_rgba)(w, h, seed, device="cuda")
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
    out = native.encode_astc(img, opt)
    torch.cuda.synchronize()
    bx, by = native.block_counts(w, h, opt)
    assert out.shape == (bx * by, 16)

    # (1) oracle, bit-exact, on three bands: top, middle, bottom (incl. the padded last block row)
    rows_b = 48
    for r0 in (0, (by // 2) - rows_b // 2, by - rows_b):
        y0, y1 = r0 * dim, min(h, (r0 + rows_b) * dim)
        want = oracle.encode_image(img[y0:y1].cpu().numpy(), block_dim=dim, **kw)
        got = out[r0 * bx:(r0 + rows_b) * bx].cpu().numpy()
        assert np.array_equal(got, want), (name, r0)

    # (2) band independence: 8 bands encoded separately concatenate to the same bytes
    from astc_encoder_b200 import sharding
    parts = []
    for b in sharding.band_plan(w, h, opt, 8):
        parts.append(native.encode_astc(img[b.y0:b.y0 + b.rows], opt))
    assert torch.equal(torch.cat(parts), out)

    # (3) header constants + endpoint order on every block (vectorised on the GPU)
    head = out[:, 0].to(torch.int64) | (out[:, 1].to(torch.int64) << 8) | (out[:, 2].to(torch.int64) << 16)
    alpha = bool(kw.get("has_alpha"))
    assert bool(((head & 0x1FFFF) == ((0x43 if alpha else 0x251) | ((12 if alpha else 8) << 13))).all())
    lo64 = torch.zeros(out.shape[0], dtype=torch.int64, device="cuda")
    for i in range(2, 11):
        lo64 |= out[:, i].to(torch.int64) << (8 * (i - 2))
    ep = [(lo64 >> (1 + 8 * i)) & 0xFF for i in range(6)]              # bits 17.. = r0 r1 g0 g1 b0 b1
    assert bool((ep[0] + ep[2] + ep[4] <= ep[1] + ep[3] + ep[5]).all())

    # (4) decode and compare with the source in the encoder's input space
    dec = native.decode_astc(out, w, h, dim)
    if kw.get("srgb"):
        lut = torch.from_numpy(native.unorm_lut(True)).cuda() * 255.0
        src = torch.stack([lut[img[..., c].long()] for c in range(3)], dim=-1)
    else:
        src = img[..., :3].float()
    if kind == "normal":
        src, decf = src[..., :2], dec[..., :2].float()
    else:
        decf = dec[..., :3].float()
    mse = float(((decf - src) ** 2).mean())
    psnr = 10 * np.log10(255.0 ** 2 / mse)
    assert psnr > (30.0 if kind == "normal" else 22.0), (name, psnr)


ALL_VARIANTS = [
    dict(), dict(has_alpha=True), dict(srgb=True), dict(has_alpha=True, srgb=True),
    dict(is_normal_map=True), dict(has_alpha=True, is_normal_map=True),
]


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("kw", ALL_VARIANTS, ids=lambda k: "-".join(sorted(k)) or "rgb")
def test_every_variant_whole_image_vs_oracle(native, oracle, dim, kw):
    """Every option set x both block sizes on a 2050 x 1030 texture (ragged in both directions for 4x4
    and 6x6; a few hundred CTAs, several blocks per thread): the WHOLE output against the oracle."""
    import torch
    from astc_encoder_b200 import synth
    w, h = 2050, 1030
    gen = synth.synth_normal if kw.get("is_normal_map") else synth.synth_rgba
    img = gen(w, h, 0xA57C2000 + dim, device="cuda")
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
    got = native.encode_astc(img, opt)
    torch.cuda.synchronize()
    okw = dict(has_alpha=bool(kw.get("has_alpha")), is_normal_map=bool(kw.get("is_normal_map")), srgb=bool(kw.get("srgb")))
    want = oracle.encode_image(img.cpu().numpy(), block_dim=dim, **okw)
    got = got.cpu().numpy()
    assert got.shape == want.shape
    bad = int((got != want).any(axis=1).sum())
    assert bad == 0, f"{bad} of {len(want)} blocks differ"


def test_texture_larger_than_4_gib(native):
    """Byte offsets beyond 2^32 in one texture (34816 x 34816 RGBA8 = 4.5 GiB in, 1.1 GiB out; the .astc header allows
    2^24 texels a side): the one-shot encode, the four bands of astc_b200_band and the fused mip chain of the same
    texture must agree -- a 32-bit offset anywhere in the address arithmetic would not."""
    import torch
    from astc_encoder_b200 import synth
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs 24 GiB of free device memory")
    side = 34816                                                       # 64 * 544
    img = synth.synth_rgba(side, side, 4242, device="cuda", rows_per_chunk=512)
    assert img.numel() > 1 << 32
    for dim in (4, 6):
        opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6)
        whole = native.encode_astc(img, opt)
        parts = []
        for g in range(4):
            y0, rows, off, nbytes = native.band(side, side, opt, 4, g)
            parts.append(native.encode_astc(img[y0:y0 + rows], opt))
            assert parts[-1].numel() == nbytes
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(parts), whole), dim
        del parts, whole
    chain = native.mip_chain(img)                                      # one fused launch, 16 levels
    assert tuple(chain[1].shape) == (side // 2, side // 2, 4)
    assert torch.equal(chain[1], native.downsample2x2(img))
    lvl = chain[1]
    for nxt in chain[2:]:
        assert torch.equal(nxt, native.downsample2x2(lvl))
        lvl = nxt
