"""CUDA path edge cases and API surface, all through the C ABI, all bit-exact vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _enc(native, img_t, opt, **kw):
    import torch
    out = native.encode_astc(img_t, opt, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _okw(opt, dim):
    return dict(block_dim=dim, has_alpha=opt.has_alpha, is_normal_map=opt.is_normal_map, srgb=opt.srgb)


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("size", [(1, 1), (2, 2), (3, 5), (4, 4), (6, 6), (7, 1), (1, 9), (129, 3), (16, 4097)], ids=str)
def test_tiny_and_ragged_sizes(native, oracle, dim, size):
    import torch
    w, h = size
    rng = np.random.default_rng(w * 131 + h)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    for opt in (native.encode_option(is4x4=dim == 4, is6x6=dim == 6),
                native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True, srgb=True),
                native.encode_option(is4x4=dim == 4, is6x6=dim == 6, is_normal_map=True)):
        got = _enc(native, torch.from_numpy(img).cuda(), opt)
        want = oracle.encode_image(img, **_okw(opt, dim))
        assert np.array_equal(got, want), (size, opt)


@pytest.mark.parametrize("dim", [4, 6])
def test_special_content(native, oracle, dim):
    """Flat, alpha-only variation (4-D PCA quirk), saturated, two-tone, single-channel ramps."""
    import torch
    rng = np.random.default_rng(11)
    n = 64
    tiles = []
    flat = np.zeros((dim, dim * n, 4), np.uint8); flat[:] = rng.integers(0, 256, (1, n, 1, 4)).repeat(dim, 2).reshape(1, dim * n, 4)
    tiles.append(flat)
    aonly = flat.copy(); aonly[..., 3] = rng.integers(0, 256, (dim, dim * n)); tiles.append(aonly)
    sat = rng.choice(np.array([0, 255], np.uint8), (dim, dim * n, 4)); tiles.append(sat)
    two = np.where(rng.random((dim, dim * n, 1)) < 0.5, rng.integers(0, 256, 4), rng.integers(0, 256, 4)).astype(np.uint8); tiles.append(two)
    ramp = np.zeros((dim, dim * n, 4), np.uint8); ramp[..., 1] = (np.arange(dim * n) * 3) % 256; ramp[..., 3] = 255; tiles.append(ramp)
    near = np.full((dim, dim * n, 4), 100, np.uint8); near[::2, ::3, 0] = 101; tiles.append(near)
    img = np.concatenate(tiles, axis=0)
    for kw in (dict(), dict(has_alpha=True), dict(srgb=True), dict(has_alpha=True, srgb=True), dict(is_normal_map=True)):
        opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
        got = _enc(native, torch.from_numpy(img).cuda(), opt)
        want = oracle.encode_image(img, **_okw(opt, dim))
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert len(bad) == 0, (kw, bad[:10])


@pytest.mark.parametrize("dim", [4, 6])
def test_pitch_and_unaligned_base(native, oracle, dim):
    """Row pitch larger than the width, and a base pointer that is only 4-byte aligned (slow path)."""
    import torch
    from astc_encoder_b200 import synth
    big = synth.synth_rgba(300, 200, 5).cuda()
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
    for x0, w in ((0, 256), (1, 255), (3, 250), (4, 96)):
        view = big[10:170, x0:x0 + w]                       # strided rows, base offset 4*x0 bytes
        got = _enc(native, view, opt)
        want = oracle.encode_image(view.cpu().numpy(), **_okw(opt, dim))
        assert np.array_equal(got, want), (x0, w)


def test_explicit_out_and_stream(native, oracle):
    import torch
    from astc_encoder_b200 import synth
    img = synth.synth_rgba(128, 64, 6).cuda()
    opt = native.encode_option(has_alpha=True)
    out = torch.zeros((native.output_size(128, 64, opt) // 16, 16), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        native.encode_astc(img, opt, out=out, stream=s)
    s.synchronize()
    assert np.array_equal(out.cpu().numpy(), oracle.encode_image(img.cpu().numpy(), block_dim=4, has_alpha=True))
    with pytest.raises(ValueError):
        native.encode_astc(img.cpu(), opt)
    with pytest.raises(ValueError):
        native.encode_astc(img, opt, out=out[:10])


def test_srgb_texture_override(native, oracle):
    """encode_astc's texture format decides sRGB decoding (main.cpp:38,214), not the option flag."""
    import torch
    from astc_encoder_b200 import synth
    img = synth.synth_rgba(64, 64, 7)
    opt = native.encode_option(srgb=True)
    lin = _enc(native, img.cuda(), opt, srgb_texture=False)
    assert np.array_equal(lin, oracle.encode_image(img.numpy(), block_dim=4))
    srgb = _enc(native, img.cuda(), native.encode_option(), srgb_texture=True)
    assert np.array_equal(srgb, oracle.encode_image(img.numpy(), block_dim=4, srgb=True))


@pytest.mark.parametrize("dim", [4, 6])
def test_batch_of_mip_chains_one_launch(native, oracle, dim):
    import torch
    from astc_encoder_b200 import synth
    chains = [synth.mip_chain(synth.synth_rgba(64, 64, synth.SEED_BATCH + i).cuda()) for i in range(3)]
    srcs = [m for c in chains for m in c]
    assert [tuple(m.shape[:2]) for m in chains[0]] == [(64 >> i, 64 >> i) for i in range(7)]
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
    before = native.launch_count()
    batch = native.Batch(srcs, opt)
    outs = batch.encode()
    torch.cuda.synchronize()
    assert native.launch_count() - before == 1
    assert batch.total_texels == sum(m.shape[0] * m.shape[1] for m in srcs)
    for m, o in zip(srcs, outs):
        want = oracle.encode_image(m.cpu().numpy(), **_okw(opt, dim))
        assert np.array_equal(o.cpu().numpy(), want), tuple(m.shape)
    batch.close()
    empty = native.Batch([], opt)
    empty.encode()
    assert empty.total_blocks == 0


def test_host_entry_point_matches_device_path(native, oracle):
    from astc_encoder_b200 import synth
    img = synth.synth_rgba(1000, 3000, 8).numpy()              # several internal bands
    for dim, kw in ((4, dict()), (6, dict(has_alpha=True, srgb=True))):
        opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
        got = native.encode_astc_host(img, opt)
        assert np.array_equal(got, oracle.encode_image(img, block_dim=dim, **kw))
    assert native.encode_astc_host(np.zeros((0, 8, 4), np.uint8), native.encode_option()).shape == (0, 16)


def test_device_decoder_matches_oracle_decoder(native, oracle):
    import torch
    from astc_encoder_b200 import synth
    img = synth.synth_rgba(250, 187, 9)
    for dim, kw in ((4, dict(has_alpha=True)), (4, dict()), (6, dict(has_alpha=True)), (6, dict())):
        opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
        blocks = native.encode_astc(img.cuda(), opt)
        dec = native.decode_astc(blocks, 250, 187, dim).cpu().numpy()
        ref, bad = oracle.decode_image(blocks.cpu().numpy(), 250, 187, dim)
        assert bad == 0 and np.array_equal(dec, ref)


def test_leaf_srgb_alpha_config0(native, oracle, leaf_rgba):
    """BASELINE config[0] as written (`leaf.png -alpha -4x4 -srgb`): only the oracle can check it."""
    import torch
    opt = native.encode_option(has_alpha=True, srgb=True)
    got = _enc(native, torch.from_numpy(leaf_rgba).cuda(), opt)
    assert np.array_equal(got, oracle.encode_image(leaf_rgba, block_dim=4, has_alpha=True, srgb=True))


def test_leaf_psnr_vs_golden(native, oracle, leaf_rgba, leaf_golden):
    import torch
    got = _enc(native, torch.from_numpy(leaf_rgba).cuda(), native.encode_option(has_alpha=True))
    dec = native.decode_astc(torch.from_numpy(got).cuda(), 1024, 1024, 4).cpu().numpy()
    dec_g, _ = oracle.decode_image(leaf_golden[4], 1024, 1024, 4)
    p, pg = oracle.psnr_per_channel(dec, leaf_rgba), oracle.psnr_per_channel(dec_g, leaf_rgba)
    assert np.all(np.abs(p - pg) <= 0.05), (p, pg)


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("passes", [1, 3, 8])
def test_block_walk_across_rows_and_passes(native, oracle, dim, passes, monkeypatch):
    """The incremental block walk (Walk in astc_kernels.cu): widths whose block rows are shorter
    than, equal to, and not a multiple of the 128-thread stride, under forced pass counts."""
    import torch
    monkeypatch.setenv("ASTC_B200_PASSES", str(passes))
    rng = np.random.default_rng(100 + passes)
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
    for w, h in ((dim * 5, dim * 300), (dim * 128, dim * 17), (dim * 129 - 1, dim * 23 + 1), (dim * 700 + 2, dim * 9)):
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        got = _enc(native, torch.from_numpy(img).cuda(), opt)
        want = oracle.encode_image(img, **_okw(opt, dim))
        assert np.array_equal(got, want), (w, h, passes)


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("passes", [1, 8])
def test_batch_walk_over_many_tiny_images(native, oracle, dim, passes, monkeypatch):
    """One launch over 150 images from 1x1 to a few hundred texels wide: a thread's next block is
    often several images further on, so the walk's image-crossing loop and re-division run a lot."""
    import torch
    monkeypatch.setenv("ASTC_B200_PASSES", str(passes))
    rng = np.random.default_rng(7 + passes)
    sizes = [(int(rng.integers(1, 40)), int(rng.integers(1, 40))) for _ in range(140)] + \
            [(333, 61), (64, 512), (dim * 128, dim * 3), (1, 1), (2, 700), (517, 3), (dim, dim), (96, 96), (1, 1), (250, 187)]
    imgs = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for w, h in sizes]
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6)
    batch = native.Batch([torch.from_numpy(i).cuda() for i in imgs], opt)
    outs = batch.encode()
    torch.cuda.synchronize()
    for img, o in zip(imgs, outs):
        want = oracle.encode_image(img, **_okw(opt, dim))
        assert np.array_equal(o.cpu().numpy(), want), img.shape
    batch.close()


@pytest.mark.parametrize("size", [(2048, 2048), (64, 32), (8, 8), (250, 187), (7, 5), (2, 2), (1, 9), (9, 1), (1, 1), (33, 2), (4099, 3)], ids=str)
def test_device_mip_downsample(native, oracle, size):
    """astc_b200_downsample2x2_device (vector and scalar paths, degenerate sizes) against the numpy
    restatement, and the device-built chain against torch's (astc_encoder_b200.synth.mip_chain)."""
    import torch
    from astc_encoder_b200 import synth
    w, h = size
    rng = np.random.default_rng(w * 131 + h)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    d = torch.from_numpy(img).cuda()
    got = native.downsample2x2(d)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), oracle.downsample2x2(img))
    chain = native.mip_chain(d)
    ref = synth.mip_chain(d)
    assert len(chain) == len(ref) and all(torch.equal(a, b) for a, b in zip(chain, ref))
    # a strided (pitched) view takes the scalar or the vector path depending on alignment
    wide = torch.zeros((h, w + 3, 4), dtype=torch.uint8, device="cuda")
    wide[:, 1:w + 1] = d
    assert torch.equal(native.downsample2x2(wide[:, 1:w + 1]), got)


def test_oracle_mufu_emulation_matches_the_device(native, oracle):
    """rcp.approx.ftz / rsqrt.approx.ftz on this GPU against the oracle's table emulation: random
    inputs over the range the encoder uses and over the whole normal range, a full sweep of the
    top 16 mantissa bits for both exponent parities, and the neighbourhood of the powers of two."""
    import torch
    rng = np.random.default_rng(11)
    parts = [np.exp(rng.uniform(np.log(1e-20), np.log(1e22), 1 << 20)).astype(np.float32),
             rng.integers(0x00800000, 0x7F000000, 1 << 20, dtype=np.uint32).view(np.float32),
             (np.uint32(0x3F800000) + (np.arange(1 << 17, dtype=np.uint32) << 7)).view(np.float32),
             (np.uint32(0x3F800000 - 64) + np.arange(128, dtype=np.uint32)).view(np.float32)]
    x = np.concatenate(parts)
    d = torch.from_numpy(x).cuda()
    got_rsq = native.mufu("rsq", d).cpu().numpy()
    assert np.array_equal(got_rsq.view(np.uint32), oracle.mufu_rsq(x).view(np.uint32))
    ok = (x > 2.0 ** -125) & (x < 2.0 ** 125)                 # reciprocals that stay normal (ftz beyond)
    got_rcp = native.mufu("rcp", d).cpu().numpy()
    assert np.array_equal(got_rcp.view(np.uint32)[ok], oracle.mufu_rcp(x).view(np.uint32)[ok])


def test_batch_and_decode_reject_malformed_tensors(native):
    """Batch applies encode_astc's tensor contract to every source and output: a permuted or
    sliced-in-x tensor, a short or non-contiguous output and a wrong dtype are errors, not garbage."""
    import torch
    opt = native.encode_option()
    good = torch.zeros((16, 16, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        native.Batch([good.permute(1, 0, 2)], opt)                       # texels not packed along x
    with pytest.raises(ValueError):
        native.Batch([good[:, ::2]], opt)                                # stride(1) != 4
    with pytest.raises(ValueError):
        native.Batch([good.to(torch.int8)], opt)
    with pytest.raises(ValueError):
        native.Batch([good.cpu()], opt)
    with pytest.raises(ValueError):
        native.Batch([good], opt, outputs=[torch.zeros(15 * 16, dtype=torch.uint8, device="cuda")])      # too small
    with pytest.raises(ValueError):
        native.Batch([good], opt, outputs=[torch.zeros((32, 16), dtype=torch.uint8, device="cuda")[::2]])  # not contiguous
    with pytest.raises(ValueError):
        native.Batch([good, good], opt, outputs=[torch.zeros((16, 16), dtype=torch.uint8, device="cuda")])
    rows = good[::2]                                                     # strided ROWS are fine
    b = native.Batch([rows], opt)
    b.encode()
    torch.cuda.synchronize()
    assert b.outputs[0].shape == (2 * 4, 16)
    b.close()
    with pytest.raises(ValueError):
        native.decode_astc(torch.zeros((3, 16), dtype=torch.uint8, device="cuda"), 16, 16, 4)              # 16 blocks needed
    with pytest.raises(ValueError):
        native.decode_astc(torch.zeros((16, 16), dtype=torch.uint8, device="cuda"), 16, 16, 5)


@pytest.mark.parametrize("dim", [4, 6])
def test_tapered_schedule_in_a_ragged_batch(native, dim):
    """A batch large enough for the tapered schedule (CTAs of 8, 4, 2 and 1 passes, csrc/astc_schedule.h), made of
    ragged textures, so that segment boundaries fall inside images, inside block rows and between images: the one
    launch must give each texture the bytes of its own single-texture encode (which runs the uniform schedule for
    the small ones and a differently cut tapered one for the large ones)."""
    import torch
    from astc_encoder_b200 import synth
    rng = np.random.default_rng(2024 + dim)
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
    srcs = []
    for i in range(36):
        w, h = (int(rng.integers(600, 1500)), int(rng.integers(600, 1500))) if i % 6 else (int(rng.integers(1, 40)), int(rng.integers(1, 40)))
        srcs.append(synth.synth_rgba(w, h, 900 + i, device="cuda"))
    batch = native.Batch(srcs, opt)
    assert batch.total_blocks > 592 * 15 * 128 * (1 if dim == 4 else 0.2)
    batch.encode()
    torch.cuda.synchronize()
    for s, o in zip(srcs, batch.outputs):
        assert torch.equal(o, native.encode_astc(s, opt)), tuple(s.shape)
    batch.close()


@pytest.mark.parametrize("size", [(2048, 2048), (64, 64), (64, 128), (4096, 64), (192, 320), (1024, 576), (250, 187), (100, 64), (1, 9), (3, 3)], ids=str)
def test_mip_chain_in_one_call(native, size):
    """astc_b200_mip_chain_device: the whole chain in one call (ONE fused launch when both sides are multiples of 64)
    must give the bytes of the level-by-level path; an arena can be reused (the fused kernel leaves its ticket zeroed)."""
    import torch
    from astc_encoder_b200 import synth
    w, h = size
    base = synth.synth_rgba(w, h, 31 + w + h, device="cuda")
    want = native.mip_chain_by_level(base)
    before = native.launch_count()
    got = native.mip_chain(base)
    launches = native.launch_count() - before
    torch.cuda.synchronize()
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
    for l, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a, b), (size, l)
    assert launches == (1 if (w % 64 == 0 and h % 64 == 0) else len(want) - 1)
    _, _, _, total = native.mip_chain_layout(w, h)
    arena = torch.full((total + 512,), 0xEE, dtype=torch.uint8, device="cuda")[:total + 256]     # garbage-filled, a little larger
    for _ in range(3):
        again = native.mip_chain(base, arena=arena)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(again, want))
    strided = synth.synth_rgba(w + 8, h, 5, device="cuda")[:, 4:4 + w]                            # pitch > 4 * w, base 16-byte aligned
    assert all(torch.equal(a, b) for a, b in zip(native.mip_chain(strided), native.mip_chain_by_level(strided)))


def test_encode_launches_can_be_captured_in_a_cuda_graph(native, oracle):
    """The encode launches carry the programmatic-stream-serialization attribute; captured into a CUDA graph (three
    launches back to back plus a mip chain and a batch) and replayed on new content they must still give the bytes of
    eager launches."""
    import torch
    from astc_encoder_b200 import synth
    opt4, opt6 = native.encode_option(has_alpha=True), native.encode_option(is6x6=True, srgb=True)
    img = synth.synth_rgba(512, 384, 77, device="cuda")
    out4, out6, out4b = native.encode_astc(img, opt4), native.encode_astc(img, opt6), native.encode_astc(img[:128], opt4)
    _, _, _, total = native.mip_chain_layout(512, 384)
    arena = torch.zeros(total, dtype=torch.uint8, device="cuda")
    chain = native.mip_chain(img, arena=arena)
    batch = native.Batch(chain, opt4)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g.capture_begin()
        native.encode_astc(img, opt4, out=out4, stream=s)
        native.encode_astc(img, opt6, out=out6, stream=s)
        native.encode_astc(img[:128], opt4, out=out4b, stream=s)
        native.mip_chain(img, stream=s, arena=arena)
        batch.encode(stream=s)
        g.capture_end()
    for seed in (5, 6):
        img.copy_(synth.synth_rgba(512, 384, seed, device="cuda"))
        for o in (out4, out6, out4b, *batch.outputs):
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
        host = img.cpu().numpy()
        assert np.array_equal(out4.cpu().numpy(), oracle.encode_image(host, block_dim=4, has_alpha=True))
        assert np.array_equal(out6.cpu().numpy(), oracle.encode_image(host, block_dim=6, srgb=True))
        assert np.array_equal(out4b.cpu().numpy(), oracle.encode_image(host[:128], block_dim=4, has_alpha=True))
        want_chain = native.mip_chain_by_level(img)
        for lv, o, w in zip(chain, batch.outputs, want_chain):
            assert torch.equal(lv, w)
            assert torch.equal(o, native.encode_astc(w, opt4))
    batch.close()


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("size", [(256, 192), (250, 187), (251, 187), (1024, 512), (5, 3)], ids=str)
def test_decoder_fast_and_generic_paths_on_random_payloads(native, oracle, dim, size):
    """decode_fast_kernel (aligned outputs: whole texel rows per store, packed weights, IDP4A interpolation) and the
    generic decode_kernel (any alignment) against the oracle decoder, on blocks with the encoder's two headers but
    RANDOM endpoint and weight bits: every trit code incl. the non-canonical ones, endpoint orders that need blue
    contraction, ragged edges.  (250 / 251 texel rows are not 16- / 8-byte multiples: the generic path.)"""
    import torch
    from astc_encoder_b200 import synth
    w, h = size
    rng = np.random.default_rng(w * 7 + h + dim)
    img = synth.synth_rgba(w, h, 12).cuda()
    for kw in (dict(has_alpha=True), dict()):
        opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
        blocks = native.encode_astc(img, opt).cpu().numpy()
        dec = native.decode_astc(torch.from_numpy(blocks).cuda(), w, h, dim).cpu().numpy()
        ref, bad = oracle.decode_image(blocks, w, h, dim)
        assert bad == 0 and np.array_equal(dec, ref), ("encoded", size, kw)
        noise = rng.integers(0, 256, blocks.shape, dtype=np.uint8)
        noise[:, 0] = blocks[:, 0]                                   # keep mode, partition count, CEM (bits 0..16)
        noise[:, 1] = blocks[:, 1]
        noise[:, 2] = (noise[:, 2] & 0xFE) | (blocks[:, 2] & 0x01)
        dec = native.decode_astc(torch.from_numpy(noise).cuda(), w, h, dim).cpu().numpy()
        ref, bad = oracle.decode_image(noise, w, h, dim)
        assert bad == 0 and np.array_equal(dec, ref), ("random payload", size, kw)
    garbage = rng.integers(0, 256, blocks.shape, dtype=np.uint8)      # anything else: decoded or the error colour, never a fault
    native.decode_astc(torch.from_numpy(garbage).cuda(), w, h, dim)
    torch.cuda.synchronize()
