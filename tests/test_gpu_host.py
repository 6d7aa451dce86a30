"""Host-buffer entry points (persistent context, single texture and batch), all through the C ABI and all
compared with the device path / the oracle.  Reference: load_tex upload + encode_astc + read_gpu
(main.cpp:46-52, astc_encode.h:87-194, astc_save.h:34-50)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_encode(native, img, opt):
    import torch
    return native.read_gpu(native.encode_astc(torch.from_numpy(np.ascontiguousarray(img)).cuda(), opt))


@pytest.mark.parametrize("dim", [4, 6])
def test_context_encode_host_sizes_grow_and_shrink(native, oracle, dim):
    """One context, textures of very different sizes one after the other (the workspace only grows), pinned-free
    (pageable numpy) memory, ragged sizes, strided rows."""
    from astc_encoder_b200 import synth
    ctx = native.Context()
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True, srgb=True)
    for (w, h) in ((64, 64), (1, 1), (2048, 1024), (250, 187), (4096, 2048), (16, 16), (3, 1000)):
        img = synth.synth_rgba(w, h, 50 + w).numpy()
        got = ctx.encode_host(img, opt)
        assert np.array_equal(got, _device_encode(native, img, opt)), (w, h)
    wide = synth.synth_rgba(300, 64, 9).numpy()
    view = wide[:, 10:270]                                              # pitch > 4 * width
    assert np.array_equal(ctx.encode_host(view, opt), oracle.encode_image(np.ascontiguousarray(view), block_dim=dim, has_alpha=True, srgb=True))
    ctx.trim()                                                          # workspace released, regrows on demand
    img = synth.synth_rgba(128, 96, 4).numpy()
    assert np.array_equal(ctx.encode_host(img, opt), oracle.encode_image(img, block_dim=dim, has_alpha=True, srgb=True))
    ctx.close()


def test_default_context_is_reused_by_encode_host(native):
    """astc_b200_encode_host (no context argument) runs on a persistent thread-local context: repeated calls
    give identical results and launch exactly one kernel per small texture."""
    from astc_encoder_b200 import synth
    opt = native.encode_option()
    img = synth.synth_rgba(256, 256, 1).numpy()
    first = native.encode_astc_host(img, opt)
    before = native.launch_count()
    for _ in range(20):
        assert np.array_equal(native.encode_astc_host(img, opt), first)
    assert native.launch_count() - before == 20


@pytest.mark.parametrize("dim,kw", [(4, dict()), (4, dict(has_alpha=True, srgb=True)), (6, dict(has_alpha=True)), (4, dict(is_normal_map=True)),
                                    (6, dict(axis_method=1))], ids=str)
def test_batch_encode_host_mip_chains(native, oracle, dim, kw):
    """Several whole mip chains + odd-sized textures in one astc_b200_context_batch_encode_host call: large levels
    copied directly, small ones through pinned staging, several upload/launch/download groups."""
    import torch
    from astc_encoder_b200 import synth
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
    gen = synth.synth_normal if kw.get("is_normal_map") else synth.synth_rgba
    images = []
    for i, size in enumerate((2048, 1024, 512)):
        images.extend(t.cpu().numpy() for t in native.mip_chain(gen(size, size, 70 + i).cuda()))
    images.append(synth.synth_rgba(250, 187, 5).numpy())
    images.append(synth.synth_rgba(3000, 2100, 6).numpy())               # a second group
    images.append(synth.synth_rgba(300, 64, 7).numpy()[:, 10:270])       # strided rows
    images.append(synth.synth_rgba(5, 3, 8).numpy())
    ctx = native.Context()
    outs = ctx.batch_encode_host(images, opt)
    okw = {k: v for k, v in kw.items()}
    for im, o in zip(images, outs):
        h, w = im.shape[:2]
        if w * h <= 512 * 512:
            want = oracle.encode_image(np.ascontiguousarray(im), block_dim=dim, **okw)
        else:
            want = _device_encode(native, im, opt)
        assert np.array_equal(o, want), (w, h)
    # again into caller-provided outputs, same context (staging and workspace reused)
    outs2 = [np.zeros_like(o) for o in outs]
    ctx.batch_encode_host(images, opt, outs=outs2)
    assert all(np.array_equal(a, b) for a, b in zip(outs, outs2))
    assert ctx.batch_encode_host([], opt) == []
    ctx.close()


def test_batch_encode_host_pinned_and_errors(native):
    import torch
    from astc_encoder_b200 import synth
    opt = native.encode_option(has_alpha=True)
    src = synth.synth_rgba(1024, 1024, 3)
    pinned = torch.empty((1024, 1024, 4), dtype=torch.uint8, pin_memory=True)
    pinned.copy_(src)
    ctx = native.Context()
    a = ctx.batch_encode_host([pinned.numpy()], opt)[0]
    b = ctx.encode_host(src.numpy(), opt)
    assert np.array_equal(a, b)
    # pinned and pageable, large and small, mixed in one batch -- and more groups than there are slots (3)
    mixed, want = [], []
    for i in range(14):
        size = (1536, 96, 2048, 40)[i % 4]
        t = synth.synth_rgba(size, size + 8 * i, 30 + i)
        if i % 3 == 0:
            p = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
            p.copy_(t)
            mixed.append(p.numpy())
        else:
            mixed.append(t.numpy())
        want.append(_device_encode(native, t.numpy(), opt))
    outs = [torch.empty(w.shape, dtype=torch.uint8, pin_memory=True).numpy() if i % 2 else np.empty_like(w) for i, w in enumerate(want)]
    ctx.batch_encode_host(mixed, opt, outs=outs)
    assert all(np.array_equal(o, w) for o, w in zip(outs, want))
    L = native.lib()
    o = opt._abi()
    assert L.astc_b200_context_encode_host(None, src.numpy().ctypes.data, 8, 8, 32, C.byref(o), a.ctypes.data) == -1
    assert L.astc_b200_context_encode_host(ctx._h, None, 8, 8, 32, C.byref(o), a.ctypes.data) == -1
    assert L.astc_b200_context_batch_encode_host(ctx._h, None, 3, C.byref(o)) == -1
    bad = opt._abi(); bad.axis_method = 7
    assert L.astc_b200_context_encode_host(ctx._h, src.numpy().ctypes.data, 8, 8, 32, C.byref(bad), a.ctypes.data) == -1
    ctx.close()


@pytest.mark.parametrize("dim", [4, 6])
def test_staged_pipeline_for_pageable_memory(native, dim):
    """Pageable (numpy) sources / destinations of >= 1 MiB go through the staged pipeline (worker threads copy bands
    into pinned slots): many bands, ragged last band, rows that are not a multiple of 16 bytes (padded device pitch),
    strided host rows, pinned-in / pageable-out and the reverse -- always the device path's bytes."""
    import torch
    from astc_encoder_b200 import synth
    ctx = native.Context()
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
    for (w, h) in ((4096, 4096), (1021, 1537), (2048, 130), (8192, 33), (515, 515)):
        img = synth.synth_rgba(w, h, 500 + w + h).numpy()
        want = _device_encode(native, img, opt)
        assert np.array_equal(ctx.encode_host(img, opt), want), (w, h)
        assert np.array_equal(native.encode_astc_host(img, opt), want), (w, h)            # default context
    img = synth.synth_rgba(3000, 1111, 9).numpy()
    want = _device_encode(native, img, opt)
    for threads in (0, 1, 6, -1):                                        # caller only ... automatic
        ctx.set_copy_threads(threads)
        assert np.array_equal(ctx.encode_host(img, opt), want), threads
    assert native.lib().astc_b200_context_set_copy_threads(ctx._h, 1000) == -1
    wide = synth.synth_rgba(1100, 700, 3).numpy()
    view = wide[:, 37:1061]                                              # pitch > 4 * width, pageable
    assert np.array_equal(ctx.encode_host(view, opt), _device_encode(native, view, opt))
    # pinned in, pageable out -- and pageable in, pinned out
    src = synth.synth_rgba(2048, 1024, 8)
    pin = torch.empty((1024, 2048, 4), dtype=torch.uint8, pin_memory=True)
    pin.copy_(src)
    want = _device_encode(native, src.numpy(), opt)
    assert np.array_equal(ctx.encode_host(pin.numpy(), opt), want)
    out_pin = torch.empty(want.shape, dtype=torch.uint8, pin_memory=True)
    ctx.encode_host(src.numpy(), opt, out=out_pin.numpy())
    assert np.array_equal(out_pin.numpy(), want)
    ctx.close()


@pytest.mark.parametrize("dim,kw", [(4, dict()), (6, dict(has_alpha=True, srgb=True)), (4, dict(is_normal_map=True))], ids=str)
def test_batch_encode_mip_chains_from_bases(native, dim, kw):
    """astc_b200_context_batch_encode_mip_chains_host: only the bases travel to the device; every level's blocks must
    equal the per-level encode of the chain made by mip_chain (fused launch for multiples of 64, per-level otherwise),
    for pinned and pageable bases, several groups, odd sizes and a 1x1 base."""
    import torch
    from astc_encoder_b200 import synth
    opt = native.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
    gen = synth.synth_normal if kw.get("is_normal_map") else synth.synth_rgba
    sizes = [(2048, 2048), (1024, 512), (250, 187), (64, 64), (1, 1), (4096, 2048), (2048, 2048), (3, 40), (4096, 4096), (128, 1024)]
    bases = []
    for i, (w, h) in enumerate(sizes):
        t = gen(w, h, 300 + i)
        if i % 3 == 1:
            p = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
            p.copy_(t)
            bases.append(p.numpy())
        else:
            bases.append(t.numpy())
    ctx = native.Context()
    for rep in range(2):                                                # the second call reuses workspace, slots and tickets
        got = ctx.batch_encode_mip_chains_host(bases, opt)
        for b, levels in zip(bases, got):
            chain = native.mip_chain(torch.from_numpy(np.ascontiguousarray(b)).cuda())
            assert len(levels) == len(chain), b.shape
            for lv, blocks in zip(chain, levels):
                assert np.array_equal(blocks, native.read_gpu(native.encode_astc(lv, opt))), (b.shape, tuple(lv.shape), rep)
    assert ctx.batch_encode_mip_chains_host([], opt) == []
    ctx.close()


def test_batch_encode_host_levels_back_to_back_in_one_buffer(native):
    """Levels that lie back to back in the caller's memory (a chain loaded from one file) are uploaded as ONE copy where
    they are contiguous in the device arena too (sizes that are multiples of 256 bytes): same bytes as separate arrays,
    from pinned and from pageable memory, with an odd-sized stranger in the middle."""
    import torch
    from astc_encoder_b200 import synth
    opt = native.encode_option(has_alpha=True)
    chains = [[t.cpu().numpy() for t in native.mip_chain(synth.synth_rgba(s, s, 40 + i).cuda())] for i, s in enumerate((1024, 2048, 512))]
    odd = synth.synth_rgba(250, 187, 9).numpy()
    flat = [lv for ch in chains[:2] for lv in ch] + [odd] + chains[2]
    total = sum(lv.nbytes for lv in flat)
    for pinned in (True, False):
        buf = torch.empty(total, dtype=torch.uint8, pin_memory=pinned).numpy()
        views, off = [], 0
        for lv in flat:
            v = buf[off:off + lv.nbytes].reshape(lv.shape)
            v[...] = lv
            views.append(v)
            off += lv.nbytes
        ctx = native.Context()
        got = ctx.batch_encode_host(views, opt)
        want = ctx.batch_encode_host(flat, opt)                           # separate arrays: nothing to merge
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        assert np.array_equal(got[0], _device_encode(native, flat[0], opt))
        ctx.close()


def test_contexts_are_independent_across_host_threads(native):
    """re-entrant per (device, stream), no global mutable state.  Six host threads, each with its own context
    (plus the thread-local default one), encode different textures concurrently -- pinned-free pageable sources, so the
    copy workers of six pools run at the same time too; every result must equal the single-threaded one."""
    import threading
    from astc_encoder_b200 import synth
    opts = [native.encode_option(), native.encode_option(is6x6=True, has_alpha=True, srgb=True), native.encode_option(is_normal_map=True)]
    jobs = []
    for i in range(6):
        w, h = 700 + 131 * i, 900 - 77 * i
        img = (synth.synth_normal if i % 3 == 2 else synth.synth_rgba)(w, h, 800 + i).numpy()
        jobs.append((img, opts[i % 3]))
    want = [_device_encode(native, img, opt) for img, opt in jobs]
    errors = []

    def work(k):
        try:
            img, opt = jobs[k]
            ctx = native.Context()
            for rep in range(8):
                if not np.array_equal(ctx.encode_host(img, opt), want[k]):
                    errors.append(("context", k, rep))
                if not np.array_equal(native.encode_astc_host(img, opt), want[k]):
                    errors.append(("default context", k, rep))
                outs = ctx.batch_encode_host([img, img[: img.shape[0] // 2]], opt)
                if not np.array_equal(outs[0], want[k]):
                    errors.append(("batch", k, rep))
            ctx.close()
        except Exception as e:                                         # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
