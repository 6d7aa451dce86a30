"""Randomised differential soak: many textures of random size and content through the CUDA path (single launches, a
batch launch and the host entry points), every variant x block size x axis method, each compared bit for bit with the
CPU oracle.   python tools/soak.py [cases] [seed]      (under gpurun)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from oracle import oracle as O


def content(rng, w, h):
    kind = rng.integers(0, 7)
    if kind == 0:
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    elif kind == 1:                                                   # smooth gradient + noise
        y, x = np.mgrid[0:h, 0:w]
        base = [(x * rng.uniform(0, 1.5) + y * rng.uniform(0, 1.5) + rng.uniform(0, 255)) % 256 for _ in range(4)]
        img = (np.stack(base, -1) + rng.normal(0, rng.uniform(0, 12), (h, w, 4))).clip(0, 255).astype(np.uint8)
    elif kind == 2:                                                   # flat, or flat with a few outliers
        img = np.tile(rng.integers(0, 256, (1, 1, 4), dtype=np.uint8), (h, w, 1))
        n = int(rng.integers(0, 8))
        img[rng.integers(0, h, n), rng.integers(0, w, n)] = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    elif kind == 3:                                                   # two-tone
        a, b = rng.integers(0, 256, (2, 4), dtype=np.uint8)
        img = np.where(rng.random((h, w, 1)) < rng.uniform(0.05, 0.95), a, b).astype(np.uint8)
    elif kind == 4:                                                   # saturated / extreme values
        img = rng.choice(np.array([0, 1, 254, 255], np.uint8), (h, w, 4))
    elif kind == 5:                                                   # one channel varies
        img = np.tile(rng.integers(0, 256, (1, 1, 4), dtype=np.uint8), (h, w, 1))
        img[..., rng.integers(0, 4)] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    else:                                                             # low-amplitude noise around a colour (near-degenerate covariance)
        img = (rng.integers(0, 256, (1, 1, 4)) + rng.integers(-2, 3, (h, w, 4))).clip(0, 255).astype(np.uint8)
    return np.ascontiguousarray(img)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261017
    rng = np.random.default_rng(seed)
    variants = [dict(), dict(has_alpha=True), dict(srgb=True), dict(has_alpha=True, srgb=True), dict(is_normal_map=True),
                dict(is_normal_map=True, has_alpha=True)]
    ctx = A.Context()
    blocks = 0
    t0 = time.time()
    for c in range(cases):
        dim = int(rng.choice([4, 6]))
        kw = dict(variants[int(rng.integers(0, len(variants)))])
        kw["axis_method"] = int(rng.integers(0, 2))
        opt = A.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
        okw = dict(block_dim=dim, has_alpha=opt.has_alpha, is_normal_map=opt.is_normal_map, srgb=opt.srgb, axis_method=opt.axis_method)
        n = int(rng.integers(1, 5))
        imgs = [content(rng, int(rng.integers(1, 400)), int(rng.integers(1, 400))) for _ in range(n)]
        want = [O.encode_image(i, **okw) for i in imgs]
        # half of the textures are strided sub-views of a larger surface (pitch > 4 * width, base at an arbitrary texel),
        # on the host and on the device
        dev = []
        for k in range(n):
            if rng.random() < 0.5:
                h, w = imgs[k].shape[:2]
                ox, px = int(rng.integers(0, 9)), int(rng.integers(0, 9))
                big = rng.integers(0, 256, (h, w + ox + px, 4), dtype=np.uint8)
                big[:, ox:ox + w] = imgs[k]
                imgs[k] = big[:, ox:ox + w]
                dev.append(torch.from_numpy(big).cuda()[:, ox:ox + w])
            else:
                imgs[k] = np.ascontiguousarray(imgs[k])
                dev.append(torch.from_numpy(imgs[k]).cuda())
        for i, d, w in zip(imgs, dev, want):
            got = A.read_gpu(A.encode_astc(d, opt))
            assert np.array_equal(got, w), ("single", c, dim, kw, i.shape)
            assert np.array_equal(ctx.encode_host(i, opt), w), ("host", c, dim, kw, i.shape)
            blocks += len(w)
        b = A.Batch(dev, opt)
        b.encode()
        torch.cuda.synchronize()
        for o, w in zip(b.outputs, want):
            assert np.array_equal(o.cpu().numpy(), w), ("batch", c, dim, kw)
        b.close()
        for o, w in zip(ctx.batch_encode_host(imgs, opt), want):
            assert np.array_equal(o, w), ("batch host", c, dim, kw)
    print(f"soak ok: {cases} cases, {blocks} blocks x 4 entry points bit-exact vs the oracle, seed {seed}, {time.time() - t0:.0f} s, {A.version()}")


if __name__ == "__main__":
    main()
