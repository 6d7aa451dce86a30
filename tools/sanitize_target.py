"""A small slice of the GPU parity suite, sized to run under compute-sanitizer (memcheck / racecheck /
synccheck / initcheck are 10-100x slower than native): every kernel variant, both block sizes, aligned +
ragged + unaligned sources, a mip-chain batch, the host band pipeline, decode, downsample, BISE.
Each result is still compared with the oracle, so a sanitizer run is also a parity run.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth
from oracle import oracle as O


def main() -> int:
    O.lib()
    checked = 0
    variants = [dict(), dict(has_alpha=True), dict(srgb=True), dict(has_alpha=True, srgb=True), dict(is_normal_map=True),
                dict(is_normal_map=True, has_alpha=True)]
    for dim in (4, 6):
        for kw in variants:
            opt = A.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw)
            for (w, h) in ((256, 192), (250, 187), (5, 3)):
                img = (synth.synth_normal if kw.get("is_normal_map") else synth.synth_rgba)(w, h, 7 + dim + w)
                got = A.read_gpu(A.encode_astc(img.cuda(), opt))
                want = O.encode_image(img.numpy(), block_dim=dim, has_alpha=opt.has_alpha, is_normal_map=opt.is_normal_map, srgb=opt.srgb)
                assert np.array_equal(got, want), (dim, kw, w, h)
                checked += len(want)
        # unaligned base + padded pitch (per-texel path)
        opt = A.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
        wide = synth.synth_rgba(200, 64, 99).cuda()
        view = wide[:, 1:190]
        got = A.read_gpu(A.encode_astc(view, opt))
        want = O.encode_image(np.ascontiguousarray(view.cpu().numpy()), block_dim=dim, has_alpha=True)
        assert np.array_equal(got, want)
        checked += len(want)
        # a mip-chain batch in one launch
        base = synth.synth_rgba(128, 128, 5).cuda()
        chain = A.mip_chain(base)                                   # ONE fused launch (128 = 2 x 64): tiles, ticket, last-CTA tail
        assert all(torch.equal(a, b) for a, b in zip(chain, A.mip_chain_by_level(base)))
        odd = synth.synth_rgba(192, 320, 6).cuda()
        assert all(torch.equal(a, b) for a, b in zip(A.mip_chain(odd), A.mip_chain_by_level(odd)))
        batch = A.Batch(chain, opt)
        outs = batch.encode()
        torch.cuda.synchronize()
        for lvl, o in zip(chain, outs):
            want = O.encode_image(lvl.cpu().numpy(), block_dim=dim, has_alpha=True)
            assert np.array_equal(o.cpu().numpy(), want)
            checked += len(want)
        batch.close()
        # host band pipeline (three streams)
        himg = synth.synth_rgba(512, 384, 3).numpy()
        got = A.encode_astc_host(himg, opt)
        want = O.encode_image(himg, block_dim=dim, has_alpha=True)
        assert np.array_equal(got, want)
        checked += len(want)
        big = synth.synth_rgba(1024, 520, 4).numpy()                # pageable and > 1 MiB: the staged pipeline (copy workers, pinned slots)
        assert np.array_equal(A.encode_astc_host(big, opt), A.read_gpu(A.encode_astc(torch.from_numpy(big).cuda(), opt)))
        dec = A.decode_astc(torch.from_numpy(want).cuda(), 512, 384, dim).cpu().numpy()
        ref, nbad = O.decode_image(want, 512, 384, dim)
        assert nbad == 0 and np.array_equal(dec, ref)
    vals = torch.randint(0, 6, (64, 16), dtype=torch.uint8, device="cuda")
    A.bise_encode(vals, 4)
    torch.cuda.synchronize()
    print(f"sanitize_target ok: {checked} blocks bit-exact vs the oracle, {A.launch_count()} launches")
    return 0


if __name__ == "__main__":
    sys.exit(main())
