"""CUDA path vs the CPU oracle, through the C ABI (ctypes).  Bar: bit-exact blocks
(SURVEY.md 8c requires >= 99.9 % identical; 100 % is expected and asserted)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

VARIANTS = [
    dict(has_alpha=False, is_normal_map=False, srgb=False),
    dict(has_alpha=True, is_normal_map=False, srgb=False),
    dict(has_alpha=False, is_normal_map=False, srgb=True),
    dict(has_alpha=True, is_normal_map=False, srgb=True),
    dict(has_alpha=False, is_normal_map=True, srgb=False),
    dict(has_alpha=True, is_normal_map=True, srgb=True),      # srgb must be ignored for normal maps
]


def _opt(native, dim, v):
    return native.encode_option(is4x4=(dim == 4), is6x6=(dim == 6), **v)


def _gpu_encode(native, img, opt):
    import torch
    src = torch.from_numpy(img).cuda()
    out = native.encode_astc(src, opt)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _mismatch_report(a, b):
    bad = np.nonzero((a != b).any(axis=1))[0]
    return f"{len(bad)} of {len(a)} blocks differ, first {bad[:8].tolist()}"


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("v", VARIANTS, ids=lambda v: "-".join(k for k, x in v.items() if x) or "rgb")
@pytest.mark.parametrize("size", [(256, 192), (250, 187), (37, 23)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_variants_match_oracle(native, oracle, dim, v, size):
    from astc_encoder_b200 import synth
    w, h = size
    if v["is_normal_map"]:
        img = synth.synth_normal(w, h, synth.SEED_CFG4).numpy()
    else:
        img = synth.synth_rgba(w, h, synth.SEED_CFG2 + dim).numpy()
    got = _gpu_encode(native, img, _opt(native, dim, v))
    want = oracle.encode_image(img, block_dim=dim, **v)
    assert got.shape == want.shape
    assert (got == want).all(), _mismatch_report(got, want)


def test_leaf_alpha_4x4_vs_oracle_and_golden(native, oracle, leaf_rgba, leaf_golden):
    xd, yd, xs, ys, gold = leaf_golden
    assert (xd, yd, xs, ys) == (4, 4, 1024, 1024)
    opt = native.encode_option(has_alpha=True)
    got = _gpu_encode(native, leaf_rgba, opt)
    want = oracle.encode_image(leaf_rgba, block_dim=4, has_alpha=True)
    assert (got == want).all(), _mismatch_report(got, want)
    same = (got == gold).all(axis=1).mean()
    assert same >= 0.995, f"only {same:.4%} of blocks match the reference's golden leaf.astc"
