"""Tiny driver for ncu: launches one encode configuration a few times.
    python tools/profile_target.py 4x4rgb16k|4x4rgb4k|4x4rgba4k|6x6rgba8k|6x6rgb8k|norm4k [iters]"""
import sys
import torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth

CFG = {
    "4x4rgb16k": (16384, 16384, synth.SEED_CFG5, A.encode_option(), False),
    "4x4rgb4k": (4096, 4096, synth.SEED_CFG2, A.encode_option(), False),
    "4x4rgba4k": (4096, 4096, synth.SEED_CFG2, A.encode_option(has_alpha=True), False),
    "6x6rgba8k": (8192, 8192, synth.SEED_CFG3, A.encode_option(is6x6=True, has_alpha=True, srgb=True), False),
    "6x6rgb8k": (8192, 8192, synth.SEED_CFG3, A.encode_option(is6x6=True), False),
    "norm4k": (4096, 4096, synth.SEED_CFG4, A.encode_option(is_normal_map=True), True),
}

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "4x4rgb16k"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    w, h, seed, opt, normal = CFG[name]
    img = (synth.synth_normal if normal else synth.synth_rgba)(w, h, seed, device="cuda")
    out = A.encode_astc(img, opt)
    for _ in range(iters):
        A.encode_astc(img, opt, out=out)
    torch.cuda.synchronize()
    print(name, "done", A.launch_count(), "launches")
