"""Throughput of the subset decoder (decode_kernel): blocks -> RGBA8.   python tools/decode_time.py   (under gpurun)"""
import sys, torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth
for dim, w in ((4, 16384), (6, 8192), (4, 4096)):
    opt = A.encode_option(is4x4=dim == 4, is6x6=dim == 6, has_alpha=True)
    img = synth.synth_rgba(w, w, 3, device="cuda")
    blocks = A.encode_astc(img, opt)
    out = A.decode_astc(blocks, w, w, dim)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        A.decode_astc(blocks, w, w, dim)
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 10
    byts = blocks.numel() + w * w * 4
    print(f"decode {dim}x{dim} {w}^2: {ms:.3f} ms (incl. the output allocation) -> {w * w / ms / 1e6:.0f} Gtexel/s, {byts / ms / 1e6:.0f} GB/s (read + write)")
