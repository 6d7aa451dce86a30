"""Ad-hoc kernel timing used during development (not the contract bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth

def run(w, h, opt, label, iters=10):
    img = synth.synth_rgba(w, h, 1234, device="cuda")
    out = A.encode_astc(img, opt)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for _ in range(3):
        A.encode_astc(img, opt, out=out)
    ev[0].record()
    for i in range(iters):
        A.encode_astc(img, opt, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    med = ts[len(ts) // 2]
    chk = int(out.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFF   # compare across experiment builds
    print(f"{label}: {w}x{h} median {med:.3f} ms best {ts[0]:.3f} ms -> {w*h/med/1e6:.1f} Gtexel/s  chk {chk:012x}", flush=True)

if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "mip"):
    if len(sys.argv) > 1 and sys.argv[1] == "short":
        run(16384, 16384, A.encode_option(), "4x4 rgb")
        run(4096, 4096, A.encode_option(has_alpha=True), "4x4 rgba")
        run(8192, 8192, A.encode_option(is6x6=True, has_alpha=True, srgb=True), "6x6 rgba srgb")
        run(8192, 8192, A.encode_option(is6x6=True), "6x6 rgb")
        run(8192, 8192, A.encode_option(is6x6=True, is_normal_map=True), "6x6 norm")
        run(4096, 4096, A.encode_option(is6x6=True, has_alpha=True), "6x6 rgba 4k")
        run(4096, 4096, A.encode_option(srgb=True), "4x4 rgb srgb")
        run(4096, 4096, A.encode_option(is_normal_map=True), "4x4 norm")
        run(16384, 16384, A.encode_option(is_normal_map=True), "4x4 norm")
        sys.exit(0)
    print(A.version(), torch.cuda.get_device_name(0))
    run(4096, 4096, A.encode_option(), "4x4 rgb")
    run(4096, 4096, A.encode_option(has_alpha=True), "4x4 rgba")
    run(4096, 4096, A.encode_option(srgb=True), "4x4 rgb srgb")
    run(4096, 4096, A.encode_option(is_normal_map=True), "4x4 norm")
    run(16384, 16384, A.encode_option(), "4x4 rgb")
    run(8192, 8192, A.encode_option(is6x6=True, has_alpha=True, srgb=True), "6x6 rgba srgb")
    run(8192, 8192, A.encode_option(is6x6=True), "6x6 rgb")


def run_mip(w, h, iters=10):
    img = synth.synth_rgba(w, h, 99, device="cuda")
    out = A.downsample2x2(img)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for _ in range(3):
        A.downsample2x2(img, out=out)
    ev[0].record()
    for i in range(iters):
        A.downsample2x2(img, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    med = ts[len(ts) // 2]
    print(f"mip 2x2: {w}x{h} median {med:.3f} ms -> {(w * h * 4 + out.numel()) / med / 1e6:.0f} GB/s (read + write)", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mip":
    run_mip(16384, 16384)
    run_mip(4096, 4096)
    run_mip(2048, 2048)
