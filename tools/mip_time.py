import sys, time, torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth
imgs = [synth.synth_rgba(2048, 2048, 100 + i, device="cuda") for i in range(32)]
for _ in range(2):
    for i in imgs: A.mip_chain_by_level(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for i in imgs: A.mip_chain_by_level(i)
b.record(); b.synchronize(); t1 = time.perf_counter()
print(f"level by level (2048^2), 11 launches per chain: {a.elapsed_time(b) / 32 * 1e3:.1f} us of GPU time per chain, {(t1 - t0) / 32 * 1e6:.1f} us wall per chain")
# preallocated outputs, C calls only
chains = [A.mip_chain_by_level(i) for i in imgs]
torch.cuda.synchronize()
a.record(); t0 = time.perf_counter()
for ch in chains:
    for l in range(len(ch) - 1): A.downsample2x2(ch[l], out=ch[l + 1])
b.record(); b.synchronize(); t1 = time.perf_counter()
print(f"level by level, preallocated: {a.elapsed_time(b) / 32 * 1e3:.1f} us GPU per chain, {(t1 - t0) / 32 * 1e6:.1f} us wall per chain")
print("ideal at 6.2 TB/s: 16 MiB read + 5.33 MiB written =", round((16 + 5.33) * 1.048576 / 6.2e3 * 1e3, 2), "us")

# the whole chain in one call (one fused launch), arenas preallocated
_, _, _, total = A.mip_chain_layout(2048, 2048)
arenas = [torch.empty(total, dtype=torch.uint8, device="cuda") for _ in imgs]
for i, ar in zip(imgs, arenas): A.mip_chain(i, arena=ar)
torch.cuda.synchronize()
a.record(); t0 = time.perf_counter()
for i, ar in zip(imgs, arenas): A.mip_chain(i, arena=ar)
b.record(); b.synchronize(); t1 = time.perf_counter()
print(f"astc_b200_mip_chain_device (ONE fused launch): {a.elapsed_time(b) / 32 * 1e3:.1f} us GPU per chain, {(t1 - t0) / 32 * 1e6:.1f} us wall per chain")
big = synth.synth_rgba(16384, 16384, 3, device="cuda")
_, _, _, total = A.mip_chain_layout(16384, 16384)
ar = torch.empty(total, dtype=torch.uint8, device="cuda")
A.mip_chain(big, arena=ar); torch.cuda.synchronize()
a.record()
for _ in range(5): A.mip_chain(big, arena=ar)
b.record(); b.synchronize()
ms = a.elapsed_time(b) / 5
print(f"16384^2 chain (14 levels) in one launch: {ms:.3f} ms = {(16384 * 16384 * 4 * (1 + 1 / 3)) / ms / 1e6:.0f} GB/s (read + write)")
# GPU time of the fused launch alone: direct C-ABI calls (no Python per-level work), 2048^2
L = A.lib()
s = torch.cuda.current_stream().cuda_stream
args = [(i.data_ptr(), 2048, 2048, 2048 * 4, ar.data_ptr(), ar.numel(), s) for i, ar in zip(imgs, arenas)]
for x in args: L.astc_b200_mip_chain_device(*x)
torch.cuda.synchronize()
a.record(); t0 = time.perf_counter()
for _ in range(4):
    for x in args: L.astc_b200_mip_chain_device(*x)
b.record(); b.synchronize(); t1 = time.perf_counter()
print(f"direct C calls: {a.elapsed_time(b) / 128 * 1e3:.1f} us GPU per 2048^2 chain, {(t1 - t0) / 128 * 1e6:.1f} us wall per call")
