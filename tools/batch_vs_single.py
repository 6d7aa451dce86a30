import sys, torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
opt = A.encode_option()
img = synth.synth_rgba(16384, 16384, 5, device="cuda")
out = A.encode_astc(img, opt)
print("single-texture kernel, 16384^2:", round(t(lambda: A.encode_astc(img, opt, out=out)), 4), "ms")
b = A.Batch([img], opt, outputs=[out])
print("batch kernel, ONE 16384^2 image:", round(t(lambda: b.encode()), 4), "ms")
# 64 x 2048^2 (same texel count), batch vs 64 single launches
imgs = [synth.synth_rgba(2048, 2048, 100 + i, device="cuda") for i in range(64)]
b2 = A.Batch(imgs, opt)
print("batch kernel, 64 x 2048^2:", round(t(lambda: b2.encode()), 4), "ms")
outs = [o for o in b2.outputs]
print("64 single launches of 2048^2:", round(t(lambda: [A.encode_astc(i, opt, out=o) for i, o in zip(imgs, outs)]), 4), "ms")
# one full mip chain set: 49 chains of 2048 (about the same texels as 64 x 2048^2 * 4/3)
chains = []
for i in range(48):
    chains.extend(A.mip_chain(imgs[i]))
b3 = A.Batch(chains, opt)
tex = sum(int(c.shape[0]) * int(c.shape[1]) for c in chains)
ms = t(lambda: b3.encode())
print(f"batch kernel, 48 mip chains ({tex/1e6:.1f} Mtexel):", round(ms, 4), "ms ->", round(tex / ms / 1e6, 1), "Gtexel/s")
# levels only >= 64
big = [c for c in chains if c.shape[0] >= 64]
b4 = A.Batch(big, opt)
tex = sum(int(c.shape[0]) * int(c.shape[1]) for c in big)
ms = t(lambda: b4.encode())
print(f"batch kernel, the levels >= 64^2 of the same chains ({tex/1e6:.1f} Mtexel):", round(ms, 4), "ms ->", round(tex / ms / 1e6, 1), "Gtexel/s")
