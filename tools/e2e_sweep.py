"""End-to-end time of astc_b200_encode_host on 16384^2 against the band size (ASTC_B200_HOST_BAND_MIB)."""
import os, sys, time
import torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from astc_encoder_b200 import synth

W = int(os.environ.get("E2E_W", "16384"))
opt = A.encode_option()
tex = synth.synth_rgba(W, W, synth.SEED_CFG5, device="cuda")
print(f"{W}x{W}", flush=True)
h_in = torch.empty((W, W, 4), dtype=torch.uint8, pin_memory=True)
h_in.copy_(tex)
h_out = torch.empty((A.output_size(W, W, opt) // 16, 16), dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
i_np, o_np = h_in.numpy(), h_out.numpy()
for mib in [int(a) for a in sys.argv[1:]] or [0, 2, 4, 8, 16, 32, 64]:
    if mib:
        os.environ["ASTC_B200_HOST_BAND_MIB"] = str(mib)
    else:
        os.environ.pop("ASTC_B200_HOST_BAND_MIB", None)      # 0 = the library's own choice
    A.encode_astc_host(i_np, opt, out=o_np)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        A.encode_astc_host(i_np, opt, out=o_np)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print(f"band {mib:3d} MiB: median {ts[len(ts)//2]:.3f} ms best {ts[0]:.3f} ms", flush=True)
