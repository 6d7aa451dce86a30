import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
ho = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True); do = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
ms = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1 GiB: {ms:.2f} ms {n/ms/1e6:.1f} GB/s")
ms = t(lambda: ho.copy_(do, non_blocking=True)); print(f"D2H 256 MiB: {ms:.2f} ms {n/4/ms/1e6:.1f} GB/s")
s2 = torch.cuda.Stream()
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
ms = t(both); print(f"H2D 1 GiB || D2H 256 MiB: {ms:.2f} ms")
# chunked 8 MiB
def chunked():
    c = 8 << 20
    for i in range(0, n, c): d[i:i+c].copy_(h[i:i+c], non_blocking=True)
ms = t(chunked); print(f"H2D 1 GiB in 8 MiB chunks: {ms:.2f} ms {n/ms/1e6:.1f} GB/s")
