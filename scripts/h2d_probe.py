"""PCIe probe on the GPU box: pinned H2D / D2H bandwidth vs transfer size, chunking and stream count."""
import torch

dev = torch.device("cuda", 0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mb in (1, 2, 4, 8, 16, 22.5, 32, 64, 128):
    n = int(mb * 1e6 / 8)
    h = torch.zeros(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device=dev)
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    ms2 = timed(lambda: h.copy_(d, non_blocking=True))
    print(f"size {mb:6.1f} MB: h2d {ms:.3f} ms {mb / ms:5.1f} GB/s | d2h {ms2:.3f} ms {mb / ms2:5.1f} GB/s")

mb = 22.5
n = int(mb * 1e6 / 8)
h = torch.zeros(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device=dev)
for parts in (1, 2, 4, 8, 16):
    step = (n + parts - 1) // parts

    def chunked():
        for p in range(parts):
            d[p * step:(p + 1) * step].copy_(h[p * step:(p + 1) * step], non_blocking=True)
    ms = timed(chunked)
    print(f"22.5 MB in {parts:2d} chunks on one stream: {ms:.3f} ms {mb / ms:5.1f} GB/s")
streams = [torch.cuda.Stream(dev) for _ in range(4)]
for parts in (2, 4):
    step = (n + parts - 1) // parts

    def multi():
        for p in range(parts):
            with torch.cuda.stream(streams[p]):
                d[p * step:(p + 1) * step].copy_(h[p * step:(p + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    import time
    for _ in range(3):
        multi()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        multi()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 100
    print(f"22.5 MB in {parts} chunks on {parts} streams: {ms:.3f} ms {mb / ms:5.1f} GB/s")
# a kernel that reads pinned host memory directly (zero-copy) instead of the copy engine
hm = torch.zeros(n, dtype=torch.float64).pin_memory()
import ctypes
cudart = torch.cuda.cudart()
ms = timed(lambda: d.copy_(hm.cuda(non_blocking=True)))
print(f"22.5 MB .cuda(): {ms:.3f} ms")
