"""How should a pageable numpy population reach the GPU?  (a) threaded copy into a pinned staging buffer,
(b) cudaHostRegister the caller's array in place, (c) plain pageable cudaMemcpy."""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

B, N = 65536, 43
a = np.random.rand(B, N)
dev = torch.device("cuda", 0)
d = torch.empty((B, N), dtype=torch.float64, device=dev)
stage = torch.empty((B, N), dtype=torch.float64).pin_memory()
snp = stage.numpy()
for nt in (1, 2, 4, 8, 16):
    ex = ThreadPoolExecutor(nt)
    step = (B + nt - 1) // nt

    def run():
        futs = [ex.submit(np.copyto, snp[lo:lo + step], a[lo:lo + step]) for lo in range(0, B, step)]
        for f in futs:
            f.result()
    run()
    t0 = time.perf_counter()
    for _ in range(20):
        run()
    dt = (time.perf_counter() - t0) / 20
    print(f"(a) staging copy, {nt:2d} threads: {dt * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(20):
    stage.copy_(torch.from_numpy(a))
print(f"(a') torch copy_ into pinned: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
rt = torch.cuda.cudart()
ptr, nbytes = a.ctypes.data, a.nbytes
for _ in range(3):
    rt.cudaHostRegister(ptr, nbytes, 0); rt.cudaHostUnregister(ptr)
t0 = time.perf_counter()
for _ in range(20):
    rt.cudaHostRegister(ptr, nbytes, 0)
    t1 = time.perf_counter()
    rt.cudaHostUnregister(ptr)
dt = (time.perf_counter() - t0) / 20
print(f"(b) cudaHostRegister + Unregister of 22.5 MB: {dt * 1e3:.3f} ms")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(torch.from_numpy(a))
torch.cuda.synchronize()
print(f"(c) pageable H2D copy: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
