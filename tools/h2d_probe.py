"""Host-side limits of the upload path on this box: pageable -> pinned memcpy rate by thread count, pinned -> device DMA
rate, pageable -> device through the driver, and the CPU quota the container really has."""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

print("affinity cpus:", len(os.sched_getaffinity(0)), "os.cpu_count:", os.cpu_count())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try:
        print(f, open(f).read().strip())
    except OSError:
        pass
n = 32 << 20
src = np.random.default_rng(0).integers(0, 255, n, dtype=np.uint8)
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
pin_np = pin.numpy()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for threads in (1, 2, 4, 8, 16, 32):
    ex = ThreadPoolExecutor(threads)
    chunk = n // threads
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        list(ex.map(lambda i: np.copyto(pin_np[i * chunk:(i + 1) * chunk], src[i * chunk:(i + 1) * chunk]), range(threads)))
        best = min(best, time.perf_counter() - t0)
    print(f"memcpy pageable->pinned 32 MiB, {threads:2d} threads: {n / best / 1e9:6.1f} GB/s ({best * 1e3:.3f} ms)")
    ex.shutdown()
torch.cuda.synchronize()
for what, s in (("pinned", pin), ("pageable", torch.from_numpy(src))):
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.copy_(s, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"H2D 32 MiB from {what}: {n / best / 1e9:6.1f} GB/s ({best * 1e3:.3f} ms)")
best = 1e9
host = torch.empty(n, dtype=torch.uint8).pin_memory()
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print(f"D2H 32 MiB to pinned: {n / best / 1e9:6.1f} GB/s ({best * 1e3:.3f} ms)")
