"""D2H bandwidth of this box into pinned memory (the e2e floor: 3.27 GB of eigenvectors per step)."""
import time, torch
n = 408 * 1000 * 1000
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
for chunks in (1, 8, 32):
    step = n // chunks
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for c in range(chunks):
            h[c * step:(c + 1) * step].copy_(d[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("chunks", chunks, "D2H %.1f ms  %.1f GB/s" % (1e3 * dt, 8 * n / dt / 1e9), flush=True)
# two pinned buffers alternately (as the pipeline does)
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for rep in range(4):
    (h if rep % 2 == 0 else h2).copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("4 x alternating buffers: %.1f ms per copy" % (1e3 * dt / 4))
# while a bandwidth-heavy kernel runs
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
s2 = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s2):
    for _ in range(40): b.copy_(a)
h.copy_(d, non_blocking=True)
torch.cuda.current_stream().synchronize(); dt = time.perf_counter() - t0
print("D2H under HBM-bound kernels: %.1f ms  %.1f GB/s" % (1e3 * dt, 8 * n / dt / 1e9))
torch.cuda.synchronize()
