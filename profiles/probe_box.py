#!/usr/bin/env python
"""One-off probe of the GPU box (run under gpurun): toolchains for oracle/_ref, host cores, measured FP64
DGEMM peak (the roofline denominator SURVEY.md 8(d) asks for: cuBLAS DGEMM 8192^3, best of 10), and the
pinned D2H rate of the host (the end-to-end floor).  Writes profiles-style JSON files into gpurun_out/."""
import json
import os
import shutil
import subprocess
import sys
import time

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
os.makedirs(out_dir, exist_ok=True)

probe = {"nproc": os.cpu_count()}
for tool in ("gfortran", "flang", "ifort", "ifx", "nvfortran", "f2c", "lfortran", "f95", "f77"):
    probe[tool] = shutil.which(tool)
try:
    probe["gcc_f951"] = subprocess.run("ls /usr/lib/gcc/x86_64-linux-gnu/*/f951 2>/dev/null", shell=True,
                                       capture_output=True, text=True).stdout.strip() or None
    probe["liblapack"] = subprocess.run("ls /usr/lib/x86_64-linux-gnu/liblapack* /usr/lib/x86_64-linux-gnu/libopenblas* 2>/dev/null",
                                        shell=True, capture_output=True, text=True).stdout.split()
    probe["cpu_model"] = subprocess.run("grep -m1 'model name' /proc/cpuinfo", shell=True, capture_output=True,
                                        text=True).stdout.strip()
    probe["mem_gb"] = round(os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30, 1)
except Exception as exc:
    probe["error"] = repr(exc)
json.dump(probe, open(os.path.join(out_dir, "box_probe.json"), "w"), indent=1)
print("probe", json.dumps(probe))

import torch

assert torch.cuda.is_available()
dev = torch.device("cuda", 0)
res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
       "how": "torch.matmul float64 (cuBLAS DGEMM), 2 N^3 flops, CUDA events, best of 10 after 3 warm-ups"}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        c = a @ b
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res["dgemm_%d_tflops" % n] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    res["dgemm_%d_ms" % n] = best
    del a, b, c
# batched TN shape of cfg5: 50 x (1000 x 1000 x 1000), A^T B
a = torch.randn(50, 1000, 1000, dtype=torch.float64, device=dev)
b = torch.randn(50, 1000, 1000, dtype=torch.float64, device=dev)
for _ in range(3):
    c = torch.bmm(a.transpose(1, 2), b)
best = 1e30
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = torch.bmm(a.transpose(1, 2), b)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["dgemm_tn_batched_50x1000_tflops"] = 50 * 2.0e9 / (best * 1e-3) / 1e12
res["dgemm_tn_batched_50x1000_ms"] = best
a1, b1 = a[0].contiguous(), b[0].contiguous()
for _ in range(3):
    c = a1.t() @ b1
best = 1e30
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = a1.t() @ b1
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["dgemm_tn_single_1000_tflops"] = 2.0e9 / (best * 1e-3) / 1e12
res["dgemm_tn_single_1000_ms"] = best
del a, b, c, a1, b1
json.dump(res, open(os.path.join(out_dir, "fp64_peak.json"), "w"), indent=1)
print("fp64", json.dumps(res))

# pinned D2H / H2D rate of this host (single GPU)
n = 1 << 28   # 2 GiB of doubles
d = torch.empty(n, dtype=torch.float64, device=dev)
h = torch.empty(n, dtype=torch.float64).pin_memory()
pc = {}
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    best = 1e30
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    pc[name + "_gbs"] = 8 * n / best / 1e9
json.dump(pc, open(os.path.join(out_dir, "pcie_n1.json"), "w"), indent=1)
print("pcie", json.dumps(pc))
