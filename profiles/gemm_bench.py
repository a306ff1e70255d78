#!/usr/bin/env python
"""FP64 contraction kernel on the cfg5 shapes: single pair (1000^3, bspatom_dipole: 64 or 128 CTAs) and the chain of
50 pairs (bspatom_dipole_chain_resident), device time of the two kernels (band x dense + TN GEMM), against the cuBLAS
figures of profiles/fp64_peak.json."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from bench import workload

atom = bsp.BspAtom(device=0)
wl = workload("cfg2", bsp, 0, 1, 1, "lin")
items, n, kd = wl["items"], 1000, 6
band = atom.MATRIX_SVT(items[0][0])
Rb = np.zeros((2 * kd + 1, n), order="F")
for d in range(kd + 1):
    Rb[kd - d, d:] = band["R"][kd - d, d:]
    if d:
        Rb[kd + d, :n - d] = band["R"][kd - d, d:]
atom.batch_upload(items); atom.batch_run()
out = {}
if len(sys.argv) > 1:
    atom.set_option("gemm_variant", float(sys.argv[1])); out["variant"] = int(sys.argv[1])
fl_pair = 2.0 * n ** 3 + 2.0 * (2 * kd + 1) * n * n
for _ in range(3):
    D = atom.dipole_chain_resident(Rb, 0, 51, n)
ms = []
for _ in range(10):
    atom.dipole_chain_resident(Rb, 0, 51, n); ms.append(atom.stats()["ms_resident_contraction"])
out["chain50_ms"] = min(ms); out["chain50_tflops"] = 50 * fl_pair / (min(ms) * 1e-3) / 1e12
ms = []
for _ in range(13):
    atom.dipole_chain_resident(Rb, 0, 2, n); ms.append(atom.stats()["ms_resident_contraction"])
out["single_ms"] = min(ms[3:]); out["single_tflops"] = fl_pair / (min(ms[3:]) * 1e-3) / 1e12
for nv in (500, 300):
    ms = []
    for _ in range(8):
        atom.dipole_chain_resident(Rb, 0, 51, nv); ms.append(atom.stats()["ms_resident_contraction"])
    f = 50 * (2.0 * n * nv * nv + 2.0 * (2 * kd + 1) * n * nv)
    out["chain50_nvec%d_tflops" % nv] = f / (min(ms[2:]) * 1e-3) / 1e12
try:
    pk = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
    out["cublas_dgemm_8192"] = pk["dgemm_8192_tflops"]; out["cublas_tn_batched_50x1000"] = pk["dgemm_tn_batched_50x1000_tflops"]
    out["cublas_tn_single_1000"] = pk["dgemm_tn_single_1000_tflops"]
    out["chain_over_cublas_peak"] = out["chain50_tflops"] / pk["dgemm_8192_tflops"]
    out["single_over_cublas_peak"] = out["single_tflops"] / pk["dgemm_8192_tflops"]
except Exception:
    pass
Cs = None
print(json.dumps(out))
