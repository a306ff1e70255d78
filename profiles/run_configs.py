#!/usr/bin/env python
"""Throughput + property checks of the BASELINE.json configs that are not the bench headline
(cfg3, cfg4, cfg5) on one B200.  Writes one JSON line per config to stdout.

    python profiles/run_configs.py [cfg3] [cfg4] [cfg5]

Parity against the oracle for these configs lives in tests/test_gpu_*.py (sampled sizes); here the
full sizes are timed and checked through size-independent properties (ascending spectrum, residual,
S-orthonormality on a sample of vectors, hydrogen levels, dipole sum rule)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp  # noqa: E402
from bspatom_b200.host import pinned_empty  # noqa: E402
from cases import band_to_dense_sym, cfg3_problems, host_basis  # noqa: E402


def timed_resident(atom, items, nvec=None, reps=3):
    atom.batch_upload(items, nvec=nvec)
    atom.batch_run()
    ms = []
    for _ in range(reps):
        atom.batch_run()
        ms.append(atom.stats()["ms_total"])
    return float(np.median(ms)), atom.stats()


def cfg3(atom):
    a, items = cfg3_problems(4096)
    ms, st = timed_resident(atom, items)
    n = a.nfun
    E = pinned_empty(len(items) * n)
    Cb = pinned_empty(len(items) * n * n)
    t0 = time.perf_counter()
    Es, Cs, info = atom.solve_batch(items, out_E=E, out_C=Cb)
    e2e = time.perf_counter() - t0
    assert not info.any(), info[info != 0][:10]
    for E_ in Es[::97]:
        assert np.all(np.diff(E_) > 0)
    # residual / orthonormality of a few problems on device-built bands
    worst_res = worst_orth = 0.0
    for i in (0, 1, 2051, 4095):
        p, l = items[i]
        band = atom.MATRIX_SVT(p)
        S = band_to_dense_sym(band["S"], n)
        H = band_to_dense_sym(band["H0"], n) + (l * (l + 1)) * band_to_dense_sym(band["Q"], n)
        Cm, Ev = Cs[i], Es[i]
        SC = S @ Cm
        worst_orth = max(worst_orth, np.abs(Cm.T @ SC - np.eye(n)).max())
        worst_res = max(worst_res, (np.abs(H @ Cm - SC * Ev).max(0) / np.maximum(1, np.abs(Ev))).max())
    return {"config": "cfg3: 4096 Yukawa/Tietz problems, N=500, k=7, all eigenpairs", "solves": len(items),
            "resident_ms": ms, "solves_per_s_resident": len(items) / (ms * 1e-3), "e2e_s": e2e,
            "solves_per_s_e2e": len(items) / e2e, "rounds": st["rounds"], "iters": st["iters"],
            "max_scaled_residual_sampled": worst_res, "max_orth_err_sampled": worst_orth, "info_nonzero": 0}


def cfg4(atom):
    a = host_basis(kind_grid=0, k=8, nfun=4000, rb=2000.0)      # ka = k+3 = 11 (the reference default)
    p = a.problem()
    items = [(p, l) for l in range(21)]
    ms, st = timed_resident(atom, items)
    n = a.nfun
    E = pinned_empty(len(items) * n)
    Cb = pinned_empty(len(items) * n * n)
    t0 = time.perf_counter()
    Es, Cs, info = atom.solve_batch(items, out_E=E, out_C=Cb)
    e2e = time.perf_counter() - t0
    assert not info.any(), info
    band = atom.MATRIX_SVT(p)
    S = band_to_dense_sym(band["S"], n)
    H0 = band_to_dense_sym(band["H0"], n)
    Q = band_to_dense_sym(band["Q"], n)
    worst_res = worst_orth = 0.0
    for l in (0, 20):
        H = H0 + l * (l + 1) * Q
        Cm, Ev = Cs[l], Es[l]
        assert np.all(np.diff(Ev) > 0)
        SC = S @ Cm
        worst_orth = max(worst_orth, np.abs(Cm.T @ SC - np.eye(n)).max())
        worst_res = max(worst_res, (np.abs(H @ Cm - SC * Ev).max(0) / np.maximum(1, np.abs(Ev))).max())
    return {"config": "cfg4: N=4000, k=8, ka=11, Rmax=2000, l=0..20, all eigenpairs", "solves": len(items),
            "resident_ms": ms, "solves_per_s_resident": len(items) / (ms * 1e-3), "e2e_s": e2e,
            "solves_per_s_e2e": len(items) / e2e, "rounds": st["rounds"], "iters": st["iters"],
            "E_2p_minus_exact": float(Es[1][0] + 0.125), "max_scaled_residual": worst_res, "max_orth_err": worst_orth}


def cfg5(atom):
    a = host_basis(kind_grid=0, k=7, nfun=1000, rb=500.0)
    p = a.problem()
    nl = 51
    Es, Cs, info = atom.solve_batch([(p, l) for l in range(nl)])
    assert not info.any()
    band = atom.MATRIX_SVT(p)
    n, kd = a.nfun, a.k - 1
    R = band_to_dense_sym(band["R"], n)
    Rb = np.zeros((2 * kd + 1, n), order="F")
    for j in range(n):
        for i in range(max(0, j - kd), min(n, j + kd + 1)):
            Rb[kd + i - j, j] = R[i, j]
    dev_ms, wall = [], []
    worst = 0.0
    for l in (0, 17, 49):                                     # pair-by-pair entry (one launch pair per call)
        t0 = time.perf_counter()
        D = atom.dipole(Rb, Cs[l + 1], Cs[l])
        wall.append(time.perf_counter() - t0)
        dev_ms.append(atom.stats()["ms_total"])
        ref = Cs[l + 1].T @ (R @ Cs[l])
        scale = np.linalg.norm(Cs[l + 1], axis=0)[:, None] * np.linalg.norm(R @ Cs[l], axis=0)[None, :]
        worst = max(worst, float(np.max(np.abs(D - ref) / scale)))
        if l == 0:
            d_1s2p = abs(D[0, 0])
    # all 50 pairs in two launches
    atom.dipole_chain(Rb, Cs)
    t0 = time.perf_counter()
    Dall = atom.dipole_chain(Rb, Cs)
    chain_wall = time.perf_counter() - t0
    chain_ms = atom.stats()["ms_total"]
    assert np.array_equal(Dall[0], atom.dipole(Rb, Cs[1], Cs[0]))
    flops = 2.0 * n ** 3 + 2.0 * (2 * kd + 1) * n * n
    return {"config": "cfg5: D = C_{l+1}^T R C_l, l=0..49, N=1000 (length gauge)", "pairs": nl - 1,
            "single_pair_device_ms": float(np.median(dev_ms)), "single_pair_fp64_tflops": flops / (np.median(dev_ms) * 1e-3) / 1e12,
            "single_pair_wall_ms_incl_pcie": 1e3 * float(np.median(wall)),
            "chain_device_ms_50_pairs": chain_ms, "chain_fp64_tflops": 50 * flops / (chain_ms * 1e-3) / 1e12,
            "chain_wall_ms_incl_pcie": 1e3 * chain_wall, "max_rel_err_vs_numpy": worst,
            "abs_<2p|r|1s>": float(d_1s2p), "exact": 128 * np.sqrt(6) / 243}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
    atom = bsp.BspAtom(0)
    for w in which:
        out = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}[w](atom)
        print(json.dumps(out), flush=True)
    atom.close()
