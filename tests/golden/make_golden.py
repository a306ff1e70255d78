"""Regenerates the golden fixtures in tests/golden/ (run from the repo root):

    python tests/golden/make_golden.py

1. shipped_truth.json -- 40-digit spectrum of the pencil the oracle assembles for the shipped input
   exec/bsp_0.inp (N=124, k=7, l=0,1,2): mpmath, L = chol(S), A = L^-1 H L^-T, mp.eigsy(A).
   The reference ships no golden vectors; this is the independent pin for
   both LAPACK dsygv (what the reference calls) and the CUDA solver.
2. shipped_band.npz -- the oracle's S, T, V, Q(=U_1/2), R, Rinv, D for the shipped input in band
   form, as a regression fixture for the oracle itself and a GPU-box-portable assembly reference.
3. hydrogen.json -- analytic levels -Z^2/(2 n^2) used by the known-answer tests.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402


def truth_spectrum(H, S, dps=40):
    import mpmath as mp

    mp.mp.dps = dps
    n = H.shape[0]
    Hm = mp.matrix(H.tolist())
    Sm = mp.matrix(S.tolist())
    L = mp.cholesky(Sm)
    Li = mp.inverse(L)
    A = Li * Hm * Li.T
    A = (A + A.T) / 2
    ev = mp.eigsy(A, eigvals_only=True)
    return sorted([mp.nstr(e, 25) for e in ev], key=lambda s: float(s))


def main():
    b = O.shipped_basis()
    m = O.matrix_svt(b, lmax=2)
    out = {"config": "exec/bsp_0.inp: KIND_GRID=2 rmax=60 ra=0 rb=500 k=7 nfun=100->124, Zatom=1", "nfun": b.nfun,
           "dps": 40, "levels": {}}
    for l in range(3):
        H = O.hamiltonian(m["T"], m["U"][:, :, l], m["V"])
        out["levels"][str(l)] = truth_spectrum(H, m["S"])
        print("l", l, out["levels"][str(l)][:3], flush=True)
    json.dump(out, open(os.path.join(HERE, "shipped_truth.json"), "w"), indent=0)
    kd = b.k - 1
    band = {k: O.dense_to_band_upper(m[k], kd) for k in ("S", "T", "V", "R", "Ri")}
    band["Q"] = O.dense_to_band_upper(m["U"][:, :, 1] / 2.0, kd)  # l=1: l(l+1)/2 = 1 -> U_1 = 2 Q
    n = b.nfun
    D = np.zeros((2 * kd + 1, n))
    for j in range(n):
        for i in range(max(0, j - kd), min(n, j + kd + 1)):
            D[kd + i - j, j] = m["D"][i, j]
    np.savez_compressed(os.path.join(HERE, "shipped_band.npz"), rt=b.rt, xg=b.xg, wg=b.wg, D=D, **band)
    json.dump({"Z": 1.0, "levels": {str(n_): -0.5 / n_ ** 2 for n_ in range(1, 30)}},
              open(os.path.join(HERE, "hydrogen.json"), "w"))


if __name__ == "__main__":
    main()
