"""Golden fixture for BASELINE cfg4 at FULL size (N=4000, k=8, ka=11, Rmax=2000, Coulomb Z=1):

    python tests/golden/make_cfg4_golden.py [l ...]          (default l = 0 20)

For each l writes tests/golden/cfg4_l{l}_dsygv.npz with
  E_dsygv   : all 4000 eigenvalues from LAPACK DSYGV(1,'V','U') -- the call of matrices.f90:248 -- on the
              oracle-assembled dense pencil (matrices.f90:68-183, 244),
  E_truth   : all 4000 eigenvalues by extended-precision Sturm bisection on the band
              (oracle.band_bisect_truth; pinned to the 40-digit table of the shipped input),
  vec_index, vectors : 8 sampled DSYGV eigenvectors (columns), C^T S C = I,
  seconds_dsygv, threads : the wall time of the DSYGV call (a CPU-baseline data point for cfg4).
Dense DSYGV at N = 4000 takes minutes per l, which is why this is a committed fixture and not part of
the test run.  The script imports only the oracle (test infrastructure) and scipy.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

N, K, RB = 4000, 8, 2000.0
VEC_INDEX = np.array([0, 1, 5, 40, 400, 2000, 3500, 3999])


def main(ls):
    b = O.make_basis(kind_grid=0, k=K, nfun=N, rb=RB)
    assert b.nfun == N and b.ka == 11
    t0 = time.time()
    m = O.matrix_svt(b, lmax=max(ls))
    print("assembly %.1f s" % (time.time() - t0), flush=True)
    for l in ls:
        H = O.hamiltonian(m["T"], m["U"][:, :, l], m["V"])
        t0 = time.time()
        w, v, info = O.dsygv(H, m["S"])
        t_d = time.time() - t0
        assert info == 0
        print("l=%d dsygv %.1f s" % (l, t_d), flush=True)
        t0 = time.time()
        wt = O.band_bisect_truth(H, m["S"], K - 1, guess=w, rel_window=1e-5)
        print("l=%d truth %.1f s; dsygv vs truth max rel %.2e" %
              (l, time.time() - t0, np.max(np.abs(w - wt) / np.maximum(np.abs(wt), 1e-2))), flush=True)
        np.savez_compressed(os.path.join(HERE, "cfg4_l%d_dsygv.npz" % l), E_dsygv=w, E_truth=wt, vec_index=VEC_INDEX,
                            vectors=np.ascontiguousarray(v[:, VEC_INDEX]), seconds_dsygv=t_d,
                            threads=os.cpu_count() or 1, l=l)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [0, 20])
