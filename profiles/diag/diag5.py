"""Accuracy vs schedule options on the bench workload (one charge, l = 0..50): residual, S-orthonormality and
eigenvalue drift for (min_iters, tau) settings, and the resident time of the full 408-solve step."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import bspatom_b200 as bsp
from bench import workload_items
from cases import band_to_dense_sym

inp, items = workload_items(bsp, 0, 8, "lin")
n = 1000
ref = None
import json
OPTS = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [[]]
for opts in OPTS:
    atom = bsp.BspAtom(0)
    for k, v in opts:
        atom.set_option(k, v)
    atom.batch_upload(items)
    atom.batch_run()
    ms = []
    for _ in range(3):
        atom.batch_run()
        ms.append(atom.stats()["ms_total"])
    st = atom.stats()
    Es, Cs, info = atom.solve_batch(items[:51])
    p = items[0][0]
    band = atom.MATRIX_SVT(p)
    S = band_to_dense_sym(band["S"], n)
    H0 = band_to_dense_sym(band["H0"], n)
    Q = band_to_dense_sym(band["Q"], n)
    worst_res = worst_orth = 0.0
    for l in (0, 1, 10, 25, 50):
        H = H0 + l * (l + 1) * Q
        Cm, Ev = np.asarray(Cs[l]), np.asarray(Es[l])
        SC = S @ Cm
        worst_orth = max(worst_orth, np.abs(Cm.T @ SC - np.eye(n)).max())
        worst_res = max(worst_res, (np.abs(H @ Cm - SC * Ev).max(0) / np.maximum(1, np.abs(Ev))).max())
    Eall = np.concatenate([np.asarray(e) for e in Es])
    if ref is None:
        ref = Eall
    drift = np.max(np.abs(Eall - ref) / np.maximum(np.abs(ref), 1e-2))
    print(opts, "ms %.2f" % np.median(ms), "rounds", st["rounds"], "iters", st["iters"], "info!=0", int(np.count_nonzero(info)),
          "res %.2e orth %.2e drift %.2e" % (worst_res, worst_orth, drift),
          {k: round(st[k], 2) for k in st if k.startswith("ms_")}, flush=True)
    atom.close()
