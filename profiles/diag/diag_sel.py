"""diagnostics: (1) cost of the selection mode (rounds, redo, stage times), (2) where |C^T S C - I| is largest over the bench batch"""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from bench import workload
atom = bsp.BspAtom(device=0)
wl = workload("cfg2", bsp, 0, 1, 8, "lin")
items = wl["items"]
n = 1000
for name, kw in (("full", {}), ("sel1.5", dict(select=bsp.Selection.from_kind_pi(1.5, 3))), ("nvec300", dict(nvec=300))):
    atom.batch_upload(items, **kw)
    atom.batch_run(); atom.batch_run()
    st = atom.stats()
    print(name, {k: round(st[k], 2) for k in ("ms_total", "rounds", "iters", "chunks_redone", "ms_eigenvalues", "ms_eigenvectors", "ms_finalize", "ms_k_round", "n_k_round", "ms_k_factor", "ms_k_back", "launches", "selected_third_solve")}, flush=True)
# orthonormality per charge and per option
for opt in (("vec_tol", 1e-12), ("vec_tol", 0.0), ("vec_tol", 2e-13)):
    atom.set_option(*opt)
    worst = []
    for iz in range(8):
        atom.batch_upload(items[iz * 51:(iz + 1) * 51])
        atom.batch_run()
        v = atom.batch_verify()
        worst.append(v["max_orthonormality_defect"])
    print(opt, "orth per charge", ["%.2e" % w for w in worst], "third solves", atom.stats()["selected_third_solve"], flush=True)
atom.set_option("vec_tol", 1e-12)
# locate the worst pencil of charge 0
w = []
for l in range(51):
    atom.batch_upload([items[l]]); atom.batch_run(); w.append(atom.batch_verify()["max_orthonormality_defect"])
lw = int(np.argmax(w)); print("worst l of Z=1:", lw, "%.2e" % w[lw], ["%.1e" % x for x in w])
Es, Cs, info = atom.solve_batch([items[lw]])
band = atom.MATRIX_SVT(items[lw][0])
from cases import band_to_dense_sym
S = band_to_dense_sym(band["S"], n)
G = np.asarray(Cs[0]).T @ S @ np.asarray(Cs[0]) - np.eye(n)
i, j = np.unravel_index(np.argmax(np.abs(G)), G.shape)
E = Es[0]
print("worst pair", i, j, "G=%.2e" % G[i, j], "E_i=%.10g E_j=%.10g gap=%.3e" % (E[i], E[j], abs(E[i] - E[j])))
