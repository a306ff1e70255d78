"""where is |C^T S C - I| large in cfg3?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from cases import cfg3_problems, band_to_dense_sym
atom = bsp.BspAtom(device=0)
a, items = cfg3_problems(4096)
n = 500
worst = []
for c0 in range(0, 4096, 256):
    atom.batch_upload(items[c0:c0 + 256]); atom.batch_run()
    worst.append(atom.batch_verify()["max_orthonormality_defect"])
print("per block of 256:", ["%.1e" % w for w in worst])
c0 = 256 * int(np.argmax(worst))
w2 = []
for i in range(c0, c0 + 256):
    atom.batch_upload([items[i]]); atom.batch_run(); w2.append(atom.batch_verify()["max_orthonormality_defect"])
order = np.argsort(w2)[::-1][:5]
for o in order:
    i = c0 + int(o)
    p, l = items[i]
    Es, Cs, info = atom.solve_batch([items[i]])
    band = atom.MATRIX_SVT(p)
    S = band_to_dense_sym(band["S"], n)
    Cm = np.asarray(Cs[0]); E = Es[0]
    G = Cm.T @ S @ Cm - np.eye(n)
    a_, b_ = np.unravel_index(np.argmax(np.abs(G)), G.shape)
    print("problem", i, "kind", p.pot_kind, "par", p.pot_par, "l", l, "defect %.2e" % w2[o], "pair", a_, b_, "E %.12g %.12g gap %.3e" % (E[a_], E[b_], abs(E[a_] - E[b_])), "info", info[0],
          "nearby E:", E[max(0, min(a_, b_) - 2):max(a_, b_) + 3])
