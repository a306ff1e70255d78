"""Find the cfg3 problem with the largest |C^T S C - I| (bench: 1.7e-9 over 4096 problems) and describe the pair."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from cases import cfg3_problems, band_to_dense_sym
atom = bsp.BspAtom(device=0)
a, items = cfg3_problems(4096)
def defect(idx):
    atom.batch_upload([items[i] for i in idx]); atom.batch_run()
    return atom.batch_verify()["max_orthonormality_defect"]
blocks = [list(range(s, s + 256)) for s in range(0, 4096, 256)]
d = [defect(b) for b in blocks]
print("per block of 256:", ["%.1e" % x for x in d], flush=True)
cur = blocks[int(np.argmax(d))]
while len(cur) > 1:
    h = len(cur) // 2
    d0, d1 = defect(cur[:h]), defect(cur[h:])
    cur = cur[:h] if d0 >= d1 else cur[h:]
i = cur[0]
p, l = items[i]
Es, Cs, info = atom.solve_batch([items[i]])
E, Cm = np.array(Es[0]), np.array(Cs[0])
band = atom.MATRIX_SVT(p)
S = band_to_dense_sym(band["S"], a.nfun)
G = Cm.T @ S @ Cm - np.eye(a.nfun)
ij = np.unravel_index(np.argmax(np.abs(G)), G.shape)
out = {"problem": int(i), "pot_kind": int(p.pot_kind), "pot_par": [float(x) for x in p.pot_par[:2]], "l": int(l),
       "defect": float(np.abs(G).max()), "pair": [int(ij[0]), int(ij[1])], "E_pair": [float(E[ij[0]]), float(E[ij[1]])],
       "gap_rel": float(abs(E[ij[0]] - E[ij[1]]) / max(abs(E[ij[0]]), 1e-300)), "stats": atom.stats(),
       "neighbours": [float(x) for x in E[max(0, min(ij) - 2): max(ij) + 3]],
       "diag_defect_max": float(np.abs(np.diag(G)).max()), "rows_over_1e-10": int((np.abs(G).max(1) > 1e-10).sum())}
print(json.dumps(out))
for opt in (("vec_tol", 1e-13), ("min_iters", 3)):
    atom.set_option(*opt)
    atom.batch_upload([items[i]]); atom.batch_run()
    print(opt, atom.batch_verify()["max_orthonormality_defect"], atom.stats()["iters"])
    atom.set_option("vec_tol", 1e-12); atom.set_option("min_iters", 2)
