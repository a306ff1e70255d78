import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from cases import cfg3_problems, band_to_dense_sym
atom = bsp.BspAtom(device=0)
a, items = cfg3_problems(4096)
n = 500
c0 = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
blk = items[c0:c0 + 256]
Es, Cs, info = atom.solve_batch(blk)
print("batch verify:", atom.batch_verify())
worst = (0, -1)
for j in range(256):
    p, l = blk[j]
    band = atom.MATRIX_SVT(p)
    S = band_to_dense_sym(band["S"], n)
    Cm = np.asarray(Cs[j])
    d = np.abs(Cm.T @ S @ Cm - np.eye(n)).max()
    if d > worst[0]: worst = (d, j)
print("host-side worst defect of the batch result: %.2e at" % worst[0], c0 + worst[1])
j = worst[1]
E1, C1, _ = atom.solve_batch([blk[j]])
print("single == batch:", np.array_equal(E1[0], Es[j]), np.array_equal(np.asarray(C1[0]), np.asarray(Cs[j])), "max |dC| %.2e" % np.abs(np.asarray(C1[0]) - np.asarray(Cs[j])).max())
p, l = blk[j]
band = atom.MATRIX_SVT(p); S = band_to_dense_sym(band["S"], n)
for name, Cm, E in (("batch", np.asarray(Cs[j]), Es[j]), ("single", np.asarray(C1[0]), E1[0])):
    G = Cm.T @ S @ Cm - np.eye(n)
    a_, b_ = np.unravel_index(np.argmax(np.abs(G)), G.shape)
    print(name, "defect %.2e pair %d %d E %.12g %.12g" % (np.abs(G).max(), a_, b_, E[a_], E[b_]), "kind", p.pot_kind, p.pot_par, "l", l)
