import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from bspatom_b200.host import pinned_empty
from cases import cfg3_problems
atom = bsp.BspAtom(device=0)
a, items = cfg3_problems(4096)
n = 500
blk = items[2560:2560 + 256]
ref = None
for rep in range(6):
    mode = "resident" if rep % 2 == 0 else "streamed"
    if mode == "resident":
        atom.batch_upload(blk); atom.batch_run()
        v = atom.batch_verify()
        E = pinned_empty(256 * n); Cb = pinned_empty(256 * n * n); info = np.zeros(256, dtype=np.int32)
        atom.batch_download(E, Cb, info)
        E, Cb = np.array(E), np.array(Cb)
    else:
        Es, Cs, info = atom.solve_batch(blk)
        v = atom.batch_verify()
        E = np.concatenate(Es); Cb = np.concatenate([np.asarray(c).ravel(order="F") for c in Cs])
    st = atom.stats()
    if ref is None: ref = (E, Cb)
    dE = np.abs(E - ref[0]).max(); dC = np.abs(Cb - ref[1]).max()
    bad = np.nonzero(np.abs(Cb - ref[1]).reshape(256, -1).max(1) > 0)[0]
    print(rep, mode, "orth %.2e res %.2e" % (v["max_orthonormality_defect"], v["max_scaled_residual"]), "vs first: dE %.2e dC %.2e differing pencils %s" % (dE, dC, bad[:8]), "sel3", st["selected_third_solve"], "iters", st["iters"], "rounds", st["rounds"], flush=True)
