import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp
from bench import workload
atom = bsp.BspAtom(device=0)
wl = workload("cfg2", bsp, 0, 1, 1, "lin")
items, n, kd = wl["items"], 1000, 6
Rb = np.zeros((2 * kd + 1, n), order="F"); Rb[kd] = 1.0
atom.batch_upload(items); atom.batch_run()
for _ in range(2):
    atom.dipole_chain_resident(Rb, 0, 51, n)
    atom.dipole_chain_resident(Rb, 0, 2, n)
print("ok")
