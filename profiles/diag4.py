"""rounds diagnostic: bench workload, poll every round (BSPATOM_DEBUG_COUNTERS=1 prints the open brackets)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import bspatom_b200 as bsp
from bench import workload_items
atom = bsp.BspAtom(0)
atom.set_option("first_check_round", 1)
atom.set_option("check_every", 1)
atom.set_option("workers", 1)
inp, items = workload_items(bsp, 0, 8, "lin")
atom.batch_upload(items)
atom.batch_run()
print(atom.stats()["rounds"], atom.stats()["iters"], flush=True)
# which pencils are slow: solve each (Z, l) group of 51 separately
for z in range(8):
    atom.batch_upload(items[z * 51:(z + 1) * 51])
    atom.batch_run()
    print("Z index", z, "rounds", atom.stats()["rounds"], flush=True)
