"""resident + e2e sensitivity to the number of chunk streams (diagnostic)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bspatom_b200 as bsp
from bench import workload_items, NFUN
atom = bsp.BspAtom(0)
inp, items = workload_items(bsp, 0, 8, "lin")
n = len(items)
E = torch.empty(n * NFUN, dtype=torch.float64).pin_memory().numpy()
Cb = torch.empty(n * NFUN * NFUN, dtype=torch.float64).pin_memory().numpy()
for workers, chunk in ((1, 0), (2, 0), (3, 0), (4, 0), (4, 51), (4, 34), (3, 34)):
    atom.set_option("workers", workers)
    atom.set_option("chunk", chunk)
    atom.batch_upload(items)
    for _ in range(2):
        atom.batch_run()
    r = []
    for _ in range(4):
        atom.batch_run(); r.append(atom.stats()["ms_total"])
    for _ in range(2):
        atom.solve_batch(items, out_E=E, out_C=Cb)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        atom.solve_batch(items, out_E=E, out_C=Cb)
        ts.append(1e3 * (time.perf_counter() - t0))
    st = atom.stats()
    print("workers", workers, "chunk", chunk, "resident ms %.1f | e2e ms/step median %.1f min %.1f" % (np.median(r), np.median(ts), min(ts)),
          "ms_total %.1f tail %.1f launches %d" % (st["ms_total"], st["wall_ms_copy_tail"], st["launches"]), flush=True)
