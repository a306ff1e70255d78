"""e2e sensitivity to the number of streaming chunks / chunk streams (diagnostic)."""
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
for workers, sc in ((2, 4), (2, 5), (2, 6), (2, 8), (2, 10), (3, 6), (3, 9)):
    atom.set_option("workers", workers)
    atom.set_option("stream_chunks", sc)
    for _ in range(2):
        atom.solve_batch(items, out_E=E, out_C=Cb)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        atom.solve_batch(items, out_E=E, out_C=Cb)
        ts.append(1e3 * (time.perf_counter() - t0))
    st = atom.stats()
    print("workers", workers, "stream_chunks", sc, "e2e ms/step median %.1f min %.1f" % (np.median(ts), min(ts)),
          "ms_total %.1f tail %.1f" % (st["ms_total"], st["wall_ms_copy_tail"]), flush=True)
