"""e2e sensitivity to host polling: solve_batch with different first_check_round (diagnostic)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import bspatom_b200 as bsp
from bench import workload_items, NFUN
atom = bsp.BspAtom(0)
inp, items = workload_items(bsp, 0, 8, "lin")
n = len(items)
E = torch.empty(n * NFUN, dtype=torch.float64).pin_memory().numpy()
Cb = torch.empty(n * NFUN * NFUN, dtype=torch.float64).pin_memory().numpy()
for fcr in (6,):
    atom.set_option("first_check_round", fcr)
    for _ in range(2):
        atom.solve_batch(items, out_E=E, out_C=Cb)
    t0 = time.perf_counter()
    for _ in range(3):
        atom.solve_batch(items, out_E=E, out_C=Cb)
    dt = (time.perf_counter() - t0) / 3
    st = atom.stats()
    print("first_check_round", fcr, "e2e ms/step %.1f" % (1e3 * dt), "ms_total %.1f rounds %d iters %d tail %.1f launches %d" %
          (st["ms_total"], st["rounds"], st["iters"], st["wall_ms_copy_tail"], st["launches"]), flush=True)
