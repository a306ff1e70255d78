"""Timeline of the double-buffered end-to-end sweep: when each solve_batch call starts / ends per handle."""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np
import bspatom_b200 as bsp
from bspatom_b200.host import pinned_empty
from bench import workload_items

inp, items = workload_items(bsp, 0, 8, "lin")
n = 1000
ns = len(items)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
opts = [kv.split("=") for kv in sys.argv[3:]]
atoms = [bsp.BspAtom(0) for _ in range(depth)]
for a in atoms:
    for k, v in opts:
        a.set_option(k, float(v))
bufs = [(pinned_empty(ns * n), pinned_empty(ns * n * n)) for _ in range(depth)]
for a, (E, Cb) in zip(atoms, bufs):
    a.solve_batch(items, out_E=E, out_C=Cb)
log = []
t00 = time.perf_counter()
def worker(w):
    for i in range(w, nb, depth):
        t0 = time.perf_counter()
        atoms[w].solve_batch(items, out_E=bufs[w][0], out_C=bufs[w][1])
        t1 = time.perf_counter()
        st = atoms[w].stats()
        log.append((i, w, 1e3 * (t0 - t00), 1e3 * (t1 - t00), st["wall_ms_upload"], st["wall_ms_run"], st["wall_ms_copy_tail"], st["ms_total"]))
th = [threading.Thread(target=worker, args=(w,)) for w in range(depth)]
for t in th: t.start()
for t in th: t.join()
tot = 1e3 * (time.perf_counter() - t00)
for r in sorted(log):
    print("batch %d handle %d start %.1f end %.1f  upload %.1f run %.1f copy_tail %.1f gpu_total %.1f" % r)
print("depth", depth, "batches", nb, "ms/step %.1f" % (tot / nb))
