"""Why is the GPU slower when results stream to the host?  Same batch: no vectors copied / copies enqueued
inline / copies enqueued after all kernels."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bspatom_b200 as bsp
from bspatom_b200.host import pinned_empty
from bench import workload_items
inp, items = workload_items(bsp, 0, 8, "lin")
n = 1000; ns = len(items)
atom = bsp.BspAtom(0)
E = pinned_empty(ns * n); Cb = pinned_empty(ns * n * n)
def run(label, **kw):
    atom.solve_batch(items, out_E=E, **kw)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); atom.solve_batch(items, out_E=E, **kw); ts.append(1e3 * (time.perf_counter() - t0))
    st = atom.stats()
    print(label, "ms %.1f" % np.median(ts), "gpu_total %.1f copy_tail %.1f" % (st["ms_total"], st["wall_ms_copy_tail"]),
          {k: round(st[k], 1) for k in ("ms_eigenvalues", "ms_eigenvectors", "ms_finalize", "ms_k_round", "ms_k_factor", "ms_k_back")}, flush=True)
for workers, sc in ((2, 8), (4, 8)):
    atom.set_option("workers", workers); atom.set_option("stream_chunks", sc)
    atom.set_option("copy_after", 0)
    run("w%d sc%d values only      " % (workers, sc), want_vectors=False)
    run("w%d sc%d vectors inline   " % (workers, sc), out_C=Cb)
    atom.set_option("copy_after", 1)
    run("w%d sc%d vectors copyafter" % (workers, sc), out_C=Cb)
