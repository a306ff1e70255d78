"""Resident step time vs chunk size / streams (no copies), and e2e vs stream_chunks."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bspatom_b200 as bsp
from bspatom_b200.host import pinned_empty
from bench import workload_items
inp, items = workload_items(bsp, 0, 8, "lin")
n = 1000; ns = len(items)
atom = bsp.BspAtom(0)
atom.batch_upload(items)
for workers in (1, 2, 4):
    for chunk in (0, 204, 102, 74, 51, 37, 26):
        atom.set_option("workers", workers); atom.set_option("chunk", chunk)
        atom.batch_run()
        ms = []
        for _ in range(3):
            atom.batch_run(); ms.append(atom.stats()["ms_total"])
        print("resident workers", workers, "chunk", chunk, "ms %.1f" % np.median(ms), flush=True)
atom.set_option("chunk", 0)
E = pinned_empty(ns * n); Cb = pinned_empty(ns * n * n)
for workers in (2, 4):
    for sc in (2, 4, 8, 16):
        atom.set_option("workers", workers); atom.set_option("stream_chunks", sc)
        atom.solve_batch(items, out_E=E, out_C=Cb)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); atom.solve_batch(items, out_E=E, out_C=Cb); ts.append(1e3 * (time.perf_counter() - t0))
        st = atom.stats()
        print("e2e workers", workers, "stream_chunks", sc, "ms %.1f" % np.median(ts), "gpu_total %.1f copy_tail %.1f" % (st["ms_total"], st["wall_ms_copy_tail"]), flush=True)
