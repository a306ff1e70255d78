"""e2e: single handle vs two alternating handles (double-buffered sweep) (diagnostic)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bspatom_b200 as bsp
from bench import workload_items, NFUN
inp, items = workload_items(bsp, 0, 8, "lin")
n = len(items)
bufs = [(torch.empty(n * NFUN, dtype=torch.float64).pin_memory().numpy(), torch.empty(n * NFUN * NFUN, dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
for depth, sc in ((1, 8), (2, 8), (2, 4), (2, 2)):
    pipe = bsp.BspAtomPipeline(0, depth=depth)
    pipe.set_option("stream_chunks", sc)
    nb = 8
    batches = [items] * nb
    oE = [bufs[i % 2][0] for i in range(nb)]
    oC = [bufs[i % 2][1] for i in range(nb)]
    pipe.solve_batches(batches[:2], oE[:2], oC[:2])
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        infos = pipe.solve_batches(batches, oE, oC)
        ts.append(1e3 * (time.perf_counter() - t0) / nb)
    assert all(not i.any() for i in infos)
    print("depth", depth, "stream_chunks", sc, "e2e ms/step", [round(t, 1) for t in ts], flush=True)
    pipe.close()
