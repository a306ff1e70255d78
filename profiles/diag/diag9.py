"""BspAtomPipeline timeline on the bench workload."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bspatom_b200 as bsp
from bspatom_b200.host import pinned_empty
from bench import workload_items
inp, items = workload_items(bsp, 0, 8, "lin")
n = 1000; ns = len(items)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5
bufs = [(pinned_empty(ns * n), pinned_empty(ns * n * n)) for _ in range(2)]
pipe = bsp.BspAtomPipeline(0, 2)
for kv in sys.argv[2:]:
    pipe.set_option(kv.split('=')[0], float(kv.split('=')[1]))
oE = [bufs[i % 2][0] for i in range(nb)]; oC = [bufs[i % 2][1] for i in range(nb)]
pipe.solve_batches([items] * 2, oE[:2], oC[:2])

for rep in range(2):
    t0 = time.perf_counter()
    pipe.solve_batches([items] * nb, oE, oC)
    dt = 1e3 * (time.perf_counter() - t0)
    ends = sorted(b for _, _, _, b in pipe.timeline)
    print(sys.argv[2:], "run", rep, "ms/step %.1f" % (dt / nb), "first %.0f" % ends[0], "period %.1f" % ((ends[-1] - ends[1]) / (len(ends) - 2)), flush=True)
