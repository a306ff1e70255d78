"""GPU: the -DBSP_DEBUG build (bspatom_b200/libbspatom_debug.so: device-side bounds asserts on every L / X / R /
check-point / list index and on the tile pipeline's issue / acquire bookkeeping) over awkward shapes: every B-spline
order, basis sizes that are not multiples of the tile, warp or block sizes, partial vector counts, device-side
selection, check-pointed solves, the fast and the full schedule.  compute-sanitizer is closed on this pool; a failed
device assert traps, the worker exits non-zero and this test fails."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import bspatom_b200 as bsp
from cases import host_basis
assert bsp.LIB_PATH.endswith("libbspatom_debug.so"), bsp.LIB_PATH
atom = bsp.BspAtom(device=0)
n_checked = 0
for k, nfun in ((3, 37), (4, 61), (5, 33), (6, 95), (7, 129), (7, 257), (8, 131), (9, 45), (10, 67), (7, 8), (3, 4)):
    a = host_basis(kind_grid=0, k=k, nfun=nfun, rb=40.0)
    items = [(a.problem(), l) for l in range(3)]
    for opts, kw in (({}, {}), ({"ckpt": 1}, {}), ({"min_iters": 3}, {}), ({}, {"nvec": max(1, nfun // 3)}),
                     ({}, {"select": bsp.Selection.from_kind_pi(0.3, 3)}), ({"vec_tol": 0.0}, {}), ({"chunk": 2}, {})):
        for o, v in opts.items():
            atom.set_option(o, v)
        Es, Cs, info = atom.solve_batch(items, **kw)
        for o in opts:
            atom.set_option(o, {"ckpt": 0, "min_iters": 2, "vec_tol": 1e-12, "chunk": 0}[o])
        assert not info.any(), (k, nfun, opts, kw, info)
        for E in Es:
            assert np.all(np.isfinite(E)) and np.all(np.diff(E) > 0)
        n_checked += 1
v = atom.batch_verify()
assert v["max_scaled_residual"] < 1e-10
print("debug build: %%d solves clean" %% n_checked)
'''


def test_debug_build_asserts_stay_silent_on_awkward_shapes():
    lib = os.path.join(ROOT, "bspatom_b200", "libbspatom_debug.so")
    if not os.path.exists(lib):
        pytest.skip("libbspatom_debug.so not built (python -m bspatom_b200.build --debug)")
    env = dict(os.environ, BSPATOM_LIB=lib)
    out = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "solves clean" in out.stdout
