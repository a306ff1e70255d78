"""numpy/scipy check of what tridiag_ab dumped: the tridiagonal has the spectrum of the band matrix, and the device
Sturm counts equal the number of eigenvalues below each shift."""
import json
import sys

import numpy as np
from scipy.linalg import eigvals_banded, eigvalsh_tridiagonal

f = open(sys.argv[1], "rb")
n, b, nmat, ndump = np.frombuffer(f.read(16), dtype=np.int32)
out = []
for _ in range(ndump if nmat > 1 else 1):
    band = np.frombuffer(f.read(8 * (b + 1) * n), dtype=np.float64).reshape(b + 1, n)
    d = np.frombuffer(f.read(8 * n), dtype=np.float64)
    e = np.frombuffer(f.read(8 * n), dtype=np.float64)
    sh = np.frombuffer(f.read(8 * n), dtype=np.float64)
    cnt = np.frombuffer(f.read(4 * n), dtype=np.int32)
    w0 = eigvals_banded(band, lower=True)
    w1 = eigvalsh_tridiagonal(d, e[: n - 1])
    err = float(np.max(np.abs(w0 - w1)) / np.max(np.abs(w0)))
    ref = np.searchsorted(w1, sh)
    out.append({"eig_err_rel_to_norm": err, "count_mismatches": int(np.count_nonzero(ref != cnt))})
print(json.dumps({"n": int(n), "b": int(b), "checked": out}))
