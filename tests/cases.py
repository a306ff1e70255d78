"""Shared test configurations and comparators.

The knots come from the PRODUCT host code (bspatom_b200.host.BspAtom.GRID); the reference values
come from the oracle (oracle/) on the same knots."""
from __future__ import annotations

import numpy as np

EPS = 2.220446049250313e-16


def host_basis(**kw):
    """sizes + knots through the product host code (no GPU involved)."""
    from bspatom_b200.host import BspInputs

    return BspInputs.from_values(**kw)


def band_to_dense_sym(ab, n):
    """LAPACK upper band (kd+1, n) -> dense symmetric."""
    kd = ab.shape[0] - 1
    A = np.zeros((n, n))
    for d in range(kd + 1):
        v = ab[kd - d, d:]
        A += np.diag(v, d)
        if d:
            A += np.diag(v, -d)
    return A


def band_to_dense_general(ab, n):
    kd = (ab.shape[0] - 1) // 2
    A = np.zeros((n, n))
    for j in range(n):
        for i in range(max(0, j - kd), min(n, j + kd + 1)):
            A[i, j] = ab[kd + i - j, j]
    return A


def rel_entry_err(a, ref):
    """max entrywise error relative to the entry (entries that are exactly 0 in ref must be 0)."""
    mask = ref != 0
    err = np.abs(a - ref)
    out = 0.0
    if mask.any():
        out = float(np.max(err[mask] / np.abs(ref[mask])))
    if (~mask).any():
        out = max(out, float(np.max(err[~mask])))
    return out


def eig_tolerance(E_ref, c_eps=64.0):
    """north_star tolerance vs LAPACK dsygv: 1e-12 relative, 1e-10 Hartree absolute for near-zero
    levels, plus the backward-error floor c*eps*|E_max| of dsygv itself (two
    backward-stable solvers differ by 3-25 eps |E_max| on these pencils)."""
    emax = np.max(np.abs(E_ref))
    return np.maximum(np.maximum(1e-12 * np.abs(E_ref), 1e-10), c_eps * EPS * emax)


def check_eigenpairs(E, Cm, H, S, res_tol=1e-9, orth_tol=1e-9):
    """size-independent properties: ascending, S-orthonormal, small generalized residual."""
    assert np.all(np.diff(E) > 0), "eigenvalues not strictly ascending"
    SC = S @ Cm
    G = Cm.T @ SC
    orth = np.abs(G - np.eye(G.shape[0])).max()
    R = H @ Cm - SC * E[None, :Cm.shape[1]]
    res = np.abs(R).max(axis=0) / np.maximum(1.0, np.abs(E[:Cm.shape[1]]))
    assert orth <= orth_tol, f"C^T S C - I = {orth:.3e}"
    assert res.max() <= res_tol, f"scaled residual {res.max():.3e}"
    return orth, res.max()


def cfg3_problems(nprob, nfun=500, k=7, rb=500.0, seed=20261018):
    """Yukawa (even i) / Tietz (odd i), draw order Z, lambda, t per i."""
    from bspatom_b200.host import Problem, POT_TIETZ, POT_YUKAWA

    a = host_basis(kind_grid=0, k=k, nfun=nfun, rb=rb)
    rng = np.random.default_rng(seed)
    items = []
    for i in range(nprob):
        Z = float(rng.integers(1, 21))
        lam = float(rng.uniform(0.0, 0.5))
        t = float(rng.uniform(0.5, 2.0))
        if i % 2 == 0:
            p = Problem(k=a.k, nfun=a.nfun, nkp=a.nkp, ka=a.ka, rt=a.rt, pot_kind=POT_YUKAWA, pot_par=(Z, lam))
        else:
            p = Problem(k=a.k, nfun=a.nfun, nkp=a.nkp, ka=a.ka, rt=a.rt, pot_kind=POT_TIETZ, pot_par=(Z, t))
        items.append((p, i % 4))
    return a, items
