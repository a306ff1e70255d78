"""ctypes wrapper over oracle/libbsp_oracle.so + host LAPACK ``dsygv``.

TEST INFRASTRUCTURE ONLY (see the header of bsp_oracle.c).  Importable from
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs;
never from ``bspatom_b200``.

PARITY UNPINNED by the reference's own tests (it has none); the
pins are analytic hydrogen, LAPACK ``dsygv`` (the routine the reference calls,
matrices.f90:248 -- here scipy's bundled OpenBLAS, since MKL is un-vendored and
un-pinned, src/Makefile:23) and the mpmath spectrum in tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbsp_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "bsp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "CC=gcc"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_SO)
        L.bsp_sizes.argtypes = [C.c_int] * 6 + [C.c_double] * 3 + [_ip]
        L.bsp_gauleg.argtypes = [C.c_double, C.c_double, _dp, _dp, C.c_int]
        L.bsp_grid.argtypes = [C.c_int] * 9 + [C.c_double] * 3 + [_dp, _dp]
        L.bsp_interv.argtypes = [_dp, C.c_int, C.c_double, _ip, _ip]
        L.bsp_bsplvb.argtypes = [C.c_int, _dp, C.c_int, C.c_double, C.c_int, _dp]
        L.bsp_bsplvb.restype = C.c_int
        L.bsp_bspall.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_double, _ip, _dp, _dp]
        L.bsp_bspall.restype = C.c_int
        L.bsp_selpot.argtypes = [C.c_int, _dp, C.c_double]
        L.bsp_selpot.restype = C.c_double
        L.bsp_rogers_params.argtypes = [C.c_double, _dp]
        L.bsp_matrix_svt.argtypes = (
            [C.c_int] * 4 + [_dp] * 4 + [C.c_int, C.c_int, _dp, _dp, C.c_int] + [_dp] * 7
        )
        L.bsp_matrix_svt.restype = C.c_int
        L.bsp_matrix_zaij.argtypes = [C.c_int] * 4 + [_dp] * 4 + [C.c_int] * 4 + [_dp, C.c_int, _dp]
        L.bsp_matrix_zaij.restype = C.c_int
        L.bsp_tormat_rvec.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
        L.bsp_hamiltonian.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.bsp_write_wf.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double, _dp, C.c_int, _dp, _dp]
        L.bsp_write_wf.restype = C.c_int
        L.bsp_dipole_dots.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp]
        L.bsp_band_bisect_ld.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, _ip, _dp, _dp, _dp]
        L.bsp_band_bisect_ld.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


@dataclass
class Basis:
    """Everything READ_INPUTS + GRID leave in MOD_BSPLINES / MOD_GRID."""

    kind_grid: int
    k: int
    ka: int
    nfun: int
    nkp: int
    nbc1: int
    nbc2: int
    nointv: int
    nintv_exp: int
    nintv_lin: int
    ra: float
    rb: float
    rmax: float
    rt: np.ndarray = field(repr=False)
    aind: np.ndarray = field(repr=False)  # (nfun, 2) Fortran order
    xg: np.ndarray = field(repr=False)
    wg: np.ndarray = field(repr=False)


def sizes(kind_grid, k, ka, nfun0, kind_bc1, kind_bc2, ra, rb, rmax):
    out = (C.c_int * 8)()
    lib().bsp_sizes(kind_grid, k, ka, nfun0, kind_bc1, kind_bc2, ra, rb, rmax, out)
    names = ["ka", "nbc1", "nbc2", "nkp", "nointv", "nfun", "nintv_exp", "nintv_lin"]
    return dict(zip(names, list(out)))


def gauleg(n, x1=-1.0, x2=1.0):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().bsp_gauleg(x1, x2, _p(x), _p(w), n)
    return x, w


def make_basis(kind_grid=0, k=7, ka=0, nfun=100, kind_bc1=0, kind_bc2=0, ra=0.0, rb=500.0, rmax=0.0) -> Basis:
    """READ_INPUTS (ReadInputs.f90:39-69) + GRID (grid.f90:1-99)."""
    s = sizes(kind_grid, k, ka, nfun, kind_bc1, kind_bc2, ra, rb, rmax)
    rt = np.zeros(s["nkp"])
    aind = np.zeros((s["nfun"], 2), order="F")
    lib().bsp_grid(kind_grid, k, s["nfun"], s["nkp"], s["nbc1"], s["nbc2"], s["nointv"],
                   s["nintv_exp"], s["nintv_lin"], ra, rb, rmax, _p(rt), _p(aind))
    xg, wg = gauleg(s["ka"])
    return Basis(kind_grid, k, s["ka"], s["nfun"], s["nkp"], s["nbc1"], s["nbc2"], s["nointv"],
                 s["nintv_exp"], s["nintv_lin"], ra, rb, rmax, rt, aind, xg, wg)


def interv(xt, x):
    left = C.c_int()
    mflag = C.c_int()
    xt = np.ascontiguousarray(xt, dtype=np.float64)
    lib().bsp_interv(_p(xt), len(xt), float(x), C.byref(left), C.byref(mflag))
    return left.value, mflag.value


def bsplvb(t, jhigh, x, left):
    t = np.ascontiguousarray(t, dtype=np.float64)
    b = np.zeros(jhigh)
    rc = lib().bsp_bsplvb(len(t), _p(t), jhigh, float(x), left, _p(b))
    if rc:
        raise FloatingPointError("FATAL ERROR - BSPLVB (bsplvb.f90:30-34)")
    return b


def bspall(b: Basis, r):
    left = C.c_int()
    bsp = np.zeros(b.k)
    dbsp = np.zeros(b.k)
    rc = lib().bsp_bspall(_p(b.rt), b.nkp, b.k, b.nfun, _p(b.aind), float(r), C.byref(left), _p(bsp), _p(dbsp))
    if rc:
        raise FloatingPointError("FATAL ERROR - BSPLVB (bsplvb.f90:30-34)")
    return left.value, bsp, dbsp


def pot_params(kind_pot, zatom=1.0, extra=()):
    par = np.zeros(8)
    if kind_pot == 1:
        lib().bsp_rogers_params(float(zatom), _p(par))
    else:
        par[0] = zatom
        for i, e in enumerate(extra):
            par[1 + i] = e
    return par


SIMONS_FUES_BL = (0.72657, 0.47095, -0.55508, -0.04008)  # ReadInputs.f90:136-139


def simons_fues_bl(lmax):
    bl = np.zeros(max(lmax, 3) + 1)
    bl[:4] = SIMONS_FUES_BL
    return bl


def selpot(kind_pot, par, r):
    par = np.ascontiguousarray(par, dtype=np.float64)
    return lib().bsp_selpot(kind_pot, _p(par), float(r))


def matrix_svt(b: Basis, lmax=0, kind_pot=0, par=None, bl=None, want_u=True, fast=None):
    """MATRIX_SVT scalar branch (matrices.f90:68-183). Dense column-major results."""
    n = b.nfun
    if par is None:
        par = pot_params(kind_pot)
    par = np.ascontiguousarray(par, dtype=np.float64)
    if kind_pot == 2 and bl is None:
        bl = simons_fues_bl(lmax)
    if bl is not None:
        bl = np.ascontiguousarray(bl, dtype=np.float64)
    if fast is None:
        fast = bool(np.all(np.diff(b.rt) >= 0))
    S = np.zeros((n, n), order="F")
    V = np.zeros((n, n), order="F")
    T = np.zeros((n, n), order="F")
    U = np.zeros((n, n, lmax + 1), order="F") if want_u else None
    R = np.zeros((n, n), order="F")
    Ri = np.zeros((n, n), order="F")
    D = np.zeros((n, n), order="F")
    rc = lib().bsp_matrix_svt(n, b.k, b.ka, b.nkp, _p(b.rt), _p(b.aind), _p(b.xg), _p(b.wg), lmax,
                              kind_pot, _p(par), _p(bl), 1 if fast else 0,
                              _p(S), _p(V), _p(T), _p(U), _p(R), _p(Ri), _p(D))
    if rc == 1:
        raise FloatingPointError("FATAL ERROR - BSPLVB (bsplvb.f90:30-34): the reference STOPs on this knot vector")
    if rc:
        raise IndexError("reference would index bsp() out of bounds (rc=%d)" % rc)
    return dict(S=S, V=V, T=T, U=U, R=R, Ri=Ri, D=D)


def matrix_zaij(b: Basis, kind_pi, zIth, ncomp_out):
    """KIND_PI >= 3 branch of MATRIX_SVT (matrices.f90:110-139, 164-175).
    zIth: complex array (nkp, ka, nlm, nm, ncomp), Fortran order; returns zAij (nfun, nfun, nlm, nm, ncomp_out), Fortran order."""
    zIth = np.asfortranarray(zIth, dtype=np.complex128)
    nkp, ka, nlm, nm, ncomp = zIth.shape
    assert nkp == b.nkp and ka == b.ka
    n = b.nfun
    zA = np.zeros((n, n, nlm, nm, ncomp_out), dtype=np.complex128, order="F")
    rc = lib().bsp_matrix_zaij(n, b.k, b.ka, b.nkp, _p(b.rt), _p(b.aind), _p(b.xg), _p(b.wg), int(kind_pi), nlm, nm, ncomp,
                               zIth.ctypes.data_as(_dp), int(ncomp_out), zA.ctypes.data_as(_dp))
    if rc:
        raise FloatingPointError("bsp_matrix_zaij rc=%d" % rc)
    return zA


def tormat_rvec(cinl, Xij):
    """TORMAT's matrix elements of r (TorusFuns.f90:127-158): rvecij(ni, li, nj, lj) = cinl(:,ni,li)^T Xij cinl(:,nj,lj), DSVMV('U').
    cinl: (nfun, n1_max, lmax+1)."""
    cinl = np.asfortranarray(cinl, dtype=np.float64)
    Xij = np.asfortranarray(Xij, dtype=np.float64)
    n, n1, nl = cinl.shape
    out = np.zeros((n1, nl, n1, nl), order="F")
    lib().bsp_tormat_rvec(n, n1, nl - 1, _p(cinl), _p(Xij), _p(out))
    return out


def hamiltonian(T, Ul, V):
    """Hij = Tij + Uij(:,:,l) + Vij (matrices.f90:244)."""
    n = T.shape[0]
    H = np.zeros((n, n), order="F")
    Ul = np.asfortranarray(Ul)
    lib().bsp_hamiltonian(n, _p(np.asfortranarray(T)), _p(Ul), _p(np.asfortranarray(V)), _p(H))
    return H


def dsygv(H, S):
    """CALL DSYGV(1,'V','U',nfun,Hij,nfun,Bij,nfun,En,WORK,LWORK,INFO) (matrices.f90:248).

    Returns (En, C, info); C(:,j) is eigenvector j, C^T S C = I."""
    from scipy.linalg import lapack

    w, v, info = lapack.dsygv(np.asfortranarray(H), np.asfortranarray(S), itype=1, jobz="V", uplo="U",
                              overwrite_a=False, overwrite_b=False)
    return w, v, info


def lower_band(A, kd):
    """row-major (kd+1, n) lower band: ab[d, i] = A[i+d, i]."""
    n = A.shape[0]
    ab = np.zeros((kd + 1, n))
    for d in range(kd + 1):
        ab[d, :n - d] = np.diagonal(A, -d)
    return np.ascontiguousarray(ab)


def band_bisect_truth(H, S, kd, idx=None, guess=None, rel_window=1e-6):
    """Third comparator: eigenvalues of the banded pencil by Sturm bisection in x87 extended precision
    straight on the band (no reduction, no eps*|E_max| floor; bsp_band_bisect_ld).  ``guess``: approximate
    eigenvalues (e.g. dsygv's) used only to start the brackets -- every bracket is verified by two counts."""
    n = H.shape[0]
    hb, sb = lower_band(np.asarray(H), kd), lower_band(np.asarray(S), kd)
    idx = np.arange(n, dtype=np.int32) if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros(len(idx))
    lo = hi = None
    if guess is not None:
        g = np.asarray(guess, dtype=np.float64)[idx]
        wdt = rel_window * np.maximum(np.abs(g), np.abs(np.asarray(guess)).max() * 1e-9) + 1e-12
        lo, hi = np.ascontiguousarray(g - wdt), np.ascontiguousarray(g + wdt)
    rc = lib().bsp_band_bisect_ld(n, kd, _p(hb), _p(sb), len(idx), idx.ctypes.data_as(_ip), _p(lo), _p(hi), _p(out))
    if rc:
        raise ValueError("bsp_band_bisect_ld rc=%d" % rc)
    return out


def dsygvx(H, S):
    """LAPACK's own bisection driver (dsygvx: dstebz + dstein after the dense reduction), all eigenvalues."""
    from scipy.linalg import eigh

    return eigh(np.asarray(H), np.asarray(S), eigvals_only=True, driver="gvx")


def solve_system(m, l):
    """One pass of the l-loop body of SOLVE_SYSTEM (matrices.f90:244-254)."""
    H = hamiltonian(m["T"], m["U"][:, :, l], m["V"])
    w, v, info = dsygv(H, m["S"])
    if info != 0:
        raise np.linalg.LinAlgError("ERROR DIAGONALIZING THE MATRIX! %d  l = %d" % (info, l))
    return w, v


def write_wf(b: Basis, ci, npts=10000):
    """WRITE_WF (Bsp_Atom.f90:101-152): returns r(0:npts), psi(0:npts)."""
    ci = np.ascontiguousarray(ci, dtype=np.float64)
    r = np.zeros(npts + 1)
    psi = np.zeros(npts + 1)
    rc = lib().bsp_write_wf(len(ci), b.k, b.nkp, _p(b.rt), b.ra, b.rb, _p(ci), npts, _p(r), _p(psi))
    if rc:
        raise FloatingPointError("WRITE_WF: bsplvb/interv failure rc=%d" % rc)
    return r, psi


def dipole_dots(A, x, cfin):
    """v = A x (DGEMV), out(n) = DDOT(cfin(:,n), v)  (PhotoIon.f90:90-105)."""
    A = np.asfortranarray(A)
    x = np.ascontiguousarray(x)
    cfin = np.asfortranarray(cfin)
    out = np.zeros(cfin.shape[1])
    lib().bsp_dipole_dots(A.shape[0], _p(A), _p(x), cfin.shape[1], _p(cfin), _p(out))
    return out


def dense_to_band_upper(A, kd):
    """LAPACK upper band storage AB(kd+1+i-j, j) = A(i,j) (1-based), shape (kd+1, n)."""
    n = A.shape[0]
    ab = np.zeros((kd + 1, n), order="F")
    for d in range(kd + 1):
        ab[kd - d, d:] = np.diagonal(A, d)
    return ab


# ---- the shipped input, exec/bsp_0.inp:8-9,12 --------------------------------
def shipped_basis() -> Basis:
    return make_basis(kind_grid=2, rmax=60.0, ra=0.0, rb=500.0, k=7, nfun=100, kind_bc1=0, kind_bc2=0)
