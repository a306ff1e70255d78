"""Host-side post-processing of SOLVE_SYSTEM: state limits, density of
states, and the two text files the companion TDSE codes read back.

Follows matrices.f90:239-240,261-265 (Enl.dat), :269-346 (selection logic, both KIND_PI branches),
:352-378 (n1_max, cinl, Eigenvec_All.dat) and the FORMATs at :391-392.  Pure bookkeeping on the
gathered eigenpairs (O(N) per l); nothing here touches the GPU."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


# ---- Fortran edit descriptors -------------------------------------------------------------
def fortran_e(v: float, w: int, d: int) -> str:
    """Ew.d with scale factor 0: sign, '0.', d digits, 'E+ee'."""
    if v == 0.0:
        body = "0." + "0" * d + "E+00"
    else:
        m, ex = ("%.*E" % (d - 1, abs(v))).split("E")
        digits = m.replace(".", "")
        ex10 = int(ex) + 1
        es = "%+03d" % ex10 if abs(ex10) <= 99 else "%+04d" % ex10
        body = "0." + digits + ("E" + es if abs(ex10) <= 99 else es)
    if v < 0 or (v == 0.0 and np.signbit(v)):
        body = "-" + body
    return body.rjust(w) if len(body) <= w else "*" * w


def fortran_g(v: float, w: int, d: int) -> str:
    """Gw.d (F2008 10.7.5.2.2): F(w-4).(d-k) followed by 4 blanks when 0.1 <= |v| < 10**d after
    rounding to d significant digits (k = decimal exponent), else Ew.d."""
    n = abs(v)
    if n == 0.0:
        body = ("%.*f" % (d - 1, 0.0)).rjust(w - 4) + "    "
        return body
    ex10 = int(("%.*E" % (d - 1, n)).split("E")[1]) + 1      # |v| = 0.ddd x 10**ex10 after rounding
    if 0 <= ex10 <= d:
        s = "%.*f" % (d - ex10, v)
        return (s.rjust(w - 4) if len(s) <= w - 4 else "*" * (w - 4)) + "    "
    return fortran_e(v, w, d)


def _t2_i4_t8_g22(i: int, e: float) -> str:
    """FORMAT(T2,I4,T8,G22.15)  matrices.f90:391"""
    return " " + "%4d" % i + "  " + fortran_g(e, 22, 15)


def write_enl(path: str, Enl: np.ndarray) -> None:
    """Enl.dat: nfun, then for l = 0..lmax the lines (i, En(i)), i = 1..nfun  (matrices.f90:239-240,264)."""
    nfun, nl = Enl.shape
    with open(path, "w") as f:
        f.write("%12d\n" % nfun)                       # WRITE(75,*) nfun  (list-directed integer)
        for l in range(nl):
            col = Enl[:, l]
            f.write("\n".join(_t2_i4_t8_g22(i + 1, float(col[i])) for i in range(nfun)) + "\n")


def write_eigenvec_all(path: str, cinl: np.ndarray) -> None:
    """Eigenvec_All.dat (matrices.f90:366-378): header nfun, n1_max, lmax; per l the line `l` and
    n1_max records FORMAT(I5,5000G20.10).  cinl has shape (nfun, n1_max, lmax+1)."""
    nfun, n1_max, nl = cinl.shape
    with open(path, "w") as f:
        f.write("%12d%12d%12d\n" % (nfun, n1_max, nl - 1))
        for l in range(nl):
            f.write("%12d\n" % l)
            for ni in range(n1_max):
                f.write("%5d" % (ni + 1) + "".join(fortran_g(float(x), 20, 10) for x in cinl[:, ni, l]) + "\n")


# ---- selection logic ----------------------------------------------------------------------
@dataclass
class StateSelection:
    """what SOLVE_SYSTEM leaves in MOD_PHOTOION for the photo-ionisation stage"""

    n0_fin: int = -1
    n1_fin: int = -1
    Emax_fin: float = -1.0
    n1_max: int = -1
    nbds: int = 0
    n01: Optional[np.ndarray] = None      # (lmax+1, 3)                 KIND_PI >= 3
    rEki: Optional[np.ndarray] = None     # (nfun, lmax+1)              KIND_PI >= 3
    ntemp: List[int] = field(default_factory=list)   # columns of Hij kept per l (ctemp)
    E_ini: Optional[np.ndarray] = None    # KIND_PI = 1, 2
    E_fin: Optional[np.ndarray] = None


def select_states(Enl: np.ndarray, kind_pi: int, l_ini: int, l_fin: int, Emax_fin: float) -> StateSelection:
    """The bookkeeping of the l-loop of SOLVE_SYSTEM (matrices.f90:269-346) and of its epilogue
    (:352-356) on the eigenvalues of all l.  Indices in the result are the reference's 1-based ones.
    State that the reference carries from one l to the next (n0_fin, n1_fin, nlim, nbds, a stale
    ntemp, Emax_fin once replaced) is carried here in the same way."""
    nfun, nl = Enl.shape
    sel = StateSelection(Emax_fin=float(Emax_fin))
    n0_fin = n1_fin = -1
    nlim = nbds = 0
    ntemp = 0
    if kind_pi >= 3:
        sel.n01 = np.zeros((nl, 3), dtype=np.int64)
        sel.rEki = np.ones((nfun, nl))
    for l in range(nl):
        En = Enl[:, l]
        if kind_pi in (1, 2):
            if l == l_ini:
                sel.E_ini = En.copy()
            elif l == l_fin:
                sel.E_fin = En.copy()
                if sel.Emax_fin == -1.0:
                    sel.Emax_fin = float(En[-1])
                neg = np.nonzero(En < 0.0)[0]
                le = np.nonzero(En <= sel.Emax_fin)[0]
                if neg.size:
                    n0_fin = int(neg[-1]) + 1
                if le.size:
                    n1_fin = int(le[-1]) + 1
                n0_fin = min(n0_fin + 1, nfun - 1)
        elif kind_pi >= 3:
            if sel.Emax_fin == -1.0:
                sel.Emax_fin = float(En[-1])
                Elim = sel.Emax_fin
            else:
                Elim = sel.Emax_fin + 0.25
                if kind_pi >= 8:
                    Elim = sel.Emax_fin
            # the search loop leaves at the first level above both limits
            above = np.nonzero((En > sel.Emax_fin) & (En > Elim))[0]
            last = int(above[0]) if above.size else nfun - 1          # 0-based index of the last level visited
            seen = En[: last + 1]
            neg = np.nonzero(seen < 0.0)[0]
            nbold = int(neg.size)
            if neg.size:
                n0_fin = int(neg[-1]) + 1
            le = np.nonzero(seen <= sel.Emax_fin)[0]
            if le.size:
                n1_fin = int(le[-1]) + 1
            lt = np.nonzero(seen <= Elim)[0]
            if lt.size:
                ntemp = int(lt[-1]) + 1
            nbds = max(nbds, nbold)
            n0_fin += 1
            n1_fin += 1
            nE0 = n0_fin
            if kind_pi >= 5:
                n0_fin = 1
            nlim = max(nlim, ntemp)
            sel.n01[l] = (n0_fin, n1_fin, nE0 - 1)
            ntemp = min(max(n1_fin + 40, nlim), nfun)
            sel.ntemp.append(ntemp)
            # density of states (matrices.f90:338-342), 1-based i = nE0+1 .. nfun-1
            i = np.arange(nE0 + 1, nfun)                       # 1-based
            if i.size:
                sel.rEki[i - 1, l] = np.sqrt(2.0 / (En[i] - En[i - 2]))
            if 1 <= nE0 < nfun:
                sel.rEki[nE0 - 1, l] = np.sqrt(1.0 / (En[nE0] - En[nE0 - 1]))
            sel.rEki[nfun - 1, l] = np.sqrt(1.0 / (En[nfun - 1] - En[nfun - 2]))
    sel.n0_fin, sel.n1_fin, sel.nbds = n0_fin, n1_fin, nbds
    sel.n1_max = n1_fin
    if kind_pi >= 3:
        sel.n1_max = min(max(int(sel.n01[:, 1].max()) + 20, nlim), nfun)
    return sel


def collect_cinl(C_per_l: List[np.ndarray], sel: StateSelection) -> np.ndarray:
    """cinl(1:nfun, 1:n1_max, l) = ctemp(1:nfun, 1:n1_max, l)  (matrices.f90:369-373); ctemp holds the
    first ntemp(l) eigenvectors of each l and zeros beyond (it is zero-initialised at l = 0)."""
    nfun = C_per_l[0].shape[0]
    nl = len(C_per_l)
    cinl = np.zeros((nfun, sel.n1_max, nl), order="F")
    for l in range(nl):
        keep = min(sel.ntemp[l], sel.n1_max, C_per_l[l].shape[1])
        cinl[:, :keep, l] = C_per_l[l][:, :keep]
    return cinl


# ---- plane-wave photo-ionisation (KIND_PI = 1, 2): TRANS_AMP factors and CROSS_SECTIONS -------------
# SURVEY.md 8(f) row f-4, dipole branch: what the reference does with the matrix elements
# <n_f l_f | A | n_0 l_0> = ci_fin^T A ci_ini (bspatom_dipole) before it writes CSs/CrossSection_*.dat.
C_AU = 137.03599913815          # Modules.f90:12
A_AU = 5.29177249e-9            # Modules.f90:12  (cm)


def three_j(j1: int, j2: int, j3: int, m1: int, m2: int, m3: int) -> float:
    """Wigner 3j symbol for integer arguments (what THREE_J, Funs_WignerSymbols.for:1-60, evaluates with
    log-factorials); Racah's formula in exact rational arithmetic, one square root at the end."""
    from fractions import Fraction
    from math import factorial as f, sqrt

    if m1 + m2 + m3 != 0 or j3 > j1 + j2 or j3 < abs(j1 - j2):
        return 0.0
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return 0.0
    tmin = max(0, j2 - j3 - m1, j1 + m2 - j3)
    tmax = min(j1 + j2 - j3, j1 - m1, j2 + m2)
    if tmax < tmin:
        return 0.0
    acc = Fraction(0)
    for t in range(tmin, tmax + 1):
        den = f(t) * f(j1 + j2 - j3 - t) * f(j1 - m1 - t) * f(j2 + m2 - t) * f(j3 - j2 + m1 + t) * f(j3 - j1 - m2 + t)
        acc += Fraction((-1) ** t, den)
    delta = Fraction(f(j1 + j2 - j3) * f(j1 - j2 + j3) * f(-j1 + j2 + j3), f(j1 + j2 + j3 + 1))
    pref = delta * f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(j3 + m3) * f(j3 - m3)
    val = acc * acc * pref                       # square of the symbol, exact
    sign = (-1) ** (j1 - j2 - m3) * (1 if acc >= 0 else -1)
    return sign * sqrt(val.numerator / val.denominator) if val.denominator < 2 ** 1000 and val.numerator < 2 ** 1000 \
        else sign * sqrt(float(val))


def dipole_angular_factors(kind_pi: int, l0: int, m0: int, lf: int, mf: int, mph: int):
    """(c0, c1, c2) of TRANS_AMP's plane-wave branch (PhotoIon.f90:67-85): the radial operator is
    A = c1*rij(:,:,1) + c2*rij(:,:,2) and the amplitude c0 * An * <f|A|i>.
    KIND_PI = 1 (length): rij1 = <B_i|r|B_j>;  KIND_PI = 2 (velocity): rij1 = <B_i|1/r|B_j>, rij2 = <B_i|B_j'>."""
    t3a = three_j(lf, 1, l0, -mf, mph, m0)
    if kind_pi == 1:
        t3b = three_j(lf, 1, l0, 0, 0, 0)
        c1 = (-1.0) ** (lf + l0 + mf) * np.sqrt(float((2 * lf + 1) * (2 * l0 + 1))) * t3a * t3b
        return 1.0, c1, 0.0
    c0 = np.sqrt(float(l0 + 1)) * t3a
    c1 = c2 = 0.0
    if lf == l0 + 1:
        c1, c2 = float(l0 + 1), -1.0
    elif lf == l0 - 1:
        c1, c2 = float(l0), 1.0
    return c0, c1, c2


def trans_amp_dipole(D: np.ndarray, E_fin: np.ndarray, n0_fin: int, n1_fin: int, c0: float) -> np.ndarray:
    """T_fi(ni) = An * c0 * <ci_fin(:,ni)| A |ci_ini>, ni = n0_fin..n1_fin (1-based, PhotoIon.f90:96-106), with the
    density-of-states normalisation An = sqrt(2 / (E_fin(ni+1) - E_fin(ni-1))).  D[ni-1] is the matrix element of
    final state ni (one column of bspatom_dipole's result).  Returns the array T_fi(n0_fin:n1_fin)."""
    # the reference reads E_fin(ni-1) and E_fin(ni+1) out of bounds when there is no bound state in l_fin (n0_fin < 2)
    # or Emax_fin was not set (n1_fin = nfun); numpy would wrap the index silently -- refuse instead
    if n0_fin < 2 or n1_fin > len(E_fin) - 1 or n1_fin < n0_fin:
        raise ValueError("trans_amp_dipole needs 2 <= n0_fin <= n1_fin <= nfun - 1 (density-of-states stencil E_fin(ni-1), "
                         "E_fin(ni+1), PhotoIon.f90:97): got n0_fin=%d n1_fin=%d nfun=%d" % (n0_fin, n1_fin, len(E_fin)))
    ni = np.arange(n0_fin, n1_fin + 1)                 # 1-based
    An = np.sqrt(2.0 / (E_fin[ni] - E_fin[ni - 2]))    # E_fin(ni+1) - E_fin(ni-1)
    return An * c0 * np.asarray(D, dtype=np.float64)[ni - 1]


def cross_sections_dipole(kind_pi: int, E0: float, E_fin: np.ndarray, T_fi: np.ndarray, n0_fin: int, n1_fin: int,
                          l0: int):
    """(Ef, sigma in Mb) of CROSS_SECTIONS for KIND_PI = 1, 2 (PhotoIon.f90:300-318, 395-414):
    sigma = M_au * (4 pi^2 / c) * 1/(2 l0 + 1) * d1 * T_fi^2, d1 = Ef - E0 (length) or 1/(Ef - E0) (velocity)."""
    ni = np.arange(n0_fin, n1_fin + 1)
    Ef = E_fin[ni - 1]
    M_au = (A_AU ** 2) * 1.0e18
    c0 = 4.0 * (np.pi ** 2) / C_AU
    c1 = 1.0 / float(2 * l0 + 1)
    d1 = (Ef - E0) if kind_pi == 1 else 1.0 / (Ef - E0)
    return Ef, M_au * c0 * c1 * d1 * np.asarray(T_fi, dtype=np.float64) ** 2


def fortran_g_e3(v: float, w: int, d: int) -> str:
    """Gw.dE3 (three exponent digits): the F sub-format is followed by e + 2 = 5 blanks."""
    n = abs(v)
    if n == 0.0:
        return ("%.*f" % (d - 1, 0.0)).rjust(w - 5) + "     "
    ex10 = int(("%.*E" % (d - 1, n)).split("E")[1]) + 1
    if 0 <= ex10 <= d:
        s = "%.*f" % (d - ex10, v)
        return (s.rjust(w - 5) if len(s) <= w - 5 else "*" * (w - 5)) + "     "
    m, _ = ("%.*E" % (d - 1, n)).split("E")
    body = ("-" if v < 0 else "") + "0." + m.replace(".", "") + "E%+04d" % ex10
    return body.rjust(w) if len(body) <= w else "*" * w


def write_cross_section(path: str, Ef: np.ndarray, sigma: np.ndarray) -> None:
    """CSs/CrossSection_Len.dat / _Vel.dat: one `FORMAT(2G20.10E3)` record per final state (PhotoIon.f90:417,461)."""
    with open(path, "w") as fh:
        for e, s in zip(Ef, sigma):
            fh.write(fortran_g_e3(float(e), 20, 10) + fortran_g_e3(float(s), 20, 10) + "\n")


# ---- CUBSPL (CubicSpline.f90): the interpolation the cross-section stage uses -------------------------
def cubspl(x0: np.ndarray, y0: np.ndarray, x1: np.ndarray) -> np.ndarray:
    """y1 = CUBSPL(x0, y0, x1) (CubicSpline.f90:1-51): clamped cubic spline through (x0, y0) with end slopes
    from the first / last pair of points (SPLINE, :55-99), evaluated at x1 (SPLINT, :103-131).
    Kept from the reference: SPLINT starts its bisection at klo = 1 on arrays indexed from 0, so abscissae
    inside the FIRST interval are extrapolated from the second one; x1 equal to an end point returns the
    tabulated value."""
    x = np.asarray(x0, dtype=np.float64)
    y = np.asarray(y0, dtype=np.float64)
    n = x.size - 1
    yp1 = (y[1] - y[0]) / (x[1] - x[0])
    ypn = (y[n] - y[n - 1]) / (x[n] - x[n - 1])
    y2 = np.zeros(n + 1)
    u = np.zeros(n + 1)
    y2[0] = -0.5
    u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1)
    for i in range(1, n):
        sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1])
        p = sig * y2[i - 1] + 2.0
        y2[i] = (sig - 1.0) / p
        u[i] = (6.0 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]))
                / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p
    qn = 0.5
    un = (3.0 / (x[n] - x[n - 1])) * (ypn - (y[n] - y[n - 1]) / (x[n] - x[n - 1]))
    y2[n] = (un - qn * u[n - 1]) / (qn * y2[n - 1] + 1.0)
    for k in range(n - 1, -1, -1):
        y2[k] = y2[k] * y2[k + 1] + u[k]
    xi = np.asarray(x1, dtype=np.float64)
    # bisection of SPLINT: klo = last k in 1..n-1 with x(k) <= xi, or 1; khi = klo + 1
    klo = np.maximum(np.searchsorted(x[1:n], xi, side="right"), 1)
    khi = klo + 1
    h = x[khi] - x[klo]
    a = (x[khi] - xi) / h
    b = (xi - x[klo]) / h
    out = a * y[klo] + b * y[khi] + ((a ** 3 - a) * y2[klo] + (b ** 3 - b) * y2[khi]) * (h ** 2) / 6.0
    out = np.where(xi == x[0], y[0], np.where(xi == x[n], y[n], out))
    return out
