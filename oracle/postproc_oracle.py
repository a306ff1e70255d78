"""Literal (loop by loop, 1-based) restatement of the bookkeeping in SOLVE_SYSTEM, matrices.f90:269-346
and :352-356, and of the record formats at :391-392.  TEST INFRASTRUCTURE ONLY (see oracle/bsp_oracle.c)."""
import math

import numpy as np


def g_edit(v, w, d):
    """Gw.d by the text of the standard: pick the F sub-format from the magnitude thresholds
    10**(k-1) - 0.5*10**(k-1-d) <= |v| < 10**k - 0.5*10**(k-d), k = 0..d; otherwise Ew.d."""
    from decimal import Decimal

    n = Decimal(repr(abs(float(v)))) if v != 0 else Decimal(0)
    exact = Decimal(abs(float(v)))                    # exact binary value
    if exact == 0:
        return ("%.*f" % (d - 1, 0.0)).rjust(w - 4) + " " * 4
    for k in range(0, d + 1):
        lo = Decimal(10) ** (k - 1) - Decimal(5) * Decimal(10) ** (k - 2 - d)
        hi = Decimal(10) ** k - Decimal(5) * Decimal(10) ** (k - 1 - d)
        if lo <= exact < hi:
            return ("%.*f" % (d - k, float(v))).rjust(w - 4) + " " * 4
    m, ex = ("%.*E" % (d - 1, abs(float(v)))).split("E")
    ex = int(ex) + 1
    s = ("-" if v < 0 else "") + "0." + m.replace(".", "") + "E" + ("+" if ex >= 0 else "-") + "%02d" % abs(ex)
    return s.rjust(w)


def solve_system_bookkeeping(Enl, kind_pi, l_ini, l_fin, emax_fin):
    """returns dict with the reference's variables after the l-loop (all 1-based)."""
    nfun, nl = Enl.shape
    lmax = nl - 1
    n0_fin = -1                                              # :234
    n1_fin = -1                                              # :235
    nlim = 0                                                 # :236
    nbds = 0                                                 # :237
    ntemp = 0
    n01 = np.zeros((lmax + 1, 3), dtype=np.int64)
    rEki = np.ones((nfun, lmax + 1))                         # :231
    ntemps = []
    E_ini = E_fin = None
    for l in range(0, lmax + 1):                             # :242
        En = [None] + [float(x) for x in Enl[:, l]]          # 1-based
        if kind_pi == 1 or kind_pi == 2:                     # :269
            if l == l_ini:                                   # :271
                E_ini = np.array(En[1:])
            elif l == l_fin:                                 # :274
                E_fin = np.array(En[1:])
                if emax_fin == -1.0:
                    emax_fin = En[nfun]                      # :276
                i = 1
                while True:                                  # :278-283
                    if En[i] < 0.0:
                        n0_fin = i
                    if En[i] <= emax_fin:
                        n1_fin = i
                    i = i + 1
                    if i > nfun:
                        break
                n0_fin = n0_fin + 1                          # :284
                n0_fin = min(n0_fin, nfun - 1)               # :285
        elif kind_pi >= 3:                                   # :292
            if emax_fin == -1.0:                             # :295-301
                emax_fin = En[nfun]
                elim = emax_fin
            else:
                elim = emax_fin + 0.25
                if kind_pi >= 8:
                    elim = emax_fin
            i = 1
            nbold = 0
            while True:                                      # :305-315
                if En[i] < 0.0:
                    n0_fin = i
                    nbold = nbold + 1
                if En[i] <= emax_fin:
                    n1_fin = i
                if En[i] <= elim:
                    ntemp = i
                if En[i] > emax_fin and En[i] > elim:
                    break
                i = i + 1
                if i > nfun:
                    break
            nbds = max(nbds, nbold)                          # :316
            n0_fin = n0_fin + 1
            n1_fin = n1_fin + 1
            nE0 = n0_fin
            if kind_pi >= 5:
                n0_fin = 1
            if ntemp > nlim:
                nlim = ntemp
            n01[l, 0] = n0_fin
            n01[l, 1] = n1_fin
            n01[l, 2] = nE0 - 1
            ntemp = max(n1_fin + 40, nlim)                   # :328
            ntemp = min(ntemp, nfun)
            ntemps.append(ntemp)
            for i in range(nE0 + 1, nfun - 1 + 1):           # :338-340
                rEki[i - 1, l] = math.sqrt(2.0 / (En[i + 1] - En[i - 1]))
            if 1 <= nE0 < nfun:
                rEki[nE0 - 1, l] = math.sqrt(1.0 / (En[nE0 + 1] - En[nE0]))          # :341
            rEki[nfun - 1, l] = math.sqrt(1.0 / (En[nfun] - En[nfun - 1]))           # :342
    n1_max = n1_fin                                          # :352
    if kind_pi >= 3:
        n1_max = max(int(n01[:, 1].max()) + 20, nlim)        # :355
        n1_max = min(n1_max, nfun)
    return dict(n0_fin=n0_fin, n1_fin=n1_fin, Emax_fin=emax_fin, n1_max=n1_max, nbds=nbds, n01=n01, rEki=rEki,
                ntemp=ntemps, E_ini=E_ini, E_fin=E_fin)


def zhvmv(zA, zx, zy):
    """ZHVMV of the reference (Modules.f90:398-425): zv = ZHEMV('U', zA) zx, zf = ZDOTU(zy, zv), written out as
    the BLAS reference loops do it: only the upper triangle of zA is read, the lower one is its conjugate and the
    imaginary part of the diagonal is ignored.  Plain loops: test infrastructure for small cases."""
    import numpy as np

    n = zA.shape[0]
    v = np.zeros(n, dtype=np.complex128)
    for j in range(n):
        t1 = zx[j]
        t2 = 0.0 + 0.0j
        for i in range(j):
            v[i] += t1 * zA[i, j]
            t2 += np.conj(zA[i, j]) * zx[i]
        v[j] += t1 * zA[j, j].real + t2
    return np.sum(np.asarray(zy) * v)      # ZDOTU: no conjugation


def trans_amp_block(zA, Cf, Ci):
    """All (bra, ket) pairs of one angular block of TRANS_AMP's general branch (PhotoIon.f90:218-232):
    T[f, i] = ZHVMV(zA, Ci[:, i], Cf[:, f]) -- vectorised form of `zhvmv` (Hermitian completion of the upper
    triangle), used where the loops would be too slow."""
    import numpy as np

    U = np.triu(zA, 1)
    Ah = U + np.conj(U).T + np.diag(np.real(np.diag(zA)))
    return np.asarray(Cf).T @ (Ah @ np.asarray(Ci))


def three_j_ref(i1, i2, i3, m1, m2, m3):
    """THREE_J as the reference evaluates it (Funs_WignerSymbols.for:1-57): log-factorial table, terms scaled by
    the smallest one, alternating sum.  Returns 0 where the reference's selection rules fail."""
    import math

    fac = [0.0] * 141                                    # FAC(I) = log((I-1)!), 1-based
    for i in range(2, 141):
        fac[i] = fac[i - 1] + math.log(float(i - 1))
    l4 = i1 + i2 + i3 + 2
    if l4 > 140 or m1 + m2 + m3 != 0:
        return 0.0
    izmax = min(i1 + i2 - i3, i1 - m1, i2 + m2) + 1
    izmin = max(0, i2 - i3 - m1, i1 + m2 - i3) + 1
    if izmax - izmin < 0:
        return 0.0
    l1, l2, l3 = i1 + i2 - i3 + 1, i3 + i1 - i2 + 1, i3 + i2 - i1 + 1
    l5, l6, l7, l8, l9, l10 = i1 + m1 + 1, i1 - m1 + 1, i2 + m2 + 1, i2 - m2 + 1, i3 + m3 + 1, i3 - m3 + 1
    if min(l1, l2, l3, l5, l6, l7, l8, l9, l10) < 1:
        return 0.0                                       # outside the triangle / |m| > j: the reference indexes FAC(0)
    abra = 0.5 * (fac[l1] + fac[l2] + fac[l3] - fac[l4] + fac[l5] + fac[l6] + fac[l7] + fac[l8] + fac[l9] + fac[l10])
    k1, k2 = i3 - i2 + m1 + 1, i3 - i1 - m2 + 1
    gros = 250.0
    ac = {}
    for ii in range(izmin, izmax + 1):
        i = ii - 1
        ac[ii] = fac[i + 1] + fac[l1 - i] + fac[l6 - i] + fac[l7 - i] + fac[k1 + i] + fac[k2 + i]
        if ac[ii] < gros:
            gros = ac[ii]
    accu = 0.0
    sig = (-1.0) ** izmin
    for ii in range(izmin, izmax + 1):
        sig = -sig
        accu += sig * math.exp(-(ac[ii] - gros))
    return (-1.0) ** (i1 - i2 - m3) * math.exp(abra - gros) * accu


def photoion_dipole_ref(kind_pi, A, ci_ini, ci_fin, E_ini_n0, E_fin, n0_fin, n1_fin, l0, m0, lf, mf, mph):
    """TRANS_AMP (PhotoIon.f90:50-107) + CROSS_SECTIONS (:300-318, :395-417) for KIND_PI = 1, 2, loop by loop.
    A: (rij1, rij2) dense operators (rij2 None for the length gauge); ci_fin: (nfun, nfun) all final vectors;
    indices 1-based as in the reference.  Returns (Ef list, T_fi list, sigma list) over ni = n0_fin..n1_fin."""
    import math

    n = len(ci_ini)
    t3a = three_j_ref(lf, 1, l0, -mf, mph, m0)
    if kind_pi == 1:
        t3b = three_j_ref(lf, 1, l0, 0, 0, 0)
        c1 = (-1.0) ** (lf + l0 + mf) * math.sqrt(float((2 * lf + 1) * (2 * l0 + 1))) * t3a * t3b
        c0 = 1.0
        Am = [[c1 * A[0][i][j] for j in range(n)] for i in range(n)]
    else:
        c0 = math.sqrt(float(l0 + 1)) * t3a
        c1 = c2 = 0.0
        if lf == l0 + 1:
            c1, c2 = float(l0 + 1), -1.0
        elif lf == l0 - 1:
            c1, c2 = float(l0), 1.0
        Am = [[c1 * A[0][i][j] + c2 * A[1][i][j] for j in range(n)] for i in range(n)]
    v = [sum(Am[i][j] * ci_ini[j] for j in range(n)) for i in range(n)]          # DGEMV
    M_au = (5.29177249e-9 ** 2) * 1.0e18
    cc0 = 4.0 * (math.pi ** 2) / 137.03599913815
    cc1 = 1.0 / float(2 * l0 + 1)
    Ef, T, sig = [], [], []
    for ni in range(n0_fin, n1_fin + 1):
        An = math.sqrt(2.0 / (E_fin[ni] - E_fin[ni - 2]))                         # E_fin(ni+1) - E_fin(ni-1)
        t = An * c0 * sum(ci_fin[i][ni - 1] * v[i] for i in range(n))            # DDOT
        e = E_fin[ni - 1]
        d1 = (e - E_ini_n0) if kind_pi == 1 else 1.0 / (e - E_ini_n0)
        Ef.append(e)
        T.append(t)
        sig.append(M_au * cc0 * cc1 * d1 * t * t)
    return Ef, T, sig


def cubspl_ref(x0, y0, x1):
    """CUBSPL / SPLINE / SPLINT (CubicSpline.f90) statement by statement; arrays indexed from 0 as in the
    reference (x0(0:n0)), including SPLINT's `klo = 1`."""
    n = len(x0) - 1
    x, y = [float(v) for v in x0], [float(v) for v in y0]
    yp1 = (y[1] - y[0]) / (x[1] - x[0])                       # :18
    ypn = (y[n] - y[n - 1]) / (x[n] - x[n - 1])               # :19
    y2 = [0.0] * (n + 1)
    u = [0.0] * (n + 1)
    if yp1 > 0.99e30:                                          # :68
        y2[0] = 0.0
        u[0] = 0.0
    else:
        y2[0] = -0.5
        u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1)
    for i in range(1, n):                                      # :77
        sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1])
        p = sig * y2[i - 1] + 2.0
        y2[i] = (sig - 1.0) / p
        u[i] = (6.0 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]))
                / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p
    if ypn > 0.99e30:                                          # :85
        qn = un = 0.0
    else:
        qn = 0.5
        un = (3.0 / (x[n] - x[n - 1])) * (ypn - (y[n] - y[n - 1]) / (x[n] - x[n - 1]))
    y2[n] = (un - qn * u[n - 1]) / (qn * y2[n - 1] + 1.0)      # :94
    for k in range(n - 1, -1, -1):                             # :96
        y2[k] = y2[k] * y2[k + 1] + u[k]
    out = []
    for xi in x1:                                              # :36
        xi = float(xi)
        if xi == x[0]:
            out.append(y[0])
        elif xi == x[n]:
            out.append(y[n])
        else:
            klo, khi = 1, n                                    # :113
            while khi - klo > 1:
                k = (khi + klo) // 2
                if x[k] > xi:
                    khi = k
                else:
                    klo = k
            h = x[khi] - x[klo]
            a = (x[khi] - xi) / h
            b = (xi - x[klo]) / h
            out.append(a * y[klo] + b * y[khi] + ((a ** 3 - a) * y2[klo] + (b ** 3 - b) * y2[khi]) * (h ** 2) / 6.0)
    return out


def chkphs_ref(b, cinl_l):
    """CHKPHS (matrices.f90:398-449) for the vectors of one l, statement by statement: cinl_l is (nfun, nvec) and is
    returned with the flipped columns.  Keeps the reference's index mapping jfun = j - (left - nbc1) (one function
    higher than WRITE_WF's when nbc1 = k - 1).  `b` is an oracle.Basis."""
    from oracle import oracle as O

    out = np.array(cinl_l, dtype=np.float64, copy=True)
    nr = 3                                                    # :418
    r0, r1 = b.ra, 0.1                                        # :419-420
    dr = (r1 - r0) / float(nr)                                # :421
    for n in range(out.shape[1]):                             # don0
        fr = [0.0] * nr
        for i in range(1, nr + 1):                            # dor0
            r = r0 + float(i) * dr                            # :430
            left, mflag = O.interv(b.rt, r)                   # :433
            bsp = O.bsplvb(b.rt, b.k, r, left)                # :434
            jmin = left - b.nbc1 + 1                          # :436
            jmax = min(jmin + b.k - 1, b.nfun)                # :437
            sumf = 0.0
            for j in range(jmin, jmax + 1):                   # doj
                jfun = j - (left - b.nbc1)                    # :441
                if j >= 1:
                    sumf = sumf + out[j - 1, n] * bsp[jfun - 1]
            fr[i - 1] = sumf
        if fr[0] < 0.0 and fr[1] < 0.0 and fr[2] < 0.0:       # :447
            out[:, n] = -out[:, n]
    return out


def writewf(b, cinl_l0, n0, n1, nbc1, npts):
    """WRITEWF (WriteWF.f90:23-61), statement by statement: fr(is), is = 0 .. n1-n0+1, for r = ra + i dr, with the
    reference's own index mapping jmin = left - nbc1 + 1, jfun = j - (left - nbc1).  cinl_l0: (nfun, nstates) of one l
    (1-based state n = column n-1).  Plain loops through the oracle's interv / bsplvb: small npts only."""
    import numpy as np

    from . import oracle as O

    k, nfun, nkp = b.k, b.nfun, b.nkp
    dr = (b.rb - b.ra) / float(npts)
    out = np.zeros((npts + 1, n1 - n0 + 2))
    rr = np.zeros(npts + 1)
    for i in range(npts + 1):
        r = b.ra + float(i) * dr
        left, _ = O.interv(b.rt, r)
        bsp = O.bsplvb(b.rt, k, r, left)
        jmin = left - nbc1 + 1
        jmax = min(jmin + k - 1, nfun)
        n = 0
        for is_ in range(0, n1 - n0 + 2):
            n = 1 if is_ == 0 else n + 1
            sumf = 0.0
            for j in range(jmin, jmax + 1):
                jfun = j - (left - nbc1)
                sumf = sumf + cinl_l0[j - 1, n - 1] * bsp[jfun - 1]
            out[i, is_] = sumf
            if is_ == 0:
                n = n0 - 1
        rr[i] = r
    return rr, out
