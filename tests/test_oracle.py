"""CPU: the oracle against its pins (analytic hydrogen, 40-digit golden spectrum, LAPACK dsygv,
structural invariants).  The reference ships no tests or golden vectors."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def test_sizes_shipped_input(oracle):
    # worked example of SURVEY.md App. A.1: nfun 100 -> 124, nkp 131, 36 exp + 84 lin intervals
    s = oracle.sizes(2, 7, 0, 100, 0, 0, 0.0, 500.0, 60.0)
    assert s == dict(ka=10, nbc1=6, nbc2=6, nkp=131, nointv=120, nfun=124, nintv_exp=36, nintv_lin=84)


def test_gauleg_is_gauss_legendre(oracle):
    for n in (4, 10, 12, 16):
        x, w = oracle.gauleg(n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        assert np.allclose(x, xr, rtol=0, atol=4e-16)
        assert np.allclose(w, wr, rtol=2e-13, atol=0)


def test_gauleg_odd_n_quirk(oracle):
    """Reference quirk (Modules.f90:134-147): for odd n the middle start value cos(pi/2) ~ 6e-17 is
    already within EPS1 of z1 = 0, the Newton loop is skipped and the middle weight is built from the
    previous node's derivative.  The default ka = k+3 is odd for even k (k=8 -> ka=11)."""
    for n in (9, 11):
        x, w = oracle.gauleg(n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        mid = n // 2
        assert np.allclose(np.delete(w, mid), np.delete(wr, mid), rtol=2e-13)
        assert abs(x[mid]) < 1e-16
        assert np.isclose(w[mid], w[mid - 1] * (1.0 - x[mid - 1] ** 2), rtol=1e-14)   # pp of node mid-1
        assert abs(w.sum() - 2.0) > 1e-2


def test_interv_semantics(oracle):
    b = oracle.shipped_basis()
    # interior point, left-continuity, end points (interv.f90:86-116)
    assert oracle.interv(b.rt, 0.0) == (b.nbc1, 0)
    assert oracle.interv(b.rt, b.rt[40]) == (41, 0)
    assert oracle.interv(b.rt, 500.0) == (b.nkp - b.nbc2, 0)
    assert oracle.interv(b.rt, 500.1) == (1, 1)
    assert oracle.interv(b.rt, -0.1) == (1, -1)


def test_bsplines_partition_of_unity_and_derivative(oracle):
    b = oracle.make_basis(kind_grid=0, k=7, nfun=60, rb=30.0)
    for r in (3.7, 11.123, 20.9):   # interior: all k splines at r belong to the basis
        left, bsp, dbsp = oracle.bspall(b, r)
        assert abs(bsp.sum() - 1.0) < 1e-14
        h = 1e-6
        _, bp, _ = oracle.bspall(b, r + h)
        _, bm, _ = oracle.bspall(b, r - h)
        assert np.allclose(dbsp, (bp - bm) / (2 * h), atol=1e-7)


def test_structure_of_assembled_matrices(oracle):
    b = oracle.shipped_basis()
    m = oracle.matrix_svt(b, lmax=2)
    n, k = b.nfun, b.k
    for nm in ("S", "T", "V", "R", "Ri"):
        A = m[nm]
        if nm in ("S", "T"):
            assert np.array_equal(A, A.T), nm                  # products commute term by term (App. B-8)
        else:                                                  # (fbra*V)*fket vs (fket*V)*fbra: rounding only;
            assert np.allclose(A, A.T, rtol=1e-14, atol=0), nm # DSYGV 'U' reads the upper triangle
        i, j = np.nonzero(A)
        assert np.max(np.abs(i - j)) == k - 1                   # half bandwidth exactly k-1
    assert np.linalg.eigvalsh(m["S"]).min() > 0
    # U_l = l(l+1) U_unit: the reference accumulates lmax+1 redundant copies (matrices.f90:148-153)
    assert np.allclose(m["U"][:, :, 2], 3.0 * m["U"][:, :, 1], rtol=1e-14, atol=0)
    assert not np.any(m["U"][:, :, 0])
    # D is the non-symmetric one: D + D^T = boundary term = 0 for functions vanishing at both ends
    assert np.abs(m["D"] + m["D"].T).max() < 1e-12


def test_fast_interval_search_equals_literal(oracle):
    b = oracle.make_basis(kind_grid=1, k=5, nfun=40, rb=50.0)
    a = oracle.matrix_svt(b, lmax=1, fast=True)
    c = oracle.matrix_svt(b, lmax=1, fast=False)
    for nm in a:
        assert np.array_equal(a[nm], c[nm]), nm


def test_hydrogen_known_answer(oracle):
    """analytic pin: level i of angular momentum l is n = i + l, E = -Z^2/(2 n^2)
    (the reference prints i+l itself, matrices.f90:263)."""
    gold = json.load(open(os.path.join(GOLD, "hydrogen.json")))["levels"]
    b = oracle.shipped_basis()
    m = oracle.matrix_svt(b, lmax=2)
    for l in range(3):
        w, v = oracle.solve_system(m, l)
        for i in range(4):
            n = i + 1 + l
            assert abs(w[i] - gold[str(n)]) < 1e-6 * abs(gold[str(n)]), (l, i)   # basis-limited
    Z = 3.0
    m = oracle.matrix_svt(oracle.make_basis(kind_grid=1, k=7, nfun=120, rb=80.0), lmax=0, par=oracle.pot_params(0, Z))
    w, _ = oracle.solve_system(m, 0)
    assert abs(w[0] + Z * Z / 2) < 1e-8


def test_golden_spectrum_shipped_input(oracle):
    """40-digit pin (tests/golden/make_golden.py): dsygv must sit inside its own backward-error
    floor c*eps*|E_max| of the truth, the low levels much closer."""
    gold = json.load(open(os.path.join(GOLD, "shipped_truth.json")))
    b = oracle.shipped_basis()
    assert gold["nfun"] == b.nfun
    m = oracle.matrix_svt(b, lmax=2)
    eps = np.finfo(float).eps
    for l in range(3):
        truth = np.array([float(s) for s in gold["levels"][str(l)]])
        w, v = oracle.solve_system(m, l)
        assert np.max(np.abs(w - truth)) < 64 * eps * truth[-1]
        H = oracle.hamiltonian(m["T"], m["U"][:, :, l], m["V"])
        assert np.abs(v.T @ m["S"] @ v - np.eye(b.nfun)).max() < 1e-12     # C^T S C = I (ITYPE=1)


def test_extended_precision_bisection_is_pinned_to_the_40_digit_table(oracle):
    """The third comparator used by the strict GPU parity tests (oracle.band_bisect_truth: Sturm bisection on
    the band in x87 extended precision) reproduces the mpmath spectrum of the shipped input to double rounding,
    for every level of l = 0, 1, 2; LAPACK's own bisection driver dsygvx is inside the strict north-star bar too."""
    gold = json.load(open(os.path.join(GOLD, "shipped_truth.json")))
    b = oracle.shipped_basis()
    m = oracle.matrix_svt(b, lmax=2)
    for l in range(3):
        truth = np.array([float(s) for s in gold["levels"][str(l)]])
        H = oracle.hamiltonian(m["T"], m["U"][:, :, l], m["V"])
        w = oracle.band_bisect_truth(H, m["S"], b.k - 1)
        assert np.max(np.abs(w - truth) / np.abs(truth)) < 2e-15
        wd, _, _ = oracle.dsygv(H, m["S"])
        w2 = oracle.band_bisect_truth(H, m["S"], b.k - 1, idx=[0, 7, 123], guess=wd)     # hinted brackets, subset
        assert np.array_equal(w2, w[[0, 7, 123]])
        strict = np.maximum(1e-12 * np.abs(truth), 1e-10)
        assert np.all(np.abs(oracle.dsygvx(H, m["S"]) - truth) <= strict)


def test_golden_band_fixture(oracle):
    z = np.load(os.path.join(GOLD, "shipped_band.npz"))
    b = oracle.shipped_basis()
    assert np.array_equal(z["rt"], b.rt)
    m = oracle.matrix_svt(b, lmax=1)
    kd = b.k - 1
    for nm in ("S", "T", "V", "R", "Ri"):
        assert np.array_equal(oracle.dense_to_band_upper(m[nm], kd), z[nm]), nm
    assert np.array_equal(oracle.dense_to_band_upper(m["U"][:, :, 1] / 2.0, kd), z["Q"])


def test_rogers_and_simons_fues_potentials(oracle):
    par = oracle.pot_params(1, zatom=20.0)            # Ca+: Z=20, Ntot=18
    assert par[1] == 18
    r = 0.7
    v = oracle.selpot(1, par, r)
    expect = -(20.0 - 18 + sum(par[2 + i] * np.exp(-par[5 + i] * r) for i in range(3))) / r
    assert abs(v - expect) < 1e-15 * abs(expect)
    assert oracle.selpot(2, oracle.pot_params(2, 1.0), 2.0) == -0.5
    b = oracle.make_basis(kind_grid=0, k=5, nfun=30, rb=20.0)
    m = oracle.matrix_svt(b, lmax=3, kind_pot=2, par=oracle.pot_params(2, 1.0))
    bl = oracle.simons_fues_bl(3)
    m0 = oracle.matrix_svt(b, lmax=3, kind_pot=0, par=oracle.pot_params(0, 1.0))
    # U_l(SF) = [l(l+1) + 2 Bl(l)] Q with Q = U_1(Coulomb)/2 ... (matrices.f90:149-152)
    Q = m0["U"][:, :, 1] / 2.0
    for l in range(4):
        assert np.allclose(m["U"][:, :, l], (l * (l + 1) + 2 * bl[l]) * Q, rtol=1e-13, atol=1e-18)


def test_write_wf_and_dipole_restatements(oracle):
    b = oracle.shipped_basis()
    m = oracle.matrix_svt(b, lmax=1)
    w0, v0 = oracle.solve_system(m, 0)
    r, psi = oracle.write_wf(b, v0[:, 0], npts=2000)
    # 1s radial function u(r) = 2 r exp(-r) (up to sign), WRITE_WF grid r_i = i*(rb-ra)/npts
    s = np.sign(psi[5])
    assert np.max(np.abs(s * psi - 2 * r * np.exp(-r))) < 2e-5
    w1, v1 = oracle.solve_system(m, 1)
    out = oracle.dipole_dots(m["R"], v0[:, 0], v1[:, :5])
    assert np.allclose(out, v1[:, :5].T @ (m["R"] @ v0[:, 0]), rtol=1e-12, atol=1e-14)
    # <2p|r|1s> = 128 sqrt(6)/243 = 1.2902663...
    assert abs(abs(out[0]) - 128 * np.sqrt(6) / 243) < 1e-6


def test_reference_enl_dat_if_built(oracle, tmp_path):
    """When oracle/_ref/Bsp_Atom_ref.x exists (oracle/build_ref.sh: needs gfortran + LAPACK, absent from the build
    image and from the pool's GPU boxes), run the REAL reference on the shipped input and pin the oracle
    restatement to its Enl.dat (matrices.f90:239-265: `nfun`, then `i, En(i)` for l = 0..lmax)."""
    import subprocess

    exe = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "Bsp_Atom_ref.x")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built: no Fortran compiler here (oracle/build_ref.sh exits 3)")
    inp = open(os.path.join(GOLD, "cfg1_shipped.inp")).read()
    subprocess.run([os.path.abspath(exe)], input=inp, text=True, cwd=tmp_path, check=True, timeout=600,
                   capture_output=True)
    toks = open(tmp_path / "Enl.dat").read().replace("D", "E").split()
    nfun = int(toks[0])
    vals = np.array([float(t) for t in toks[2::2]]).reshape(-1, nfun)      # (lmax+1, nfun)
    b = oracle.shipped_basis()
    assert nfun == b.nfun and vals.shape[0] == 3
    m = oracle.matrix_svt(b, lmax=2)
    eps = np.finfo(float).eps
    for l in range(3):
        w, _ = oracle.solve_system(m, l)
        # two LAPACK builds (MKL/reference LAPACK there, OpenBLAS here) agree within dsygv's backward error
        assert np.max(np.abs(vals[l] - w)) <= 64 * eps * np.abs(w).max() + 1e-13 * np.abs(w).max()


def test_zaij_restatement_reduces_to_scalar_branch():
    """the complex branch of the restatement (matrices.f90:110-139) against the pinned scalar branch: zIth = 1 gives
    sumc / sumd of :141-142 (Rinv, D), zIth = r gives sumr of :144 -- bitwise, the arithmetic is the same"""
    from oracle import oracle as O

    b = O.make_basis(kind_grid=2, k=7, nfun=60, rb=500.0, rmax=40.0)
    m = O.matrix_svt(b, lmax=0)
    ones = np.ones((b.nkp, b.ka, 2, 1, 1), dtype=np.complex128, order="F")
    z = O.matrix_zaij(b, 3, ones, 2)
    assert np.array_equal(z[:, :, 1, 0, 0].real, m["Ri"]) and np.array_equal(z[:, :, 0, 0, 1].real, m["D"])
    assert not np.any(z.imag)
    rtab = np.zeros((b.nkp, b.ka, 1, 1, 2), dtype=np.complex128, order="F")
    for ibet in range(b.nkp - 1):
        rtab[ibet, :, 0, 0, 0] = (b.rt[ibet + 1] + b.rt[ibet]) / 2.0 + b.xg * ((b.rt[ibet + 1] - b.rt[ibet]) / 2.0)
    rtab[:, :, 0, 0, 1] = 1j * rtab[:, :, 0, 0, 0]
    zx = O.matrix_zaij(b, 5, rtab, 2)
    assert np.array_equal(zx[:, :, 0, 0, 0].real, m["R"]) and np.array_equal(zx[:, :, 0, 0, 1].imag, m["R"])
    # KIND_PI = 4 allocates four components and only fills two (matrices.f90:164-173)
    z4 = O.matrix_zaij(b, 4, np.ones((b.nkp, b.ka, 1, 1, 3), dtype=np.complex128, order="F"), 4)
    assert not np.any(z4[..., 2:]) and np.array_equal(z4[:, :, 0, 0, 0], z[:, :, 0, 0, 0])


def test_tormat_restatement():
    """DSVMV('U') loops (TorusFuns.f90:130-150, Modules.f90:427-452) against numpy on the mirrored upper triangle"""
    from oracle import oracle as O

    rng = np.random.default_rng(1)
    n, n1, nl = 40, 5, 3
    X = rng.standard_normal((n, n))
    cinl = np.asfortranarray(rng.standard_normal((n, n1, nl)))
    Xs = np.triu(X) + np.triu(X, 1).T
    ref = np.einsum("ias,ij,jbt->asbt", cinl, Xs, cinl)
    got = O.tormat_rvec(cinl, X)
    assert np.max(np.abs(got - ref)) < 1e-12 * np.abs(ref).max()


def test_matrix_svt_against_scipy_bsplines():
    """Independent pin of the restated MATRIX_SVT (matrices.f90:68-183): the same integrals from third-party code --
    scipy.interpolate.BSpline for the basis functions and their derivatives (B_j lives on the knots rt(j..j+k)),
    numpy's Gauss-Legendre nodes -- with the reference's own quadrature rule (ka points per knot interval; ka = k+3 = 10
    is even here, so gauleg's odd-ka quirk does not enter).  S, T, r and B_i B_j' are polynomial and exact at ka points;
    V = -Z/r and the centrifugal term are not: they are compared at the SAME ka-point rule, which is what the reference
    computes."""
    from scipy.interpolate import BSpline

    from oracle import oracle as O

    for kw in (dict(kind_grid=0, k=7, nfun=40, rb=20.0), dict(kind_grid=2, k=7, nfun=60, rb=500.0, rmax=40.0)):
        b = O.make_basis(**kw)
        k, n = b.k, b.nfun
        assert b.ka % 2 == 0
        m = O.matrix_svt(b, lmax=2)
        xg, wg = np.polynomial.legendre.leggauss(b.ka)
        assert np.allclose(np.sort(b.xg), xg, rtol=0, atol=1e-15)
        funs = [BSpline.basis_element(b.rt[j:j + k + 1], extrapolate=False) for j in range(n)]
        ders = [f.derivative() for f in funs]
        mats = {nm: np.zeros((n, n)) for nm in ("S", "T", "V", "R", "Ri", "D", "U1", "U2")}
        for ibet in range(b.nkp - 1):
            ta, tb = b.rt[ibet], b.rt[ibet + 1]
            if not tb > ta:
                continue
            r = (tb + ta) / 2 + xg * (tb - ta) / 2
            dr = wg * (tb - ta) / 2
            js = [j for j in range(n) if b.rt[j] <= ta and tb <= b.rt[j + k]]
            F = {j: np.nan_to_num(funs[j](r)) for j in js}
            dF = {j: np.nan_to_num(ders[j](r)) for j in js}
            for i in js:
                for j in js:
                    mats["S"][i, j] += np.sum(F[i] * F[j] * dr)
                    mats["T"][i, j] += np.sum(dF[i] * 0.5 * dF[j] * dr)
                    mats["V"][i, j] += np.sum(F[i] * (-1.0 / r) * F[j] * dr)
                    mats["R"][i, j] += np.sum(F[i] * r * F[j] * dr)
                    mats["Ri"][i, j] += np.sum(F[i] / r * F[j] * dr)
                    mats["D"][i, j] += np.sum(F[i] * dF[j] * dr)
                    mats["U1"][i, j] += np.sum(F[i] * (2.0 / (2.0 * r * r)) * F[j] * dr)
                    mats["U2"][i, j] += np.sum(F[i] * (6.0 / (2.0 * r * r)) * F[j] * dr)
        ref = dict(S=m["S"], T=m["T"], V=m["V"], R=m["R"], Ri=m["Ri"], D=m["D"], U1=m["U"][:, :, 1], U2=m["U"][:, :, 2])
        for nm, got in mats.items():
            scale = np.abs(ref[nm]).max(axis=1, keepdims=True)
            err = np.max(np.abs(got - ref[nm]) / scale)
            assert err < 1e-13, (kw, nm, err)        # measured: 5e-15 .. 3e-14
        assert not np.any(m["U"][:, :, 0])


def test_write_wf_and_zaij_against_scipy_bsplines():
    """WRITE_WF (Bsp_Atom.f90:118-146) and the KIND_PI >= 3 branch of MATRIX_SVT (matrices.f90:110-139) restatements
    against scipy.interpolate.BSpline: psi(r) = sum_j c_j B_j(r), and zAij = sum fbra W (fket | dfket) dr with a complex
    table W on the quadrature grid."""
    from scipy.interpolate import BSpline

    from oracle import oracle as O

    b = O.make_basis(kind_grid=0, k=6, nfun=30, rb=15.0)
    k, n = b.k, b.nfun
    rng = np.random.default_rng(4)
    c = rng.standard_normal(n)
    r, psi = O.write_wf(b, c, npts=333)
    ref = sum(c[j] * np.nan_to_num(BSpline.basis_element(b.rt[j:j + k + 1], extrapolate=False)(r)) for j in range(n))
    inner = (r > b.rt[0]) & (r < b.rt[-1])
    assert np.max(np.abs(psi[inner] - ref[inner])) < 1e-13 * np.abs(ref).max()
    # complex table: smooth in r, different per block / component
    nlm, nm, ncomp = 2, 1, 2
    xg, wg = np.polynomial.legendre.leggauss(b.ka)
    if not np.allclose(np.sort(b.xg), xg, atol=1e-15):
        pytest.skip("odd ka: gauleg's middle-weight quirk, not numpy's rule")
    order = np.argsort(b.xg)                      # the reference's node order is gauleg's, not ascending
    z = np.zeros((b.nkp, b.ka, nlm, nm, ncomp), dtype=np.complex128, order="F")
    rr = np.zeros((b.nkp, b.ka))
    for ibet in range(b.nkp - 1):
        rr[ibet] = (b.rt[ibet + 1] + b.rt[ibet]) / 2 + b.xg * (b.rt[ibet + 1] - b.rt[ibet]) / 2
        for il in range(nlm):
            for cc in range(ncomp):
                z[ibet, :, il, 0, cc] = np.exp(1j * (0.3 + 0.2 * il) * rr[ibet]) * (1.0 + 0.5 * cc) / (1.0 + rr[ibet])
    zA3 = O.matrix_zaij(b, 3, z, 2)
    zA5 = O.matrix_zaij(b, 5, z, 2)
    funs = [BSpline.basis_element(b.rt[j:j + k + 1], extrapolate=False) for j in range(n)]
    ders = [f.derivative() for f in funs]
    ref3 = np.zeros((n, n, nlm, 2), dtype=np.complex128)
    ref5 = np.zeros((n, n, nlm, 2), dtype=np.complex128)
    for ibet in range(b.nkp - 1):
        ta, tb = b.rt[ibet], b.rt[ibet + 1]
        if not tb > ta:
            continue
        rq = rr[ibet]
        dr = b.wg * (tb - ta) / 2
        js = [j for j in range(n) if b.rt[j] <= ta and tb <= b.rt[j + k]]
        F = {j: np.nan_to_num(funs[j](rq)) for j in js}
        dF = {j: np.nan_to_num(ders[j](rq)) for j in js}
        for i in js:
            for j in js:
                for il in range(nlm):
                    w1, w2 = z[ibet, :, il, 0, 0], z[ibet, :, il, 0, 1]
                    ref3[i, j, il, 0] += np.sum(F[i] * (w1 / rq) * F[j] * dr)
                    ref3[i, j, il, 1] += np.sum(F[i] * w1 * dF[j] * dr)
                    ref5[i, j, il, 0] += np.sum(F[i] * w1 * F[j] * dr)
                    ref5[i, j, il, 1] += np.sum(F[i] * w2 * F[j] * dr)
    for got, ref_ in ((zA3[:, :, :, 0, :], ref3), (zA5[:, :, :, 0, :], ref5)):
        for il in range(nlm):
            for cc in range(2):
                scale = np.abs(ref_[:, :, il, cc]).max(axis=1, keepdims=True)
                assert np.max(np.abs(got[:, :, il, cc] - ref_[:, :, il, cc]) / scale) < 1e-13
    del order
