"""CPU: SOLVE_SYSTEM bookkeeping + writers against the literal oracle restatement
(bit-exact: integer / index work and formatted text)."""
import os

import numpy as np
import pytest

from bspatom_b200 import postproc as P
from oracle import postproc_oracle as PO


def spectra(oracle, lmax=3, nfun=60):
    b = oracle.make_basis(kind_grid=1, k=7, nfun=nfun, rb=60.0)
    m = oracle.matrix_svt(b, lmax=lmax)
    E, Cs = [], []
    for l in range(lmax + 1):
        w, v = oracle.solve_system(m, l)
        E.append(w)
        Cs.append(v)
    return np.stack(E, axis=1), Cs


@pytest.mark.parametrize("kind_pi,emax", [(1, 1.5), (2, -1.0), (3, 1.5), (3, -1.0), (5, 0.7), (8, 2.0)])
def test_selection_matches_literal_restatement(oracle, kind_pi, emax):
    Enl, Cs = spectra(oracle)
    got = P.select_states(Enl, kind_pi, l_ini=0, l_fin=1, Emax_fin=emax)
    ref = PO.solve_system_bookkeeping(Enl, kind_pi, 0, 1, emax)
    assert (got.n0_fin, got.n1_fin, got.n1_max, got.nbds) == (ref["n0_fin"], ref["n1_fin"], ref["n1_max"], ref["nbds"])
    assert got.Emax_fin == ref["Emax_fin"]
    if kind_pi >= 3:
        assert np.array_equal(got.n01, ref["n01"])
        assert np.array_equal(got.rEki, ref["rEki"])
        assert got.ntemp == ref["ntemp"]
        cinl = P.collect_cinl(Cs, got)
        assert cinl.shape == (Enl.shape[0], ref["n1_max"], Enl.shape[1])
        assert np.array_equal(cinl[:, :5, 2], Cs[2][:, :5])
    else:
        assert np.array_equal(got.E_ini, ref["E_ini"]) and np.array_equal(got.E_fin, ref["E_fin"])


def test_shipped_input_state_limits(oracle):
    """cfg1 with KIND_PI=1: hydrogen, l_fin = 2, Emax_fin = 1.5 (the shipped VARS_TISE values)."""
    b = oracle.shipped_basis()
    m = oracle.matrix_svt(b, lmax=2)
    Enl = np.stack([oracle.solve_system(m, l)[0] for l in range(3)], axis=1)
    got = P.select_states(Enl, 1, l_ini=0, l_fin=2, Emax_fin=1.5)
    ref = PO.solve_system_bookkeeping(Enl, 1, 0, 2, 1.5)
    assert (got.n0_fin, got.n1_fin) == (ref["n0_fin"], ref["n1_fin"])
    assert got.n0_fin == int(np.sum(Enl[:, 2] < 0)) + 1            # first continuum level
    assert Enl[got.n1_fin - 1, 2] <= 1.5 < Enl[got.n1_fin, 2]


def test_g_edit_descriptor():
    assert P.fortran_g(-0.499999999965092, 22, 15) == "-0.499999999965092    "
    assert P.fortran_g(586683.159432870, 22, 15) == "  586683.159432870    "
    assert P.fortran_g(1.0, 22, 15) == "  1.00000000000000    "
    assert P.fortran_g(0.0, 22, 15) == "  0.00000000000000    "
    assert P.fortran_g(4.2125693676679726e9, 22, 15) == "  4212569367.66797    "
    assert P.fortran_g(1.2345678901234567e16, 22, 15) == " 0.123456789012346E+17"
    assert P.fortran_g(-2.5e-3, 22, 15) == "-0.250000000000000E-02"
    assert P.fortran_g(0.123456789, 20, 10) == "    0.1234567890    "
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.standard_normal(300) * 10.0 ** rng.integers(-20, 20, 300),
                           [0.1, 0.09999999999999999, 1e15, 999999999999999.5, 9.9999999999999995e14, 1e-1 - 1e-17]])
    for v in vals:
        for w, d in ((22, 15), (20, 10)):
            assert P.fortran_g(float(v), w, d) == PO.g_edit(float(v), w, d), (v, w, d)


def test_writers_round_trip(tmp_path, oracle):
    Enl, Cs = spectra(oracle, lmax=1, nfun=30)
    p = os.path.join(tmp_path, "Enl.dat")
    P.write_enl(p, Enl)
    lines = open(p).read().splitlines()
    assert int(lines[0]) == 30 and len(lines) == 1 + 2 * 30
    assert lines[1].startswith("    1  ") and len(lines[1]) == 29          # T2,I4,T8,G22.15
    back = np.array([float(s[7:]) for s in lines[1:]]).reshape(2, 30).T
    assert np.allclose(back, Enl, rtol=1e-14, atol=0)       # 15 significant digits
    sel = P.select_states(Enl, 3, 0, 1, 0.5)
    cinl = P.collect_cinl(Cs, sel)
    q = os.path.join(tmp_path, "Eigenvec_All.dat")
    P.write_eigenvec_all(q, cinl)
    lines = open(q).read().splitlines()
    assert [int(x) for x in lines[0].split()] == [30, sel.n1_max, 1]
    assert int(lines[1]) == 0 and len(lines[2]) == 5 + 20 * 30            # I5,5000G20.10
    row = np.array([float(lines[2][5 + 20 * i: 25 + 20 * i]) for i in range(30)])
    assert np.allclose(row, cinl[:, 0, 0], rtol=1e-9, atol=1e-99)


def test_zhvmv_restatement_matches_blas_zhemv():
    """oracle.zhvmv / trans_amp_block (general branch of TRANS_AMP, PhotoIon.f90:218-232) against the BLAS the
    reference calls: ZHEMV('U') + ZDOTU on a NON-Hermitian complex matrix with a complex diagonal."""
    import numpy as np
    from scipy.linalg import blas

    from oracle import postproc_oracle as po

    rng = np.random.default_rng(5)
    n = 23
    zA = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Cf = rng.standard_normal((n, 4))
    Ci = rng.standard_normal((n, 3))
    T = po.trans_amp_block(zA, Cf, Ci)
    for f in range(4):
        for i in range(3):
            zx = Ci[:, i].astype(np.complex128)
            zy = Cf[:, f].astype(np.complex128)
            v = blas.zhemv(1.0, np.asfortranarray(zA), zx, lower=0)
            ref = blas.zdotu(zy, v)
            assert abs(po.zhvmv(zA, zx, zy) - ref) < 1e-12 * (1 + abs(ref))
            assert abs(T[f, i] - ref) < 1e-12 * (1 + abs(ref))


def test_three_j_exact_vs_reference_restatement_and_closed_forms():
    import numpy as np

    from bspatom_b200 import postproc as pp
    from oracle import postproc_oracle as po

    # closed forms: (l+1 1 l; 0 0 0) = (-1)^(l+1) sqrt((l+1) / ((2l+1)(2l+3)))
    for l in range(0, 12):
        exact = (-1) ** (l + 1) * np.sqrt((l + 1) / ((2 * l + 1) * (2 * l + 3)))
        assert abs(pp.three_j(l + 1, 1, l, 0, 0, 0) - exact) < 1e-15
    assert abs(pp.three_j(1, 1, 0, 0, 0, 0) + 1 / np.sqrt(3)) < 1e-15
    assert pp.three_j(2, 1, 0, 0, 0, 0) == 0.0 and pp.three_j(1, 1, 1, 0, 0, 0) == 0.0
    # the reference's log-factorial evaluation on a sweep of arguments
    for j1 in range(0, 7):
        for j3 in range(abs(j1 - 1), j1 + 2):
            for m1 in range(-j1, j1 + 1):
                for m2 in (-1, 0, 1):
                    m3 = -m1 - m2
                    if abs(m3) > j3:
                        continue
                    a, b = pp.three_j(j1, 1, j3, m1, m2, m3), po.three_j_ref(j1, 1, j3, m1, m2, m3)
                    assert abs(a - b) < 1e-13, (j1, j3, m1, m2, a, b)


def test_plane_wave_cross_sections_match_the_loop_restatement(tmp_path):
    """TRANS_AMP factors + CROSS_SECTIONS of the dipole branch (PhotoIon.f90:50-107, 300-318, 395-417) on a real
    hydrogen pencil from the oracle: vectorised host code vs the loop-by-loop restatement, both gauges."""
    import numpy as np

    from bspatom_b200 import postproc as pp
    from oracle import oracle as O
    from oracle import postproc_oracle as po

    b = O.make_basis(kind_grid=0, k=7, nfun=60, rb=40.0)
    m = O.matrix_svt(b, lmax=1)
    E0v, C0 = O.solve_system(m, 0)
    E1v, C1 = O.solve_system(m, 1)
    Enl = np.stack([E0v, E1v], axis=1)
    for kind_pi in (1, 2):
        sel = pp.select_states(Enl, kind_pi, 0, 1, 1.5)
        assert 1 <= sel.n0_fin <= sel.n1_fin < b.nfun
        ops = (m["R"], None) if kind_pi == 1 else (m["Ri"], m["D"])
        c0, c1, c2 = pp.dipole_angular_factors(kind_pi, 0, 0, 1, 0, 0)
        A = c1 * ops[0] + (c2 * ops[1] if ops[1] is not None else 0.0)
        D = C1.T @ (A @ C0[:, 0])                                   # what bspatom_dipole returns, column n0 = 1
        T = pp.trans_amp_dipole(D, sel.E_fin, sel.n0_fin, sel.n1_fin, c0)
        Ef, sig = pp.cross_sections_dipole(kind_pi, float(sel.E_ini[0]), sel.E_fin, T, sel.n0_fin, sel.n1_fin, 0)
        rEf, rT, rsig = po.photoion_dipole_ref(kind_pi, [o.tolist() if o is not None else None for o in ops],
                                               C0[:, 0].tolist(), C1.tolist(), float(sel.E_ini[0]), sel.E_fin.tolist(),
                                               sel.n0_fin, sel.n1_fin, 0, 0, 1, 0, 0)
        assert np.array_equal(Ef, np.array(rEf))
        assert np.allclose(T, rT, rtol=1e-11, atol=1e-13 * np.abs(rT).max())
        assert np.allclose(sig, rsig, rtol=1e-11, atol=1e-13 * np.abs(rsig).max())
        assert np.all(sig >= 0.0) and sig.max() > 0.0
        path = tmp_path / ("cs%d.dat" % kind_pi)
        pp.write_cross_section(str(path), Ef, sig)
        lines = path.read_text().splitlines()
        assert len(lines) == len(Ef) and all(len(ln) == 40 for ln in lines)
        back = np.array([[float(x) for x in ln.split()] for ln in lines])
        assert np.allclose(back[:, 0], Ef, rtol=1e-9) and np.allclose(back[:, 1], sig, rtol=1e-9)


def test_g20_10e3_edit_descriptor():
    from bspatom_b200.postproc import fortran_g_e3

    assert fortran_g_e3(1.0, 20, 10) == "    1.000000000     "
    assert fortran_g_e3(0.0, 20, 10) == "    0.000000000     "
    assert fortran_g_e3(-123.456, 20, 10) == "   -123.4560000     "
    assert fortran_g_e3(1.5e-7, 20, 10) == "   0.1500000000E-006"
    assert fortran_g_e3(-2.5e123, 20, 10) == "  -0.2500000000E+124"
    assert all(len(fortran_g_e3(v, 20, 10)) == 20 for v in (3.14159, 1e10, 9.9999999999e9, 0.1, 0.0999))


def test_cubspl_matches_the_statement_by_statement_restatement():
    """CUBSPL (CubicSpline.f90): vectorised host version vs the literal restatement, bit for bit on the same
    operations order; cubic reproduction away from the first interval; the klo = 1 quirk of SPLINT."""
    import numpy as np

    from bspatom_b200.postproc import cubspl
    from oracle.postproc_oracle import cubspl_ref

    rng = np.random.default_rng(3)
    x0 = np.cumsum(rng.uniform(0.05, 0.4, 41))
    y0 = np.sin(1.3 * x0) * np.exp(-0.1 * x0)
    x1 = np.concatenate(([x0[0], x0[-1], x0[7]], rng.uniform(x0[0], x0[-1], 200)))
    got = cubspl(x0, y0, x1)
    ref = np.array(cubspl_ref(x0, y0, x1))
    assert np.allclose(got, ref, rtol=0, atol=1e-15 * np.abs(ref).max())
    assert got[0] == y0[0] and got[1] == y0[-1] and abs(got[2] - y0[7]) < 1e-15
    inner = x1 > x0[1]
    assert np.max(np.abs(got[inner] - np.sin(1.3 * x1[inner]) * np.exp(-0.1 * x1[inner]))) < 5e-3
    # a straight line is reproduced everywhere, including the extrapolated first interval
    xl = np.linspace(0.0, 2.0, 11)
    assert np.allclose(cubspl(xl, 3.0 * xl - 1.0, np.array([0.05, 0.15, 1.234])), 3.0 * np.array([0.05, 0.15, 1.234]) - 1.0,
                       atol=1e-13)
    # the quirk: inside the first interval the second interval's cubic is used (differs from a proper spline)
    xq = np.array([0.0, 1.0, 2.0, 3.0, 4.0])
    yq = np.array([0.0, 1.0, 0.0, 1.0, 0.0])
    v = cubspl(xq, yq, np.array([0.5]))[0]
    assert abs(v - cubspl_ref(xq, yq, [0.5])[0]) < 1e-15


def test_trans_amp_dipole_refuses_out_of_range_stencil():
    """n0_fin < 2 or n1_fin = nfun: the reference reads E_fin out of bounds (PhotoIon.f90:97); numpy would wrap silently"""
    from bspatom_b200 import postproc

    E = np.linspace(-0.5, 3.0, 30)
    D = np.ones(30)
    with pytest.raises(ValueError):
        postproc.trans_amp_dipole(D, E, 0, 10, 1.0)
    with pytest.raises(ValueError):
        postproc.trans_amp_dipole(D, E, 3, 30, 1.0)
    assert postproc.trans_amp_dipole(D, E, 2, 29, 1.0).shape == (28,)


def test_writewf_restatement_against_write_wf():
    """WRITEWF's literal index mapping (WriteWF.f90:38,51: left - nbc1) = WRITE_WF's (Bsp_Atom.f90:133,138: left - k) on
    coefficients shifted by k - nbc1, plus c(1) times the dropped first B-spline in the first knot interval"""
    from oracle import oracle as O
    from oracle import postproc_oracle as PO

    b = O.make_basis(kind_grid=0, k=5, nfun=40, rb=20.0)
    m = O.matrix_svt(b, lmax=0)
    w, v = O.solve_system(m, 0)
    nbc1 = int(np.sum(b.rt == b.rt[0]))
    assert nbc1 == b.k - 1
    rr, out = PO.writewf(b, v, 2, 4, nbc1, 97)
    t1 = b.rt[nbc1]
    for j, n in enumerate([1, 2, 3, 4]):
        c = np.zeros(b.nfun)
        c[: b.nfun - 1] = v[1:, n - 1]
        _, psi = O.write_wf(b, c, npts=97)
        first = rr < t1
        psi[first] += ((t1 - rr[first]) / (t1 - b.ra)) ** (b.k - 1) * v[0, n - 1]
        assert first.sum() >= 2 and np.max(np.abs(out[:, j] - psi)) < 1e-14
