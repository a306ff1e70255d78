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
