"""GPU: the KIND_PI >= 3 branch of MATRIX_SVT (complex band matrices zAij from a tabulated zIth,
matrices.f90:110-139, 164-175) and TORMAT's all-pairs matrix elements of r (TorusFuns.f90:127-158), through the
C-ABI, against the oracle's statement-by-statement restatements."""
import numpy as np
import pytest

import bspatom_b200 as bsp
from cases import host_basis

pytestmark = pytest.mark.gpu


def make(oracle, kind_grid=0, k=7, nfun=100, rb=500.0, rmax=0.0, ka=0):
    a = host_basis(kind_grid=kind_grid, k=k, nfun=nfun, rb=rb, rmax=rmax, ka=ka)
    b = oracle.make_basis(kind_grid=a.KIND_GRID, k=a.k, ka=a.ka, nfun=nfun, ra=a.ra, rb=a.rb, rmax=a.rmax)
    assert np.array_equal(b.rt, a.rt)
    p = a.problem()
    p.xg, p.wg = b.xg, b.wg
    return a, b, p


def synthetic_zith(b, nlm, nm, ncomp, seed):
    """a smooth complex table on the quadrature grid (the real one comes from ZINT_TH on the host)"""
    rng = np.random.default_rng(seed)
    z = np.zeros((b.nkp, b.ka, nlm, nm, ncomp), dtype=np.complex128, order="F")
    for ibet in range(b.nkp - 1):
        f1, f2 = (b.rt[ibet + 1] + b.rt[ibet]) / 2.0, (b.rt[ibet + 1] - b.rt[ibet]) / 2.0
        r = f1 + b.xg * f2
        for il in range(nlm):
            for jl in range(nm):
                for c in range(ncomp):
                    a, ph, q = rng.uniform(0.5, 2.0), rng.uniform(0, 6.28), rng.uniform(0.01, 0.3)
                    z[ibet, :, il, jl, c] = a * np.exp(1j * (q * r + ph)) * np.exp(-0.01 * r) * (1.0 + 0.1 * c)
    return z


def band_of(zA_dense, k):
    n = zA_dense.shape[0]
    out = np.zeros((2 * k - 1,) + zA_dense.shape[1:], dtype=np.complex128, order="F")
    for j in range(n):
        for i in range(max(0, j - k + 1), min(n, j + k)):
            out[k - 1 + i - j, j] = zA_dense[i, j]
    return out


@pytest.mark.parametrize("kind_pi,ncomp,ncomp_out", [(3, 1, 2), (4, 3, 4), (5, 2, 2), (8, 4, 4)])
@pytest.mark.parametrize("shape", [dict(kind_grid=2, k=7, nfun=100, rmax=60.0), dict(kind_grid=0, k=8, nfun=213, rb=90.0),
                                   dict(kind_grid=0, k=4, nfun=57, rb=30.0)])
def test_zaij_parity(atom, oracle, kind_pi, ncomp, ncomp_out, shape):
    a, b, p = make(oracle, **shape)
    nlm, nm = 3, 2
    z = synthetic_zith(b, nlm, nm, ncomp, seed=kind_pi)
    ref = oracle.matrix_zaij(b, kind_pi, z, ncomp_out)
    got = atom.MATRIX_SVT_Z(kind_pi, z, ncomp_out=ncomp_out, prob=p)
    assert got.shape == (2 * a.k - 1, a.nfun, nlm, nm, ncomp_out)
    refb = band_of(ref, a.k)
    # bar: 1e-13 relative to the largest entry of the band row (the d-term B_i A B_j' has sign cancellation, like D)
    for il in range(nlm):
        for jl in range(nm):
            for c in range(ncomp_out):
                g, r = got[:, :, il, jl, c], refb[:, :, il, jl, c]
                if kind_pi in (3, 4) and c >= 2:
                    assert not np.any(g) and not np.any(r)      # zsume / zsumf never accumulated (matrices.f90:170-173)
                    continue
                colmax = np.abs(r).max(axis=0, keepdims=True)
                assert colmax.min() > 0
                assert np.max(np.abs(g - r) / colmax) < 1e-13, (il, jl, c)


def test_zaij_reduces_to_the_scalar_branch(atom, oracle):
    """zIth = 1: component 1 of KIND_PI = 3 is Rinv = int B_i B_j / r and component 2 is D = int B_i B_j'
    (matrices.f90:141-142); zIth = r with KIND_PI = 5 gives Xij = int B_i r B_j (:144)."""
    a, b, p = make(oracle, kind_grid=0, k=7, nfun=120, rb=60.0)
    band = atom.MATRIX_SVT(p)
    ones = np.ones((b.nkp, b.ka, 1, 1, 1), dtype=np.complex128, order="F")
    zA = atom.MATRIX_SVT_Z(3, ones, prob=p)[:, :, 0, 0, :]
    k = a.k
    assert np.max(np.abs(zA.imag)) == 0.0
    assert np.max(np.abs(zA[:k, :, 0].real - band["Rinv"])) <= 1e-13 * np.abs(band["Rinv"]).max()
    assert np.max(np.abs(zA[:, :, 1].real - band["D"])) <= 1e-13 * np.abs(band["D"]).max()
    rtab = np.zeros((b.nkp, b.ka, 1, 1, 2), dtype=np.complex128, order="F")
    for ibet in range(b.nkp - 1):
        rtab[ibet, :, 0, 0, 0] = (b.rt[ibet + 1] + b.rt[ibet]) / 2.0 + b.xg * ((b.rt[ibet + 1] - b.rt[ibet]) / 2.0)
    zX = atom.MATRIX_SVT_Z(5, rtab, prob=p)[:, :, 0, 0, 0]
    assert np.max(np.abs(zX[:k].real - band["R"])) <= 1e-13 * np.abs(band["R"]).max()


def test_zaij_feeds_trans_amp_hermitian(atom, oracle):
    """the assembled block goes straight into the structured-light contraction (PhotoIon.f90:218-232)"""
    from oracle import postproc_oracle as PO

    a, b, p = make(oracle, kind_grid=0, k=6, nfun=90, rb=45.0)
    z = synthetic_zith(b, 1, 1, 1, seed=11)
    zA = atom.MATRIX_SVT_Z(3, z, prob=p)[:, :, 0, 0, 0]          # general band, component c
    k = a.k
    rng = np.random.default_rng(5)
    Cf, Ci = rng.standard_normal((a.nfun, 7)), rng.standard_normal((a.nfun, 9))
    T = atom.trans_amp_hermitian(np.asfortranarray(zA[:k, :]), Cf, Ci)
    dense = oracle.matrix_zaij(b, 3, z, 2)[:, :, 0, 0, 0]
    for i in range(7):
        for j in range(9):
            ref = PO.zhvmv(dense, Ci[:, j], Cf[:, i])
            assert abs(T[i, j] - ref) <= 1e-12 * (np.abs(dense).sum(0).max() * np.linalg.norm(Ci[:, j]) * np.linalg.norm(Cf[:, i]))


def test_tormat_rvecij(atom, oracle):
    """rvecij(ni,li,nj,lj) = cinl(:,ni,li)^T Xij cinl(:,nj,lj) for all pairs (TorusFuns.f90:130-150) in one contraction"""
    a, b, p = make(oracle, kind_grid=0, k=7, nfun=140, rb=70.0)
    lmax, n1 = 2, 25
    m = oracle.matrix_svt(b, lmax=lmax)
    cinl = np.zeros((a.nfun, n1, lmax + 1), order="F")
    for l in range(lmax + 1):
        cinl[:, :, l] = oracle.solve_system(m, l)[1][:, :n1]
    ref = oracle.tormat_rvec(cinl, m["R"])
    band = atom.MATRIX_SVT(p)
    got = atom.TORMAT_RVEC(cinl, band["R"])
    assert got.shape == ref.shape == (n1, lmax + 1, n1, lmax + 1)
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 1e-12 * scale
    # <2p|r|1s> of hydrogen = 128 sqrt(6) / 243
    assert abs(abs(got[0, 0, 0, 1]) - 128 * np.sqrt(6) / 243) < 1e-4
