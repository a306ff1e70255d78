"""GPU: device-side state selection (bsp_problem.sel_mode) against the reference's bookkeeping
(matrices.f90:296-334, restated in bspatom_b200.postproc.select_states / oracle.postproc_oracle), and the
device-resident consumers of the eigenvectors (dipole chain, wavefunction synthesis)."""
import numpy as np
import pytest

import bspatom_b200 as bsp
from bspatom_b200 import postproc
from cases import check_eigenpairs, host_basis

pytestmark = pytest.mark.gpu


def reference_ntemp(Es, kind_pi, Emax_fin):
    """columns of Hij the reference keeps per l (ctemp(:, 1:ntemp, l)) from the eigenvalues of all l"""
    Enl = np.stack(Es, axis=1)
    return postproc.select_states(Enl, kind_pi, 0, Enl.shape[1] - 1, Emax_fin).ntemp


@pytest.mark.parametrize("kind_pi,Emax_fin", [(3, 0.02), (8, 0.1), (5, 40.0)])
def test_selection_equals_reference_bookkeeping_shipped_input(atom, oracle, kind_pi, Emax_fin):
    """cfg1 knots, l = 0..2: the eigenvector counts chosen on the device equal ntemp of the reference's l loop;
    every eigenvalue is still returned to rounding; the selected eigenpairs meet the residual / orthonormality bars;
    fewer eigenvector bytes cross PCIe."""
    a = host_basis(kind_grid=2, k=7, nfun=100, rb=500.0, rmax=60.0)
    p = a.problem()
    items = [(p, l) for l in range(3)]
    Ef, Cf, info_f = atom.solve_batch(items)
    full_bytes = atom.stats()["c_bytes_copied"]
    sel = bsp.Selection.from_kind_pi(Emax_fin, kind_pi)
    Es, Cs, info = atom.solve_batch(items, select=sel)
    nsel = atom.selection()
    assert not info.any() and not info_f.any()
    assert list(nsel) == list(reference_ntemp(Es, kind_pi, Emax_fin)), (nsel, reference_ntemp(Es, kind_pi, Emax_fin))
    assert atom.stats()["c_bytes_copied"] == 8 * a.nfun * int(nsel.sum()) <= full_bytes
    b = oracle.shipped_basis()
    m = oracle.matrix_svt(b, lmax=2)
    for l in range(3):
        assert np.max(np.abs(Es[l] - Ef[l]) / np.maximum(np.abs(Ef[l]), 1e-2)) < 1e-13
        H = oracle.hamiltonian(m["T"], m["U"][:, :, l], m["V"])
        Cm = np.asarray(Cs[l])[:, :nsel[l]]
        check_eigenpairs(Es[l], Cm, H, m["S"], res_tol=1e-11, orth_tol=1e-10)
        ov = np.abs(np.sum(Cm * (m["S"] @ np.asarray(Cf[l])[:, :nsel[l]]), axis=0))
        assert ov.min() > 1 - 1e-9


def test_selection_groups_running_maximum_and_chunk_boundaries(atom):
    """two charges x l = 0..5 in one batch: the running maximum nlim restarts with every Problem (group), also when
    the library is forced to cut the batch into small chunks (groups are never split)."""
    a = host_basis(kind_grid=0, k=7, nfun=300, rb=150.0)
    p1 = a.problem()
    p2 = bsp.Problem(k=a.k, nfun=a.nfun, nkp=a.nkp, ka=a.ka, rt=a.rt, pot_kind=bsp.POT_COULOMB, pot_par=(2.0,))
    items = [(p1, l) for l in range(6)] + [(p2, l) for l in range(6)]
    sel = bsp.Selection.from_kind_pi(0.8, 3)
    out = {}
    for chunk in (0, 4):
        atom.set_option("chunk", chunk)
        try:
            Es, Cs, info = atom.solve_batch(items, select=sel)
        finally:
            atom.set_option("chunk", 0)
        assert not info.any()
        nsel = atom.selection()
        want = list(reference_ntemp(Es[:6], 3, 0.8)) + list(reference_ntemp(Es[6:], 3, 0.8))
        assert list(nsel) == want, (chunk, list(nsel), want)
        out[chunk] = (Es, Cs, nsel)
    for e0, e4 in zip(out[0][0], out[4][0]):
        assert np.array_equal(e0, e4)
    for c0, c4, n in zip(out[0][1], out[4][1], out[0][2]):
        assert np.array_equal(np.asarray(c0)[:, :n], np.asarray(c4)[:, :n])
    # the counts are not all equal: a real selection happened, and nlim keeps l > 0 at the l = 0 count
    assert len(set(out[0][2][:6])) >= 1 and out[0][2].max() < a.nfun


def test_selection_off_and_mixed(atom):
    """sel_mode = 0 problems in the same batch keep their nvec; Emax_fin above the spectrum selects everything"""
    a = host_basis(kind_grid=0, k=7, nfun=120, rb=60.0)
    p = a.problem()
    Es, Cs, info = atom.solve_batch([(p, 0), (p, 1)], select=bsp.Selection(1e9))
    assert list(atom.selection()) == [120, 120] and not info.any()
    E0, C0, _ = atom.solve_batch([(p, 0), (p, 1)], nvec=7)
    assert list(atom.selection()) == [7, 7]


def general_band(A, kd):
    n = A.shape[0]
    ab = np.zeros((2 * kd + 1, n), order="F")
    for j in range(n):
        for i in range(max(0, j - kd), min(n, j + kd + 1)):
            ab[kd + i - j, j] = A[i, j]
    return ab


def test_resident_dipole_chain_and_wavefunction(atom, oracle):
    """the consumers of the eigenvectors on the blocks the solver left in HBM equal the host-buffer entry points
    fed with the downloaded vectors (bit for bit: same kernels), with and without a device-side selection"""
    a = host_basis(kind_grid=0, k=7, nfun=200, rb=100.0)
    b = oracle.make_basis(kind_grid=0, k=7, nfun=200, rb=100.0)
    m = oracle.matrix_svt(b, lmax=0, want_u=False)
    Rb = general_band(m["R"], 6)
    p = a.problem()
    items = [(p, l) for l in range(4)]
    Es, Cs, info = atom.solve_batch(items)
    assert not info.any()
    D_res = atom.dipole_chain_resident(Rb, 0, 4, 150)
    r_res, psi_res = atom.wavefunction_resident(2, 3, 5, a.ra, a.rb, npts=2000)
    D_host = atom.dipole_chain(Rb, [np.asarray(c)[:, :150] for c in Cs])
    assert np.array_equal(D_res, D_host)
    atom.adopt(a)
    r_h, psi_h = atom.WRITE_WF(np.asarray(Cs[2])[:, 3:8], npts=2000)
    assert np.array_equal(r_res, r_h) and np.array_equal(psi_res, psi_h)
    # with a selection: only the selected columns are resident
    Es, Cs, info = atom.solve_batch(items, select=bsp.Selection.from_kind_pi(0.5, 3))
    nsel = atom.selection()
    nv = int(nsel.min())
    D_sel = atom.dipole_chain_resident(Rb, 0, 4, nv)
    ref = [np.asarray(Cs[l + 1])[:, :nv].T @ (m["R"] @ np.asarray(Cs[l])[:, :nv]) for l in range(3)]
    for l in range(3):
        assert np.max(np.abs(D_sel[l] - ref[l])) < 1e-12 * np.abs(ref[l]).max()
    with pytest.raises(bsp.BspAtomError):
        atom.dipole_chain_resident(Rb, 0, 4, int(nsel.min()) + 1 if nsel.min() < 200 else 201)


def test_solve_system_outputs_with_device_selection(oracle, tmp_path):
    """SURVEY.md 8(f) row f-2 end to end on the GPU path: SOLVE_SYSTEM (KIND_PI = 3, Emax_fin = 0.05: ntemp = 96 of 124) with the state
    selection on the device, then Enl.dat / Eigenvec_All.dat -- against (a) the same with every eigenvector computed and
    the selection on the host and (b) the oracle's dsygv eigenpairs pushed through the restated bookkeeping
    (matrices.f90:269-378)."""
    from oracle import postproc_oracle as PO

    def run(device_select, sub):
        atom = bsp.BspAtom(device=0)
        inp = bsp.BspInputs.from_values(kind_grid=2, k=7, nfun=100, rb=500.0, rmax=60.0, lmax=2)
        atom.adopt(inp)
        atom.KIND_PI, atom.l_ini, atom.l_fin, atom.Emax_fin = 3, 0, 1, 0.05
        atom.SOLVE_SYSTEM(device_select=device_select)
        d = tmp_path / sub
        d.mkdir()
        atom.write_outputs(str(d))
        out = (atom.Enl.copy(), atom.sel, atom.selection().copy(), open(d / "Enl.dat").read().splitlines(),
               open(d / "Eigenvec_All.dat").read().splitlines())
        atom.close()
        return out

    Enl_d, sel_d, ns_d, enl_d, vec_d = run(True, "dev")
    Enl_h, sel_h, ns_h, enl_h, vec_h = run(False, "host")
    nfun = Enl_d.shape[0]
    # the device computed exactly the ntemp(l) columns the reference keeps; the host run computed all
    assert list(ns_d) == list(sel_d.ntemp) and list(ns_h) == [nfun] * 3
    assert list(sel_d.ntemp) == list(sel_h.ntemp) and sel_d.n1_max == sel_h.n1_max and np.array_equal(sel_d.n01, sel_h.n01)
    assert max(sel_d.ntemp) < nfun                      # the selection does cut work on this input
    # eigenvalues: identical where a vector was computed (Rayleigh quotients), closed brackets elsewhere
    tol = np.maximum(1e-12 * np.abs(Enl_h), 1e-10)
    assert np.all(np.abs(Enl_d - Enl_h) <= tol)
    # files: same structure, same numbers to the printed digits for the kept states
    assert len(enl_d) == len(enl_h) == 1 + 3 * nfun and vec_d[0] == vec_h[0] and len(vec_d) == len(vec_h)
    for a, b in zip(vec_d[1:], vec_h[1:]):
        if len(a) < 30:
            assert a == b
        else:
            va = np.array([float(a[5 + 20 * i: 25 + 20 * i]) for i in range(nfun)])
            vb = np.array([float(b[5 + 20 * i: 25 + 20 * i]) for i in range(nfun)])
            assert np.max(np.abs(va - vb)) <= 1e-9 * max(1.0, np.abs(vb).max())
    # against the oracle: dsygv spectra through the restated bookkeeping
    bo = oracle.shipped_basis()
    m = oracle.matrix_svt(bo, lmax=2)
    Eo = np.stack([oracle.solve_system(m, l)[0] for l in range(3)], axis=1)
    ref = PO.solve_system_bookkeeping(Eo, 3, 0, 1, 0.05)
    assert list(ref["ntemp"]) == list(sel_d.ntemp) and ref["n1_max"] == sel_d.n1_max
    tol = np.maximum(np.maximum(1e-12 * np.abs(Eo), 1e-10), 64 * np.finfo(float).eps * np.abs(Eo).max())
    assert np.all(np.abs(Enl_d - Eo) <= tol)
