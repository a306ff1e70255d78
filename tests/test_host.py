"""CPU: host logic of the product (input parsing, sizes, knots) and the C-ABI library surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import bspatom_b200 as bsp
from bspatom_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIPPED_INPUT = """!		INPUT FOR BSP_ATOM PROGRAM
&VARS_BSP KIND_GRID=2 rmax=60.0D0 ra=0.0D0 rb=500.0D0
 k=7 nfun=100 KIND_BC1=0 KIND_BC2=0 &end
! VARS_TISE
&VARS_TISE n0_ini=1 l_ini=0 m_ini=0 l_fin=2 Emax_fin=1.50D0 Zatom=1.0D0 &end
&VARS_FIELD KIND_PI=0 I0=1.0D15 Eph=0.75D0 w0=1.0D0 b0=0.0D0 nEpts=-200 Eref=0.125D0
 nthpts=250 moam=1 mph=-1 KIND_SCP=0 ncyc=10 &end
"""  # same values as exec/bsp_0.inp:8-9,12,21-22 (the GPU box has no /root/reference)


def test_read_inputs_shipped():
    a = bsp.BspInputs().READ_INPUTS(SHIPPED_INPUT)
    # ReadInputs.f90:39-69 worked example; lmax raised to l_fin (ReadInputs.f90:87)
    assert (a.k, a.ka, a.nfun, a.nkp, a.nointv, a.nintv_exp, a.nintv_lin) == (7, 10, 124, 131, 120, 36, 84)
    assert (a.nbc1, a.nbc2, a.lmax, a.KIND_PI, a.Zatom, a.Emax_fin) == (6, 6, 2, 0, 1.0, 1.5)


def test_namelist_parser_details():
    nml = bsp.parse_namelists("! c\n&VARS_BSP k=8, nfun=40 ,ra=1.5d0 rb = 2.0D1 /\n&VARS_TISE Zatom=2 lmax=3 &END")
    assert nml["VARS_BSP"]["k"] == 8 and nml["VARS_BSP"]["nfun"] == 40
    assert nml["VARS_BSP"]["ra"] == 1.5 and nml["VARS_BSP"]["rb"] == 20.0
    assert nml["VARS_TISE"]["Zatom"] == 2 and nml["VARS_TISE"]["lmax"] == 3
    assert nml["VARS_TISE"]["Emax_fin"] == -1.0          # default, ReadInputs.f90:81


@pytest.mark.parametrize("kw", [dict(kind_grid=0, nfun=1000), dict(kind_grid=1, nfun=300),
                                dict(kind_grid=2, nfun=782, rmax=70.0), dict(kind_grid=2, nfun=808, rmax=60.0),
                                dict(kind_grid=2, nfun=408, rmax=60.0), dict(kind_grid=0, nfun=4000, k=8, rb=2000.0)])
def test_grid_matches_oracle_bitwise(oracle, kw):
    kw = dict(dict(k=7, rb=500.0), **kw)
    a = bsp.BspInputs.from_values(**kw)
    b = oracle.make_basis(**kw)
    assert (a.nfun, a.nkp, a.ka, a.nbc1, a.nbc2) == (b.nfun, b.nkp, b.ka, b.nbc1, b.nbc2)
    assert np.array_equal(a.rt, b.rt)


def test_knot_end_quirk_is_preserved():
    """App. B-1: KIND_GRID=2 overwrites the first knot of the rb block; nfun0=808 gives rb+1ulp
    (non-monotone knots), nfun0=408 gives rb-1ulp.  The host must hand these over verbatim."""
    a = bsp.BspInputs.from_values(kind_grid=2, k=7, nfun=808, rb=500.0, rmax=60.0)
    assert a.nfun == 1000 and a.rt[a.nkp - a.nbc2] > 500.0 and a.rt[a.nkp - a.nbc2 + 1] == 500.0
    a = bsp.BspInputs.from_values(kind_grid=2, k=7, nfun=408, rb=500.0, rmax=60.0)
    assert a.nfun == 504 and a.rt[a.nkp - a.nbc2] < 500.0


def test_rogers_table_matches_oracle(oracle):
    a = bsp.BspInputs.from_values(kind_pot=1, zatom=20.0, nfun=50)
    assert np.array_equal(a.pot_par, oracle.pot_params(1, zatom=20.0))
    a = bsp.BspInputs.from_values(kind_pot=2, zatom=1.0, nfun=50, lmax=5)
    assert np.array_equal(a.Bl, oracle.simons_fues_bl(5))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "bspatom.h")).read()
    declared = set(re.findall(r"\b(bspatom_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bspatom_version() == 100


def test_struct_layout_matches_header(tmp_path):
    """the ctypes mirror of struct bsp_problem against the C compiler's view of include/bspatom.h: size and the
    offset of every field"""
    import subprocess

    fields = [f for f, _ in _lib.BspProblem._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bspatom.h"\nint main(void){\n'
                   'printf("%zu\\n", sizeof(bsp_problem));\n' +
                   "".join('printf("%%zu\\n", offsetof(bsp_problem, %s));\n' % f for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.BspProblem)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.BspProblem, f).offset == off, f


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(bsp.BspAtomError, match="no CUDA device"):
        bsp.BspAtom(device=0)
    # the LAPACK-shaped entry reports a device error through info, it does not compute on the CPU
    n = 8
    H = np.eye(n)
    w, A, B, info = bsp.dsygv(H, np.eye(n))
    assert info < -100


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bspatom_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "bsp_oracle" not in txt, f


def _build_c_driver(tmp_path):
    import subprocess

    exe = str(tmp_path / "c_driver")
    libdir = os.path.join(ROOT, "bspatom_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c_driver.c"), "-L", libdir, "-lbspatom", "-Wl,-rpath," + libdir, "-lm"])
    return exe


def test_c_driver_compiles_and_fails_loudly_without_a_device(tmp_path):
    """a plain C program binds the header and the library; without a GPU bspatom_create returns
    BSPATOM_ENODEVICE and nothing is computed (no CPU fallback)"""
    import subprocess

    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_extra.py::test_c_driver")
    from bspatom_b200 import build as _b

    _b.build()
    r = subprocess.run([_build_c_driver(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stdout
    # struct bsp_problem: the C compiler's layout = the ctypes mirror (= the BIND(C) type of INTEGRATION.md)
    P = _lib.BspProblem
    want = "sizeof(bsp_problem) = %d, offsets rt %d pot_par %d v_tab %d l %d ul_extra %d nvec %d sel_mode %d sel_ecut_a %d sel_ecut_b %d" % (
        ctypes.sizeof(P), P.rt.offset, P.pot_par.offset, P.v_tab.offset, P.l.offset, P.ul_extra.offset, P.nvec.offset,
        P.sel_mode.offset, P.sel_ecut_a.offset, P.sel_ecut_b.offset)
    assert want in r.stdout, r.stdout
