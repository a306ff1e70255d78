#!/usr/bin/env bash
# oracle/build_ref.sh -- builds the REAL reference (carlosmwh1985/BspAtom, Fortran) into oracle/_ref/
# from the sources where they lie (default /root/reference/src), for use as the parity pin and CPU baseline.
# TEST INFRASTRUCTURE ONLY.  Outputs only into oracle/_ref/ (git-ignored); no reference source is copied
# into the tracked tree: the two portability edits below are applied on the fly into oracle/_ref/build/.
#
# Needs: gfortran + LAPACK/BLAS (liblapack/libblas, or OpenBLAS via LAPACK_LIBS).  Neither exists in the build
# image nor on the pool's GPU boxes (profiles/box_probe_r2.json: every Fortran compiler absent, no
# liblapack/libopenblas), so on those this script prints why and exits 3 -- parity then stays pinned by
# oracle/bsp_oracle.c + LAPACK dsygv + the extended-precision table (see DESIGN.md "Oracle").
#
# Portability edits against the ifort/MKL build of src/Makefile:18-23:
#   * Bsp_Atom.f90:59   INQUIRE(DIRECTORY=...) is an Intel extension  -> INQUIRE(FILE='CSs/.', ...)
#   * CubicSpline.f90:125 PAUSE (deleted feature)                       -> accepted by -std=legacy
#   * Funs_WignerSymbols.for:21 arithmetic IF                           -> accepted by -std=legacy
#   * LIBS = -lpthread -lm -mkl                                         -> ${LAPACK_LIBS:--llapack -lblas}
#
# After a successful build:  (cd oracle/_ref && ./Bsp_Atom_ref.x < /root/reference/exec/bsp_0.inp)
# writes Enl.dat for the shipped input (cfg1), which tests/test_oracle.py::test_reference_enl_dat_if_built
# compares with the oracle restatement.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${BSPATOM_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
FC="${FC:-gfortran}"
LAPACK_LIBS="${LAPACK_LIBS:--llapack -lblas}"

if [ ! -d "$SRC" ]; then echo "build_ref: reference sources not found at $SRC" >&2; exit 3; fi
if ! command -v "$FC" >/dev/null 2>&1; then
    echo "build_ref: no Fortran compiler ($FC) on this machine -- the reference is 100% Fortran; oracle/_ref not built" >&2
    exit 3
fi
mkdir -p "$OUT/build"
cd "$OUT/build"
# same grouping / order as src/Makefile:28-44 (Modules first: every unit USEs it)
ORDER="Modules.f90 Bsp_Atom.f90 ReadInputs.f90 matrices.f90 PhotoIon.f90 WriteWF.f90 grid.f90 CubicSpline.f90 \
bsplvb.f90 interv.f90 Ang_Ints.f90 Ang_Ints_Aux.f90 TorusFuns.f90 TorusFunsInts.f90 Funs_AssLegendre.f90 \
Funs_AssLaguerre.f90 Funs_SphHarms.f90 Funs_Bessel.f90 Funs_WignerSymbols.for"
OBJS=""
for f in $ORDER; do
    case "$f" in
        Bsp_Atom.f90) sed "s/INQUIRE( *DIRECTORY='CSs'/INQUIRE( FILE='CSs\/.'/" "$SRC/$f" > "$f" ;;
        *) ln -sf "$SRC/$f" "$f" ;;
    esac
    o="${f%.*}.o"
    "$FC" -O3 -std=legacy -ffree-line-length-none -fno-fast-math -c "$f" -o "$o"
    OBJS="$OBJS $o"
done
# shellcheck disable=SC2086
"$FC" -o "$OUT/Bsp_Atom_ref.x" $OBJS $LAPACK_LIBS -lpthread -lm
echo "build_ref: built $OUT/Bsp_Atom_ref.x"
