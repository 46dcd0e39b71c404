#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds oracle/_ref/ from the UNMODIFIED reference sources where they
# lie under $DOT_REFERENCE (default /root/reference).  Nothing is copied into the repo: only
# object files / binaries land in oracle/_ref/ (git-ignored, but shipped to the GPU box).
# Recipe = SURVEY.md Appendix A: g++ -O3 -DNDEBUG -std=gnu++17 -mavx2 -mfma, OpenMP shim for
# tbb::parallel_for, CHOLMOD 3.0.12 / AMD / CAMD / COLAMD / CCOLAMD / METIS 5.1.0 straight from
# the vendored sources, BLAS/LAPACK = the OpenBLAS 0.3.15 .so bundled with opencv_python_headless.
# The reference's own CMake build is NOT run (it downloads TBB and needs GLFW/OpenGL).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${DOT_REFERENCE:-/root/reference}"
OUT="$HERE/../_ref"
OBJ="$OUT/obj"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$OBJ/metis" "$OBJ/ss" "$OBJ/dot"

BLASDIR="$(python3 - <<'EOF'
import glob, os, sys, sysconfig
sp = sysconfig.get_paths()["purelib"]
c = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so"))
print(os.path.dirname(c[0]) if c else "")
EOF
)"
if [ -z "$BLASDIR" ]; then echo "build_ref: no OpenBLAS found" >&2; exit 1; fi
BLASLIB="$(ls "$BLASDIR"/libopenblas*.so | head -1)"

M="$REF/SuiteSparse/metis-5.1.0"
SS="$REF/SuiteSparse"

# ---------- METIS 5.1.0 (IDXTYPEWIDTH 64, REALTYPEWIDTH 32 as vendored) ----------
metis_cc() {
  local src="$1" tag="$2"
  local o="$OBJ/metis/${tag}_$(basename "${src%.c}").o"
  [ "$o" -nt "$src" ] || gcc -O2 -w -fPIC -DLINUX -D_FILE_OFFSET_BITS=64 -DNDEBUG -DNDEBUG2 -DHAVE_EXECINFO_H -DHAVE_GETLINE \
      -std=c99 -D_GNU_SOURCE -I"$M/GKlib" -I"$M/include" -I"$M/libmetis" -c "$src" -o "$o"
}
export -f metis_cc; export OBJ M
ls "$M"/GKlib/*.c | xargs -P "$JOBS" -I{} bash -c 'metis_cc {} gk'
ls "$M"/libmetis/*.c | xargs -P "$JOBS" -I{} bash -c 'metis_cc {} lm'
rm -f "$OUT/libmetis.a"; ar rcs "$OUT/libmetis.a" "$OBJ"/metis/*.o

# ---------- SuiteSparse: config + AMD/CAMD/COLAMD/CCOLAMD (int) + CHOLMOD (int) ----------
SSINC="-I$SS/SuiteSparse_config -I$SS/AMD/Include -I$SS/CAMD/Include -I$SS/COLAMD/Include -I$SS/CCOLAMD/Include -I$SS/CHOLMOD/Include -I$M/include"
ss_cc() {
  local src="$1" tag="$2"; shift 2
  local o="$OBJ/ss/${tag}_$(basename "${src%.c}").o"
  [ "$o" -nt "$src" ] || gcc -O3 -w -fPIC -fexceptions -DNDEBUG $SSINC "$@" -c "$src" -o "$o"
}
export -f ss_cc; export SS SSINC
ss_cc "$SS/SuiteSparse_config/SuiteSparse_config.c" cfg
for pkg in AMD CAMD COLAMD CCOLAMD; do
  ls "$SS/$pkg"/Source/*.c | xargs -P "$JOBS" -I{} bash -c "ss_cc {} $pkg -DDINT"
done
for sub in Core Check Cholesky MatrixOps Modify Partition Supernodal; do
  ls "$SS"/CHOLMOD/$sub/cholmod_*.c | xargs -P "$JOBS" -I{} bash -c "ss_cc {} ch_$sub -DDINT"
done
rm -f "$OUT/libsuitesparse_dot.a"; ar rcs "$OUT/libsuitesparse_dot.a" "$OBJ"/ss/*.o

# ---------- DOT sources (unmodified) + headless driver ----------
S="$REF/src"
INC="-I$HERE/shim -I$S -I$S/Energy -I$S/Energy/Physics_Elasticity -I$S/Utils -I$S/Utils/SVD -I$S/LinSysSolver -I$S/TimeStepper \
 -I$S/Utils/SVD_EFTYCHIOS -I$REF/libigl/include -I$REF/libigl/external/eigen -I$SS/CHOLMOD/Include -I$SS/SuiteSparse_config -I$M/include"
CXXF="-std=gnu++17 -O3 -DNDEBUG -fopenmp -mavx2 -mfma -pthread -DUSE_AVX_IMPLEMENTATION -w"
dot_cc() {
  local src="$1"
  local o="$OBJ/dot/$(basename "${src%.cpp}").o"
  [ "$o" -nt "$src" ] || g++ $CXXF $INC -c "$src" -o "$o"
}
export -f dot_cc; export INC CXXF
printf '%s\n' "$S/Energy/Energy.cpp" "$S/Energy/Physics_Elasticity/FixedCoRotEnergy.cpp" "$S/Energy/Physics_Elasticity/StableNHEnergy.cpp" \
  "$S/Mesh.cpp" "$S/Config.cpp" "$S/AnimScripter.cpp" "$S/Utils/IglUtils.cpp" "$S/LinSysSolver/CHOLMODSolver.cpp" \
  "$S/TimeStepper/Optimizer.cpp" "$S/TimeStepper/ADMMDDTimeStepper.cpp" "$S/TimeStepper/DOTTimeStepper.cpp" "$S/TimeStepper/LBFGSTimeStepper.cpp" \
  "$S/Utils/SVD_EFTYCHIOS/Singular_Value_Decomposition_Helper.cpp" "$S/Utils/SVD_EFTYCHIOS/PTHREAD_QUEUE.cpp" \
  "$HERE/driver.cpp" | xargs -P "$JOBS" -I{} bash -c 'dot_cc {}'

g++ -fopenmp "$OBJ"/dot/*.o "$OUT/libsuitesparse_dot.a" "$OUT/libmetis.a" "$BLASLIB" -Wl,--disable-new-dtags -Wl,-rpath,"$BLASDIR" -lpthread -lm -o "$OUT/dot_ref"
echo "$BLASDIR" > "$OUT/blasdir.txt"

# ---------- drop-in build: the SAME unmodified reference steppers over libdotgpu (integration/dropin) ----------
# integration/dropin precedes src/LinSysSolver on the include path, so every `#include "CHOLMODSolver.hpp"` of the
# reference resolves to the libdotgpu-backed class; CHOLMODSolver.cpp is not compiled.  Needs dot_b200/libdotgpu.so.
REPO="$(cd "$HERE/../.." && pwd)"
if [ -f "$REPO/dot_b200/libdotgpu.so" ]; then
  mkdir -p "$OBJ/dot_gpu"
  GINC="-I$REPO/integration/dropin -I$REPO/include $INC"
  dotgpu_cc() {
    local src="$1"
    local o="$OBJ/dot_gpu/$(basename "${src%.cpp}").o"
    [ "$o" -nt "$src" ] && [ "$o" -nt "$REPO/integration/dropin/CHOLMODSolver.hpp" ] && [ "$o" -nt "$REPO/integration/dropin/GpuEnergy.hpp" ] && [ "$o" -nt "$REPO/integration/dropin/GpuDOTStepper.hpp" ] && [ "$o" -nt "$REPO/include/dotgpu.h" ] \
      || g++ $CXXF -DDOTGPU_DROPIN $GINC -c "$src" -o "$o"
  }
  export -f dotgpu_cc; export GINC REPO
  printf '%s\n' "$S/Energy/Energy.cpp" "$S/Energy/Physics_Elasticity/FixedCoRotEnergy.cpp" "$S/Energy/Physics_Elasticity/StableNHEnergy.cpp" \
    "$S/Mesh.cpp" "$S/Config.cpp" "$S/AnimScripter.cpp" "$S/Utils/IglUtils.cpp" \
    "$S/TimeStepper/Optimizer.cpp" "$S/TimeStepper/ADMMDDTimeStepper.cpp" "$S/TimeStepper/DOTTimeStepper.cpp" "$S/TimeStepper/LBFGSTimeStepper.cpp" \
    "$S/Utils/SVD_EFTYCHIOS/Singular_Value_Decomposition_Helper.cpp" "$S/Utils/SVD_EFTYCHIOS/PTHREAD_QUEUE.cpp" \
    "$HERE/driver.cpp" | xargs -P "$JOBS" -I{} bash -c 'dotgpu_cc {}'
  g++ -fopenmp "$OBJ"/dot_gpu/*.o "$OUT/libmetis.a" -L"$REPO/dot_b200" -ldotgpu -Wl,--disable-new-dtags \
    -Wl,-rpath,'$ORIGIN/../../dot_b200' -lpthread -lm -o "$OUT/dot_ref_gpu"
  echo "build_ref: built $OUT/dot_ref_gpu (reference steppers over libdotgpu)"
fi
echo "build_ref: built $OUT/dot_ref (BLAS: $BLASLIB)"
