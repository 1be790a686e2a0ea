#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds oracle/_ref/libref_lu.so from the reference sources
# WHERE THEY LIE under /root/reference (nothing is copied into the repo: the one patched header lives in a mktemp
# directory under /tmp for the duration of the build; oracle/_ref/ holds only the .so and is git-ignored):
#   src/conflux/lu/{conflux_opt,layout}.cpp + libs/costa layout descriptors, -DCONFLUX_WITH_VALIDATION,
#   thread-backed MPI stub (oracle/mpi_stub), OpenBLAS 0.3.15 + LAPACKE from the opencv wheel.
# The ONE deviation from verbatim: a sed-inserted line in a scratch copy of conflux_opt.hpp that copies
# pivotBuff -> A00Buff when the tournament has zero rounds (Px == 1); without it the reference returns
# NaN at Px == 1 (conflux_opt.hpp:294-310,778).
# The reference's own build system (cmake + MKL + MPI) cannot run in this image: no MPI, no MKL.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${CONFLUX_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/conflux/lu" ]; then
  echo "[build_ref] $REF absent (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
SP="$(python -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')"
BLAS="$(ls "$SP"/opencv_python_headless.libs/libopenblasp-r0-*.so | head -1)"
GF_DIR="$SP/scipy.libs"
mkdir -p "$OUT"
rm -rf "$OUT/patched"                       # (older recipe kept the scratch copy here)
SCRATCH="$(mktemp -d /tmp/cflx_ref_patch.XXXXXX)"   # scratch copy of ONE header, outside the repo, deleted on exit
trap 'rm -rf "$SCRATCH"' EXIT
mkdir -p "$SCRATCH/conflux/lu"
# scratch copy with the Px==1 fix (inserted right before the rounds loop of tournament_rounds)
sed 's|^    for (int r = 0; r < n_rounds; ++r) {$|    if (n_rounds == 0) { parallel_mcopy(v, v, \&pivotBuff[0], v, \&A00Buff[0], v); } /* oracle fix: Px==1 */\n    for (int r = 0; r < n_rounds; ++r) {|' \
  "$REF/src/conflux/lu/conflux_opt.hpp" > "$SCRATCH/conflux/lu/conflux_opt.hpp"
grep -q "oracle fix: Px==1" "$SCRATCH/conflux/lu/conflux_opt.hpp" || { echo "[build_ref] patch did not apply" >&2; exit 1; }
COSTA="$REF/libs/costa/src"
g++ -O2 -DNDEBUG -std=c++17 -fopenmp -fPIC -shared -DCONFLUX_WITH_VALIDATION -w \
  -I"$HERE/mpi_stub" -I"$SCRATCH" -I"$REF/src" -I"$COSTA" \
  "$HERE/ref_driver.cpp" "$HERE/mpi_stub/mpi_threads.cpp" \
  "$REF/src/conflux/lu/conflux_opt.cpp" "$REF/src/conflux/lu/layout.cpp" \
  "$COSTA/costa/layout.cpp" \
  "$COSTA"/costa/grid2grid/{block,grid2D,interval,scalapack_layout,ranks_reordering}.cpp \
  "$BLAS" -Wl,--disable-new-dtags -Wl,-rpath,"$(dirname "$BLAS")" -Wl,-rpath,"$GF_DIR" -lpthread \
  -o "$OUT/libref_lu.so"
echo "[build_ref] built $OUT/libref_lu.so against $BLAS"
