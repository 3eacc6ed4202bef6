#!/bin/bash
# Builds the UNMODIFIED BEEP reference (sources compiled where they lie under
# $BEEP_REFERENCE, default /root/reference) into oracle/_ref/:
#   oracle/_ref/obj/*.o      every src/*.cc the reference's own CMake builds
#                            (src/CMakeLists.txt:1-5: all but indbased_code.cc,
#                            runtests.cc, old_code.cc) with CMakeLists.txt:27 flags
#   oracle/_ref/BEEP         the reference executable (single-rank MPI shim)
#   oracle/_ref/ref_harness  adapter + golden-vector generator (ref_harness.cc)
#   oracle/_ref/ref_mvn      golden vectors of the ABC-SMC weight kernel (ref_mvn.cc)
#   oracle/_ref/ref_decisions golden vectors of the resampling decisions (ref_decisions.cc)
# The reference's CMake cannot be used here (find_package(MPI REQUIRED) and a
# network FetchContent); nothing is copied into the repo.  TEST INFRASTRUCTURE.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${BEEP_REFERENCE:-/root/reference}"
OUT="${BEEP_REF_OUT:-$ROOT/oracle/_ref}"
OPT="${BEEP_REF_OPT:--O3}"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF/src not found (reference absent) - nothing to do"; exit 0; fi
mkdir -p "$OUT/obj"
CXXFLAGS="$OPT -g0 -w -std=c++11 -mfma -I$ROOT/host/ref_adapter/mpi_shim -I$REF/toml11 -I$REF/src"
ls "$REF"/src/*.cc | grep -v -E "indbased_code|runtests|old_code" | \
  xargs -P "$(nproc)" -I{} sh -c 'f="{}"; b=$(basename "$f" .cc); o="'"$OUT"'/obj/$b.o"; if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then g++ '"$CXXFLAGS"' -c "$f" -o "$o"; fi'
# utils.cc once more with the two integer samplers renamed, so that the harness can
# supply poisson_sample / binomial_sample (link-time injection; sources untouched)
g++ $CXXFLAGS -Dpoisson_sample=beepref_poisson_sample -Dbinomial_sample=beepref_binomial_sample \
    -c "$REF/src/utils.cc" -o "$OUT/obj/utils_renamed.oo"
g++ -o "$OUT/BEEP" "$OUT"/obj/*.o
OBJS=$(ls "$OUT"/obj/*.o | grep -v -E "/(main|utils)\.o$")
g++ $CXXFLAGS -ffp-contract=off -I"$ROOT/include" -c "$HERE/ref_harness.cc" -o "$OUT/ref_harness.oo"
g++ -o "$OUT/ref_harness" "$OUT/ref_harness.oo" $OBJS "$OUT/obj/utils_renamed.oo"
g++ $CXXFLAGS -ffp-contract=off -c "$HERE/ref_mvn.cc" -o "$OUT/ref_mvn.oo"
g++ -o "$OUT/ref_mvn" "$OUT/ref_mvn.oo" $OBJS "$OUT/obj/utils.o"
# the decisions harness: utils.cc once more with ran() renamed, so that the harness supplies it (link-time injection)
g++ $CXXFLAGS -Dran=beepref_ran -c "$REF/src/utils.cc" -o "$OUT/obj/utils_ran.oo"
g++ $CXXFLAGS -ffp-contract=off -I"$ROOT/include" -c "$HERE/ref_decisions.cc" -o "$OUT/ref_decisions.oo"
g++ -o "$OUT/ref_decisions" "$OUT/ref_decisions.oo" $OBJS "$OUT/obj/utils_ran.oo"
echo "build_ref: built $OUT/BEEP, $OUT/ref_harness and $OUT/ref_mvn"
