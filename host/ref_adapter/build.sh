#!/bin/bash
# Builds host/_ref/BEEP_b200: the reference's own host code (UNMODIFIED sources compiled where they lie under
# $BEEP_REFERENCE, default /root/reference -- every src/*.cc the reference's CMake builds except main.cc) with
# the adapter beep_b200_main.cc in place of main.cc, linked against the engine's C-ABI library.
# The reference's CMake needs MPI and a network fetch (src/CMakeLists.txt:13-25), so the objects are compiled
# directly with its flags (CMakeLists.txt:27,34) over a single-rank mpi.h (mpi_shim/: the engine's ranks talk
# NCCL, the reference's Mpi object is left with one rank).  Nothing is copied into the repo; host/_ref/ is
# git-ignored and travels to the GPU box as a binary.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${BEEP_REFERENCE:-/root/reference}"
OUT="$ROOT/host/_ref"
if [ ! -d "$REF/src" ]; then echo "ref_adapter/build.sh: $REF/src not found (reference absent) - nothing to do"; exit 0; fi
mkdir -p "$OUT/obj"
CXXFLAGS="-O3 -g0 -w -std=c++11 -mfma -I$HERE/mpi_shim -I$REF/toml11 -I$REF/src"
ls "$REF"/src/*.cc | grep -v -E "indbased_code|runtests|old_code|/main\.cc" | \
  xargs -P "$(nproc)" -I{} sh -c 'f="{}"; b=$(basename "$f" .cc); o="'"$OUT"'/obj/$b.o"; if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then g++ '"$CXXFLAGS"' -c "$f" -o "$o"; fi'
# mode=pmcmc: the reference's PMCMC driver runs as compiled, with PMCMC::obs_prob (pmcmc.cc:180) and Mpi::particle_sample
# (mpi.cc:331) supplied by the adapter.  pmcmc.cc is compiled once more as position-independent code (calls to its own
# global members then go through the symbol instead of being bound or inlined at compile time) and the two symbols are
# made weak in COPIES of the objects, so the adapter's strong definitions win at link time.  Sources untouched.
mkdir -p "$OUT/objw"
g++ $CXXFLAGS -fPIC -fno-inline-functions -c "$REF/src/pmcmc.cc" -o "$OUT/objw/pmcmc.o"
objcopy --weaken-symbol=_ZN5PMCMC8obs_probERSt6vectorIdSaIdEE "$OUT/objw/pmcmc.o"
objcopy --weaken-symbol=_ZN3Mpi15particle_sampleEjRKSt6vectorIS0_IjSaIjEESaIS2_EERS0_I5StateSaIS7_EERK16ObservationModel "$OUT/obj/mpi.o" "$OUT/objw/mpi.o"
# mode=map: ML::generate_samples (ml.cc:389) is supplied by the adapter the same way
g++ $CXXFLAGS -fPIC -fno-inline-functions -c "$REF/src/ml.cc" -o "$OUT/objw/ml.o"
objcopy --weaken-symbol=_ZN2ML16generate_samplesERSt6vectorI11ParamSampleSaIS1_EERKS0_IdSaIdEERKS0_IS6_SaIS6_EEdR10Generation "$OUT/objw/ml.o"
OBJS=$(ls "$OUT"/obj/*.o | grep -v -E "/(pmcmc|mpi|ml)\.o$")
g++ $CXXFLAGS -std=c++14 -I"$ROOT/include" -I"$HERE" -c "$HERE/beep_b200_main.cc" -o "$OUT/beep_b200_main.oo"
g++ -o "$OUT/BEEP_b200" "$OUT/beep_b200_main.oo" $OBJS "$OUT/objw/pmcmc.o" "$OUT/objw/mpi.o" "$OUT/objw/ml.o" -L"$ROOT/beep_b200/lib" -lbeep_b200 -Wl,-rpath,'$ORIGIN/../../beep_b200/lib' -lpthread
echo "ref_adapter/build.sh: built $OUT/BEEP_b200"
