#!/bin/bash
# Builds the UNMODIFIED reference mpi/mpi_lsbsort.cpp into oracle/_ref/mpi_lsbsort_shim
# (ranks = threads, SHIM_RANKS=<R>) straight from /root/reference; nothing is copied.
# TEST INFRASTRUCTURE ONLY.  Skips quietly (exit 0) when /root/reference is absent
# (the GPU box): the prebuilt binary travels with the gpurun snapshot.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
ref="${LSB_REFERENCE_DIR:-/root/reference}/mpi/mpi_lsbsort.cpp"
if [ ! -f "$ref" ]; then echo "build_ref: $ref not present, keeping prebuilt oracle/_ref"; exit 0; fi
pa_inc="$(python -c 'import pyarrow, os; print(os.path.join(os.path.dirname(pyarrow.__file__), "include"))')"
mkdir -p "$here/_ref"
g++ -O3 -std=c++17 -pthread -w -I "$here/shim" -I "$pa_inc" -DREF_SRC="\"$ref\"" \
    "$here/shim/ref_driver.cpp" -o "$here/_ref/mpi_lsbsort_shim"
# no -DNDEBUG, as in the reference's own build line (mpi/README.md:18): its asserts stay live.
echo "built $here/_ref/mpi_lsbsort_shim"
