// pcg_random.hpp -- satisfies `#include <pcg_random.hpp>` of the reference
// (mpi/mpi_lsbsort.cpp:15) with the copy of imneme/pcg-cpp that pyarrow vendors
// (namespace arrow_vendored).  build_ref.sh puts pyarrow's include dir on the path.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <arrow/vendored/pcg/pcg_random.hpp>
using namespace arrow_vendored;
