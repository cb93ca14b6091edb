// ref_driver.cpp -- runs the UNMODIFIED reference main() once per "rank" thread.
// TEST INFRASTRUCTURE ONLY.  REF_SRC is passed by build_ref.sh and points at
// /root/reference/mpi/mpi_lsbsort.cpp where it lies (never copied into the repo).
#include <thread>
#include <vector>
#include <cstdlib>
#include "mpi.h"

#define main lsb_reference_main
#include REF_SRC
#undef main

int main(int argc, char** argv) {
  const char* e = std::getenv("SHIM_RANKS");
  int ranks = e ? std::atoi(e) : 1;
  if (ranks < 1) ranks = 1;
  shim::world().size = ranks;
  shim::world().slots.resize(ranks);
  std::vector<std::thread> th;
  std::vector<int> rc(ranks, 0);
  for (int r = 0; r < ranks; r++)
    th.emplace_back([&, r] {
      shim::my_rank = r;
      rc[r] = lsb_reference_main(argc, argv);
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < ranks; r++)
    if (rc[r]) return rc[r];
  return 0;
}
