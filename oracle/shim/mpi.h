// mpi.h -- ranks-as-threads stand-in for the handful of MPI calls that
// /root/reference/mpi/mpi_lsbsort.cpp makes.  TEST INFRASTRUCTURE ONLY: it lets the
// UNMODIFIED reference translation unit be compiled into oracle/_ref/ in a container
// with no MPI, so the oracle restatement can be pinned against the real thing.
// Each "rank" is a std::thread; collectives are publish-pointer / barrier / copy.
#pragma once
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;  // == size of one element in bytes
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_SUM 0
#define MPI_SUCCESS 0

namespace shim {
struct Slot {
  const void* sbuf = nullptr;
  const int* scounts = nullptr;
  const int* sdispls = nullptr;
  long long scalar = 0;
};
struct World {
  int size = 1;
  std::vector<Slot> slots;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long generation = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    unsigned long gen = generation;
    if (++waiting == size) {
      waiting = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};
inline World& world() {
  static World w;
  return w;
}
inline thread_local int my_rank = 0;
}  // namespace shim

inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = shim::my_rank; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = shim::world().size; return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { shim::world().barrier(); return MPI_SUCCESS; }
inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* out) { *out = n * t; return MPI_SUCCESS; }
inline int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }

inline int MPI_Alltoallv(const void* sbuf, const int* scounts, const int* sdispls, MPI_Datatype st,
                         void* rbuf, const int* rcounts, const int* rdispls, MPI_Datatype rt, MPI_Comm) {
  auto& w = shim::world();
  const int me = shim::my_rank;
  w.slots[me].sbuf = sbuf;
  w.slots[me].scounts = scounts;
  w.slots[me].sdispls = sdispls;
  w.barrier();
  for (int p = 0; p < w.size; p++) {
    const shim::Slot& s = w.slots[p];
    if (s.scounts[me] != rcounts[p]) {
      std::fprintf(stderr, "shim MPI_Alltoallv: count mismatch %d<-%d (%d vs %d)\n", me, p, rcounts[p], s.scounts[me]);
      std::abort();
    }
    std::memcpy(static_cast<char*>(rbuf) + (size_t)rdispls[p] * rt,
                static_cast<const char*>(s.sbuf) + (size_t)s.sdispls[me] * st, (size_t)rcounts[p] * rt);
  }
  w.barrier();
  return MPI_SUCCESS;
}

// only the (long long, MPI_SUM, count 1) form is used; rank 0's output is left untouched
inline int MPI_Exscan(const void* sbuf, void* rbuf, int, MPI_Datatype, MPI_Op, MPI_Comm) {
  auto& w = shim::world();
  const int me = shim::my_rank;
  w.slots[me].scalar = *static_cast<const long long*>(sbuf);
  w.barrier();
  if (me > 0) {
    long long acc = 0;
    for (int p = 0; p < me; p++) acc += w.slots[p].scalar;
    *static_cast<long long*>(rbuf) = acc;
  }
  w.barrier();
  return MPI_SUCCESS;
}

inline int MPI_Gather(const void* sbuf, long long scount, MPI_Datatype st, void* rbuf, long long,
                      MPI_Datatype, int root, MPI_Comm) {
  auto& w = shim::world();
  const int me = shim::my_rank;
  w.slots[me].sbuf = sbuf;
  w.barrier();
  if (me == root)
    for (int p = 0; p < w.size; p++)
      std::memcpy(static_cast<char*>(rbuf) + (size_t)p * scount * st, w.slots[p].sbuf, (size_t)scount * st);
  w.barrier();
  return MPI_SUCCESS;
}
