// p2p_bench.cu -- how fast can a kernel store 16-byte elements into a peer GPU over NVLink as a
// function of the contiguous run length?  (design evidence for the exchange step; 2 GPUs, one process)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_bench tools/p2p_bench.cu && ./p2p_bench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct __align__(16) Elt { uint64_t k, v; };

// thread i writes element i of the source to dst[perm(run(i)) * R + i % R]: runs of R elements land at
// pseudo-random run slots; `skew` shifts every run start by that many elements (mis-alignment)
__global__ void scatter_runs(const Elt* __restrict__ src, Elt* dst, int64_t n, int log2R, int64_t nruns, int skew) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t run = i >> log2R, within = i & ((1 << log2R) - 1);
    const int64_t slot = (run * 2654435761LL + 12345) & (nruns - 1);  // nruns is a power of two; odd multiplier = permutation
    Elt e = src[i];
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + (slot << log2R) + within + skew), "l"(e.k), "l"(e.v) : "memory");
  }
}

// the sort's pattern: tiles of T elements, each tile appends T/256 elements to each of 256 frontiers
// that are n/256 apart (destination regions of one segment); `half`: only odd bins go to dst2
__global__ void scatter_frontier(const Elt* __restrict__ src, Elt* dst, Elt* dst2, int64_t n, int T) {
  const int run = T / 256;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t region = n / 256;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t tile = i / T;
    const int p = (int)(i - tile * T);
    const int bin = p / run, off = p - bin * run;
    if (bin >= 256) continue;
    Elt e = src[i];
    Elt* base = (bin & 1) ? dst2 : dst;
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(base + bin * region + tile * run + off), "l"(e.k), "l"(e.v) : "memory");
  }
}

// the same pattern with `nb` frontiers (nb = 65536: a single-step 16-bit scatter): tiles of nb * run elements in
// input order, each appends `run` elements to every one of nb frontiers that are n / nb apart.  run = 1 is the
// "one-element appends into 65 536 tile-ordered frontiers" case: nothing coalesces inside a warp, every 128-byte
// output line is completed by 8 consecutive tiles, and 65 536 x 128 B = 8 MiB of lines are open in L2 at a time.
__global__ void scatter_frontier_nb(const Elt* __restrict__ src, Elt* dst, int64_t n, int nb, int run) {
  const int64_t T = (int64_t)nb * run;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t region = n / nb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t tile = i / T;
    const int64_t p = i - tile * T;
    const int64_t bin = p / run;
    const int off = (int)(p - bin * run);
    Elt e = src[i];
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + bin * region + tile * run + off), "l"(e.k), "l"(e.v) : "memory");
  }
}

// same, with L2 eviction-priority hints: mode bit0 = loads evict_first, bit1 = stores evict_last,
// bit2 = stores evict_first
__global__ void scatter_frontier_hint(const Elt* __restrict__ src, Elt* dst, int64_t n, int T, int mode) {
  const int run = T / 256;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t region = n / 256;
  uint64_t pol_first, pol_last;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t tile = i / T;
    const int p = (int)(i - tile * T);
    const int bin = p / run, off = p - bin * run;
    if (bin >= 256) continue;
    Elt e;
    if (mode & 1) asm volatile("ld.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(e.k), "=l"(e.v) : "l"(src + i), "l"(pol_first));
    else e = src[i];
    Elt* out = dst + bin * region + tile * run + off;
    if (mode & 2) asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(out), "l"(e.k), "l"(e.v), "l"(pol_last) : "memory");
    else if (mode & 4) asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(out), "l"(e.k), "l"(e.v), "l"(pol_first) : "memory");
    else asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(out), "l"(e.k), "l"(e.v) : "memory");
  }
}

// same pattern, but every warp store instruction covers a 512-byte ALIGNED window of the destination
// (lanes outside the run idle), so a 128-byte line inside a run is never split between two instructions
__global__ void scatter_frontier_aligned(const Elt* __restrict__ src, Elt* dst, int64_t n, int T) {
  const int run = T / 256;
  const int64_t region = n / 256;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nruns = (n / T) * 256;
  for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < nruns; r += nwarps) {
    const int64_t tile = r >> 8;
    const int bin = (int)(r & 255);
    const int64_t s0 = tile * T + (int64_t)bin * run;        // first source element of the run
    const int64_t g0 = bin * region + tile * run;            // first destination element
    const int64_t g1 = g0 + run;
    for (int64_t w = g0 & ~31LL; w < g1; w += 32) {
      const int64_t g = w + lane;
      if (g >= g0 && g < g1) {
        Elt e = src[s0 + (g - g0)];
        asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + g), "l"(e.k), "l"(e.v) : "memory");
      }
    }
  }
}

int main(int argc, char** argv) {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (argc > 2 && !strcmp(argv[2], "local")) {
    // single GPU: how fast can HBM take the sort's write pattern (256 frontiers, T/256-element runs)
    // next to a sequential read?  (upper bound for one partition launch)
    const int64_t n = 1LL << atoi(argv[1]);
    Elt *src, *dst;
    CK(cudaMalloc(&src, n * sizeof(Elt)));
    CK(cudaMalloc(&dst, (n + 64) * sizeof(Elt)));
    CK(cudaMemset(src, 1, n * sizeof(Elt)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const int Ts[] = {2048, 2816, 3840, 4096, 5120, 5632, 6144, 7680, 8192, 8448, 11264, 12288, 16384, 16640};
    for (int T : Ts) {
      float ms = 0;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a));
        scatter_frontier<<<148 * 8, 512>>>(src, dst, dst, n, T);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
      }
      printf("local frontier pattern T=%d (run %d el = %d B): %.2f ms, %.0f GB/s read+write\n", T, T / 256, T / 16, ms,
             n * 32.0 / ms / 1e6);
    }
    for (int nb : {256, 4096, 65536})
      for (int run : {1, 2, 4, 8, 10, 16, 22}) {
        if ((int64_t)nb * run > n) continue;
        float ms = 0;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaEventRecord(a));
          scatter_frontier_nb<<<148 * 8, 512>>>(src, dst, n, nb, run);
          CK(cudaEventRecord(b));
          CK(cudaEventSynchronize(b));
          CK(cudaEventElapsedTime(&ms, a, b));
        }
        printf("%d tile-ordered frontiers, %d-element (%d B) appends: %.2f ms, %.0f GB/s read+write\n", nb, run, run * 16, ms,
               n * 32.0 / ms / 1e6);
      }
    for (int T : {2816, 5632, 8448, 11264}) {
      float ms = 0;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a));
        scatter_frontier_aligned<<<148 * 8, 512>>>(src, dst, n, T);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
      }
      printf("destination-aligned warp windows, T=%d (run %d el): %.2f ms, %.0f GB/s read+write\n", T, T / 256, ms,
             (n / T) * T * 32.0 / ms / 1e6);
    }
    for (int mode = 0; mode < 8; mode++) {
      if ((mode & 6) == 6) continue;
      float ms = 0;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a));
        scatter_frontier_hint<<<148 * 8, 512>>>(src, dst, n, 5632, mode);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
      }
      printf("T=5632 hints: loads %s, stores %s: %.2f ms, %.0f GB/s read+write\n", (mode & 1) ? "evict_first" : "default",
             (mode & 2) ? "evict_last" : (mode & 4) ? "evict_first" : "default", ms, n * 32.0 / ms / 1e6);
    }
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(a));
      CK(cudaMemcpyAsync(dst, src, n * sizeof(Elt), cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("cudaMemcpy D2D: %.2f ms, %.0f GB/s read+write\n", ms, n * 32.0 / ms / 1e6);
    return 0;
  }
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const int64_t n = 1LL << (argc > 1 ? atoi(argv[1]) : 27);
  Elt *src, *local, *remote;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&remote, (n + 64) * sizeof(Elt)));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, n * sizeof(Elt)));
  CK(cudaMalloc(&local, (n + 64) * sizeof(Elt)));
  CK(cudaMemset(src, 1, n * sizeof(Elt)));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  {
    float ms = 0;
    for (int T = 2048; T <= 8192; T += 3584) {
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a));
        scatter_frontier<<<148 * 8, 512>>>(src, local, remote, n, T);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
      }
      printf("frontier pattern T=%d (run %d el), half of the bins to the peer: %.2f ms, %.0f GB/s to peer, %.0f GB/s total\n", T, T / 256,
             ms, n * 8.0 / ms / 1e6, n * 16.0 / ms / 1e6);
    }
  }
  printf("%-8s %-6s %12s %12s\n", "run(el)", "skew", "local GB/s", "peer GB/s");
  for (int skew = 0; skew <= 3; skew += 3)
    for (int lr = (argc > 2 ? 3 : 0); lr <= (argc > 2 ? 5 : 14); lr += (lr < 8 ? 1 : 3)) {
      float ms[2];
      for (int t = 0; t < 2; t++) {
        Elt* dst = t ? remote : local;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaEventRecord(a));
          scatter_runs<<<148 * 8, 512>>>(src, dst, n, lr, n >> lr, skew);
          CK(cudaEventRecord(b));
          CK(cudaEventSynchronize(b));
          CK(cudaEventElapsedTime(&ms[t], a, b));
        }
      }
      printf("%-8d %-6d %12.0f %12.0f\n", 1 << lr, skew, n * 16.0 / ms[0] / 1e6, n * 16.0 / ms[1] / 1e6);
    }
  return 0;
}
