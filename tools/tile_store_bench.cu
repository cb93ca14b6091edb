// tile_store_bench.cu -- the store phase of partition_kernel in isolation: tiles of 5632 elements are
// staged in shared memory (already in bin order), then appended to 256 frontiers as runs of 22 +- 6
// elements.  What does the lane -> destination mapping of the store instructions cost?
//   mode 0: slot p of the sorted tile goes to thread p % 512 (partition_kernel's mapping): a 128-byte
//           line inside a run is split between two store instructions whenever a warp's 32-slot
//           window ends inside it
//   mode 1: one 8-lane group per destination LINE (128 B aligned), lanes outside the run idle
//   mode 2: one 16-lane group per 256-byte aligned chunk
//   mode 3: one warp per 512-byte aligned window
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tile_store_bench tools/tile_store_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct __align__(16) Elt { uint64_t k, v; };
constexpr int T = 5632, THREADS = 512, RUN = 22;
__constant__ int c_a[32];     // zero-sum jitter, doubled
__constant__ int c_cum[33];   // prefix sums of c_a

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Smem {
  Elt tile[T];
  uint64_t bar;
  long long gbase[256];   // destination index of the bin's first element of this tile
  int binstart[257];      // first slot of the bin in the sorted tile
  int ustart[257];        // first store unit of the bin (modes 1-3)
  unsigned char sbin[T];  // bin of a slot (mode 0; the real kernel reads it from the key)
  unsigned char ubin[4096];
  int wsum[16];
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) tile_store(const Elt* __restrict__ src, Elt* dst, long long ntiles, long long region) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int LOG_UNIT = MODE == 1 ? 3 : MODE == 2 ? 4 : 5;  // elements per store unit (log2)
  constexpr int UNIT = 1 << LOG_UNIT;
  {  // one CTA per tile, the tile pulled in by one bulk copy on the TMA engine (as partition_kernel does)
    const long long t = blockIdx.x;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s.bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s.bar)), "r"(T * 16) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s.tile)),
                   "l"(src + t * T), "r"(T * 16), "r"(smem_u32(&s.bar))
                   : "memory");
    }
    int nunits = 0;
    if (tid < 256) {
      const int b = tid, c = (int)((7 * t) & 15);
      const int len = RUN + c_a[(b + c) & 15];
      const int in = RUN * b + (c_cum[c + (b & 15)] - c_cum[c]);   // whole 16-bin cycles sum to zero
      long long g = (long long)b * region + RUN * t;
      for (int j = 0; j < (int)(t & 15); j++) g += c_a[(b + 7 * j) & 15];
      s.binstart[b] = in;
      s.gbase[b] = g;
      if (b == 255) s.binstart[256] = T;
      if (MODE == 0) {
        for (int j = 0; j < len; j++) s.sbin[in + j] = (unsigned char)b;
      } else {
        nunits = (int)(((g + len - 1) >> LOG_UNIT) - (g >> LOG_UNIT)) + 1;
        int incl = nunits;
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        if (lane == 31) s.wsum[warp] = incl;
        s.ustart[b] = incl - nunits;  // + warp offset below
      }
    }
    __syncthreads();
    {
      unsigned ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&s.bar)) : "memory");
    }
    if (MODE != 0) {
      if (tid < 256) {
        int off = 0;
        for (int w = 0; w < warp; w++) off += s.wsum[w];
        const int u0 = s.ustart[tid] + off;
        s.ustart[tid] = u0;
        for (int j = 0; j < nunits; j++) s.ubin[u0 + j] = (unsigned char)tid;
        if (tid == 255) s.ustart[256] = u0 + nunits;
      }
      __syncthreads();
    }
    if (MODE == 0) {
      for (int p = tid; p < T; p += THREADS) {
        const int b = s.sbin[p];
        const Elt e = s.tile[p];
        asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + s.gbase[b] + (p - s.binstart[b])), "l"(e.k), "l"(e.v) : "memory");
      }
    } else {
      const int total = s.ustart[256];
      const int grp = tid >> LOG_UNIT, j = tid & (UNIT - 1);
      for (int u = grp; u < total; u += THREADS / UNIT) {
        const int b = s.ubin[u];
        const long long g0 = s.gbase[b];
        const int len = s.binstart[b + 1] - s.binstart[b];
        const long long g = (((g0 >> LOG_UNIT) + (u - s.ustart[b])) << LOG_UNIT) + j;
        if (g >= g0 && g < g0 + len) {
          const Elt e = s.tile[s.binstart[b] + (int)(g - g0)];
          asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + g), "l"(e.k), "l"(e.v) : "memory");
        }
      }
    }
  }
}

int main(int argc, char** argv) {
  const long long n = 1LL << (argc > 1 ? atoi(argv[1]) : 30);
  const long long ntiles = n / T, region = n / 256;
  int a[16] = {-6, 5, -3, 4, 0, -5, 6, -1, 2, -4, 3, -2, 1, -1, 0, 1}, a2[32], cum[33];
  int sum = 0;
  for (int i = 0; i < 16; i++) sum += a[i];
  if (sum) { printf("jitter must sum to zero (%d)\n", sum); return 1; }
  cum[0] = 0;
  for (int i = 0; i < 32; i++) { a2[i] = a[i & 15]; cum[i + 1] = cum[i] + a2[i]; }
  CK(cudaMemcpyToSymbol(c_a, a2, sizeof(a2)));
  CK(cudaMemcpyToSymbol(c_cum, cum, sizeof(cum)));
  Elt *src, *dst;
  CK(cudaMalloc(&src, n * sizeof(Elt)));
  CK(cudaMalloc(&dst, (n + 4096) * sizeof(Elt)));
  CK(cudaMemset(src, 1, n * sizeof(Elt)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const char* names[4] = {"slot p -> thread p % 512 (partition_kernel)", "8 lanes per 128-byte destination line",
                          "16 lanes per 256-byte destination chunk", "one warp per 512-byte destination window"};
#define RUN_MODE(M)                                                                                              \
  {                                                                                                              \
    CK(cudaFuncSetAttribute(tile_store<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));     \
    float ms = 0;                                                                                                \
    for (int rep = 0; rep < 4; rep++) {                                                                          \
      CK(cudaEventRecord(e0));                                                                                   \
      tile_store<M><<<(unsigned)ntiles, THREADS, sizeof(Smem)>>>(src, dst, ntiles, region);                               \
      CK(cudaEventRecord(e1));                                                                                   \
      CK(cudaEventSynchronize(e1));                                                                              \
      CK(cudaEventElapsedTime(&ms, e0, e1));                                                                     \
    }                                                                                                            \
    printf("mode %d, %s: %.2f ms, %.0f GB/s read+write\n", M, names[M], ms, ntiles * T * 32.0 / ms / 1e6);       \
  }
  RUN_MODE(0) RUN_MODE(1) RUN_MODE(2) RUN_MODE(3)
  CK(cudaGetLastError());
  return 0;
}
