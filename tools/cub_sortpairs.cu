// cub_sortpairs.cu -- library yardstick for BASELINE configs[1]: cub::DeviceRadixSort::SortPairs on
// 2^k (uint64 key, uint64 value) pairs (SoA, as CUB takes them), same 128 B/element of algorithmic
// traffic per 4-pass sort.  Context only: CUB's answer is the same stable sort, but its layout (two
// arrays) and digit schedule (8 x 8 bits, onesweep) are its own.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/cub_sortpairs tools/cub_sortpairs.cu
#include <cub/cub.cuh>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void fill(uint64_t* k, uint64_t* v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ULL + 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 29; x *= 0x9E3779B97F4A7C15ULL; x ^= x >> 32;
    k[i] = x;
    v[i] = (uint64_t)i;
  }
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 30;
  const int64_t n = 1LL << lg;
  uint64_t *k0, *k1, *v0, *v1;
  CK(cudaMalloc(&k0, n * 8)); CK(cudaMalloc(&k1, n * 8)); CK(cudaMalloc(&v0, n * 8)); CK(cudaMalloc(&v1, n * 8));
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, n));
  void* tmp;
  CK(cudaMalloc(&tmp, tmp_bytes));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    fill<<<148 * 8, 512>>>(k0, v0, n);
    CK(cudaEventRecord(a));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  printf("cub::DeviceRadixSort::SortPairs u64/u64 n=2^%d: %.3f ms = %.0f M elements/s (temp %.1f MiB)\n", lg, best,
         n / best / 1e3, tmp_bytes / 1048576.0);
  return 0;
}
