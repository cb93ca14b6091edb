// p2p_ipc_bench.cu -- same frontier-pattern peer-store test as p2p_bench.cu, but across TWO PROCESSES:
// the child owns GPU1's buffer and shares it either with legacy CUDA IPC (cudaIpcGetMemHandle) or with
// the VMM API (cuMemCreate + POSIX fd over SCM_RIGHTS).  Design evidence for how peers must be mapped.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_ipc_bench tools/p2p_ipc_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("%s: %s\n", #x, s); exit(1); } } while (0)

struct __align__(16) Elt { uint64_t k, v; };

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

static void send_fd(int sock, int fd) {
  char c = 0, ctl[CMSG_SPACE(sizeof(int))];
  iovec io = {&c, 1};
  msghdr m = {};
  m.msg_iov = &io; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof(ctl);
  cmsghdr* h = CMSG_FIRSTHDR(&m);
  h->cmsg_level = SOL_SOCKET; h->cmsg_type = SCM_RIGHTS; h->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(h), &fd, sizeof(int));
  if (sendmsg(sock, &m, 0) < 0) { perror("sendmsg"); exit(1); }
}
static int recv_fd(int sock) {
  char c, ctl[CMSG_SPACE(sizeof(int))];
  iovec io = {&c, 1};
  msghdr m = {};
  m.msg_iov = &io; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof(ctl);
  if (recvmsg(sock, &m, 0) < 0) { perror("recvmsg"); exit(1); }
  int fd;
  memcpy(&fd, CMSG_DATA(CMSG_FIRSTHDR(&m)), sizeof(int));
  return fd;
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 28;
  const bool vmm = argc > 2 && !strcmp(argv[2], "vmm");
  const int64_t n = 1LL << lg;
  const size_t bytes = (size_t)n * sizeof(Elt);
  int sv[2];
  socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
  pid_t pid = fork();
  if (pid == 0) {  // child: owns the remote buffer on GPU 1
    CK(cudaSetDevice(1));
    CK(cudaFree(0));
    if (vmm) {
      CUmemAllocationProp prop = {};
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = 1;
      prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t gran = 0;
      CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
      if (argc > 3 && (size_t)atoll(argv[3]) > gran) gran = (size_t)atoll(argv[3]);
      size_t sz = (bytes + gran - 1) / gran * gran;
      CUmemGenericAllocationHandle h;
      CU(cuMemCreate(&h, sz, &prop, 0));
      int fd = -1;
      CU(cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
      write(sv[1], &sz, sizeof(sz));
      send_fd(sv[1], fd);
    } else {
      void* p;
      CK(cudaMalloc(&p, bytes));
      cudaIpcMemHandle_t h;
      CK(cudaIpcGetMemHandle(&h, p));
      write(sv[1], &h, sizeof(h));
    }
    char done;
    read(sv[1], &done, 1);
    _exit(0);
  }
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  Elt* remote = nullptr;
  if (vmm) {
    size_t sz;
    read(sv[0], &sz, sizeof(sz));
    int fd = recv_fd(sv[0]);
    CUmemGenericAllocationHandle h;
    CU(cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va;
    const size_t align = argc > 3 ? (size_t)atoll(argv[3]) : 0;
    CU(cuMemAddressReserve(&va, sz, align, 0, 0));
    printf("va=%llx align=%zu ", (unsigned long long)va, align);
    CU(cuMemMap(va, sz, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = 0;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(va, sz, &acc, 1));
    remote = (Elt*)va;
  } else {
    cudaIpcMemHandle_t h;
    read(sv[0], &h, sizeof(h));
    CK(cudaIpcOpenMemHandle((void**)&remote, h, cudaIpcMemLazyEnablePeerAccess));
  }
  Elt *src, *local;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&local, bytes));
  CK(cudaMemset(src, 1, bytes));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(a));
    scatter_frontier<<<148 * 8, 512>>>(src, local, remote, n, 5632);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms, a, b));
  }
  printf("%s n=2^%d frontier pattern (run 22 el), half of the bins to the peer process: %.2f ms, %.0f GB/s to peer\n",
         vmm ? "VMM+fd   " : "legacyIPC", lg, ms, n * 8.0 / ms / 1e6);
  char done = 1;
  write(sv[0], &done, 1);
  waitpid(pid, nullptr, 0);
  return 0;
}
