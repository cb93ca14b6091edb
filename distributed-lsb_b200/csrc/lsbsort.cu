// lsbsort.cu -- C ABI (include/lsbsort.h) over the sm_100a kernels in lsb_kernels.cuh.
//
// Host-side restatement of the reference's pass structure, mpi/mpi_lsbsort.cpp:481-585
// (globalShuffle / mySort), for device-resident shards.  A reference pass on a digit of
// up to 16 bits is executed as one or two stable 8-bit counting-sort steps:
//
//   G == 1 : [low sub-digit] A -> B, [high sub-digit] B -> A, both over the whole shard,
//            with every 256-bin histogram of the whole sort taken in ONE up-front read.
//   G  > 1 : the shard is cut into V parts (virtual ranks g*V+q).  Counts of the full digit per
//            part are known before the pass starts (first pass: counted; later: produced by the
//            previous pass's exchange kernel), NCCL reduce-scatter/all-gather + digit-major /
//            rank-minor exclusive scan (== :327-479) give every part its global offsets, then per
//            part: [low sub-digit] part -> scratch, [high sub-digit] scratch -> part (both local),
//            and an exchange kernel that stores every run straight into the owning GPU's other
//            shard over NVLink (== :530-576) while the next part is being sorted.
//
// Two stable steps (low bits, then high bits) are exactly one stable step on the full
// digit, so the array after each pass is bit-identical to the reference's.
#include "../../include/lsbsort.h"
#include "lsb_kernels.cuh"

#include <nccl.h>  // types only: the library itself is bound at run time, see NcclApi
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace lsb;

namespace {

struct SubPass {
  int shift;
  int bits;
};

struct PassPlan {
  int shift;    // first bit of the digit
  int bits;     // width of the digit (radix_bits, or the remainder for the last pass)
  int lo_bits;  // low sub-digit width (0 if the digit fits one step)
  int hi_bits;
};

}  // namespace

struct lsb_ctx {
  lsb_config cfg;
  int G = 1, my = 0;
  int64_t n = 0, per = 0, here = 0, first = 0, per_stream = 0;
  int npasses = 0;
  cudaStream_t stream = nullptr;
  Elt* buf[2] = {nullptr, nullptr};  // A, B
  int cur = 0;                        // which of buf[] holds the data
  Elt* peer[2][LSB_MAX_GPUS] = {};    // peer[b][g]: shard b of GPU g (own pointer for g == my)
  bool peer_open[2][LSB_MAX_GPUS] = {};
  uint64_t* lookback = nullptr;
  size_t lookback_tiles = 0;
  uint32_t* tile_counters = nullptr;  // [64]
  int next_counter = 0;
  int gen = 0;
  int variant = 0;   // partition tile shape (PartCfgA..F, 6 = persistent kernel), LSB_PT_VARIANT
  int num_sms = 148;
  int tile = 0;      // elements per partition tile
  unsigned long long* hist = nullptr;        // [HIST_MAX_SUB][256]
  int64_t* scan_out = nullptr;               // [HIST_MAX_SUB][257]
  unsigned long long* counts_local = nullptr;  // [65536]
  unsigned long long* counts_all = nullptr;    // [G][65536]
  int64_t* mybase = nullptr;                 // [65536]
  int64_t* localbase = nullptr;              // [65536]
  unsigned long long* next_hist = nullptr;     // [G][2][256] counted by the exchange kernel
  unsigned long long* next_hist_all = nullptr; // [G][G][2][256]
  int hist_ready_digit = -1;                 // digit whose sub-digit histograms already sit in hist[]
  // pipelined pass (virtual ranks): shard cut into V parts
  bool pipelined = false;
  int V = 4;
  int64_t vpart = 0;                         // elements per part
  Elt* scratch[2] = {nullptr, nullptr};      // part-sized scratch for the local sort of one part
  unsigned* dense_local = nullptr;           // [V][65536] counts of the full digit per part (first pass)
  unsigned* next_dense = nullptr;            // [G][V][65536] counted by the exchange kernel for the next pass
  unsigned* dense_mine = nullptr;            // [V][65536] after the reduce-scatter
  unsigned* c_all = nullptr;                 // [G*V][65536] counts of every virtual rank
  int64_t* mybase_v = nullptr;               // [V][65536]
  int64_t* localbase_v = nullptr;            // [V][65536]
  int64_t* bases_v = nullptr;                // [V][2][257]
  int64_t* seg_start_v = nullptr;            // [V][2] = {0, m_q}
  uint32_t* seg_tiles_v = nullptr;           // [V][2] = {0, tiles of part q}
  int dense_ready_digit = -1;                // digit whose dense counts already sit in c_all
  cudaStream_t xstream = nullptr;            // exchange stream (highest priority)
  cudaEvent_t ev_sorted[8] = {}, ev_xs[8] = {}, ev_x[8] = {};
  double exchange_ms_acc = 0;
  uint32_t* seg_tile_start = nullptr;        // [257]
  int64_t* one_seg_start = nullptr;          // {0, here}
  uint32_t* one_seg_tiles = nullptr;         // {0, ceil(here/TILE)}
  unsigned long long* small = nullptr;       // [64] scratch: sent[8], verify[5], barrier word, gather
  unsigned long long* small_all = nullptr;   // [G][16]
  unsigned long long* host_small = nullptr;  // pinned [8*16 + 64]
  ncclComm_t comm = nullptr;
  bool comm_ready = false;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  std::vector<cudaEvent_t> phase_ev;
  std::vector<int> phase_kind;  // 0 hist, 1 scan/collective, 2 partition
  int64_t launches = 0;
  int64_t part_elems = 0;
  int skipped = 0;
  unsigned long long* host_hist = nullptr;  // pinned [HIST_MAX_SUB][256]
  std::string err;
};

namespace {

thread_local std::string g_create_err;

// NCCL is bound with dlopen on first multi-GPU use instead of at link time: a process that
// also hosts PyTorch must end up with ONE libnccl.so.2 (torch's bundled 2.28 needs symbols the
// system 2.27 lacks), so if one is already loaded we take that, else the loader's default.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load() {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!handle) { error = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define LSB_SYM(field, name)                                            \
    *reinterpret_cast<void**>(&field) = dlsym(handle, name);           \
    if (!field) { error = std::string("dlsym ") + name; handle = nullptr; return false; }
    LSB_SYM(GetUniqueId, "ncclGetUniqueId");
    LSB_SYM(CommInitRank, "ncclCommInitRank");
    LSB_SYM(CommDestroy, "ncclCommDestroy");
    LSB_SYM(AllGather, "ncclAllGather");
    LSB_SYM(AllReduce, "ncclAllReduce");
    LSB_SYM(ReduceScatter, "ncclReduceScatter");
    LSB_SYM(GetErrorString, "ncclGetErrorString");
#undef LSB_SYM
    return true;
  }
};
NcclApi g_nccl;

int fail(lsb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return code;
}

#define CU(c, expr)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(c, LSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));     \
  } while (0)

#define NC(c, expr)                                                                          \
  do {                                                                                       \
    ncclResult_t r_ = (expr);                                                                \
    if (r_ != ncclSuccess)                                                                   \
      return fail(c, LSB_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r_));  \
  } while (0)

inline int64_t div_ceil(int64_t a, int64_t b) { return (a + b - 1) / b; }

PassPlan plan_pass(const lsb_ctx* c, int digit) {
  PassPlan p;
  p.shift = c->cfg.radix_bits * digit;
  p.bits = std::min<int>(c->cfg.radix_bits, 64 - p.shift);
  p.lo_bits = p.bits > 8 ? p.bits / 2 : 0;  // balanced split: fewer bins per step = longer runs per bin
  p.hi_bits = p.bits - p.lo_bits;
  return p;
}

int phase_mark(lsb_ctx* c, int kind) {
  if (!(c->cfg.flags & LSB_FLAG_PHASE_EVENTS)) return LSB_OK;
  cudaEvent_t ev;
  CU(c, cudaEventCreate(&ev));
  CU(c, cudaEventRecord(ev, c->stream));
  c->phase_ev.push_back(ev);
  c->phase_kind.push_back(kind);
  return LSB_OK;
}

int begin_call(lsb_ctx* c) {
  for (auto ev : c->phase_ev) cudaEventDestroy(ev);
  c->phase_ev.clear();
  c->phase_kind.clear();
  c->launches = 0;
  c->part_elems = 0;
  c->skipped = 0;
  c->next_counter = 0;
  CU(c, cudaMemsetAsync(c->tile_counters, 0, 256 * sizeof(uint32_t), c->stream));
  c->exchange_ms_acc = 0;
  CU(c, cudaEventRecord(c->ev_start, c->stream));
  return phase_mark(c, -1);
}

int end_call(lsb_ctx* c, lsb_stats* st, int passes, int subpasses) {
  CU(c, cudaEventRecord(c->ev_stop, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  if (!st) return LSB_OK;
  memset(st, 0, sizeof(*st));
  float ms = 0;
  CU(c, cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop));
  st->device_ms = ms;
  st->passes = passes;
  st->subpasses = subpasses;
  st->skipped = c->skipped;
  st->elements = c->here;
  st->kernel_launches = c->launches;
  int sp = 0;
  for (size_t i = 1; i < c->phase_ev.size(); i++) {
    float d = 0;
    CU(c, cudaEventElapsedTime(&d, c->phase_ev[i - 1], c->phase_ev[i]));
    switch (c->phase_kind[i]) {
      case 0: st->hist_ms += d; break;
      case 1: st->scan_ms += d; break;
      case 3: st->exchange_ms += d; break;
      case 2:
        st->partition_ms += d;
        if (sp < LSB_MAX_SUBPASSES) st->subpass_ms[sp] = d;
        sp++;
        break;
    }
  }
  st->partition_launches = subpasses;
  st->partition_elements = c->part_elems;
  if (c->exchange_ms_acc > 0) st->exchange_ms = c->exchange_ms_acc;
  if (c->G > 1) {
    CU(c, cudaMemcpy(c->host_small, c->small, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int g = 0; g < c->G; g++) st->sent[g] = (int64_t)c->host_small[g];
  } else {
    st->sent[0] = c->here;
  }
  return LSB_OK;
}

// ---- kernel launch helpers -----------------------------------------------------------

int launch_hist(lsb_ctx* c, const Elt* src, const SubPass* subs, int nsub) {
  // hist[] holds nsub 256-bin histograms; more than HIST_MAX_SUB sub-digits => several reads
  CU(c, cudaMemsetAsync(c->hist, 0, sizeof(unsigned long long) * 256 * HIST_MAX_SUB, c->stream));
  if (c->here > 0) {
    HistArgs a;
    memset(&a, 0, sizeof(a));
    a.src = src;
    a.m = c->here;
    a.nsub = nsub;
    for (int s = 0; s < nsub; s++) {
      a.shift[s] = subs[s].shift;
      a.mask[s] = (1u << subs[s].bits) - 1;
    }
    a.out = c->hist;
    int grid = (int)std::min<int64_t>(148 * 4, div_ceil(c->here, HIST_THREADS));
    bool bytes = true;  // sub-digits are exactly bytes 0..nsub-1 of the key (radix 8 and 16)
    for (int s = 0; s < nsub; s++) bytes = bytes && subs[s].shift == 8 * s && subs[s].bits == 8;
#define LSB_HIST(N, B) hist_kernel<N, B><<<grid, HIST_THREADS, 0, c->stream>>>(a)
    if (bytes && nsub == 8) LSB_HIST(8, true);
    else switch (nsub) {
      case 1: LSB_HIST(1, false); break;
      case 2: LSB_HIST(2, false); break;
      case 3: LSB_HIST(3, false); break;
      case 4: LSB_HIST(4, false); break;
      case 5: LSB_HIST(5, false); break;
      case 6: LSB_HIST(6, false); break;
      case 7: LSB_HIST(7, false); break;
      case 8: LSB_HIST(8, false); break;
      case 9: LSB_HIST(9, false); break;
      case 10: LSB_HIST(10, false); break;
      case 11: LSB_HIST(11, false); break;
      case 12: LSB_HIST(12, false); break;
      case 13: LSB_HIST(13, false); break;
      case 14: LSB_HIST(14, false); break;
      case 15: LSB_HIST(15, false); break;
      default: LSB_HIST(16, false); break;
    }
#undef LSB_HIST
    c->launches++;
  }
  CU(c, cudaGetLastError());
  int rc = phase_mark(c, 0);
  if (rc) return rc;
  scan256_kernel<<<nsub, 256, 0, c->stream>>>(c->hist, c->scan_out);
  c->launches++;
  CU(c, cudaGetLastError());
  return phase_mark(c, 1);
}

int launch_partition(lsb_ctx* c, const Elt* src, int shift, int bits, int seg_bits, const int64_t* seg_start,
                     const uint32_t* seg_tile_start, const int64_t* bases, int dst_buf, bool global_dst,
                     int full_shift = -1, int full_bits = 0, int64_t m_override = -1, Elt* dst_override = nullptr) {
  if (c->gen > 126) {  // tags exhausted: wipe the look-back words and start over
    CU(c, cudaMemsetAsync(c->lookback, 0, c->lookback_tiles * 256 * sizeof(uint64_t), c->stream));
    c->gen = 0;
  }
  if (c->next_counter >= 256) return fail(c, LSB_ERR_STATE, "too many partition launches in one call");
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.src = src;
  a.shift = shift;
  a.mask = (1u << bits) - 1;
  a.seg_bits = seg_bits;
  a.seg_start = seg_start;
  a.seg_tile_start = seg_tile_start;
  a.bases = bases;
  a.lookback = c->lookback;
  a.tile_counter = c->tile_counters + c->next_counter++;
  a.tag_agg = (uint64_t)(2 * c->gen + 1) << 56;
  a.tag_inc = (uint64_t)(2 * c->gen + 2) << 56;
  c->gen++;
  if (global_dst) {
    a.per = c->per;
    a.world = c->G;
    for (int g = 0; g < c->G; g++) a.dst[g] = c->peer[dst_buf][g];
  } else {
    a.per = INT64_MAX / 16;
    a.world = 1;
    a.dst[0] = dst_override ? dst_override : c->buf[dst_buf];
  }
  const int64_t m_launch = m_override >= 0 ? m_override : c->here;
  const int64_t max_tiles = div_ceil(m_launch, c->tile) + (seg_bits ? (1 << seg_bits) : 0);
  c->part_elems += m_launch;
  if (m_launch > 0) {
    const bool runs = full_shift >= 0;
    if (runs) {
      a.run_counts = c->counts_local;
      a.full_shift = full_shift;
      a.full_mask = (1u << full_bits) - 1;
    }
    static const int extra_smem = getenv("LSB_PT_EXTRA_SMEM") ? atoi(getenv("LSB_PT_EXTRA_SMEM")) : 0;  // experiment: force 1 CTA/SM
#define LSB_PART(CFG)                                                                                         \
    if (runs) partition_kernel<CFG, true><<<(unsigned)max_tiles, CFG::THREADS, CFG::SMEM + extra_smem, c->stream>>>(a);     \
    else partition_kernel<CFG, false><<<(unsigned)max_tiles, CFG::THREADS, CFG::SMEM + extra_smem, c->stream>>>(a)
    if (c->variant == 6) {  // persistent kernel: one CTA per SM, single-segment inputs only
      if (seg_bits) return fail(c, LSB_ERR_STATE, "persistent partition kernel needs a single segment");
      const unsigned grid = (unsigned)std::min<int64_t>(c->num_sms, div_ceil(m_launch, c->tile));
      if (runs) partition_persistent_kernel<PersistCfgA, true><<<grid, PersistCfgA::THREADS, PersistCfgA::SMEM, c->stream>>>(a);
      else partition_persistent_kernel<PersistCfgA, false><<<grid, PersistCfgA::THREADS, PersistCfgA::SMEM, c->stream>>>(a);
    } else
    switch (c->variant) {
      case 0: LSB_PART(PartCfgA); break;
      case 1: LSB_PART(PartCfgB); break;
      case 2: LSB_PART(PartCfgC); break;
      case 3: LSB_PART(PartCfgD); break;
      case 5: LSB_PART(PartCfgF); break;
      default: LSB_PART(PartCfgE); break;
    }
#undef LSB_PART
    c->launches++;
  }
  CU(c, cudaGetLastError());
  return phase_mark(c, 2);
}

// exchange-free barrier across GPUs, ordered on the sort stream
int stream_barrier(lsb_ctx* c) {
  if (c->G == 1) return LSB_OK;
  NC(c, g_nccl.AllReduce(c->small + 32, c->small + 32, 1, ncclUint64, ncclSum, c->comm, c->stream));
  return LSB_OK;
}

// counts of a full digit of this shard -> all-gather -> digit-major/rank-minor scan -> mybase[]
int global_offsets(lsb_ctx* c, int nb) {
  const unsigned long long* all = c->counts_local;
  if (c->G > 1) {
    NC(c, g_nccl.AllGather(c->counts_local, c->counts_all, (size_t)nb, ncclUint64, c->comm, c->stream));
    all = c->counts_all;
  }
  CU(c, cudaMemsetAsync(c->small, 0, 8 * sizeof(unsigned long long), c->stream));
  GlobalScanArgs s;
  s.counts = all;
  s.nb = nb;
  s.G = c->G;
  s.my = c->my;
  s.per = c->per;
  s.mybase = c->mybase;
  s.sent = c->small;
  global_scan_kernel<<<1, 1024, 0, c->stream>>>(s);
  c->launches++;
  CU(c, cudaGetLastError());
  return LSB_OK;
}

// one reference pass, multi-GPU shape (also correct for G == 1), exchange by direct scatter:
// the high sub-digit step stores its runs straight into the peers (LSB_FLAG_DIRECT_SCATTER)
int pass_global_direct(lsb_ctx* c, int digit, int* subpasses) {
  const PassPlan p = plan_pass(c, digit);
  const int other = c->cur ^ 1;
  const Elt* src2 = c->buf[c->cur];
  int dst_buf = other;
  const int64_t* seg_start = c->one_seg_start;
  const uint32_t* seg_tiles = c->one_seg_tiles;
  int rc;
  if (p.lo_bits > 0) {
    SubPass lo{p.shift, p.lo_bits};
    if ((rc = launch_hist(c, c->buf[c->cur], &lo, 1))) return rc;
    // scan_out[0..256] = segment starts of the shard once grouped by the low bits
    if ((rc = launch_partition(c, c->buf[c->cur], p.shift, p.lo_bits, 0, c->one_seg_start, c->one_seg_tiles,
                               c->scan_out, other, false)))
      return rc;
    (*subpasses)++;
    seg_tiles_kernel<<<1, 256, 0, c->stream>>>(c->scan_out, 1 << p.lo_bits, c->tile, c->seg_tile_start);
    c->launches++;
    CU(c, cudaGetLastError());
    src2 = c->buf[other];
    dst_buf = c->cur;
    seg_start = c->scan_out;
    seg_tiles = c->seg_tile_start;
  }
  const int nb = 1 << p.bits;
  CU(c, cudaMemsetAsync(c->counts_local, 0, sizeof(unsigned long long) * nb, c->stream));
  if (c->here > 0) {
    SegCountArgs sc;
    sc.src = src2;
    sc.m = c->here;
    sc.seg_start = seg_start;
    sc.lo_bits = p.lo_bits;
    sc.shift_hi = p.shift + p.lo_bits;
    sc.mask_hi = (1u << p.hi_bits) - 1;
    sc.out = c->counts_local;
    int grid = (int)std::min<int64_t>(148 * 4, div_ceil(c->here, HIST_THREADS));
    seg_count_kernel<<<grid, HIST_THREADS, 0, c->stream>>>(sc);
    c->launches++;
    CU(c, cudaGetLastError());
  }
  if ((rc = phase_mark(c, 0))) return rc;
  // the all-gather doubles as the barrier "every GPU is done reading the shard that is
  // about to be overwritten by its peers"
  if ((rc = global_offsets(c, nb))) return rc;
  if ((rc = phase_mark(c, 1))) return rc;
  if ((rc = launch_partition(c, src2, p.shift + p.lo_bits, p.hi_bits, p.lo_bits, seg_start, seg_tiles, c->mybase,
                             dst_buf, true)))
    return rc;
  (*subpasses)++;
  // peers' stores into my shard must have landed before anything reads it
  if ((rc = stream_barrier(c))) return rc;
  c->cur = dst_buf;
  return LSB_OK;
}

// one reference pass, multi-GPU shape (also correct for G == 1):
//   1. ONE read counts both sub-digits of the shard (256 bins each);
//   2. local stable step on the low bits, local stable step on the high bits: the shard is sorted
//      by the full digit (== localShuffle, :213-247); the second step also emits the shard's
//      counts of the full digit from the runs it writes (== counts, :226-229);
//   3. count all-gather + digit-major/rank-minor scan (== :327-479);
//   4. exchange kernel: every run goes to its global position in the owning GPU's shard (== :530-576).
int pass_global_pipelined(lsb_ctx* c, int digit, int* subpasses, bool fuse_next);

int pass_global(lsb_ctx* c, int digit, int* subpasses, bool fuse_next) {
  if (c->cfg.flags & LSB_FLAG_DIRECT_SCATTER) return pass_global_direct(c, digit, subpasses);
  if (c->pipelined) return pass_global_pipelined(c, digit, subpasses, fuse_next);
  const PassPlan p = plan_pass(c, digit);
  const int nb = 1 << p.bits;
  int rc;
  SubPass subs[2];
  int ns = 0;
  if (p.lo_bits) subs[ns++] = {p.shift, p.lo_bits};
  subs[ns++] = {p.shift + p.lo_bits, p.hi_bits};
  if (c->hist_ready_digit == digit) {  // the previous pass's exchange already counted this pass's sub-digits
    scan256_kernel<<<ns, 256, 0, c->stream>>>(c->hist, c->scan_out);
    c->launches++;
    CU(c, cudaGetLastError());
  } else if ((rc = launch_hist(c, c->buf[c->cur], subs, ns))) {
    return rc;
  }
  c->hist_ready_digit = -1;
  CU(c, cudaMemsetAsync(c->counts_local, 0, sizeof(unsigned long long) * nb, c->stream));
  for (int s = 0; s < ns; s++) {
    const bool last = (s == ns - 1);
    if ((rc = launch_partition(c, c->buf[c->cur], subs[s].shift, subs[s].bits, 0, c->one_seg_start, c->one_seg_tiles,
                               c->scan_out + (size_t)s * 257, c->cur ^ 1, false, last ? p.shift : -1, p.bits)))
      return rc;
    c->cur ^= 1;
    (*subpasses)++;
  }
  // local offsets of every digit: exclusive scan of this shard's counts in digit order
  {
    GlobalScanArgs s;
    s.counts = c->counts_local;
    s.nb = nb;
    s.G = 1;
    s.my = 0;
    s.per = INT64_MAX / 16;
    s.mybase = c->localbase;
    s.sent = nullptr;
    global_scan_kernel<<<1, 1024, 0, c->stream>>>(s);
    c->launches++;
    CU(c, cudaGetLastError());
  }
  // queued after the local sort, the all-gather is also the barrier "every GPU is done reading the
  // buffer its peers are about to overwrite"
  if ((rc = global_offsets(c, nb))) return rc;
  if ((rc = phase_mark(c, 1))) return rc;
  const int xbuf = c->cur ^ 1;
  SubPass nsubs[2];
  int next_ns = 0;
  if (fuse_next && digit + 1 < c->npasses) {
    const PassPlan q = plan_pass(c, digit + 1);
    if (q.lo_bits) nsubs[next_ns++] = {q.shift, q.lo_bits};
    nsubs[next_ns++] = {q.shift + q.lo_bits, q.hi_bits};
    CU(c, cudaMemsetAsync(c->next_hist, 0, sizeof(unsigned long long) * 512 * c->G, c->stream));
  }
  if (c->here > 0) {
    ExchArgs x;
    memset(&x, 0, sizeof(x));
    x.src = c->buf[c->cur];
    x.m = c->here;
    x.shift = p.shift;
    x.mask = (uint32_t)(nb - 1);
    x.localbase = c->localbase;
    x.mybase = c->mybase;
    x.per = c->per;
    x.world = c->G;
    for (int g = 0; g < c->G; g++) x.dst[g] = c->peer[xbuf][g];
    if (next_ns) {
      x.next_nsub = next_ns;
      for (int s = 0; s < next_ns; s++) {
        x.next_shift[s] = nsubs[s].shift;
        x.next_mask[s] = (1u << nsubs[s].bits) - 1;
      }
      x.next_hist = c->next_hist;
    }
    static const int ex_mult = getenv("LSB_EX_GRID") ? atoi(getenv("LSB_EX_GRID")) : 8;  // CTAs per SM worth of grid
    const int grid = (int)std::min<int64_t>((int64_t)c->num_sms * ex_mult, div_ceil(c->here, (int64_t)EX_THREADS * EX_U));
    exchange_kernel<<<grid, EX_THREADS, 0, c->stream>>>(x);
    c->launches++;
    CU(c, cudaGetLastError());
  }
  if ((rc = phase_mark(c, 3))) return rc;
  if (next_ns) {
    // all-gather of the per-destination counts: also the barrier "all peers' stores into my shard landed"
    if (c->G > 1) {
      NC(c, g_nccl.AllGather(c->next_hist, c->next_hist_all, (size_t)512 * c->G, ncclUint64, c->comm, c->stream));
    } else {
      CU(c, cudaMemcpyAsync(c->next_hist_all, c->next_hist, sizeof(unsigned long long) * 512, cudaMemcpyDeviceToDevice, c->stream));
    }
    next_hist_reduce_kernel<<<1, 512, 0, c->stream>>>(c->next_hist_all, c->G, c->my, c->hist);
    c->launches++;
    CU(c, cudaGetLastError());
    c->hist_ready_digit = digit + 1;
    if ((rc = phase_mark(c, 1))) return rc;
  } else if ((rc = stream_barrier(c))) {  // peers' stores into my shard must have landed before anything reads it
    return rc;
  }
  c->cur = xbuf;
  return LSB_OK;
}

// one reference pass, multi-GPU, pipelined over V parts of the shard (virtual ranks g*V+q):
//   counts of the full digit of every part are known up front (first pass: counted here; later
//   passes: produced by the previous pass's exchange kernel), so
//     part q:   local low step, local high step (compute stream)  ->  exchange (exchange stream)
//   and the exchange of part q overlaps the local sort of part q+1.
int pass_global_pipelined(lsb_ctx* c, int digit, int* subpasses, bool fuse_next) {
  const PassPlan p = plan_pass(c, digit);
  const int nb = 1 << p.bits;
  const int V = c->V;
  int rc;
  auto part_len = [&](int q) { return std::max<int64_t>(0, std::min<int64_t>(c->vpart, c->here - (int64_t)q * c->vpart)); };

  if (c->dense_ready_digit != digit) {  // nobody has counted this digit yet
    CU(c, cudaMemsetAsync(c->dense_local, 0, sizeof(unsigned) * (size_t)V * nb, c->stream));
    for (int q = 0; q < V; q++) {
      const int64_t m = part_len(q);
      if (m <= 0) continue;
      dense_count32_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(c->buf[c->cur] + (int64_t)q * c->vpart, m, p.shift,
                                                               (uint32_t)(nb - 1), c->dense_local + (size_t)q * nb);
      c->launches++;
    }
    CU(c, cudaGetLastError());
    if (c->G > 1) {
      NC(c, g_nccl.AllGather(c->dense_local, c->c_all, (size_t)V * nb, ncclUint32, c->comm, c->stream));
    } else {
      CU(c, cudaMemcpyAsync(c->c_all, c->dense_local, sizeof(unsigned) * (size_t)V * nb, cudaMemcpyDeviceToDevice, c->stream));
    }
    if ((rc = phase_mark(c, 0))) return rc;
  }
  c->dense_ready_digit = -1;

  CU(c, cudaMemsetAsync(c->small, 0, 8 * sizeof(unsigned long long), c->stream));
  {
    VrScanArgs v;
    v.counts = c->c_all;
    v.nb = nb;
    v.GV = c->G * V;
    v.first_vr = c->my * V;
    v.V = V;
    v.G = c->G;
    v.per = c->per;
    v.totals = c->counts_local;                          // [nb] u64 scratch
    v.digit_base = c->localbase;                         // [nb] i64 scratch
    v.mybase = c->mybase_v;
    v.sent = c->small;
    vr_totals_kernel<<<(nb + 255) / 256, 256, 0, c->stream>>>(v);
    GlobalScanArgs gs;
    gs.counts = c->counts_local;
    gs.nb = nb;
    gs.G = 1;
    gs.my = 0;
    gs.per = INT64_MAX / 16;
    gs.mybase = c->localbase;
    gs.sent = nullptr;
    global_scan_kernel<<<1, 1024, 0, c->stream>>>(gs);
    vr_place_kernel<<<(nb + 255) / 256, 256, 0, c->stream>>>(v);
    c->launches += 2;
    PartPrepArgs pp;
    pp.counts = c->c_all + (size_t)c->my * V * nb;
    pp.nb = nb;
    pp.lo_bits = p.lo_bits;
    pp.hi_bits = p.hi_bits;
    pp.localbase = c->localbase_v;
    pp.bases = c->bases_v;
    part_prep_kernel<<<V, 1024, 0, c->stream>>>(pp);
    c->launches += 2;
    CU(c, cudaGetLastError());
  }
  if ((rc = phase_mark(c, 1))) return rc;

  int next_nb = 0, next_shift = 0;
  const bool has_next = fuse_next && digit + 1 < c->npasses;
  if (has_next) {
    const PassPlan q = plan_pass(c, digit + 1);
    next_nb = 1 << q.bits;
    next_shift = q.shift;
    CU(c, cudaMemsetAsync(c->next_dense, 0, sizeof(unsigned) * (size_t)c->G * V * next_nb, c->stream));
  }
  static const int ex_mult = getenv("LSB_EX_GRID_PIPE") ? atoi(getenv("LSB_EX_GRID_PIPE")) : 1;
  const int xbuf = c->cur ^ 1;
  int last_x = -1;
  for (int q = 0; q < V; q++) {
    const int64_t m = part_len(q);
    if (m <= 0) continue;
    Elt* part = c->buf[c->cur] + (int64_t)q * c->vpart;
    const Elt* sorted = part;
    const int64_t* segs = c->seg_start_v + 2 * q;
    const uint32_t* tiles = c->seg_tiles_v + 2 * q;
    if (p.lo_bits) {  // low step into the scratch, high step back in place
      if ((rc = launch_partition(c, part, p.shift, p.lo_bits, 0, segs, tiles, c->bases_v + ((size_t)q * 2 + 0) * 257, 0, false,
                                 -1, 0, m, c->scratch[0])))
        return rc;
      if ((rc = launch_partition(c, c->scratch[0], p.shift + p.lo_bits, p.hi_bits, 0, segs, tiles,
                                 c->bases_v + ((size_t)q * 2 + 1) * 257, 0, false, -1, 0, m, part)))
        return rc;
      (*subpasses) += 2;
    } else {  // a single step: its output lives in an alternating scratch until it has been exchanged
      if (q >= 2 && last_x >= 0) CU(c, cudaStreamWaitEvent(c->stream, c->ev_x[q - 2], 0));
      if ((rc = launch_partition(c, part, p.shift, p.hi_bits, 0, segs, tiles, c->bases_v + ((size_t)q * 2 + 1) * 257, 0, false,
                                 -1, 0, m, c->scratch[q & 1])))
        return rc;
      sorted = c->scratch[q & 1];
      (*subpasses)++;
    }
    CU(c, cudaEventRecord(c->ev_sorted[q], c->stream));
    CU(c, cudaStreamWaitEvent(c->xstream, c->ev_sorted[q], 0));
    CU(c, cudaEventRecord(c->ev_xs[q], c->xstream));
    ExchVrArgs x;
    memset(&x, 0, sizeof(x));
    x.src = sorted;
    x.m = m;
    x.shift = p.shift;
    x.mask = (uint32_t)(nb - 1);
    x.localbase = c->localbase_v + (size_t)q * nb;
    x.mybase = c->mybase_v + (size_t)q * nb;
    x.per = c->per;
    x.world = c->G;
    for (int g = 0; g < c->G; g++) x.dst[g] = c->peer[xbuf][g];
    x.has_next = has_next;
    x.next_shift = next_shift;
    x.next_mask = (uint32_t)(next_nb - 1);
    x.V = V;
    x.next_nb = next_nb;
    x.part = c->vpart;
    x.next_dense = c->next_dense;
    const int grid = (int)std::min<int64_t>((int64_t)c->num_sms * ex_mult, div_ceil(m, (int64_t)EX_THREADS * EX_U));
    exchange_vr_kernel<<<grid, EX_THREADS, 0, c->xstream>>>(x);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev_x[q], c->xstream));
    last_x = q;
  }
  if (last_x >= 0) CU(c, cudaStreamWaitEvent(c->stream, c->ev_x[last_x], 0));
  if ((rc = phase_mark(c, 4))) return rc;
  if (has_next) {
    // sum the G contributions to each of my parts, then share every virtual rank's counts: the two
    // collectives are also the barrier "all peers' stores into my shard have landed"
    if (c->G > 1) {
      NC(c, g_nccl.ReduceScatter(c->next_dense, c->dense_mine, (size_t)V * next_nb, ncclUint32, ncclSum, c->comm, c->stream));
      NC(c, g_nccl.AllGather(c->dense_mine, c->c_all, (size_t)V * next_nb, ncclUint32, c->comm, c->stream));
    } else {
      CU(c, cudaMemcpyAsync(c->c_all, c->next_dense, sizeof(unsigned) * (size_t)V * next_nb, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->dense_ready_digit = digit + 1;
    if ((rc = phase_mark(c, 1))) return rc;
  } else if ((rc = stream_barrier(c))) {
    return rc;
  }
  c->cur = xbuf;
  return LSB_OK;
}

// passes [d0, d1) on one GPU: one histogram read for all sub-digits, then the partitions
int passes_single(lsb_ctx* c, int d0, int d1, int* subpasses) {
  std::vector<SubPass> subs;
  for (int d = d0; d < d1; d++) {
    const PassPlan p = plan_pass(c, d);
    if (p.lo_bits) subs.push_back({p.shift, p.lo_bits});
    subs.push_back({p.shift + p.lo_bits, p.hi_bits});
  }
  int rc;
  for (size_t s0 = 0; s0 < subs.size(); s0 += HIST_MAX_SUB) {
    const int ns = (int)std::min<size_t>(HIST_MAX_SUB, subs.size() - s0);
    if ((rc = launch_hist(c, c->buf[c->cur], subs.data() + s0, ns))) return rc;
    const bool may_skip = !(c->cfg.flags & LSB_FLAG_NO_SKIP) && c->here > 0;
    if (may_skip) {  // a digit that is constant over the shard makes its stable pass the identity
      CU(c, cudaMemcpyAsync(c->host_hist, c->hist, sizeof(unsigned long long) * 256 * ns, cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
    }
    for (int s = 0; s < ns; s++) {
      const SubPass& sp = subs[s0 + s];
      if (may_skip) {
        bool constant = false;
        for (int b = 0; b < 256; b++) constant = constant || c->host_hist[s * 256 + b] == (unsigned long long)c->here;
        if (constant) { c->skipped++; continue; }
      }
      if ((rc = launch_partition(c, c->buf[c->cur], sp.shift, sp.bits, 0, c->one_seg_start, c->one_seg_tiles,
                                 c->scan_out + (size_t)s * 257, c->cur ^ 1, false)))
        return rc;
      c->cur ^= 1;
      (*subpasses)++;
    }
  }
  return LSB_OK;
}

int check_ready(lsb_ctx* c) {
  if (!c) return LSB_ERR_ARG;
  if (c->G > 1 && !c->comm_ready) return fail(c, LSB_ERR_STATE, "world_size > 1: call lsb_comm_init first");
  return LSB_OK;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int lsb_abi_version(void) { return LSB_ABI_VERSION; }

const char* lsb_status_string(int s) {
  switch (s) {
    case LSB_OK: return "ok";
    case LSB_ERR_ARG: return "bad argument";
    case LSB_ERR_CUDA: return "CUDA error";
    case LSB_ERR_NCCL: return "NCCL error";
    case LSB_ERR_STATE: return "wrong state";
    case LSB_ERR_NOMEM: return "out of memory";
    case LSB_ERR_VERIFY: return "verification failed";
  }
  return "unknown";
}

const char* lsb_last_error(const lsb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int lsb_create(lsb_ctx** out, const lsb_config* cfg) {
  if (!out || !cfg) return fail(nullptr, LSB_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->n < 0 || cfg->world_size < 1 || cfg->world_size > LSB_MAX_GPUS || cfg->world_rank < 0 ||
      cfg->world_rank >= cfg->world_size || cfg->radix_bits < 1 || cfg->radix_bits > 16 || cfg->ranks < 0)
    return fail(nullptr, LSB_ERR_ARG, "bad lsb_config (n >= 0, 1 <= world_size <= 8, 1 <= radix_bits <= 16)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, LSB_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LSB_ERR_ARG, "device ordinal out of range");

  lsb_ctx* c = new lsb_ctx();
  c->cfg = *cfg;
  if (c->cfg.ranks == 0) c->cfg.ranks = cfg->world_size;
  if (c->cfg.and_draws < 1) c->cfg.and_draws = 1;
  c->G = cfg->world_size;
  c->my = cfg->world_rank;
  c->n = cfg->n;
  // DistributedArray::create, mpi/mpi_lsbsort.cpp:144-149
  c->per = std::max<int64_t>(div_ceil(c->n, c->G), 1);
  c->here = c->per;
  if (c->per * c->my + c->here > c->n) c->here = c->n - c->per * c->my;
  if (c->here < 0) c->here = 0;
  c->first = c->per * c->my;
  c->per_stream = std::max<int64_t>(div_ceil(c->n, c->cfg.ranks), 1);
  c->npasses = (64 + c->cfg.radix_bits - 1) / c->cfg.radix_bits;

#define CUC(expr)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      int code_ = (e_ == cudaErrorMemoryAllocation) ? LSB_ERR_NOMEM : LSB_ERR_CUDA;      \
      fail(nullptr, code_, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
      lsb_destroy(c);                                                                    \
      return code_;                                                                      \
    }                                                                                    \
  } while (0)

  CUC(cudaSetDevice(cfg->device));
  CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUC(cudaEventCreate(&c->ev_start));
  CUC(cudaEventCreate(&c->ev_stop));
  const size_t shard_bytes = (size_t)std::max<int64_t>(c->per, 1) * sizeof(Elt);
  CUC(cudaMalloc(&c->buf[0], shard_bytes));
  CUC(cudaMalloc(&c->buf[1], shard_bytes));
  c->peer[0][c->my] = c->buf[0];
  c->peer[1][c->my] = c->buf[1];
  {
    const char* v = getenv("LSB_PT_VARIANT");
    c->variant = v ? atoi(v) : 4;  // PartCfgE measured fastest on B200 (profiles/)
    if (c->variant < 0 || c->variant > 6) c->variant = 4;
    if (c->variant == 6 && (cfg->flags & LSB_FLAG_DIRECT_SCATTER)) c->variant = 4;  // segmented input
    const int tiles[7] = {PartCfgA::TILE, PartCfgB::TILE, PartCfgC::TILE, PartCfgD::TILE, PartCfgE::TILE, PartCfgF::TILE,
                          PersistCfgA::TILE};
    c->tile = tiles[c->variant];
  }
  c->lookback_tiles = (size_t)div_ceil(c->per, c->tile) + 256 + 1;
  CUC(cudaMalloc(&c->lookback, c->lookback_tiles * 256 * sizeof(uint64_t)));
  CUC(cudaMemsetAsync(c->lookback, 0, c->lookback_tiles * 256 * sizeof(uint64_t), c->stream));
  CUC(cudaMalloc(&c->tile_counters, 256 * sizeof(uint32_t)));
  CUC(cudaMalloc(&c->hist, sizeof(unsigned long long) * 256 * HIST_MAX_SUB));
  CUC(cudaMalloc(&c->scan_out, sizeof(int64_t) * 257 * HIST_MAX_SUB));
  CUC(cudaMalloc(&c->counts_local, sizeof(unsigned long long) * 65536));
  CUC(cudaMalloc(&c->counts_all, sizeof(unsigned long long) * 65536 * c->G));
  CUC(cudaMalloc(&c->mybase, sizeof(int64_t) * 65536));
  CUC(cudaMalloc(&c->localbase, sizeof(int64_t) * 65536));
  CUC(cudaMalloc(&c->next_hist, sizeof(unsigned long long) * 512 * LSB_MAX_GPUS));
  CUC(cudaMalloc(&c->next_hist_all, sizeof(unsigned long long) * 512 * LSB_MAX_GPUS * LSB_MAX_GPUS));
  CUC(cudaMalloc(&c->seg_tile_start, sizeof(uint32_t) * 257));
  CUC(cudaMalloc(&c->one_seg_start, sizeof(int64_t) * 2));
  CUC(cudaMalloc(&c->one_seg_tiles, sizeof(uint32_t) * 2));
  {
    const bool multi = c->G > 1 || (cfg->flags & LSB_FLAG_TWO_LEVEL);
    c->pipelined = multi && !(cfg->flags & (LSB_FLAG_NO_PIPELINE | LSB_FLAG_DIRECT_SCATTER));
    const char* v = getenv("LSB_VPARTS");
    c->V = v ? atoi(v) : 8;
    if (c->V < 1 || c->V > 8) c->V = 8;
    c->vpart = std::max<int64_t>(div_ceil(c->per, c->V), 1);
  }
  if (c->pipelined) {
    const int V = c->V;
    CUC(cudaMalloc(&c->scratch[0], (size_t)(c->vpart + 64) * sizeof(Elt)));
    CUC(cudaMalloc(&c->scratch[1], (size_t)(c->vpart + 64) * sizeof(Elt)));
    CUC(cudaMalloc(&c->dense_local, sizeof(unsigned) * 65536 * V));
    CUC(cudaMalloc(&c->dense_mine, sizeof(unsigned) * 65536 * V));
    CUC(cudaMalloc(&c->next_dense, sizeof(unsigned) * 65536 * V * c->G));
    CUC(cudaMalloc(&c->c_all, sizeof(unsigned) * 65536 * V * c->G));
    CUC(cudaMalloc(&c->mybase_v, sizeof(int64_t) * 65536 * V));
    CUC(cudaMalloc(&c->localbase_v, sizeof(int64_t) * 65536 * V));
    CUC(cudaMalloc(&c->bases_v, sizeof(int64_t) * 257 * 2 * V));
    CUC(cudaMalloc(&c->seg_start_v, sizeof(int64_t) * 2 * V));
    CUC(cudaMalloc(&c->seg_tiles_v, sizeof(uint32_t) * 2 * V));
    int64_t segs[16];
    uint32_t tls[16];
    for (int q = 0; q < V; q++) {
      const int64_t m = std::max<int64_t>(0, std::min<int64_t>(c->vpart, c->here - (int64_t)q * c->vpart));
      segs[2 * q] = 0;
      segs[2 * q + 1] = m;
      tls[2 * q] = 0;
      tls[2 * q + 1] = (uint32_t)div_ceil(m, c->tile);
    }
    CUC(cudaMemcpyAsync(c->seg_start_v, segs, sizeof(int64_t) * 2 * V, cudaMemcpyHostToDevice, c->stream));
    CUC(cudaMemcpyAsync(c->seg_tiles_v, tls, sizeof(uint32_t) * 2 * V, cudaMemcpyHostToDevice, c->stream));
    CUC(cudaStreamSynchronize(c->stream));
    int lo_prio = 0, hi_prio = 0;
    CUC(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CUC(cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, hi_prio));
    for (int q = 0; q < 8; q++) {
      CUC(cudaEventCreateWithFlags(&c->ev_sorted[q], cudaEventDisableTiming));
      CUC(cudaEventCreateWithFlags(&c->ev_xs[q], cudaEventDisableTiming));
      CUC(cudaEventCreateWithFlags(&c->ev_x[q], cudaEventDisableTiming));
    }
  }
  CUC(cudaMalloc(&c->small, sizeof(unsigned long long) * 64));
  CUC(cudaMalloc(&c->small_all, sizeof(unsigned long long) * 16 * LSB_MAX_GPUS));
  CUC(cudaMemsetAsync(c->small, 0, sizeof(unsigned long long) * 64, c->stream));
  CUC(cudaHostAlloc(&c->host_hist, sizeof(unsigned long long) * 256 * HIST_MAX_SUB, cudaHostAllocDefault));
  CUC(cudaHostAlloc(&c->host_small, sizeof(unsigned long long) * (16 * LSB_MAX_GPUS + 64), cudaHostAllocDefault));
  const int64_t seg[2] = {0, c->here};
  const uint32_t tl[2] = {0, (uint32_t)div_ceil(c->here, c->tile)};
  CUC(cudaMemcpyAsync(c->one_seg_start, seg, sizeof(seg), cudaMemcpyHostToDevice, c->stream));
  CUC(cudaMemcpyAsync(c->one_seg_tiles, tl, sizeof(tl), cudaMemcpyHostToDevice, c->stream));
#define LSB_SET_ATTR(CFG)                                                                                   \
  CUC(cudaFuncSetAttribute(partition_kernel<CFG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM + 8192)); \
  CUC(cudaFuncSetAttribute(partition_kernel<CFG, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));    \
  CUC(cudaFuncSetAttribute(partition_kernel<CFG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM));  \
  CUC(cudaFuncSetAttribute(partition_kernel<CFG, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  LSB_SET_ATTR(PartCfgA)
  LSB_SET_ATTR(PartCfgB)
  LSB_SET_ATTR(PartCfgC)
  LSB_SET_ATTR(PartCfgD)
  LSB_SET_ATTR(PartCfgE)
  LSB_SET_ATTR(PartCfgF)
  CUC(cudaFuncSetAttribute(partition_persistent_kernel<PersistCfgA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistCfgA::SMEM));
  CUC(cudaFuncSetAttribute(partition_persistent_kernel<PersistCfgA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistCfgA::SMEM));
  CUC(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
#undef LSB_SET_ATTR
  CUC(cudaStreamSynchronize(c->stream));
#undef CUC
  *out = c;
  return LSB_OK;
}

void lsb_destroy(lsb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int b = 0; b < 2; b++)
    for (int g = 0; g < LSB_MAX_GPUS; g++)
      if (c->peer_open[b][g]) cudaIpcCloseMemHandle(c->peer[b][g]);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  for (auto ev : c->phase_ev) cudaEventDestroy(ev);
  cudaFree(c->buf[0]);
  cudaFree(c->buf[1]);
  cudaFree(c->lookback);
  cudaFree(c->tile_counters);
  cudaFree(c->hist);
  cudaFree(c->scan_out);
  cudaFree(c->counts_local);
  cudaFree(c->counts_all);
  cudaFree(c->mybase);
  cudaFree(c->localbase);
  cudaFree(c->next_hist);
  cudaFree(c->next_hist_all);
  cudaFree(c->scratch[0]);
  cudaFree(c->scratch[1]);
  cudaFree(c->dense_local);
  cudaFree(c->dense_mine);
  cudaFree(c->next_dense);
  cudaFree(c->c_all);
  cudaFree(c->mybase_v);
  cudaFree(c->localbase_v);
  cudaFree(c->bases_v);
  cudaFree(c->seg_start_v);
  cudaFree(c->seg_tiles_v);
  for (int q = 0; q < 8; q++) {
    if (c->ev_sorted[q]) cudaEventDestroy(c->ev_sorted[q]);
    if (c->ev_xs[q]) cudaEventDestroy(c->ev_xs[q]);
    if (c->ev_x[q]) cudaEventDestroy(c->ev_x[q]);
  }
  if (c->xstream) cudaStreamDestroy(c->xstream);
  cudaFree(c->seg_tile_start);
  cudaFree(c->one_seg_start);
  cudaFree(c->one_seg_tiles);
  cudaFree(c->small);
  cudaFree(c->small_all);
  if (c->host_small) cudaFreeHost(c->host_small);
  if (c->host_hist) cudaFreeHost(c->host_hist);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_stop) cudaEventDestroy(c->ev_stop);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int lsb_comm_unique_id(void* id_out) {
  if (!id_out) return LSB_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == LSB_COMM_ID_BYTES, "ncclUniqueId size");
  if (!g_nccl.load()) return fail(nullptr, LSB_ERR_NCCL, g_nccl.error);
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, LSB_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  memcpy(id_out, &id, sizeof(id));
  return LSB_OK;
}

int lsb_comm_init(lsb_ctx* c, const void* id_bytes) {
  if (!c || !id_bytes) return LSB_ERR_ARG;
  if (c->G == 1) { c->comm_ready = true; return LSB_OK; }
  if (c->comm_ready) return fail(c, LSB_ERR_STATE, "communicator already initialised");
  if (!g_nccl.load()) return fail(c, LSB_ERR_NCCL, g_nccl.error);
  CU(c, cudaSetDevice(c->cfg.device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  NC(c, g_nccl.CommInitRank(&c->comm, c->G, id, c->my));
  // exchange CUDA IPC handles of both shards through the communicator itself
  struct Handles { cudaIpcMemHandle_t h[2]; };
  static_assert(sizeof(Handles) == 128, "ipc handle size");
  Handles mine;
  CU(c, cudaIpcGetMemHandle(&mine.h[0], c->buf[0]));
  CU(c, cudaIpcGetMemHandle(&mine.h[1], c->buf[1]));
  unsigned char* d_h = nullptr;
  CU(c, cudaMalloc(&d_h, sizeof(Handles) * (c->G + 1)));
  CU(c, cudaMemcpyAsync(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  NC(c, g_nccl.AllGather(d_h, d_h + sizeof(Handles), sizeof(Handles), ncclUint8, c->comm, c->stream));
  std::vector<Handles> all(c->G);
  CU(c, cudaMemcpyAsync(all.data(), d_h + sizeof(Handles), sizeof(Handles) * c->G, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaFree(d_h));
  for (int g = 0; g < c->G; g++) {
    if (g == c->my) continue;
    for (int b = 0; b < 2; b++) {
      void* p = nullptr;
      CU(c, cudaIpcOpenMemHandle(&p, all[g].h[b], cudaIpcMemLazyEnablePeerAccess));
      c->peer[b][g] = reinterpret_cast<Elt*>(p);
      c->peer_open[b][g] = true;
    }
  }
  c->comm_ready = true;
  return LSB_OK;
}

int lsb_barrier(lsb_ctx* c) {
  int rc = check_ready(c);
  if (rc) return rc;
  CU(c, cudaSetDevice(c->cfg.device));
  if ((rc = stream_barrier(c))) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

int lsb_shard_info(const lsb_ctx* c, int64_t* per, int64_t* here, int64_t* first_global) {
  if (!c) return LSB_ERR_ARG;
  if (per) *per = c->per;
  if (here) *here = c->here;
  if (first_global) *first_global = c->first;
  return LSB_OK;
}

int lsb_num_passes(const lsb_ctx* c) { return c ? c->npasses : LSB_ERR_ARG; }

int lsb_digit_bits(const lsb_ctx* c, int digit) {
  if (!c || digit < 0 || digit >= c->npasses) return LSB_ERR_ARG;
  return plan_pass(c, digit).bits;
}

int lsb_generate(lsb_ctx* c) {
  if (!c) return LSB_ERR_ARG;
  c->hist_ready_digit = -1;
  CU(c, cudaSetDevice(c->cfg.device));
  c->cur = 0;
  if (c->here > 0) {
    GenArgs a;
    a.dst = c->buf[0];
    a.first_global = c->first;
    a.count = c->here;
    a.per_stream = c->per_stream;
    a.seed_base = c->cfg.seed_base;
    a.key_mask = c->cfg.key_mask;
    a.and_draws = c->cfg.and_draws;
    u128 m, p;
    pcg_jump_coeffs((u128)32 * (u128)c->cfg.and_draws, m, p);
    a.row_mult_hi = (uint64_t)(m >> 64);
    a.row_mult_lo = (uint64_t)m;
    a.row_plus_hi = (uint64_t)(p >> 64);
    a.row_plus_lo = (uint64_t)p;
    const int64_t warps = div_ceil(c->here, 32 * GEN_ROWS);
    const int64_t blocks = div_ceil(warps * 32, GEN_THREADS);
    generate_kernel<<<(unsigned)blocks, GEN_THREADS, 0, c->stream>>>(a);
    CU(c, cudaGetLastError());
  }
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

int lsb_upload(lsb_ctx* c, const lsb_elt* host, int64_t off, int64_t count) {
  if (c) c->hist_ready_digit = -1;
  if (!c || (!host && count) || off < 0 || count < 0 || off + count > c->per) return fail(c, LSB_ERR_ARG, "lsb_upload: range");
  CU(c, cudaSetDevice(c->cfg.device));
  if (count) CU(c, cudaMemcpyAsync(c->buf[c->cur] + off, host, (size_t)count * sizeof(Elt), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

int lsb_download(lsb_ctx* c, lsb_elt* host, int64_t off, int64_t count) {
  if (!c || (!host && count) || off < 0 || count < 0 || off + count > c->per) return fail(c, LSB_ERR_ARG, "lsb_download: range");
  CU(c, cudaSetDevice(c->cfg.device));
  if (count) CU(c, cudaMemcpyAsync(host, c->buf[c->cur] + off, (size_t)count * sizeof(Elt), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

int lsb_device_ptr(lsb_ctx* c, void** ptr) {
  if (!c || !ptr) return LSB_ERR_ARG;
  *ptr = c->buf[c->cur];
  return LSB_OK;
}

int lsb_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return LSB_ERR_ARG;
  cudaError_t e = cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(nullptr, LSB_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return LSB_OK;
}

int lsb_host_free(void* ptr) {
  if (!ptr) return LSB_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? LSB_OK : LSB_ERR_CUDA;
}

int lsb_sort(lsb_ctx* c, lsb_stats* st) {
  int rc = check_ready(c);
  if (rc) return rc;
  CU(c, cudaSetDevice(c->cfg.device));
  if ((rc = begin_call(c))) return rc;
  c->hist_ready_digit = -1;
  c->dense_ready_digit = -1;
  int subpasses = 0;
  if (c->G == 1 && !(c->cfg.flags & LSB_FLAG_TWO_LEVEL)) {
    if ((rc = passes_single(c, 0, c->npasses, &subpasses))) return rc;
  } else {
    for (int d = 0; d < c->npasses; d++)
      if ((rc = pass_global(c, d, &subpasses, true))) return rc;
  }
  return end_call(c, st, c->npasses, subpasses);
}

int lsb_pass(lsb_ctx* c, int digit, lsb_stats* st) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (digit < 0 || digit >= c->npasses) return fail(c, LSB_ERR_ARG, "lsb_pass: digit out of range");
  CU(c, cudaSetDevice(c->cfg.device));
  if ((rc = begin_call(c))) return rc;
  c->hist_ready_digit = -1;
  c->dense_ready_digit = -1;
  int subpasses = 0;
  if (c->G == 1 && !(c->cfg.flags & LSB_FLAG_TWO_LEVEL)) rc = passes_single(c, digit, digit + 1, &subpasses);
  else rc = pass_global(c, digit, &subpasses, false);
  if (rc) return rc;
  return end_call(c, st, 1, subpasses);
}

int lsb_sort_host(lsb_ctx* c, const lsb_elt* host_in, lsb_elt* host_out, int64_t count, lsb_stats* st) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (count != c->here || (count && (!host_in || !host_out))) return fail(c, LSB_ERR_ARG, "lsb_sort_host: count must equal this shard's size");
  CU(c, cudaSetDevice(c->cfg.device));
  c->cur = 0;
  if (count) CU(c, cudaMemcpyAsync(c->buf[0], host_in, (size_t)count * sizeof(Elt), cudaMemcpyHostToDevice, c->stream));
  if ((rc = lsb_sort(c, st))) return rc;
  if (count) CU(c, cudaMemcpyAsync(host_out, c->buf[c->cur], (size_t)count * sizeof(Elt), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

static int dense_counts(lsb_ctx* c, int digit, int* nb_out) {
  const PassPlan p = plan_pass(c, digit);
  const int nb = 1 << p.bits;
  CU(c, cudaMemsetAsync(c->counts_local, 0, sizeof(unsigned long long) * nb, c->stream));
  if (c->here > 0) {
    dense_count_kernel<<<148 * 8, 256, 0, c->stream>>>(c->buf[c->cur], c->here, p.shift, (uint32_t)(nb - 1), c->counts_local);
    CU(c, cudaGetLastError());
  }
  *nb_out = nb;
  return LSB_OK;
}

int lsb_histogram(lsb_ctx* c, int digit, int64_t* host_counts) {
  if (!c || !host_counts || digit < 0 || digit >= c->npasses) return fail(c, LSB_ERR_ARG, "lsb_histogram: argument");
  CU(c, cudaSetDevice(c->cfg.device));
  int nb = 0, rc = dense_counts(c, digit, &nb);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(host_counts, c->counts_local, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

int lsb_starts(lsb_ctx* c, int digit, int64_t* host_starts) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (!host_starts || digit < 0 || digit >= c->npasses) return fail(c, LSB_ERR_ARG, "lsb_starts: argument");
  CU(c, cudaSetDevice(c->cfg.device));
  int nb = 0;
  if ((rc = dense_counts(c, digit, &nb))) return rc;
  if ((rc = global_offsets(c, nb))) return rc;
  CU(c, cudaMemcpyAsync(host_starts, c->mybase, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return LSB_OK;
}

static int local_verify(lsb_ctx* c) {
  // small[40..44] = checksum[4], violations
  CU(c, cudaMemsetAsync(c->small + 40, 0, 8 * sizeof(unsigned long long), c->stream));
  if (c->here > 0) {
    int grid = (int)std::min<int64_t>(148 * 8, div_ceil(c->here, 256));
    verify_kernel<<<grid, 256, 0, c->stream>>>(c->buf[c->cur], c->here, c->small + 40);
    CU(c, cudaGetLastError());
  }
  return LSB_OK;
}

int lsb_checksum(lsb_ctx* c, uint64_t out[4]) {
  if (!c || !out) return LSB_ERR_ARG;
  CU(c, cudaSetDevice(c->cfg.device));
  int rc = local_verify(c);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->host_small, c->small + 40, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 4; i++) out[i] = c->host_small[i];
  return LSB_OK;
}

int lsb_verify_device(lsb_ctx* c, lsb_verify* out) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (!out) return LSB_ERR_ARG;
  CU(c, cudaSetDevice(c->cfg.device));
  if ((rc = local_verify(c))) return rc;
  // record = {checksum[4], violations, here, first.key, first.val, last.key, last.val}
  unsigned long long* rec = c->small + 48;  // 10 words used, 16 reserved
  CU(c, cudaMemcpyAsync(rec, c->small + 40, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
  const unsigned long long here = (unsigned long long)c->here;
  CU(c, cudaMemcpyAsync(rec + 5, &here, sizeof(here), cudaMemcpyHostToDevice, c->stream));
  if (c->here > 0) {
    CU(c, cudaMemcpyAsync(rec + 6, c->buf[c->cur], sizeof(Elt), cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaMemcpyAsync(rec + 8, c->buf[c->cur] + (c->here - 1), sizeof(Elt), cudaMemcpyDeviceToDevice, c->stream));
  }
  const unsigned long long* all = rec;
  if (c->G > 1) {
    NC(c, g_nccl.AllGather(rec, c->small_all, 16, ncclUint64, c->comm, c->stream));
    all = c->small_all;
  }
  CU(c, cudaMemcpyAsync(c->host_small, all, sizeof(unsigned long long) * 16 * c->G, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  memset(out, 0, sizeof(*out));
  bool have_prev = false;
  uint64_t pk = 0, pv = 0;
  for (int g = 0; g < c->G; g++) {
    const unsigned long long* r = c->host_small + 16 * g;
    out->checksum[0] += r[0];
    out->checksum[1] ^= r[1];
    out->checksum[2] ^= r[2];
    out->checksum[3] += r[3];
    out->order_violations += (int64_t)r[4];
    out->elements += (int64_t)r[5];
    if (r[5] == 0) continue;
    if (have_prev && (pk > r[6] || (pk == r[6] && pv >= r[7]))) out->order_violations++;
    pk = r[8];
    pv = r[9];
    have_prev = true;
  }
  if (out->order_violations) return fail(c, LSB_ERR_VERIFY, "shards are not strictly increasing in (key,val)");
  return LSB_OK;
}

}  // extern "C"
