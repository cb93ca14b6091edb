/*
 * lsbsort.h -- C ABI of the B200-native distributed LSD radix sort.
 *
 * This is the drop-in boundary for the one hot path of ronawho/distributed-lsb:
 *   mySort(A, B)              mpi/mpi_lsbsort.cpp:580-585
 *   globalShuffle(A, B, d)    mpi/mpi_lsbsort.cpp:481-577
 * plus the two things main() does around it that define the workload and the answer:
 *   the pcg64(rank) generator mpi/mpi_lsbsort.cpp:650-656
 *   the verifier              mpi/mpi_lsbsort.cpp:710-739
 * The reference has no plugin/FFI interface (everything is one translation unit), so
 * the seam is where main() calls mySort (mpi/mpi_lsbsort.cpp:691).  INTEGRATION.md shows
 * the few lines a maintainer adds to mpi_lsbsort.cpp to call through this header.
 *
 * Process model: one process (= one MPI rank of the reference) per GPU.  The
 * reference's DistributedArray<SortElement> (mpi/mpi_lsbsort.cpp:90-161) becomes a
 * device-resident pair of shards A/B owned by the context: shard g holds the global
 * indices [g*per, g*per+here), per = ceil(n/G), here clamped exactly as
 * DistributedArray::create does (:144-149).
 *
 * Element ABI: 16 bytes, { uint64 key; uint64 val; }, key first, little endian --
 * identical to SortElement (mpi/mpi_lsbsort.cpp:29-32).
 *
 * Errors: every call returns LSB_OK (0) or a negative lsb_status; the text of the
 * CUDA/NCCL failure is kept in the context (lsb_last_error).  The reference aborts on
 * error instead ("errors returned by MPI calls do not need to be handled", :17-19).
 * There is NO CPU fallback: without a CUDA device lsb_create fails with LSB_ERR_CUDA.
 *
 * Threading: a context is not thread-safe, but contexts share no state: different contexts (also on one GPU)
 * may be used from different threads at the same time.  Calls are synchronous on return.
 */
#ifndef LSBSORT_H
#define LSBSORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSB_ABI_VERSION 1
#define LSB_MAX_GPUS 8
#define LSB_MAX_SUBPASSES 32
#define LSB_MAX_PARTS 16   /* parts per shard of the multi-GPU pass (virtual ranks) */
#define LSB_COMM_ID_BYTES 128

typedef enum {
  LSB_OK = 0,
  LSB_ERR_ARG = -1,     /* bad argument / configuration                         */
  LSB_ERR_CUDA = -2,    /* CUDA runtime failure (text in lsb_last_error)        */
  LSB_ERR_NCCL = -3,    /* NCCL failure                                         */
  LSB_ERR_STATE = -4,   /* call made in the wrong state (e.g. comm not set up)  */
  LSB_ERR_NOMEM = -5,   /* device or host allocation failed                     */
  LSB_ERR_VERIFY = -6   /* lsb_verify_device found the shard(s) unsorted        */
} lsb_status;

/* one 16-byte record == SortElement, mpi/mpi_lsbsort.cpp:29-32 */
typedef struct lsb_elt {
  uint64_t key;
  uint64_t val;
} lsb_elt;

typedef struct lsb_ctx lsb_ctx;

/*
 * Knobs of the reference driver (mpi/mpi_lsbsort.cpp:594-611) plus what `mpirun -n R`
 * and `#define RADIX` (:21) fix at launch/compile time there.
 */
typedef struct lsb_config {
  int64_t n;            /* --n: total number of elements over all shards (:594-598)        */
  int32_t ranks;        /* R of `mpirun -n R`: number of pcg64 streams of the generator;
                           stream r = pcg64(seed_base + r) fills global slots
                           [r*ceil(n/R), (r+1)*ceil(n/R)) (:144-149,:650-656). 0 => world_size */
  int32_t world_size;   /* G: number of GPUs == number of processes                        */
  int32_t world_rank;   /* g: which shard this process owns                                */
  int32_t device;       /* CUDA device ordinal for this process                            */
  int32_t radix_bits;   /* RADIX (:21): 1..16; the reference uses 16 (4 passes)            */
  int32_t and_draws;    /* 1 = reference; k>1: key = AND of k successive draws (skew)      */
  uint64_t seed_base;   /* 0 = reference (pcg64(myRank), :650)                             */
  uint64_t key_mask;    /* ~0 = reference; 0xFFFFFF = "only low 24 bits random" skew       */
  uint32_t flags;       /* LSB_FLAG_*                                                      */
  uint32_t reserved;
} lsb_config;

#define LSB_FLAG_NONE 0u
#define LSB_FLAG_PHASE_EVENTS 1u /* record a CUDA event after every kernel of lsb_sort      */
#define LSB_FLAG_TWO_LEVEL 2u    /* run the multi-GPU pass shape (virtual ranks: per-part counts,
                                    digit-major / rank-minor scan, part sort + exchange kernel)
                                    even when world_size == 1                                  */
#define LSB_FLAG_ONE_PASS 4u     /* sort a 9..16-bit digit with the one-pass kernel: every element moves through
                                    HBM once per reference pass (L2-resident supertiles) instead of twice (two
                                    stable 8-bit counting-sort steps, the default); same permutation          */
#define LSB_FLAG_NO_SKIP 8u      /* do not skip passes whose digit is constant over the shard
                                    (a stable pass on a constant digit is the identity; the
                                    reference always runs it; chpl passes nBits for the same
                                    purpose, chpl/arkouda-radix-sort.chpl:70,78)              */

/* what lsb_sort / lsb_pass measured, device time from CUDA events on the sort stream */
typedef struct lsb_stats {
  double device_ms;       /* whole call: first kernel start -> last kernel end             */
  int32_t passes;         /* reference passes run (N_DIGITS, :22)                          */
  int32_t subpasses;      /* scatter-kernel launches (2 per 16-bit pass and part; 1 with ONE_PASS) */
  int32_t skipped;        /* steps (ONE_PASS: passes) skipped because their digit was constant */
  int32_t reserved0;
  int64_t elements;       /* elements of THIS shard that took part                         */
  double hist_ms;         /* count kernels (needs LSB_FLAG_PHASE_EVENTS, else 0)           */
  double scan_ms;         /* scans + collectives on counts                                 */
  double partition_ms;    /* partition (scatter) kernels                                   */
  double exchange_ms;     /* G > 1: exchange kernels (NVLink stores)                       */
  double subpass_ms[LSB_MAX_SUBPASSES]; /* per partition-kernel launch                     */
  int64_t sent[LSB_MAX_GPUS]; /* last pass: elements this shard sent to each GPU (:553-554) */
  int64_t partition_launches;
  int64_t partition_elements; /* elements moved by all partition launches of the call          */
  int64_t kernel_launches; /* all kernels launched by the call                             */
} lsb_stats;

typedef struct lsb_verify {
  int64_t order_violations;  /* i with (key,val)[i-1] >= (key,val)[i], within and across shards */
  int64_t elements;          /* global element count seen                                   */
  uint64_t checksum[4];      /* global multiset hash, see lsb_checksum                      */
} lsb_verify;

/* ---- lifetime ------------------------------------------------------------------ */

int lsb_abi_version(void);

/* DistributedArray<SortElement>::create for A and B (:138-161,:638-639): allocates both
 * device shards and every scratch table once.  The reference re-allocates its count and
 * send/recv buffers inside every pass (:489-490,:511-516,:538-543); here nothing is
 * allocated after lsb_create. */
int lsb_create(lsb_ctx** out, const lsb_config* cfg);
void lsb_destroy(lsb_ctx* ctx);
const char* lsb_last_error(const lsb_ctx* ctx);
const char* lsb_status_string(int status);

/* Experiment knob (profiling sweeps only; the defaults are what bench.py measures): sets a
 * process-wide tunable read by the NEXT lsb_create.  Keys: one-pass kernel -- "op_cfg" (tile
 * shape: 0 = 512 threads x 11 elements, 1 = 256 x 11), "op_t1" (tiles per supertile, 1..256),
 * "op_nx" (supertile scratch slots), "op_lead" (supertiles between a tile and its use),
 * "op_hints" (L2 eviction hints, bit mask), "op_persist" (MiB of persisting L2),
 * "op_ctas_mgpu" (one-pass CTAs per SM while an exchange kernel shares the GPU), "timeout_ms"
 * (watchdog); multi-GPU pass -- "vparts" (parts per shard), "vramp" (size ratio of neighbouring
 * parts x 100), "ex_ctas", "ex_threads", "ex_u" (exchange kernel shape); 8-bit scatter kernel --
 * "pt_direct" (1 = tile index from blockIdx and the tile load issued first, 0 = tiles handed out
 * by a ticket counter), "pt_chunks" (log2 of the bulk copies a tile arrives in, 0..4), "pt_variant" (1 = the CTA of tile t
 * also bulk-prefetches tile t + pt_pf_tiles into L2 [default], 0 = no prefetch), "pt_pf_tiles"
 * (prefetch distance in tiles, 0 = half the SM count).
 * Unknown key / bad value: LSB_ERR_ARG. */
int lsb_tune(const char* key, int value);

/* ---- multi-GPU wiring (MPI_Init / MPI_COMM_WORLD, :588,:613-616) ----------------- */

/* rank 0 calls lsb_comm_unique_id and hands the 128 bytes to every process (any
 * out-of-band channel); then every process calls lsb_comm_init.  It creates the NCCL
 * communicator used for the per-pass count all-gather (replaces the count transpose
 * MPI_Alltoallv + MPI_Exscan, :327-414) and maps every peer's A/B shards over
 * NVLink (CUDA IPC) so the scatter kernel can store straight into the destination
 * shard (replaces pack + MPI_Alltoallv + unpack, :530-576).  Not needed when G == 1. */
int lsb_comm_unique_id(void* id_out /* LSB_COMM_ID_BYTES */);
int lsb_comm_init(lsb_ctx* ctx, const void* id /* LSB_COMM_ID_BYTES */);

/* MPI_Barrier(MPI_COMM_WORLD) (:688,:693): every GPU has finished everything queued on
 * its sort stream.  Collective; a plain device synchronise when G == 1. */
int lsb_barrier(lsb_ctx* ctx);

/* ---- data ---------------------------------------------------------------------- */

/* per = numElementsPerRank (:103), here = numElementsHere (:104), first = r*per */
int lsb_shard_info(const lsb_ctx* ctx, int64_t* per, int64_t* here, int64_t* first_global);

/* the generator loop (:650-656) for this shard, on the device, bit-exact with pcg64 */
int lsb_generate(lsb_ctx* ctx);

/* caller-supplied data instead of lsb_generate: copy `count` elements from host memory
 * into this shard starting at local index local_off (A.localPart()[local_off..]) */
int lsb_upload(lsb_ctx* ctx, const lsb_elt* host, int64_t local_off, int64_t count);
/* copy elements of the current (sorted or not) shard back to host memory */
int lsb_download(lsb_ctx* ctx, lsb_elt* host, int64_t local_off, int64_t count);

/* device pointer of the shard that currently holds the data (after lsb_sort: the result) */
int lsb_device_ptr(lsb_ctx* ctx, void** ptr);

/* pinned host memory helpers for the host-buffer path */
int lsb_host_alloc(void** ptr, int64_t bytes);
int lsb_host_free(void* ptr);

/* ---- the hot path ---------------------------------------------------------------- */

/* mySort (:580-585): all ceil(64/radix_bits) passes, least significant digit first;
 * on return the shard holds its slice of the globally, stably sorted array.
 * Collective: every process must call it.  stats may be NULL. */
int lsb_sort(lsb_ctx* ctx, lsb_stats* stats);

/* globalShuffle(A, B, digit) (:481-577): one stable pass on digit `digit`. Collective. */
int lsb_pass(lsb_ctx* ctx, int digit, lsb_stats* stats);

/* host-buffer form of mySort for one process: upload `count` elements (must equal
 * `here`), sort, download into host_out.  This is the end-to-end entry point a caller
 * with data in host memory uses; copies are inside the call. Collective. */
int lsb_sort_host(lsb_ctx* ctx, const lsb_elt* host_in, lsb_elt* host_out, int64_t count,
                  lsb_stats* stats);

/* ---- test hooks into the pass (same tables the reference computes) ---------------- */

/* counts[d], d < 2^bits(digit): the localShuffle count loop (:226-229) on this shard */
int lsb_histogram(lsb_ctx* ctx, int digit, int64_t* host_counts);
/* starts[d]: global output index of the first of MY elements whose digit is d, i.e.
 * GlobalStarts[d*R + myRank] after copyCountsToGlobalCounts + exclusiveScan +
 * copyStartsFromGlobalStarts (:519-525; order defined at :350). Collective. */
int lsb_starts(lsb_ctx* ctx, int digit, int64_t* host_starts);

/* number of passes N_DIGITS = ceil(64/radix_bits) (:22; ceil as chpl:78-79) and the
 * width in bits of digit `digit` (radix_bits, or the remainder for the last one) */
int lsb_num_passes(const lsb_ctx* ctx);
int lsb_digit_bits(const lsb_ctx* ctx, int digit);

/* ---- verification (:710-739, without gathering to rank 0) -------------------------- */

/* order-independent multiset hash of this shard (no communication):
 * out[0] = sum mix(key,val), out[1] = xor mix(key,val), out[2] = xor key, out[3] = sum val */
int lsb_checksum(lsb_ctx* ctx, uint64_t out[4]);

/* Collective.  Checks that (key,val) is strictly increasing inside every shard and
 * across shard boundaries, and returns the global multiset hash.  Because val is the
 * unique original index (:654), "strictly increasing in (key,val) and same multiset as
 * the input" is equivalent to "equals std::stable_sort by key of the input" (:722-726).
 * Returns LSB_ERR_VERIFY if order_violations != 0. */
int lsb_verify_device(lsb_ctx* ctx, lsb_verify* out);

#ifdef __cplusplus
}
#endif
#endif /* LSBSORT_H */
