/*
 * lsb_oracle.c -- CPU restatement of the reference's distributed LSD radix sort.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object, and only as the checker / baseline.
 * The product path (distributed-lsb_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 *   - against the PCG64 known-answer values and sort golden vectors recorded in
 *     SURVEY.md section 8(c) (generated from the pcg-cpp header that pyarrow
 *     vendors and cross-checked against the unmodified reference's --print),
 *   - against the unmodified /root/reference/mpi/mpi_lsbsort.cpp itself, built
 *     by oracle/build_ref.sh into oracle/_ref/ over a ranks-as-threads mpi.h
 *     shim (fixtures committed in tests/golden/ref_print_*.txt).
 *
 * What is restated (all citations relative to /root/reference/):
 *   element layout ............ mpi/mpi_lsbsort.cpp:29-32   (16 B: u64 key, u64 val)
 *   block distribution ........ mpi/mpi_lsbsort.cpp:138-161 (per = ceil(n/R), here clamped >= 0)
 *   digit extraction .......... mpi/mpi_lsbsort.cpp:203-205
 *   per-rank count ............ mpi/mpi_lsbsort.cpp:226-229
 *   digit-major/rank-minor scan mpi/mpi_lsbsort.cpp:350,385-414
 *   stable placement .......... mpi/mpi_lsbsort.cpp:241-246,546-575
 *   pass loop ................. mpi/mpi_lsbsort.cpp:580-585
 *   generator ................. mpi/mpi_lsbsort.cpp:650-656
 *   result definition ......... mpi/mpi_lsbsort.cpp:722-736 (std::stable_sort by key)
 *
 * Third-party arithmetic absent from /root/reference: pcg64 from imneme/pcg-cpp
 * (un-pinned `git clone`, mpi/getpcg.sh:3).  pcg64 = setseq_xsl_rr_128_64 with the
 * default stream: 128-bit LCG, advance-then-output, XSL-RR output function.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct { uint64_t key, val; } lsbo_elt;

/* ------------------------------------------------------------------ PCG64 */

#define PCG_MULT_HI 2549297995355413924ULL
#define PCG_MULT_LO 4865540595714422341ULL
#define PCG_INC_HI  6364136223846793005ULL
#define PCG_INC_LO  1442695040888963407ULL

static inline u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }

/* pcg-cpp: engine(seed): state = bump(seed + increment); bump(s) = s*MULT + INC */
u128 lsbo_pcg64_seed(uint64_t seed) {
  const u128 mult = mk128(PCG_MULT_HI, PCG_MULT_LO), inc = mk128(PCG_INC_HI, PCG_INC_LO);
  return ((u128)seed + inc) * mult + inc;
}

/* advance first, then XSL-RR on the new state */
static inline uint64_t pcg64_next(u128 *st) {
  const u128 mult = mk128(PCG_MULT_HI, PCG_MULT_LO), inc = mk128(PCG_INC_HI, PCG_INC_LO);
  u128 s = *st * mult + inc;
  *st = s;
  uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
  unsigned rot = (unsigned)(s >> 122);
  uint64_t x = hi ^ lo;
  return (x >> rot) | (x << ((-rot) & 63));
}

/* Brown's O(log delta) jump-ahead (pcg-cpp `advance`) */
u128 lsbo_pcg64_advance(u128 st, u128 delta) {
  u128 cur_mult = mk128(PCG_MULT_HI, PCG_MULT_LO), cur_plus = mk128(PCG_INC_HI, PCG_INC_LO);
  u128 acc_mult = 1, acc_plus = 0;
  while (delta > 0) {
    if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * st + acc_plus;
}

/* first `count` outputs of pcg64(seed), after skipping `skip` outputs */
void lsbo_pcg64_stream(uint64_t seed, uint64_t skip, uint64_t *out, int64_t count) {
  u128 st = lsbo_pcg64_seed(seed);
  if (skip) st = lsbo_pcg64_advance(st, skip);
  for (int64_t i = 0; i < count; i++) out[i] = pcg64_next(&st);
}

/* --------------------------------------------------------- distribution */

static inline int64_t div_ceil(int64_t x, int64_t y) { return (x + y - 1) / y; }

int64_t lsbo_per_rank(int64_t n, int ranks) { return div_ceil(n, ranks); }

int64_t lsbo_here(int64_t n, int ranks, int r) {
  int64_t per = div_ceil(n, ranks), here = per;
  if (per * r + here > n) here = n - per * r;
  if (here < 0) here = 0;
  return here;
}

/* ------------------------------------------------------------ generator */
/*
 * Fills all ranks*per slots (the reference fills padding slots on the last
 * rank too, mpi/mpi_lsbsort.cpp:650-656); only the first n take part in the sort.
 *   key_mask : applied to every key (config 4(i) skew: 0xFFFFFF); ~0 = reference
 *   and_draws: k >= 1; key = AND of k successive draws (config 4(ii) entropy
 *              reduction); 1 = reference.  Element j of a stream uses draws
 *              [k*j, k*j+k).
 */
void lsbo_generate(lsbo_elt *out, int64_t n, int ranks, uint64_t seed_base,
                   uint64_t key_mask, int and_draws) {
  int64_t per = div_ceil(n, ranks);
  if (and_draws < 1) and_draws = 1;
  for (int r = 0; r < ranks; r++) {
    u128 st = lsbo_pcg64_seed(seed_base + (uint64_t)r);
    for (int64_t i = 0; i < per; i++) {
      uint64_t k = pcg64_next(&st);
      for (int d = 1; d < and_draws; d++) k &= pcg64_next(&st);
      out[r * per + i].key = k & key_mask;
      out[r * per + i].val = (uint64_t)(r * per + i);
    }
  }
}

/* --------------------------------------------------------------- digits */

int lsbo_num_passes(int radix_bits) { return (64 + radix_bits - 1) / radix_bits; }

static inline uint32_t digit_of(uint64_t key, int radix_bits, int pass) {
  int shift = radix_bits * pass;
  uint64_t mask = (radix_bits >= 64) ? ~0ULL : ((1ULL << radix_bits) - 1);
  return (uint32_t)((key >> shift) & mask);
}

/* ------------------------------------------------------------- one pass */
/*
 * One globalShuffle over the *global* array (ranks simulated): src -> dst.
 *   counts    [ranks][nb]   per-rank digit counts          (:226-229)
 *   starts    [nb][ranks]   exclusive scan, digit-major, rank-minor (:350,:407-413)
 *   sendcounts[ranks][ranks] elements src rank -> dst rank (:553-554)
 * Any of the three output tables may be NULL.
 */
int lsbo_pass(const lsbo_elt *src, lsbo_elt *dst, int64_t n, int ranks,
              int radix_bits, int pass, int64_t *counts_out, int64_t *starts_out,
              int64_t *sendcounts_out) {
  const int64_t nb = (int64_t)1 << radix_bits;
  const int64_t per = div_ceil(n, ranks);
  int64_t *counts = (int64_t *)calloc((size_t)(nb * ranks), sizeof(int64_t));
  int64_t *next = (int64_t *)malloc((size_t)(nb * ranks) * sizeof(int64_t));
  if (!counts || !next) { free(counts); free(next); return -1; }

  for (int r = 0; r < ranks; r++) {
    int64_t here = lsbo_here(n, ranks, r);
    const lsbo_elt *a = src + r * per;
    int64_t *c = counts + (int64_t)r * nb;
    for (int64_t i = 0; i < here; i++) c[digit_of(a[i].key, radix_bits, pass)]++;
  }
  int64_t sum = 0;
  for (int64_t d = 0; d < nb; d++)
    for (int r = 0; r < ranks; r++) {
      next[d * ranks + r] = sum;
      sum += counts[(int64_t)r * nb + d];
    }
  if (counts_out) memcpy(counts_out, counts, (size_t)(nb * ranks) * sizeof(int64_t));
  if (starts_out) memcpy(starts_out, next, (size_t)(nb * ranks) * sizeof(int64_t));
  if (sendcounts_out) memset(sendcounts_out, 0, sizeof(int64_t) * ranks * ranks);

  for (int r = 0; r < ranks; r++) {
    int64_t here = lsbo_here(n, ranks, r);
    const lsbo_elt *a = src + r * per;
    for (int64_t i = 0; i < here; i++) {
      uint32_t d = digit_of(a[i].key, radix_bits, pass);
      int64_t g = next[(int64_t)d * ranks + r]++;
      dst[g] = a[i];
      if (sendcounts_out) sendcounts_out[r * ranks + (int)(g / per)]++;
    }
  }
  free(counts);
  free(next);
  return 0;
}

/* mySort: all passes, LSB first; result back in `a` (scratch `b`, both >= n). */
int lsbo_sort(lsbo_elt *a, lsbo_elt *b, int64_t n, int ranks, int radix_bits) {
  int np = lsbo_num_passes(radix_bits);
  lsbo_elt *s = a, *d = b;
  for (int p = 0; p < np; p++) {
    if (lsbo_pass(s, d, n, ranks, radix_bits, p, 0, 0, 0)) return -1;
    lsbo_elt *t = s; s = d; d = t;
  }
  if (s != a) memcpy(a, s, (size_t)n * sizeof(lsbo_elt));
  return 0;
}

/* ------------------------------------- definitional result: stable sort */
/* bottom-up merge sort by key only == std::stable_sort(key <), :722-726 */
int lsbo_stable_sort(lsbo_elt *a, int64_t n) {
  if (n < 2) return 0;
  lsbo_elt *tmp = (lsbo_elt *)malloc((size_t)n * sizeof(lsbo_elt));
  if (!tmp) return -1;
  lsbo_elt *s = a, *d = tmp;
  for (int64_t w = 1; w < n; w *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * w) {
      int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int64_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi) d[k++] = (s[j].key < s[i].key) ? s[j++] : s[i++];
      while (i < mid) d[k++] = s[i++];
      while (j < hi) d[k++] = s[j++];
    }
    lsbo_elt *t = s; s = d; d = t;
  }
  if (s != a) memcpy(a, s, (size_t)n * sizeof(lsbo_elt));
  free(tmp);
  return 0;
}

/* ------------------------------------------------------------ checksums */

uint64_t lsbo_fnv1a64(const void *data, int64_t nbytes) {
  const unsigned char *p = (const unsigned char *)data;
  uint64_t h = 1469598103934665603ULL;
  for (int64_t i = 0; i < nbytes; i++) { h ^= p[i]; h *= 1099511628211ULL; }
  return h;
}

/* order-independent multiset hash; must match lsb_checksum() in include/lsbsort.h:
 *   out[0] = sum of mix(key,val), out[1] = xor of mix(key,val),
 *   out[2] = xor of keys,         out[3] = sum of vals  (all mod 2^64)        */
static inline uint64_t mix64(uint64_t k, uint64_t v) {
  uint64_t x = k * 0x9E3779B97F4A7C15ULL + (v ^ 0xD6E8FEB86659FD93ULL);
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL;
  x ^= x >> 29; x *= 0x9E3779B97F4A7C15ULL;
  x ^= x >> 32;
  return x;
}

void lsbo_checksum(const lsbo_elt *a, int64_t n, uint64_t out[4]) {
  uint64_t s = 0, x = 0, xk = 0, sv = 0;
  for (int64_t i = 0; i < n; i++) {
    uint64_t m = mix64(a[i].key, a[i].val);
    s += m; x ^= m; xk ^= a[i].key; sv += a[i].val;
  }
  out[0] = s; out[1] = x; out[2] = xk; out[3] = sv;
}

/* number of i in [1,n) with (key,val)[i-1] >= (key,val)[i]; 0 == strictly increasing */
int64_t lsbo_order_violations(const lsbo_elt *a, int64_t n) {
  int64_t bad = 0;
  for (int64_t i = 1; i < n; i++) {
    if (a[i - 1].key > a[i].key || (a[i - 1].key == a[i].key && a[i - 1].val >= a[i].val)) bad++;
  }
  return bad;
}
