// cco_kernels.cuh -- hand-written sm_100a kernels of the CCO train hot path.
//
// Reference semantics (what each stage replaces) -- Apache Mahout 0.13.0 SimilarityAnalysis,
// called from /root/reference/src/main/scala/URAlgorithm.scala:323-329,343-346:
//   input validation           -> k_check_rows; canonicalisation slow path k_expand_keys / k_unique_* / k_rowptr_from_keys
//   H2 sampleDownAndBinarize   -> k_downsample_count / k_downsample_write (row ranges: whole matrix or a rank's user block)
//   H3 numNonZeroElementsPerColumn -> k_col_histogram (raw, before the allreduce), k_downsample_count or
//                                 k_col_histogram_u32 (post-sample)
//   `drmA.t`                   -> k_transpose_scatter
//   scheduling                 -> k_row_work, k_bin_bounds, k_partition_rows; per-column LLR constants k_col_terms
//   H4 A'^T B' counts, H5 LLR, H6 top-k -> k_rows<GROUP, DENSE> (one fused kernel, nothing materialised)
//   result assembly            -> k_len_to_i64, k_compact_rows; small device->host results k_mail_bytes
//
// Everything here is integer/byte work plus scalar fp64; no tensor cores.  The accumulator, the
// candidate buffer, the select histogram and the LLR tables of a row live in shared memory; B' and the per-column
// terms are gathered from L2/HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cco {

// ------------------------------------------------------------------------------------------------
// device views
// ------------------------------------------------------------------------------------------------
struct RowArgs {
  // A'^T: users of every primary item
  const uint32_t *at_ptr;
  const int32_t *at_users;
  // B' (CSR over users)
  const uint32_t *b_ptr;
  const int32_t *b_col;
  const int32_t *marg_a;  // colA, downsampled, per primary item
  const int32_t *marg_b;  // colB, downsampled, per column of B'
  const uint2 *ext;       // experiment (tools/experiments/cco_rows2.cuh): (start, len) of B'[u] per (item, user) pair; unused
  int32_t max_marg_b;     // largest colB (bounds every co-occurrence count together with rowA)
  // schedule: items sorted by estimated work, descending; bin b = rows_sorted[bin_bounds[b], bin_bounds[b+1])
  const int32_t *rows_sorted;
  const uint32_t *row_work;  // by item: w_a = sum_{u in a} degB'(u), saturated at 2^32-1
  const int32_t *bin_bounds;
  int32_t bin;
  int32_t n_cols_b;
  long long n_users;  // N
  int32_t self;       // A'^T A': skip the diagonal
  int32_t top_k;
  int32_t has_min_llr;
  double min_llr;
  uint32_t flags;
  int32_t count_bits;  // packed hash word = (key << count_bits) | count
  int32_t slots;       // hash/dense table words in shared memory
  int32_t cap;         // max distinct keys per pass for hashed rows (load-factor bound)
  int32_t tsize_x16;   // table words per expected distinct key, in sixteenths (32 = load factor 1/2)
  int32_t cbuf;        // candidate buffer entries per group (power of two, >= top_k + GROUP)
  int32_t caux;        // scratch entries for the out-of-place compaction of the radix select (0 for warps)
  int32_t keep_max;    // M: a prune keeps between top_k and max(M, top_k) candidates
  int32_t final_max;   // the final sort runs on at most this many candidates (next_pow2(top_k))
  int32_t group_smem_bytes;  // shared memory of one group (multiple of 16)
  const struct ColTerm *col_terms;  // per column of B': {columnEntropy, xLogX(colB - 1), colB}
  // outputs, strided
  int32_t out_stride;
  int32_t *out_col;
  double *out_llr;
  int32_t *out_cnt;
  int32_t *out_len;
  unsigned long long *stat_distinct;
  unsigned long long *stat_evaluated;  // cells whose fp64 LLR was actually evaluated (after the dominance filter)
  int *err_flag;     // set to 1 if a hash table overflowed (result invalid)
  int32_t emit_all;  // debug: write every non-zero cell (col,count), no LLR/top-k
};

constexpr uint32_t kEmpty = 0xffffffffu;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}
// sampler of include/cco_b200.h "Sampler" (bit-identical to oracle/cco_oracle.c orc_hash64/orc_u01)
__device__ __forceinline__ double sample_u01(int32_t seed, uint32_t u, uint32_t j) {
  uint64_t x = mix64(((uint64_t)(uint32_t)seed << 32) | (uint64_t)u);
  uint64_t h = mix64(x + (uint64_t)j * 0x9e3779b97f4a7c15ULL);
  return __dmul_rn((double)(h >> 11), 0x1.0p-53);
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// ------------------------------------------------------------------------------------------------
// LogLikelihood (Mahout mahout-math LogLikelihood.java).  __dmul_rn/__dsub_rn keep
// nvcc from contracting x*log(x) - ... into FMAs so the evaluation order matches the JVM's.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double xlogx(long long x) {
  return x == 0 ? 0.0 : __dmul_rn((double)x, log((double)x));
}
// counts inside the row kernel are < 2^31 (n_rows < 2^31 is validated): the 32-bit conversion is exact and cheaper
__device__ __forceinline__ double xlogx_u32(uint32_t x) {
  return x == 0 ? 0.0 : __dmul_rn((double)x, log((double)x));
}
__device__ __forceinline__ double entropy2(long long a, long long b, bool varargs) {
  if (varargs) return __dsub_rn(xlogx(a + b), __dadd_rn(__dadd_rn(0.0, xlogx(a)), xlogx(b)));
  return __dsub_rn(__dsub_rn(xlogx(a + b), xlogx(a)), xlogx(b));
}
__device__ __forceinline__ double entropy4(long long a, long long b, long long c, long long d, bool varargs) {
  if (varargs) {
    double r = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(0.0, xlogx(a)), xlogx(b)), xlogx(c)), xlogx(d));
    return __dsub_rn(xlogx(a + b + c + d), r);
  }
  return __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(xlogx(a + b + c + d), xlogx(a)), xlogx(b)), xlogx(c)), xlogx(d));
}
__device__ __forceinline__ double llr_cells(long long k11, long long k12, long long k21, long long k22, bool varargs) {
  double row_e = entropy2(k11 + k12, k21 + k22, varargs);
  double col_e = entropy2(k11 + k21, k12 + k22, varargs);
  double mat_e = entropy4(k11, k12, k21, k22, varargs);
  double s = __dadd_rn(row_e, col_e);
  if (s < mat_e) return 0.0;  // round off error
  return __dmul_rn(2.0, __dsub_rn(s, mat_e));
}
// fused form used by the row kernel: row entropy hoisted per row (bit-identical sub-expression)
__device__ __forceinline__ double llr_hoisted(long long k11, long long ra, long long cb, long long n, double row_e,
                                              bool varargs) {
  long long k12 = ra - k11, k21 = cb - k11, k22 = n - ra - cb + k11;
  double col_e = entropy2(cb, n - cb, varargs);
  double mat_e = entropy4(k11, k12, k21, k22, varargs);
  double s = __dadd_rn(row_e, col_e);
  if (s < mat_e) return 0.0;
  return __dmul_rn(2.0, __dsub_rn(s, mat_e));
}

__global__ void k_debug_llr(long long n, const long long *k11, const long long *k12, const long long *k21,
                            const long long *k22, uint32_t flags, double *out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = llr_cells(k11[i], k12[i], k21[i], k22[i], (flags & CCO_FLAG_ENTROPY_VARARGS) != 0);
}

// ------------------------------------------------------------------------------------------------
// Row-parallel passes over a CSR matrix: SG lanes cooperate on one user row.
// ------------------------------------------------------------------------------------------------
constexpr int kSG = 8;  // lanes per user row in the preparation passes (avg row ~10-30 entries)

// Row-parallel passes give a row to a sub-group of kSG lanes.  A user with thousands of entries would keep one sub-group
// busy long after the rest of the grid has drained (Zipf users: the top row of C3 has ~6 K entries, of C4 ~40 K) -- and
// that tail does not shrink when the users are sharded over GPUs.  Rows above kHeavyRow entries are therefore listed once
// (k_list_heavy_rows) and handled by a second launch of the same kernel with a whole warp per listed row.
constexpr int kHeavyRow = 256;
__global__ void k_list_heavy_rows(long long n_rows, const long long *__restrict__ rp, int32_t *__restrict__ list, int *__restrict__ n_list) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x)
    if (rp[r + 1] - rp[r] > kHeavyRow) list[atomicAdd(n_list, 1)] = (int32_t)r;
}
// the rows one launch walks: every light row (list == nullptr) or the listed heavy ones
#define CCO_ROW_LOOP_BEGIN(SG)                                                                         \
  const int lane = threadIdx.x % SG;                                                                   \
  long long it = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / SG;                              \
  const long long stride = (long long)gridDim.x * blockDim.x / SG;                                     \
  const long long n_it = list ? (long long)*n_list : n_rows;                                           \
  for (; it < n_it; it += stride) {                                                                    \
    const long long row = list ? (long long)list[it] : it;
#define CCO_ROW_LOOP_END }

// flags[0] |= malformed (row_ptr not monotone / outside [q_lo, q_hi] / column out of range), flags[1] |= not canonical
template <int SG>
__global__ void k_check_rows(long long n_rows, int32_t n_cols, const long long *__restrict__ rp, const int32_t *__restrict__ col,
                             long long q_lo, long long q_hi, const int32_t *__restrict__ list, const int *__restrict__ n_list, int *flags) {
  int bad = 0, unsorted = 0;
  CCO_ROW_LOOP_BEGIN(SG)
    long long s = rp[row], e = rp[row + 1];
    if (e < s || s < q_lo || e > q_hi) { bad = 1; continue; }   // never dereference an offset outside the uploaded block
    if (!list && e - s > kHeavyRow) continue;
    for (long long q = s + lane; q < e; q += SG) {
      int32_t c = col[q];
      if (c < 0 || c >= n_cols) bad = 1;
      if (q > s && col[q - 1] >= c) unsorted = 1;
    }
  CCO_ROW_LOOP_END
  if (bad) atomicOr(&flags[0], 1);
  if (unsorted) atomicOr(&flags[1], 1);
}

// raw column counts c_j of rows [row_begin,row_end) (numNonZeroElementsPerColumn of the raw matrix)
// The counters are REPLICATED (n_copies arrays, copy_stride apart, chosen by CTA): with Zipf-skewed items 8 % of all
// entries hit one column, and its atomics serialise in one L2 slice (~0.3 ms per 12.5 M-entry matrix at C3);
// k_sum_copies folds the copies back into copy 0.
__global__ void k_col_histogram(long long row_begin, long long row_end, const long long *__restrict__ rp,
                                const int32_t *__restrict__ col, int32_t n_cols, int32_t *__restrict__ counts, int n_copies,
                                long long copy_stride) {
  // element-parallel over the contiguous slice rp[row_begin]..rp[row_end]; warp-uniform trip count
  const long long s = rp[row_begin], e = rp[row_end];
  const int lane = threadIdx.x & 31;
  int32_t *mine = counts + (long long)(blockIdx.x % n_copies) * copy_stride;
  for (long long q0 = s + blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); q0 < e;
       q0 += (long long)gridDim.x * blockDim.x) {
    const long long q = q0 + lane;
    const bool act = q < e;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if (act) {
      int32_t c = col[q];
      // warp-aggregate lanes hitting the same column (Zipf-hot columns)
      unsigned peers = __match_any_sync(am, c);
      // (an out-of-range id of a not yet validated matrix is skipped here and reported by k_check_rows)
      if ((__ffs(peers) - 1) == lane && (uint32_t)c < (uint32_t)n_cols) atomicAdd(&mine[c], __popc(peers));
    }
  }
}
__global__ void k_sum_copies(long long n, int n_copies, long long copy_stride, int32_t *__restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int32_t acc = counts[i];
    for (int k = 1; k < n_copies; ++k) acc += counts[i + k * copy_stride];
    counts[i] = acc;
  }
}

// column histogram of the entries [*lo, *hi) of a column-index array (bounds read on the device, 32-bit offsets)
__global__ void k_col_histogram_u32(const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi,
                                    const int32_t *__restrict__ col, int32_t *__restrict__ counts) {
  const long long s = *lo, e = *hi;
  const int lane = threadIdx.x & 31;
  for (long long q0 = s + blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); q0 < e;
       q0 += (long long)gridDim.x * blockDim.x) {
    const long long q = q0 + lane;
    const bool act = q < e;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if (act) {
      const int32_t c = col[q];
      const unsigned peers = __match_any_sync(am, c);
      if ((__ffs(peers) - 1) == lane) atomicAdd(&counts[c], __popc(peers));
    }
  }
}

__device__ __forceinline__ double row_sample_rate(long long d, int32_t m, bool intdiv) {
  if (d <= 0) return 1.0;
  const long long md = d < m ? d : (long long)m;
  return intdiv ? (double)(md / d) : __ddiv_rn((double)md, (double)d);
}
// The sampler keeps (u, j) iff u01 <= min(rowRate, colRate) with u01 = (h >> 11) * 2^-53 (include/cco_b200.h "Sampler";
// bit-identical to oracle/cco_oracle.c orc_downsample).  u01 is a 53-bit integer scaled by a power of two, so the
// comparison is exactly  (h >> 11) <= floor(rate * 2^53): rates become integer thresholds -- one per column (k_col_thresholds,
// once per train) and one per row -- and an entry costs one mix64 and one integer compare, no fp64 division.  A rate of
// 1 (row and column within m) gets the threshold 2^53: u01 < 1 always passes, the hash is not even computed.
constexpr unsigned long long kKeepAlways = 1ULL << 53;
__device__ __forceinline__ unsigned long long rate_threshold(double rate) {
  if (rate >= 1.0) return kKeepAlways;
  return (unsigned long long)floor(__dmul_rn(rate, 0x1.0p53));   // exact: a power-of-two scaling, then the integer part
}
__global__ void k_col_thresholds(int32_t n_cols, const int32_t *__restrict__ raw_counts, int32_t m, unsigned long long *__restrict__ thr) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_cols; j += gridDim.x * blockDim.x) {
    const int32_t c = raw_counts[j];
    thr[j] = c <= m ? kKeepAlways : rate_threshold(__ddiv_rn((double)m, (double)c));
  }
}
__device__ __forceinline__ bool keep_entry_thr(unsigned long long t_row, unsigned long long t_col, uint64_t x_row, uint32_t j) {
  const unsigned long long t = t_row < t_col ? t_row : t_col;
  if (t >= kKeepAlways) return true;
  const uint64_t h = mix64(x_row + (uint64_t)j * 0x9e3779b97f4a7c15ULL);
  return (h >> 11) <= t;
}

// pass 1 of sampleDownAndBinarize: kept entries per row + post-sample column marginals.
// The matrix handed in is a block of n_rows user rows (the whole matrix on one GPU, this rank's user block otherwise);
// row_base = global index of its first user: the sampler hashes GLOBAL user ids and kept_per_row is indexed globally.
template <int SG>
__global__ void k_downsample_count(long long n_rows, long long row_base, const long long *__restrict__ rp, const int32_t *__restrict__ col,
                                   int32_t n_cols, long long q_lo, long long q_hi, const unsigned long long *__restrict__ col_thr, int32_t m,
                                   int32_t seed, uint32_t flags, const int32_t *__restrict__ list, const int *__restrict__ n_list,
                                   uint32_t *__restrict__ kept_per_row, int32_t *__restrict__ new_counts, uint8_t *__restrict__ keep_flag) {
  const unsigned sg_mask = SG == 32 ? 0xffffffffu : (((1u << SG) - 1u) << ((threadIdx.x & 31) / SG * SG));
  const bool intdiv = (flags & CCO_FLAG_ROWRATE_INTDIV) != 0;
  // all lanes of a sub-group share `row`, so loop trip counts are sub-group uniform
  CCO_ROW_LOOP_BEGIN(SG)
    long long s = rp[row], e = rp[row + 1], d = e - s;
    if (!list && d > kHeavyRow) continue;
    if (d < 0 || s < q_lo || e > q_hi) continue;   // malformed row_ptr: k_check_rows reports it
    const unsigned long long t_row = rate_threshold(row_sample_rate(d, m, intdiv));
    const uint32_t g = (uint32_t)(row_base + row);
    const uint64_t x_row = mix64(((uint64_t)(uint32_t)seed << 32) | (uint64_t)g);
    uint32_t kept = 0;
    for (long long q0 = s; q0 < e; q0 += SG) {
      long long q = q0 + lane;
      bool keep = false;
      int32_t j = 0;
      if (q < e) {
        j = col[q];
        // (ids outside [0, n_cols) belong to a malformed matrix: dropped here, reported by k_check_rows)
        keep = (uint32_t)j < (uint32_t)n_cols && keep_entry_thr(t_row, col_thr[j], x_row, (uint32_t)j);
        keep_flag[q - q_lo] = keep ? 1 : 0;   // pass 2 compacts by these decisions instead of hashing again
      }
      if (keep && new_counts) atomicAdd(&new_counts[j], 1);
      kept += __popc(__ballot_sync(sg_mask, keep) & sg_mask);
    }
    if (lane == 0) kept_per_row[g] = kept;
  CCO_ROW_LOOP_END
}

// pass 2: ordered compaction by the recorded decisions (ascending columns are preserved).  new_ptr is the GLOBAL row
// pointer of the sampled matrix; out_base (nullable) points at the entry the output buffer starts at (this rank's block
// offset when the block is written into a send buffer, null = 0 when it is written in place).
template <int SG>
__global__ void k_downsample_write(long long n_rows, long long row_base, const long long *__restrict__ rp, const int32_t *__restrict__ col,
                                   long long q_lo, long long q_hi, const uint8_t *__restrict__ keep_flag,
                                   const int32_t *__restrict__ list, const int *__restrict__ n_list,
                                   const uint32_t *__restrict__ new_ptr, const uint32_t *__restrict__ out_base, int32_t *__restrict__ new_col) {
  const int sg_shift = SG == 32 ? 0 : (threadIdx.x & 31) / SG * SG;
  const unsigned sg_mask = SG == 32 ? 0xffffffffu : (((1u << SG) - 1u) << sg_shift);
  const uint32_t base = out_base ? *out_base : 0u;
  CCO_ROW_LOOP_BEGIN(SG)
    long long s = rp[row], e = rp[row + 1], d = e - s;
    if (!list && d > kHeavyRow) continue;
    if (d < 0 || s < q_lo || e > q_hi) continue;
    uint32_t w = new_ptr[row_base + row] - base;
    for (long long q0 = s; q0 < e; q0 += SG) {
      long long q = q0 + lane;
      bool keep = false;
      int32_t j = 0;
      if (q < e) {
        j = col[q];
        keep = keep_flag[q - q_lo] != 0;
      }
      unsigned b = (__ballot_sync(sg_mask, keep) & sg_mask) >> sg_shift;
      if (keep) new_col[w + __popc(b & ((1u << lane) - 1u))] = j;
      w += __popc(b);
    }
  CCO_ROW_LOOP_END
}

// multi-GPU: every rank sampled its user block into a padded send buffer; after the all-gather the W padded blocks
// (cap entries apart) are packed into the contiguous column array of the sampled matrix.  Block q holds users
// [q * S, min((q + 1) * S, U)): its length is new_ptr[end] - new_ptr[begin], read here -- no host round trip.
__global__ void k_pack_blocks(int world, long long S, long long U, long long cap, const uint32_t *__restrict__ new_ptr,
                              const int32_t *__restrict__ gathered, int32_t *__restrict__ new_col) {
  for (int q = blockIdx.y; q < world; q += gridDim.y) {
    const long long u0 = min((long long)q * S, U), u1 = min(u0 + S, U);
    const uint32_t lo = new_ptr[u0], hi = new_ptr[u1];
    const int32_t *src = gathered + (long long)q * cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < hi - lo; i += gridDim.x * blockDim.x) new_col[lo + i] = src[i];
  }
}

// `drmA.t`: scatter users into per-item lists (order inside a list is irrelevant to the integer counts)
__global__ void k_transpose_scatter(long long n_rows, const uint32_t *__restrict__ rp, const int32_t *__restrict__ col,
                                    uint32_t *__restrict__ cursor, int32_t *__restrict__ users) {
  const int lane = threadIdx.x % kSG;
  long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / kSG;
  const long long stride = (long long)gridDim.x * blockDim.x / kSG;
  for (; row < n_rows; row += stride) {
    uint32_t s = rp[row], e = rp[row + 1];
    for (uint32_t q = s + lane; q < e; q += kSG) users[atomicAdd(&cursor[col[q]], 1u)] = (int32_t)row;
  }
}

// w_a = sum over users of item a of degB'(u)  (= products of output row a), saturating; also P
__global__ void k_row_work(int32_t n_items, const uint32_t *__restrict__ at_ptr, const int32_t *__restrict__ at_users,
                           const uint32_t *__restrict__ b_ptr, uint32_t *__restrict__ row_work,
                           unsigned long long *__restrict__ work64, int32_t *__restrict__ item_ids,
                           uint2 *__restrict__ ext) {
  const int lane = threadIdx.x % kSG;
  const unsigned sg_mask = ((1u << kSG) - 1u) << ((threadIdx.x & 31) / kSG * kSG);
  int item = (blockIdx.x * blockDim.x + threadIdx.x) / kSG;
  const int stride = gridDim.x * blockDim.x / kSG;
  for (; item < n_items; item += stride) {
    uint32_t s = at_ptr[item], e = at_ptr[item + 1];
    unsigned long long w = 0;
    for (uint32_t q = s + lane; q < e; q += kSG) {
      int32_t u = at_users[q];
      const uint32_t bs = b_ptr[u], bl = b_ptr[u + 1] - bs;
      w += bl;
      if (ext) ext[q] = make_uint2(bs, bl);   // the row kernel streams these instead of gathering b_ptr twice per user
    }
#pragma unroll
    for (int o = kSG / 2; o > 0; o >>= 1) w += __shfl_xor_sync(sg_mask, w, o);
    if (lane == 0) {
      row_work[item] = w > 0xffffffffULL ? 0xffffffffu : (uint32_t)w;
      work64[item] = w;
      item_ids[item] = item;
    }
  }
}

// bin boundaries inside the work-descending row list: bin b holds rows whose distinct-cell bound
// D = min(w, n_cols_b) satisfies thresholds[b-1] >= D > thresholds[b]  (thresholds descending)
struct BinThresholds {
  uint32_t t[12];  // by value in the launch parameters: no host->device copy, no host sync
};
__global__ void k_bin_bounds(int32_t n_rows, const uint32_t *__restrict__ sorted_work, int32_t n_bins,
                             const BinThresholds thresholds, int32_t *__restrict__ bounds) {
  int b = threadIdx.x;
  if (b > n_bins) return;
  if (b == 0) { bounds[0] = 0; return; }
  // first index whose work <= thresholds[b-1]  (sorted descending).  The last bin ends at the first row WITHOUT work:
  // rows of other ranks (masked to zero by k_mask_work) and empty rows need no kernel, their out_len is preset to 0.
  uint32_t t = b == n_bins ? 0u : thresholds.t[b - 1];
  int lo = 0, hi = n_rows;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sorted_work[mid] > t) lo = mid + 1; else hi = mid;
  }
  bounds[b] = lo;
}

// ------------------------------------------------------------------------------------------------
// Per-column constants of B' for the fused LLR: columnEntropy = entropy(cb, N - cb) depends only on the
// column, so it is evaluated once per column (same operations, same bits) instead of once per cell.
// ------------------------------------------------------------------------------------------------
struct __align__(16) ColTerm {
  double col_e;     // entropy(cb, N - cb)
  double x_cbm1;    // xLogX(cb - 1): the k21 term of every k11 == 1 cell of this column
  int32_t cb;
  int32_t pad[3];
};
__global__ void k_col_terms(int32_t n_cols, const int32_t *__restrict__ marg, long long n_users, uint32_t flags,
                            ColTerm *__restrict__ out) {
  const bool varargs = (flags & CCO_FLAG_ENTROPY_VARARGS) != 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cols; i += gridDim.x * blockDim.x) {
    long long cb = marg[i];
    ColTerm t;
    t.col_e = entropy2(cb, n_users - cb, varargs);
    t.x_cbm1 = cb >= 1 ? xlogx(cb - 1) : 0.0;
    t.cb = (int32_t)cb;
    t.pad[0] = t.pad[1] = t.pad[2] = 0;
    out[i] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// The fused row kernel.  A GROUP of threads (one warp, or a whole CTA of 256 / 1024 threads) owns one
// primary item a at a time and keeps everything for that row in shared memory:
//   count  : for u in users(a): for b in B'[u]: table[b]++     32-user chunks per warp, products flattened
//                                                               over lanes by a warp prefix-sum + shuffle search
//   compact: occupied table words -> dense per-warp lists (in place)
//   score  : LLR(k11, colA[a], colB[b], N) in fp64, in registers, 2 logs per cell (the other xLogX terms
//            are per-row / per-column / small-integer tables holding bit-identical values)
//   select : running top-k under the total order (llr desc, col asc): threshold-pruned candidate buffer
// DENSE: table indexed by b directly (n_cols_b <= slots); otherwise a packed open-addressing hash
// (key << count_bits | count), multi-pass over hash partitions when the row's distinct-cell bound
// exceeds the table capacity.  Nothing of A'^T B' is ever written to HBM except the kept top-k.
// ------------------------------------------------------------------------------------------------
// candidate entry, 16 bytes so that the sort moves it with one LDS.128 / STS.128:
//   x,y = low/high word of the fp64 LLR bit pattern (positive doubles order like unsigned integers),
//   z = column, w = k11
__device__ __forceinline__ bool cand_better(const uint4 &p, const uint4 &q) {
  return p.y > q.y || (p.y == q.y && (p.x > q.x || (p.x == q.x && p.z < q.z)));
}

template <int GROUP>
__device__ __forceinline__ void group_sync() {
  if (GROUP == 32) __syncwarp(); else __syncthreads();
}

// bitonic sort of the candidate buffer, best first; pads [n, n2) with key 0 (never valid: LLR > 0)
template <int GROUP>
__device__ void sort_candidates(uint4 *tk, int n, int gtid) {
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + gtid; i < n2; i += GROUP) tk[i] = make_uint4(0u, 0u, 0xffffffffu, 0u);
  group_sync<GROUP>();
  for (int k2 = 2; k2 <= n2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = gtid; t < (n2 >> 1); t += GROUP) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool up = (i & k2) == 0;
        const uint4 ei = tk[i], ep = tk[p];
        const bool swap = up ? cand_better(ep, ei) : cand_better(ei, ep);
        if (swap) { tk[i] = ep; tk[p] = ei; }
      }
      group_sync<GROUP>();
    }
  }
}

// ---- candidate reduction: MSB-first 8-bit radix select on the composite key (llr hi, llr lo, ~col) ----------
// Keeps every candidate >= a threshold T chosen so that  top_k <= kept <= max(M, top_k)  (exactly top_k when the
// key is fully resolved; keys are unique because columns are).  Cost: a few histogram passes over the buffer
// instead of a full sort.  Returns the kept count; the threshold entry goes to thr (same layout as a candidate).
__device__ __forceinline__ uint32_t cand_word(const uint4 &e, int wi) { return wi == 0 ? e.y : (wi == 1 ? e.x : ~e.z); }

template <int GROUP>
__device__ int reduce_candidates(uint4 *tk, uint4 *aux, int n, int k, int M, int *hist, int *ctrl, int gtid) {
  // ctrl[16..18] = prefix words, ctrl[24] = D, ctrl[25] = c_gt, ctrl[26] = c_D, ctrl[27] = output cursor
  volatile int *vc = ctrl;
  uint32_t pre0 = 0u, pre1 = 0u, pre2 = 0u;
  int nb = 0, above = 0, kept = n;
  while (true) {
    for (int i = gtid; i < 256; i += GROUP) hist[i] = 0;
    group_sync<GROUP>();
    const int wi = nb >> 5, sh = 24 - (nb & 31);
    for (int i = gtid; i < n; i += GROUP) {
      const uint4 e = tk[i];
      const uint32_t w0 = e.y, w1 = e.x, w2 = ~e.z;
      bool match;
      if (nb == 0) match = true;
      else if (nb < 32) match = (w0 >> (32 - nb)) == (pre0 >> (32 - nb));
      else if (nb == 32) match = w0 == pre0;
      else if (nb < 64) match = w0 == pre0 && (w1 >> (64 - nb)) == (pre1 >> (64 - nb));
      else if (nb == 64) match = w0 == pre0 && w1 == pre1;
      else match = w0 == pre0 && w1 == pre1 && (w2 >> (96 - nb)) == (pre2 >> (96 - nb));
      // digit = byte (sh / 8) of the key word, extracted with PRMT: ptxas 12.9 turned `(word >> 24) & 255` of the
      // peeled nb == 0 pass into an index by the WHOLE word in one inlining context (compute-sanitizer: invalid shared atomic)
      if (match) atomicAdd(&hist[__byte_perm(cand_word(e, wi), 0u, 0x4440u | (uint32_t)(sh >> 3))], 1);
    }
    group_sync<GROUP>();
    if (gtid < 32) {
      // lane l owns digits 255-8l .. 248-8l (descending); find the digit holding the (k-above)-th best
      int c[8], ssum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * gtid - j]; ssum += c[j]; }
      int incl = ssum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (gtid >= d) incl += v;
      }
      const int excl = incl - ssum, need = k - above;
      if (excl < need && need <= incl) {
        int run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (run < need && run + c[j] >= need) { ctrl[24] = 255 - 8 * gtid - j; ctrl[25] = run; ctrl[26] = c[j]; }
          run += c[j];
        }
      }
    }
    group_sync<GROUP>();
    const int D = vc[24], c_gt = vc[25], c_d = vc[26];
    if (wi == 0) pre0 |= (uint32_t)D << sh; else if (wi == 1) pre1 |= (uint32_t)D << sh; else pre2 |= (uint32_t)D << sh;
    nb += 8;
    kept = above + c_gt + c_d;
    if (kept <= M || nb == 96) break;
    above += c_gt;
    group_sync<GROUP>();
  }
  // keep e iff key(e) >= prefix (low bits zero)
  if (gtid == 0) ctrl[27] = 0;
  group_sync<GROUP>();
  if (GROUP == 32) {
    // single warp: in-place batch compaction (reads of a batch complete before its writes)
    int w = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + gtid;
      uint4 e = make_uint4(0u, 0u, 0u, 0u);
      bool keep = false;
      if (i < n) {
        e = tk[i];
        const uint32_t w0 = e.y, w1 = e.x, w2 = ~e.z;
        keep = w0 > pre0 || (w0 == pre0 && (w1 > pre1 || (w1 == pre1 && w2 >= pre2)));
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) tk[w + __popc(m & ((1u << gtid) - 1u))] = e;
      w += __popc(m);
      __syncwarp();
    }
  } else {
    for (int i = gtid; i < n; i += GROUP) {
      const uint4 e = tk[i];
      const uint32_t w0 = e.y, w1 = e.x, w2 = ~e.z;
      const bool keep = w0 > pre0 || (w0 == pre0 && (w1 > pre1 || (w1 == pre1 && w2 >= pre2)));
      if (keep) aux[atomicAdd(&ctrl[27], 1)] = e;
    }
    group_sync<GROUP>();
    for (int i = gtid; i < kept; i += GROUP) tk[i] = aux[i];
  }
  if (gtid == 0) {
    ctrl[0] = kept;
    ctrl[4] = (int)pre1; ctrl[5] = (int)pre0; ctrl[6] = (int)~pre2; ctrl[7] = 0;  // threshold as a candidate entry
    ctrl[1] = 1;
  }
  group_sync<GROUP>();
  return kept;
}

constexpr int kCutBins = 512;   // level-1 integer cut: u16 colB bins; they alias the 1 KB radix-select histogram (dead until the score loop)
constexpr int kDomLevels = 15;  // dominance filter keeps cfail[1..15] in ctrl[41..55]
constexpr int kX12N = 31;  // x12tab[j] = xLogX(ra - j) for j < 31; x12tab[31] = xLogX(N - ra)

template <int GROUP, bool DENSE>
__device__ __forceinline__ void accumulate(uint32_t *table, uint32_t tsize, uint32_t b, int cbits, uint32_t n_pass,
                                           uint32_t pass, int *err_flag) {
  if (DENSE) {
    atomicAdd(&table[b], 1u);
    return;
  }
  if (n_pass > 1 && ((b * 0x85ebca6bu) >> 12) % n_pass != pass) return;
  uint32_t slot = __umulhi(b * 0x9e3779b1u, tsize);
  const uint32_t want = b << cbits;
  uint32_t probes = 0;
  while (true) {
    const uint32_t w = *reinterpret_cast<volatile uint32_t *>(&table[slot]);
    if ((w >> cbits) == b && w != kEmpty) { atomicAdd(&table[slot], 1u); return; }
    if (w == kEmpty) {
      const uint32_t old = atomicCAS(&table[slot], kEmpty, want | 1u);
      if (old == kEmpty) return;
      if ((old >> cbits) == b) { atomicAdd(&table[slot], 1u); return; }
    }
    slot = (slot + 1 == tsize) ? 0 : slot + 1;
    if (++probes > tsize) { atomicOr(err_flag, 1); return; }
  }
}

template <int GROUP, bool DENSE>
__global__ void __launch_bounds__(GROUP == 32 ? 256 : GROUP) k_rows(const RowArgs a) {
  const int GROUPS = GROUP == 32 ? (int)(blockDim.x >> 5) : 1;  // warp-owned rows: several independent warps per CTA
  constexpr int NW = GROUP / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int gid = tid / GROUP, gtid = tid % GROUP, gw = gtid >> 5;
  unsigned char *base = smem_raw + (size_t)gid * a.group_smem_bytes;
  uint4 *tk = reinterpret_cast<uint4 *>(base);
  uint4 *aux = tk + a.cbuf;                                  // a.caux entries (0 for warp-owned rows)
  double *x12tab = reinterpret_cast<double *>(aux + a.caux);
  double *x11tab = x12tab + 32;
  int *ctrl = reinterpret_cast<int *>(x11tab + 32);  // [0] ncand [1] have_thr [4..7] threshold entry [16..27] select state [40..55] dominance frontier [64..64+NW) per-warp list sizes
  int *hist = ctrl + 128;                                     // 256 bins of the radix select
  uint32_t *wqueue = reinterpret_cast<uint32_t *>(hist + 256);  // NW * 64 queued cells awaiting evaluation
  uint32_t *h1 = reinterpret_cast<uint32_t *>(hist);   // kCutBins/2 words: u16 histogram of colB over the strongly positive k11 == 1 cells
  uint32_t *table = wqueue + NW * 64;
  volatile int *vctrl = ctrl;

  const int row_begin = a.bin_bounds[a.bin], row_end = a.bin_bounds[a.bin + 1];
  const bool varargs = (a.flags & CCO_FLAG_ENTROPY_VARARGS) != 0;
  const int cbits = a.count_bits;
  const uint32_t cmask = (1u << cbits) - 1u;
  const int prune_limit = a.cbuf - GROUP;
  const long long N = a.n_users;
  const double xN = xlogx(N);
  unsigned long long distinct_local = 0, evaluated_local = 0;
  if (gtid < 32) x11tab[gtid] = xlogx((long long)gtid);

  for (int ri = row_begin + blockIdx.x * GROUPS + gid; ri < row_end; ri += gridDim.x * GROUPS) {
    const int item = a.rows_sorted[ri];
    const uint32_t u_begin = a.at_ptr[item], u_end = a.at_ptr[item + 1];
    const long long ra = a.marg_a[item];
    // table sized to the row: load factor <= 1/2 of the distinct-cell bound D = min(w, n_cols_b)
    uint32_t n_pass = 1, tsize = (uint32_t)a.n_cols_b;
    if (!DENSE) {
      const uint32_t w = a.row_work[item];
      const uint32_t dbound = w < (uint32_t)a.n_cols_b ? w : (uint32_t)a.n_cols_b;
      n_pass = (dbound + (uint32_t)a.cap - 1u) / (uint32_t)a.cap;
      if (n_pass == 0) n_pass = 1;
      tsize = n_pass > 1 ? (uint32_t)a.slots
                         : (uint32_t)min((unsigned long long)a.slots,
                                         max(((unsigned long long)dbound * (unsigned long long)a.tsize_x16) >> 4, 64ull));
      tsize = min((uint32_t)a.slots, (tsize + 32u * NW - 1u) / (32u * NW) * (32u * NW));
    }
    group_sync<GROUP>();  // previous row fully done with shared memory
    if (gtid < 32) {
      const long long v = (gtid < kX12N) ? ra - gtid : N - ra;
      x12tab[gtid] = v >= 0 ? xlogx(v) : 0.0;
    }
    if (gtid == 0) { ctrl[0] = 0; ctrl[1] = 0; }
    if (gtid < 16) ctrl[40 + gtid] = 0x7fffffff;
    int emitted = 0;

    for (uint32_t pass = 0; pass < n_pass; ++pass) {
      // ---- clear --------------------------------------------------------------------------------------
      for (uint32_t i = gtid; i < tsize; i += GROUP) table[i] = DENSE ? 0u : kEmpty;
      group_sync<GROUP>();
      // ---- count: each warp takes 32-user chunks; products of a chunk are flattened over the lanes -----
      // users are dealt to the group's warps in equal chunks of <= 32 so short rows still use every warp
      const uint32_t deg = u_end - u_begin;
      const uint32_t per = NW == 1 ? 32u : min(32u, max(1u, (deg + NW - 1) / NW));
      for (uint32_t c0 = u_begin + gw * per; c0 < u_end; c0 += NW * per) {
        const uint32_t i = c0 + lane;
        uint32_t s = 0, len = 0;
        if (lane < per && i < u_end) {
          const int32_t u = a.at_users[i];
          s = a.b_ptr[u];
          len = a.b_ptr[u + 1] - s;
        }
        uint32_t off = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, off, d);
          if (lane >= d) off += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, off, 31);
        off -= len;  // exclusive
        for (uint32_t p0 = 0; p0 < total; p0 += 64) {
          // two products per lane per trip (two independent gathers in flight)
          uint32_t bb[2];
          bool act[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t p = p0 + h * 32 + lane;
            int j = 0;
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
              const int c = j + st;
              const uint32_t v = __shfl_sync(0xffffffffu, off, c);
              if (v <= p) j = c;
            }
            const uint32_t sj = __shfl_sync(0xffffffffu, s, j), oj = __shfl_sync(0xffffffffu, off, j);
            act[h] = p < total;
            bb[h] = act[h] ? (uint32_t)a.b_col[sj + (p - oj)] : 0u;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (act[h]) accumulate<GROUP, DENSE>(table, tsize, bb[h], cbits, n_pass, pass, a.err_flag);
        }
      }
      group_sync<GROUP>();
      // ---- compact: each warp packs the occupied words of its own table segment, in place ----------------
      const uint32_t seg = (((tsize + NW - 1) / NW) + 31u) & ~31u;
      const uint32_t seg_lo = min((uint32_t)gw * seg, tsize), seg_hi = min(seg_lo + seg, tsize);
      uint32_t n_mine = 0;
      for (uint32_t pos = seg_lo; pos < seg_hi; pos += 32) {
        const uint32_t idx = pos + lane;
        uint32_t w = DENSE ? 0u : kEmpty;
        if (idx < seg_hi) w = table[idx];
        const bool valid = DENSE ? (w != 0u) : (w != kEmpty);
        const uint32_t word = DENSE ? ((idx << cbits) | w) : w;
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        __syncwarp();
        if (valid) table[seg_lo + n_mine + __popc(m & ((1u << lane) - 1u))] = word;
        n_mine += __popc(m);
        __syncwarp();
      }
      if (lane == 0) distinct_local += n_mine;
      if (a.emit_all) {
        // debug: every non-zero cell of the row (col, count), unordered
        int basepos = 0;
        if (lane == 0) basepos = atomicAdd(&ctrl[0], (int)n_mine);
        basepos = __shfl_sync(0xffffffffu, basepos, 0);
        for (uint32_t q = lane; q < n_mine; q += 32) {
          const uint32_t word = table[seg_lo + q];
          const size_t o = (size_t)item * a.out_stride + emitted + basepos + q;
          a.out_col[o] = (int32_t)(word >> cbits);
          a.out_cnt[o] = (int32_t)(word & cmask);
        }
        group_sync<GROUP>();
        emitted += vctrl[0];
        group_sync<GROUP>();
        if (gtid == 0) ctrl[0] = 0;
        group_sync<GROUP>();
        continue;
      }
      // ---- level-1 integer cut (exact) ---------------------------------------------------------------
      // On the strongly positive side (2*rowA*colB < k11*N) the LLR of k11 == 1 cells is strictly decreasing in colB, so
      // the smallest colB c1 with >= top_k such cells at or below it bounds the row's k-th best from below: k11 == 1 cells
      // with colB > c1 can never be kept and are dropped by an integer compare in the filter stage.
      int cut1 = 0x7fffffff;
      if (a.row_work[item] < 65536u) {   // u16 bins cannot overflow
        for (int i = gtid; i < kCutBins / 2; i += GROUP) h1[i] = 0u;
        group_sync<GROUP>();
        for (uint32_t q0 = 0; q0 < n_mine; q0 += 32) {
          const uint32_t q = q0 + lane;
          if (q < n_mine) {
            const uint32_t word = table[seg_lo + q];
            const uint32_t b = word >> cbits;
            if ((word & cmask) == 1u && !(a.self && (int)b == item)) {
              const uint32_t cb = (uint32_t)a.marg_b[b];
              if (cb < (uint32_t)kCutBins && 2ull * (unsigned long long)ra * cb < (unsigned long long)N)
                atomicAdd(&h1[cb >> 1], 1u << (16u * (cb & 1u)));
            }
          }
        }
        group_sync<GROUP>();
        if (gtid < 32) {
          // lane l owns bins [16 l, 16 l + 16): 8 words
          uint32_t sum = 0;
          for (int wi = 0; wi < 8; ++wi) { const uint32_t v = h1[gtid * 8 + wi]; sum += (v & 0xffffu) + (v >> 16); }
          uint32_t incl = sum;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (gtid >= d) incl += v;
          }
          const uint32_t excl = incl - sum;
          int found = 0x7fffffff;
          if (excl < (uint32_t)a.top_k && incl >= (uint32_t)a.top_k) {
            uint32_t run = excl;
            for (int wi = 0; wi < 8 && found == 0x7fffffff; ++wi) {
              const uint32_t v = h1[gtid * 8 + wi];
              run += v & 0xffffu;
              if (run >= (uint32_t)a.top_k) { found = gtid * 16 + 2 * wi; break; }
              run += v >> 16;
              if (run >= (uint32_t)a.top_k) { found = gtid * 16 + 2 * wi + 1; break; }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
          if (gtid == 0) ctrl[9] = found;
        }
        group_sync<GROUP>();
        cut1 = vctrl[9];
      }
      // ---- score + select -----------------------------------------------------------------------------------
      const double x_ra = x12tab[0], x_nra = x12tab[kX12N];
      const double row_e = varargs ? __dsub_rn(xN, __dadd_rn(__dadd_rn(0.0, x_ra), x_nra))
                                   : __dsub_rn(__dsub_rn(xN, x_ra), x_nra);
      // Two stages per warp so that the fp64 evaluation always runs on full warps:
      //   filter  : 32 cells at a time through the exact dominance filter (integer work only); survivors are queued
      //   evaluate: 32 queued cells at a time -> LLR -> threshold test -> candidate buffer
      // One evaluation batch per round, then (CTA-owned rows) one barrier that also decides whether any warp has work.
      uint32_t *wq = wqueue + gw * 64;
      uint32_t pos = 0;
      int qn = 0;
      while (true) {
        while (qn < 32 && pos < n_mine) {
          const uint32_t q = pos + lane;
          bool surv = false;
          uint32_t word = 0;
          if (q < n_mine) {
            word = table[seg_lo + q];
            const uint32_t b = word >> cbits, k11 = word & cmask;
            if (!(a.self && (int)b == item)) {
              // Dominance filter (exact, DESIGN.md "dominance"): for fixed rowA and N, on the positively associated
              // side (rowA*cb < k11*N) the LLR grows with k11 and shrinks with cb, so every evaluated cell (k, c) that
              // fails strictly on LLR proves that all cells (k' <= k, c' >= c) fail too; cfail[k'] = smallest such c.
              const long long cb = a.marg_b[b];
              const bool pos_side = (unsigned long long)ra * (unsigned long long)cb < (unsigned long long)k11 * (unsigned long long)N;
              surv = !(pos_side && k11 <= (uint32_t)kDomLevels && (int)cb >= vctrl[40 + k11]);
              if (k11 == 1u && (int)cb > cut1 && 2ull * (unsigned long long)ra * (unsigned long long)cb < (unsigned long long)N)
                surv = false;   // beyond the level-1 integer cut
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, surv);
          if (surv) wq[qn + __popc(m & ((1u << lane) - 1u))] = word;
          qn += __popc(m);
          pos += 32;
          __syncwarp();
        }
        const int take = qn < 32 ? qn : 32;
        bool pass_ok = false;
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        if (lane < take) {
          const uint32_t word = wq[qn - take + lane];
          const uint32_t b = word >> cbits, k11 = word & cmask;
          const ColTerm ct = a.col_terms[b];
          const long long cb = ct.cb;
          const bool pos_side = (unsigned long long)ra * (unsigned long long)cb < (unsigned long long)k11 * (unsigned long long)N;
          const uint32_t kf = k11 < (uint32_t)kDomLevels ? k11 : (uint32_t)kDomLevels;
          ++evaluated_local;
          const long long k21 = cb - k11, k22 = N - ra - cb + k11;
          const double x11 = k11 < 32 ? x11tab[k11] : xlogx_u32(k11);
          const double x12 = k11 < kX12N ? x12tab[k11] : xlogx_u32((uint32_t)(ra - k11));
          const double x21 = k11 == 1 ? ct.x_cbm1 : xlogx_u32((uint32_t)k21), x22 = xlogx_u32((uint32_t)k22);
          double mat_e;
          if (varargs)
            mat_e = __dsub_rn(xN, __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(0.0, x11), x12), x21), x22));
          else
            mat_e = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(xN, x11), x12), x21), x22);
          const double sre = __dadd_rn(row_e, ct.col_e);
          const double v = (sre < mat_e) ? 0.0 : __dmul_rn(2.0, __dsub_rn(sre, mat_e));
          const bool min_ok = !a.has_min_llr || v >= a.min_llr;
          pass_ok = v > 0.0 && min_ok;
          // (cells whose LLR rounds to 0 are cancellation noise: they teach nothing)
          bool strict_fail = v > 0.0 && !min_ok;
          const unsigned long long key = (unsigned long long)__double_as_longlong(v);
          e = make_uint4((uint32_t)key, (uint32_t)(key >> 32), b, k11);
          if (pass_ok && vctrl[1]) {
            const uint4 thr = make_uint4((uint32_t)vctrl[4], (uint32_t)vctrl[5], (uint32_t)vctrl[6], (uint32_t)vctrl[7]);
            pass_ok = !cand_better(thr, e);
            strict_fail = e.y < thr.y || (e.y == thr.y && e.x < thr.x);
          }
          if (strict_fail && pos_side)
            for (uint32_t kk = kf; kk >= 1 && (int)cb < vctrl[40 + kk]; --kk) atomicMin(&ctrl[40 + kk], (int)cb);
        }
        qn -= take;
        const unsigned m = __ballot_sync(0xffffffffu, pass_ok);
        int basepos_round = 0;
        if (m) {
          int basepos = 0;
          if (lane == 0) basepos = atomicAdd(&ctrl[0], __popc(m));
          basepos = __shfl_sync(0xffffffffu, basepos, 0);
          if (pass_ok) tk[basepos + __popc(m & ((1u << lane) - 1u))] = e;
          basepos_round = basepos;
        }
        const bool more = qn > 0 || pos < n_mine;
        // Both decisions of this round are taken from barrier results (CTA-uniform by construction).  Re-reading ctrl[0]
        // after the barrier raced with warps that had already looped back and appended (or with warp 0's single-warp
        // select storing the kept count): warps of one CTA could disagree on `n > prune_limit` and pair different
        // barriers.  `mine` = the counter right after this warp's own append (0 if it appended nothing): the counter only
        // grows inside a round, so the largest `mine` is its final value.
        int mine = 0;
        if (m) mine = basepos_round + __popc(m);
        bool any_more, need_prune;
        if (GROUP == 32) {
          __syncwarp();
          any_more = more;
          need_prune = mine > prune_limit;
        } else {
          any_more = __syncthreads_or(more ? 1 : 0) != 0;
          need_prune = __syncthreads_or(mine > prune_limit ? 1 : 0) != 0;
        }
        if (need_prune) {
          const int n = vctrl[0];   // nobody appends until every warp has left this branch
          if (GROUP > 32 && n <= 512) {
            // small buffer: one warp runs the whole select (no CTA barriers inside), the others wait once
            if (gw == 0) reduce_candidates<32>(tk, aux, n, a.top_k, a.keep_max, hist, ctrl, lane);
            group_sync<GROUP>();
          } else {
            reduce_candidates<GROUP>(tk, aux, n, a.top_k, a.keep_max, hist, ctrl, gtid);
          }
        }
        if (!any_more) break;
      }
      group_sync<GROUP>();
    }
    // ---- final select + write -------------------------------------------------------------------------------
    if (a.emit_all) {
      if (gtid == 0) a.out_len[item] = emitted;
    } else {
      int n = vctrl[0];
      if (n > 0) {
        if (GROUP > 32 && n <= 512) {
          if (gw == 0) {
            int m = n;
            if (m > a.final_max) m = reduce_candidates<32>(tk, aux, m, a.top_k, a.final_max, hist, ctrl, lane);
            sort_candidates<32>(tk, m, lane);
            if (lane == 0) ctrl[0] = m;
          }
          group_sync<GROUP>();
          n = vctrl[0];
        } else {
          if (n > a.final_max) n = reduce_candidates<GROUP>(tk, aux, n, a.top_k, a.final_max, hist, ctrl, gtid);
          sort_candidates<GROUP>(tk, n, gtid);
        }
        const int keep = n < a.top_k ? n : a.top_k;
        for (int i = gtid; i < keep; i += GROUP) {
          const size_t o = (size_t)item * a.out_stride + i;
          const uint4 e = tk[i];
          a.out_col[o] = (int32_t)e.z;
          a.out_llr[o] = __longlong_as_double((long long)(((unsigned long long)e.y << 32) | e.x));
          a.out_cnt[o] = (int32_t)e.w;
        }
        if (gtid == 0) a.out_len[item] = keep;
      } else if (gtid == 0) {
        a.out_len[item] = 0;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) distinct_local += __shfl_xor_sync(0xffffffffu, distinct_local, o);
  if (lane == 0 && distinct_local) atomicAdd(a.stat_distinct, distinct_local);
  for (int o = 16; o > 0; o >>= 1) evaluated_local += __shfl_xor_sync(0xffffffffu, evaluated_local, o);
  if (lane == 0 && evaluated_local) atomicAdd(a.stat_evaluated, evaluated_local);
}

// packed output: gather the strided per-row results into CSR order (out_ptr = exclusive scan of the masked row lengths:
// rows outside this rank's range have length 0 there and are skipped)
__global__ void k_compact_rows(int32_t n_items, int32_t stride, const long long *__restrict__ out_ptr,
                               const int32_t *__restrict__ col, const double *__restrict__ llr, const int32_t *__restrict__ cnt,
                               int32_t *__restrict__ p_col, double *__restrict__ p_llr, int32_t *__restrict__ p_cnt) {
  // one warp per row
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int item = warp; item < n_items; item += nwarps) {
    const long long o = out_ptr[item];
    const int n = (int)(out_ptr[item + 1] - o);
    size_t src = (size_t)item * stride;
    for (int i = lane; i < n; i += 32) {
      p_col[o + i] = col[src + i];
      if (p_llr) p_llr[o + i] = llr[src + i];
      if (p_cnt) p_cnt[o + i] = cnt[src + i];
    }
  }
}

// ---- canonicalisation slow path (unsorted / duplicated input rows) -------------------------------
__global__ void k_expand_keys(long long n_rows, const long long *__restrict__ rp, const int32_t *__restrict__ col,
                              unsigned long long *__restrict__ keys) {
  const int lane = threadIdx.x % kSG;
  long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / kSG;
  const long long stride = (long long)gridDim.x * blockDim.x / kSG;
  const long long q_base = rp[0];   // a rank's user block keeps the caller's absolute offsets
  for (; row < n_rows; row += stride) {
    long long s = rp[row], e = rp[row + 1];
    for (long long q = s + lane; q < e; q += kSG) keys[q - q_base] = ((unsigned long long)row << 32) | (uint32_t)col[q];
  }
}
__global__ void k_unique_flags(long long n, const unsigned long long *__restrict__ keys, uint32_t *__restrict__ flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void k_unique_scatter(long long n, const unsigned long long *__restrict__ keys,
                                 const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                 unsigned long long *__restrict__ out_keys, int32_t *__restrict__ out_col) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (flag[i]) { out_keys[pos[i]] = keys[i]; out_col[pos[i]] = (int32_t)(keys[i] & 0xffffffffULL); }
}
// row_ptr[r] = first index whose key >= (r << 32)
__global__ void k_rowptr_from_keys(long long n_rows, long long n_unique, const unsigned long long *__restrict__ keys,
                                   long long *__restrict__ rp) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r <= n_rows; r += (long long)gridDim.x * blockDim.x) {
    unsigned long long t = (unsigned long long)r << 32;
    long long lo = 0, hi = n_unique;
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (keys[mid] < t) lo = mid + 1; else hi = mid;
    }
    rp[r] = lo;
  }
}

// small device -> host results go through a mapped pinned "mailbox" written by this kernel, not through the D2H copy
// engine: a few-byte cudaMemcpyAsync would queue behind the multi-megabyte indicator copies of the previous
// indicator (one DMA FIFO per direction) and stall the launch pipeline behind them.
__global__ void k_mail_bytes(unsigned char *__restrict__ dst_mapped, const unsigned char *__restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst_mapped[i] = src[i];
  __threadfence_system();
}

// device twin of cco_partition_rows (cco_api.cu): contiguous item ranges of equal (products + 1 per row)
__global__ void k_partition_rows(const long long *__restrict__ work_prefix, int32_t n_items, int32_t world, int32_t *bounds) {
  const int r = threadIdx.x;
  if (r > world) return;
  if (r == 0) { bounds[0] = 0; return; }
  if (r == world) { bounds[r] = n_items; return; }
  const long long total = work_prefix[n_items] + n_items;
  const long long target = (long long)((__int128)total * r / world);
  int lo = 0, hi = n_items;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (work_prefix[mid] + mid < target) lo = mid + 1; else hi = mid;
  }
  bounds[r] = lo;
}

// ---- ingest (Preparator.prepare on integer-tokenised events) ---------------------------------------
// per-user event counts of the primary type (duplicates count: Preparator.scala:129-132)
__global__ void k_ingest_count_users(long long n, const long long *__restrict__ user, int32_t *__restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&counts[user[i]], 1);
}
// flag[u] = 1 iff user u stays in the dictionary
__global__ void k_ingest_user_flags(long long n_users_raw, const int32_t *__restrict__ counts, int32_t need,
                                    uint32_t *__restrict__ flag) {
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n_users_raw; u += (long long)gridDim.x * blockDim.x)
    flag[u] = counts[u] >= need ? 1u : 0u;
}
// map[i] = flag[i] ? pos[i] : -1
__global__ void k_ingest_make_map(long long n, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                  int32_t *__restrict__ map) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    map[i] = flag[i] ? (int32_t)pos[i] : -1;
}
// items that still have an event of a surviving user (Preparator.scala:184)
__global__ void k_ingest_item_flags(long long n, const long long *__restrict__ user, const int32_t *__restrict__ item,
                                    const int32_t *__restrict__ user_map, uint32_t *__restrict__ item_flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (user_map[user[i]] >= 0) item_flag[item[i]] = 1u;
}
// key = (new user << 32 | new item) for surviving events, ~0 for dropped ones (they sort to the end)
__global__ void k_ingest_keys(long long n, const long long *__restrict__ user, const int32_t *__restrict__ item,
                              const int32_t *__restrict__ user_map, const int32_t *__restrict__ item_map,
                              unsigned long long *__restrict__ keys, unsigned long long *__restrict__ n_kept) {
  unsigned long long kept = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t r = user_map[user[i]];
    if (r >= 0) {
      keys[i] = ((unsigned long long)(uint32_t)r << 32) | (uint32_t)item_map[item[i]];
      ++kept;
    } else {
      keys[i] = ~0ULL;
    }
  }
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_kept, kept);
}

__global__ void k_fill_u32(long long n, uint32_t v, uint32_t *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = v;
}

// ---- synthetic event streams (bench.py / tests; SURVEY.md 8d spec, numpy twin in synth.py) ---------------------------
// inclusive normalised CDF over ranks -> first rank whose CDF value exceeds u (numpy searchsorted side="right"), clipped
__device__ __forceinline__ int32_t cdf_upper_bound(const double *__restrict__ cdf, int32_t n, double u) {
  int32_t lo = 0, hi = n;
  while (lo < hi) {
    const int32_t mid = (int32_t)(((uint32_t)lo + (uint32_t)hi) >> 1);
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo < n ? lo : n - 1;
}
__global__ void k_synth_events(long long n_events, unsigned long long seed, const double *__restrict__ user_cdf,
                               const int32_t *__restrict__ user_perm, int32_t n_users, const double *__restrict__ item_cdf,
                               const int32_t *__restrict__ item_perm, int32_t n_items, long long *__restrict__ user,
                               int32_t *__restrict__ item) {
  const uint64_t base = mix64(seed);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_events; e += (long long)gridDim.x * blockDim.x) {
    const uint64_t h1 = mix64(base + (uint64_t)(e + 1) * 0x9e3779b97f4a7c15ULL);
    const uint64_t h2 = mix64(h1 ^ 0x6a09e667f3bcc909ULL);
    const double u1 = __dmul_rn((double)(h1 >> 11), 0x1.0p-53), u2 = __dmul_rn((double)(h2 >> 11), 0x1.0p-53);
    user[e] = user_perm[cdf_upper_bound(user_cdf, n_users, u1)];
    item[e] = item_perm[cdf_upper_bound(item_cdf, n_items, u2)];
  }
}

__global__ void k_max_i32(long long n, const int32_t *__restrict__ x, int32_t *__restrict__ out) {
  int32_t m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, x[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// rank partition on the device: work of the rows outside this rank's [bounds[rank], bounds[rank + 1]) becomes 0, so they
// sort behind every row with work and fall out of the last bin; no bound ever travels to the host before the kernels run
__global__ void k_mask_work(int32_t n_items, const uint32_t *__restrict__ row_work, const int32_t *__restrict__ bounds, int rank,
                            uint32_t *__restrict__ masked) {
  const int32_t lo = bounds ? bounds[rank] : 0, hi = bounds ? bounds[rank + 1] : n_items;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x)
    masked[i] = (i >= lo && i < hi) ? row_work[i] : 0u;
}

// kept-cell counts of this rank's rows as int64 (0 outside its range), input of the exclusive scan that gives row_ptr
__global__ void k_len_to_i64(int32_t n_items, const int32_t *__restrict__ len, const int32_t *__restrict__ bounds, int rank,
                             long long *__restrict__ out) {
  const int32_t lo = bounds ? bounds[rank] : 0, hi = bounds ? bounds[rank + 1] : n_items;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n_items; i += gridDim.x * blockDim.x)
    out[i] = (i >= lo && i < hi) ? len[i] : 0;
}

// what the host needs from one indicator, in one mailbox record: [0] row_lo [1] row_hi [2] kept cells [3] products of the
// range [4] distinct cells [5] evaluated cells [6] hash-overflow flag
__global__ void k_indicator_record(int32_t n_items, const int32_t *__restrict__ bounds, int rank, const long long *__restrict__ out_ptr,
                                   const long long *__restrict__ work_prefix, const unsigned long long *__restrict__ distinct_eval,
                                   const int *__restrict__ err, long long *__restrict__ rec) {
  if (threadIdx.x || blockIdx.x) return;
  const int32_t lo = bounds ? bounds[rank] : 0, hi = bounds ? bounds[rank + 1] : n_items;
  rec[0] = lo;
  rec[1] = hi;
  rec[2] = out_ptr[n_items];
  rec[3] = work_prefix[hi] - work_prefix[lo];
  rec[4] = (long long)distinct_eval[0];
  rec[5] = (long long)distinct_eval[1];
  rec[6] = *err;
}

// group (single-process multi-GPU) mode: rebase this rank's row pointers by the cells of the ranks before it
__global__ void k_add_i64(long long n, long long v, long long *__restrict__ x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] += v;
}

}  // namespace cco
