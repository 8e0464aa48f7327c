// cco_api.cu -- C ABI (include/cco_b200.h) and host orchestration of the sm_100a CCO model builder.
//
// Replaces Mahout's SimilarityAnalysis.cooccurrencesIDSs / crossOccurrenceDownsampled as called from
// /root/reference/src/main/scala/URAlgorithm.scala:323-329,343-346.  No CPU fallback: every compute
// entry fails with CCO_E_CUDA when no CUDA device is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <mutex>
#include <vector>

#include "../../include/cco_b200.h"
#include "cco_kernels.cuh"

namespace cco {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CK(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return set_error(_e == cudaErrorMemoryAllocation ? CCO_E_OOM : CCO_E_CUDA, "%s: %s (%s:%d)", #expr, \
                       cudaGetErrorString(_e), __FILE__, __LINE__);                                   \
  } while (0)
#define CKR(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r != CCO_OK) return _r; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily (only multi-GPU contexts need it)
// ------------------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;
static std::mutex g_nccl_mu;
static int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.h) return CCO_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_error(CCO_E_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                         \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                    \
  if (!g_nccl.field) return set_error(CCO_E_NCCL, "libnccl: missing symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(Broadcast, "ncclBroadcast")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.h = h;
  return CCO_OK;
}
constexpr int kNcclInt32 = 2, kNcclUint32 = 3, kNcclSum = 0;  // ncclInt32, ncclUint32, ncclSum (nccl.h enum values)

}  // namespace cco

using namespace cco;

// ------------------------------------------------------------------------------------------------
// context / result objects
// ------------------------------------------------------------------------------------------------
struct PinnedBuf {
  void *p;
  size_t cap;
  bool used;
};

struct cco_ctx {
  int device = 0, rank = 0, world = 1;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t tev[2] = {};
  cudaEvent_t copy_ev[2] = {};
  cudaStream_t bin_stream[8] = {};
  cudaEvent_t bin_ev[9] = {};
  std::vector<PinnedBuf> pinned;
  std::mutex mu;
  ncclComm_t comm = nullptr;
  int launches = 0;
  // mailbox for small device -> host results (mapped pinned memory written by k_mail_bytes)
  unsigned char *mail_h = nullptr, *mail_d = nullptr;
  size_t mail_used = 0;
  struct MailItem { void *dst; size_t off, n; };
  std::vector<MailItem> mail_pending;

  void *pinned_get(size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (size_t i = 0; i < pinned.size(); ++i)
      if (!pinned[i].used && pinned[i].cap >= bytes && (best < 0 || pinned[i].cap < pinned[best].cap)) best = (int)i;
    if (best >= 0) {
      pinned[best].used = true;
      return pinned[best].p;
    }
    void *p = nullptr;
    size_t cap = (bytes + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    pinned.push_back({p, cap, true});
    return p;
  }
  void pinned_put(void *p) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto &b : pinned)
      if (b.p == p) b.used = false;
  }
};

struct ResultMat {
  int64_t row_begin = 0, row_end = 0;
  int32_t n_cols = 0;
  int64_t *row_ptr = nullptr;
  int32_t *col = nullptr;
  double *llr = nullptr;
  int32_t *cnt = nullptr;
};
struct cco_dataset {
  cco_ctx *ctx = nullptr;
  int n_mats = 0;
  long long n_users = 0;
  std::vector<long long> n_cols, nnz;
  std::vector<long long *> rp;   // device, int64 [n_users+1]
  std::vector<int32_t *> col;    // device
  std::vector<cudaEvent_t> ready;  // per matrix: host->device copy finished (copy stream)
  bool h2d_pending = false;        // uploaded asynchronously: ms_h2d is read when the train joins
  float ms_h2d = 0;
};

struct cco_result {
  cco_ctx *ctx = nullptr;
  std::vector<ResultMat> mats;
  cco_stats_t stats;
};

namespace cco {

// per-call device arena on top of the stream-ordered allocator
struct Arena {
  cudaStream_t s;
  std::vector<void *> ptrs;
  explicit Arena(cudaStream_t st) : s(st) {}
  ~Arena() {
    for (void *p : ptrs) cudaFreeAsync(p, s);
  }
  template <typename T>
  int alloc(T **out, size_t n) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(n * sizeof(T), 16);
    cudaError_t e = cudaMallocAsync(&p, bytes, s);
    if (e != cudaSuccess) return set_error(CCO_E_OOM, "cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(p);
    *out = (T *)p;
    return CCO_OK;
  }
  void release(void *p) {
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == p) {
        cudaFreeAsync(p, s);
        ptrs.erase(ptrs.begin() + i);
        return;
      }
  }
};

static inline int grid_for(long long work_items, int block, int sm_count, int waves = 8) {
  long long g = (work_items + block - 1) / block;
  long long cap = (long long)sm_count * waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

constexpr size_t kMailBytes = 1 << 16;
// enqueue "copy n bytes from device to *dst_host"; the value is there after mail_wait()
static int mail_fetch(cco_ctx *c, void *dst_host, const void *src_dev, size_t n) {
  size_t off = (c->mail_used + 7) & ~(size_t)7;
  if (off + n > kMailBytes) return set_error(CCO_E_CUDA, "internal: mailbox overflow");
  k_mail_bytes<<<1, 128, 0, c->stream>>>(c->mail_d + off, (const unsigned char *)src_dev, (int)n);
  c->mail_pending.push_back({dst_host, off, n});
  c->mail_used = off + n;
  return CCO_OK;
}
static int mail_wait(cco_ctx *c) {
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  for (auto &m : c->mail_pending) memcpy(m.dst, c->mail_h + m.off, m.n);
  c->mail_pending.clear();
  c->mail_used = 0;
  return CCO_OK;
}

struct DevRaw {  // a matrix as uploaded (int64 row_ptr like the host)
  long long n_rows = 0;
  int32_t n_cols = 0;
  long long nnz = 0;
  long long *rp = nullptr;
  int32_t *col = nullptr;
};
struct DevMat {  // after canonicalise + downsample
  long long n_rows = 0;
  int32_t n_cols = 0;
  uint32_t *rp = nullptr;  // [n_rows+1]
  int32_t *col = nullptr;
  int32_t *marg = nullptr;  // post-sample column counts
};

static int exclusive_sum_u32(cco_ctx *c, Arena &ar, const uint32_t *in, uint32_t *out, long long n) {
  size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, c->stream));
  void *tmp;
  CKR(ar.alloc((char **)&tmp, tb));
  CK(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, c->stream));
  ar.release(tmp);
  return CCO_OK;
}
static int exclusive_sum_i64(cco_ctx *c, Arena &ar, const long long *in, long long *out, long long n) {
  size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, c->stream));
  void *tmp;
  CKR(ar.alloc((char **)&tmp, tb));
  CK(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, c->stream));
  ar.release(tmp);
  return CCO_OK;
}

// canonicalisation slow path: sort (row,col) keys, drop duplicates, rebuild row_ptr
static int canonicalize_device(cco_ctx *c, Arena &ar, DevRaw &m) {
  if (m.nnz == 0) return CCO_OK;
  unsigned long long *k0, *k1;
  CKR(ar.alloc(&k0, m.nnz));
  CKR(ar.alloc(&k1, m.nnz));
  k_expand_keys<<<grid_for(m.n_rows * kSG, 256, c->sm_count), 256, 0, c->stream>>>(m.n_rows, m.rp, m.col, k0);
  c->launches++;
  int row_bits = 1;
  while ((1LL << row_bits) < m.n_rows) ++row_bits;
  cub::DoubleBuffer<unsigned long long> db(k0, k1);
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, m.nnz, 0, 32 + row_bits, c->stream));
  void *tmp;
  CKR(ar.alloc((char **)&tmp, tb));
  CK(cub::DeviceRadixSort::SortKeys(tmp, tb, db, m.nnz, 0, 32 + row_bits, c->stream));
  ar.release(tmp);
  unsigned long long *sorted = db.Current(), *other = db.Alternate();
  uint32_t *flag, *pos;
  CKR(ar.alloc(&flag, m.nnz + 1));
  CKR(ar.alloc(&pos, m.nnz + 1));
  CK(cudaMemsetAsync(flag + m.nnz, 0, 4, c->stream));
  k_unique_flags<<<grid_for(m.nnz, 256, c->sm_count), 256, 0, c->stream>>>(m.nnz, sorted, flag);
  c->launches++;
  CKR(exclusive_sum_u32(c, ar, flag, pos, m.nnz + 1));
  uint32_t n_unique = 0;
  CK(cudaMemcpyAsync(&n_unique, pos + m.nnz, 4, cudaMemcpyDeviceToHost, c->stream));
  k_unique_scatter<<<grid_for(m.nnz, 256, c->sm_count), 256, 0, c->stream>>>(m.nnz, sorted, flag, pos, other, m.col);
  CK(cudaStreamSynchronize(c->stream));
  k_rowptr_from_keys<<<grid_for(m.n_rows + 1, 256, c->sm_count), 256, 0, c->stream>>>(m.n_rows, n_unique, other, m.rp);
  c->launches += 2;
  m.nnz = n_unique;
  CK(cudaGetLastError());
  ar.release(flag);
  ar.release(pos);
  ar.release(k0);
  ar.release(k1);
  return CCO_OK;
}

// sampleDownAndBinarize of one uploaded matrix (raw column counts already final in raw_counts)
static int downsample_device(cco_ctx *c, Arena &ar, const DevRaw &raw, const int32_t *raw_counts, int32_t m,
                             int32_t seed, uint32_t flags, DevMat *out) {
  out->n_rows = raw.n_rows;
  out->n_cols = raw.n_cols;
  uint32_t *kept;
  CKR(ar.alloc(&kept, raw.n_rows + 1));
  CKR(ar.alloc(&out->rp, raw.n_rows + 1));
  CKR(ar.alloc(&out->marg, std::max<int32_t>(raw.n_cols, 1)));
  CKR(ar.alloc(&out->col, std::max<long long>(raw.nnz, 1)));
  CK(cudaMemsetAsync(out->marg, 0, sizeof(int32_t) * std::max<int32_t>(raw.n_cols, 1), c->stream));
  CK(cudaMemsetAsync(kept + raw.n_rows, 0, 4, c->stream));
  int g = grid_for(raw.n_rows * kSG, 256, c->sm_count);
  k_downsample_count<<<g, 256, 0, c->stream>>>(0, raw.n_rows, raw.rp, raw.col, raw_counts, m, seed, flags, kept, out->marg);
  CKR(exclusive_sum_u32(c, ar, kept, out->rp, raw.n_rows + 1));
  k_downsample_write<<<g, 256, 0, c->stream>>>(0, raw.n_rows, raw.rp, raw.col, raw_counts, m, seed, flags, out->rp, out->col);
  c->launches += 2;
  CK(cudaGetLastError());
  ar.release(kept);
  return CCO_OK;
}

// Multi-GPU form of sampleDownAndBinarize: rank r samples only its block of users, the per-row kept counts are
// all-gathered (so every rank derives the same row_ptr), each rank writes its block of the compacted column array at
// its global offset and the blocks are exchanged with grouped broadcasts over NVLink.  The post-sample column
// marginals are then a local histogram of the gathered matrix.  All matrices go through phase 1 before the single
// host synchronisation (block offsets), then ONE NCCL group moves every block of every matrix.
static int downsample_sharded_all(cco_ctx *c, Arena &ar, const std::vector<DevRaw> &raw, const int32_t *raw_counts,
                                  const std::vector<long long> &col_off, const cco_indicator_params_t *params, int32_t seed,
                                  uint32_t flags, std::vector<DevMat> &dm) {
  cudaStream_t s = c->stream;
  const int W = c->world, r = c->rank, n_mats = (int)raw.size();
  const long long U = raw[0].n_rows;
  const long long S = (U + W - 1) / W;
  const long long u_lo = std::min<long long>((long long)r * S, U), u_hi = std::min<long long>(u_lo + S, U);
  const int g = grid_for(std::max<long long>(u_hi - u_lo, 1) * kSG, 256, c->sm_count);
  std::vector<std::vector<uint32_t>> offs(n_mats, std::vector<uint32_t>((size_t)W + 1, 0));
  for (int i = 0; i < n_mats; ++i) {
    DevMat *out = &dm[i];
    out->n_rows = U;
    out->n_cols = raw[i].n_cols;
    uint32_t *kept;
    CKR(ar.alloc(&kept, (size_t)(W * S + 1)));
    CKR(ar.alloc(&out->rp, U + 1));
    CKR(ar.alloc(&out->marg, std::max<int32_t>(raw[i].n_cols, 1)));
    CKR(ar.alloc(&out->col, std::max<long long>(raw[i].nnz, 1)));
    CK(cudaMemsetAsync(out->marg, 0, sizeof(int32_t) * std::max<int32_t>(raw[i].n_cols, 1), s));
    CK(cudaMemsetAsync(kept, 0, sizeof(uint32_t) * (size_t)(W * S + 1), s));
    const int32_t m = params[i].max_interactions;
    const int32_t *rc_i = raw_counts + col_off[i];
    if (u_hi > u_lo) {
      k_downsample_count<<<g, 256, 0, s>>>(u_lo, u_hi, raw[i].rp, raw[i].col, rc_i, m, seed, flags, kept, nullptr);
      c->launches++;
    }
    int rc = g_nccl.AllGather(kept + (size_t)r * S, kept, (size_t)S, kNcclUint32, c->comm, s);
    if (rc != 0) return set_error(CCO_E_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(rc));
    CKR(exclusive_sum_u32(c, ar, kept, out->rp, U + 1));
    if (u_hi > u_lo) {
      k_downsample_write<<<g, 256, 0, s>>>(u_lo, u_hi, raw[i].rp, raw[i].col, rc_i, m, seed, flags, out->rp, out->col);
      c->launches++;
    }
    for (int q = 0; q <= W; ++q) CKR(mail_fetch(c, &offs[i][q], out->rp + std::min<long long>((long long)q * S, U), 4));
    ar.release(kept);
  }
  CKR(mail_wait(c));
  g_nccl.GroupStart();
  for (int i = 0; i < n_mats; ++i)
    for (int q = 0; q < W; ++q) {
      const size_t cnt = offs[i][q + 1] - offs[i][q];
      if (cnt == 0) continue;
      int rc = g_nccl.Broadcast(dm[i].col + offs[i][q], dm[i].col + offs[i][q], cnt, kNcclInt32, q, c->comm, s);
      if (rc != 0) {
        g_nccl.GroupEnd();
        return set_error(CCO_E_NCCL, "ncclBroadcast: %s", g_nccl.GetErrorString(rc));
      }
    }
  int rc = g_nccl.GroupEnd();
  if (rc != 0) return set_error(CCO_E_NCCL, "ncclGroupEnd: %s", g_nccl.GetErrorString(rc));
  for (int i = 0; i < n_mats; ++i)
    if (offs[i][W] > 0) {
      k_col_histogram_u32<<<grid_for(offs[i][W], 256, c->sm_count), 256, 0, s>>>(dm[i].rp, dm[i].rp + U, dm[i].col, dm[i].marg);
      c->launches++;
    }
  CK(cudaGetLastError());
  return CCO_OK;
}

// ---- row-kernel configurations -----------------------------------------------------------------

struct BinCfg {
  int group;    // threads that own one row: 32 (warp), 256 or 1024 (whole CTA)
  int slots;    // table words per group
  int cap;      // distinct keys a hashed table may hold per pass
  int cbuf;     // candidate buffer entries per group
  int caux, keep_max, final_max;
  bool dense;
  size_t region;  // shared-memory bytes per group
  size_t smem;    // per CTA
  int ctas_per_sm;
};

static int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

template <int GROUP>
static int launch_rows_t(cco_ctx *c, const RowArgs &a, BinCfg &cfg, cudaStream_t st) {
  constexpr int CTA = GROUP == 32 ? 64 : GROUP;   // warp-owned rows: two independent warps per CTA (fine-grained smem packing)
  int occ = 1;
  void (*kern)(const RowArgs) = cfg.dense ? k_rows<GROUP, true> : k_rows<GROUP, false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA, cfg.smem));
  // 4 waves of CTAs over the work-sorted row list: a CTA that draws cheap rows retires early and the hardware
  // scheduler backfills, which balances the tail better than one persistent wave (tools/tune_rows.py: -6 % at C3)
  kern<<<c->sm_count * std::max(occ, 1) * 4, CTA, cfg.smem, st>>>(a);
  cfg.ctas_per_sm = occ;
  c->launches++;
  CK(cudaGetLastError());
  return CCO_OK;
}
static int launch_rows(cco_ctx *c, const RowArgs &a, BinCfg &cfg, cudaStream_t st) {
  switch (cfg.group) {
    case 1024: return launch_rows_t<1024>(c, a, cfg, st);
    case 512: return launch_rows_t<512>(c, a, cfg, st);
    case 256: return launch_rows_t<256>(c, a, cfg, st);
    case 128: return launch_rows_t<128>(c, a, cfg, st);
    case 32: return launch_rows_t<32>(c, a, cfg, st);
  }
  return set_error(CCO_E_INVALID_ARG, "internal: bad bin config");
}

static BinCfg make_cfg(cco_ctx *c, int group, int want_slots, int top_k, int n_cols_b) {
  BinCfg f;
  const int groups = group == 32 ? 2 : 1;
  f.group = group;
  f.final_max = next_pow2(top_k);
  f.cbuf = next_pow2(top_k + std::max(group, 128) + (group == 32 ? 64 : 0));
  if (group == 32 && top_k + 32 <= 96) f.cbuf = 128;  // small top_k: a 128-entry buffer doubles the warps per SM (-4 % at C3)
  f.keep_max = std::max(f.final_max, (f.cbuf - group) / 2);
  f.caux = group == 32 ? 0 : f.keep_max;
  size_t fixed = (size_t)(f.cbuf + f.caux) * 16 + 2 * 256 + 512 + 1024 + (size_t)(group / 32) * 256 + 2048;  // candidates, x12/x11 tables, ctrl, histogram, queues, level-1 cut histogram
  size_t avail = (c->smem_optin - 1024) / groups;  // slack for static shared memory
  int max_slots = (int)((avail - fixed) / 4) & ~1023;
  f.slots = std::min(want_slots, max_slots);
  f.cap = f.slots / 2;  // load factor <= 1/2: 2/3 and 3/4 are 8 % and 17 % slower (probe chains), tools/tune_rows.py
  f.dense = n_cols_b <= f.slots;
  f.region = (fixed + (size_t)f.slots * 4 + 15) & ~(size_t)15;
  f.smem = f.region * groups;
  f.ctas_per_sm = 1;
  return f;
}

struct IndicatorOut {
  int64_t row_begin = 0, row_end = 0;
  int64_t nnz = 0;
  int64_t products = 0, distinct = 0, evaluated = 0;
};

// One indicator: rows [row_lo,row_hi) of A'^T B'.  Leaves the packed result in pinned host memory.
static int run_indicator(cco_ctx *c, Arena &ar, const uint32_t *at_ptr, const int32_t *at_users, int32_t n_items_a,
                         const int32_t *marg_a, int32_t max_marg_a, const DevMat &B, long long n_users, bool self,
                         const cco_indicator_params_t &prm, uint32_t flags, bool emit_all, int rank, int world,
                         ResultMat *rm, IndicatorOut *io, float *ms_rows) {
  cudaStream_t s = c->stream;
  const int32_t n_cols_b = B.n_cols;
  // 1. work per output row + schedule ------------------------------------------------------------
  uint32_t *row_work, *sorted_work;
  unsigned long long *work64;
  long long *work_prefix;
  int32_t *ids, *rows_sorted;
  CKR(ar.alloc(&row_work, n_items_a + 1));
  CKR(ar.alloc(&work64, n_items_a + 1));
  CKR(ar.alloc(&work_prefix, n_items_a + 1));
  CKR(ar.alloc(&ids, n_items_a + 1));
  CK(cudaMemsetAsync(work64 + n_items_a, 0, 8, s));
  k_row_work<<<grid_for((long long)n_items_a * kSG, 256, c->sm_count), 256, 0, s>>>(n_items_a, at_ptr, at_users, B.rp,
                                                                                 row_work, work64, ids);
  c->launches++;
  CKR(exclusive_sum_i64(c, ar, (const long long *)work64, work_prefix, (long long)n_items_a + 1));
  // rank partition: contiguous item ranges balanced by work prefix (identical on every rank)
  int32_t row_lo = 0, row_hi = n_items_a;
  if (world > 1) {
    int32_t *d_pb;
    std::vector<int32_t> bounds((size_t)world + 1);
    CKR(ar.alloc(&d_pb, world + 1));
    k_partition_rows<<<1, 64, 0, s>>>(work_prefix, n_items_a, world, d_pb);   // world <= 63 ranks per job
    c->launches++;
    CKR(mail_fetch(c, bounds.data(), d_pb, sizeof(int32_t) * ((size_t)world + 1)));
    CKR(mail_wait(c));
    row_lo = bounds[rank];
    row_hi = bounds[rank + 1];
  }
  const int32_t n_my = row_hi - row_lo;
  io->row_begin = row_lo;
  io->row_end = row_hi;
  long long hp2[2] = {0, 0};   // filled by the mailbox before the final sync of this indicator
  CKR(mail_fetch(c, &hp2[0], work_prefix + row_lo, 8));
  CKR(mail_fetch(c, &hp2[1], work_prefix + row_hi, 8));
  CKR(ar.alloc(&sorted_work, n_my + 1));
  CKR(ar.alloc(&rows_sorted, n_my + 1));
  if (n_my > 0) {
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, row_work + row_lo, sorted_work, ids + row_lo, rows_sorted,
                                                 n_my, 0, 32, s));
    void *tmp;
    CKR(ar.alloc((char **)&tmp, tb));
    CK(cub::DeviceRadixSort::SortPairsDescending(tmp, tb, row_work + row_lo, sorted_work, ids + row_lo, rows_sorted, n_my,
                                                 0, 32, s));
    ar.release(tmp);
  }
  // 2. bins -----------------------------------------------------------------------------------------
  const int k_eff = emit_all ? 1 : prm.top_k;
  const bool warp_ok = k_eff + 32 <= 256;  // warp-owned rows keep a 256-entry candidate buffer
  // Work bins, largest rows first.  {threads that own a row, table words, largest row work w the bin takes}.
  // Bin 0 is the multi-pass bin (same config as bin 1).  Rows up to 1024 products are WARP-owned: no CTA barrier
  // anywhere in their count / compact / score / select pipeline (profiles/r01_k_rows_ncu_full.md: barriers cost the
  // CTA-owned bins 35-45 % of their warp time); larger rows need the table and the parallelism of a whole CTA.
  struct BinSpec { int group, slots; uint32_t max_w; };
  std::vector<BinSpec> spec = {{1024, 1 << 20, 0xffffffffu}, {1024, 1 << 20, 0xffffffffu}, {512, 16384, 8192u}, {256, 8192, 4096u}};
  if (warp_ok) {
    // rows of 1025..2048 products: a 128-thread CTA shares one 4096-word table (9 CTAs/SM) -- a warp-owned 4096-word
    // table leaves only 10 warps per SM (tools/tune_rows.py: -4 % at C3); up to 1024 products rows are warp-owned
    spec.push_back({128, 4096, 2048u});
    spec.push_back({32, 2048, 1024u});
    spec.push_back({32, 1024, 512u});
    spec.push_back({32, 512, 256u});
  } else {
    spec.push_back({128, 4096, 2048u});
  }
  const int kBins = (int)spec.size();
  std::vector<BinCfg> cfgs(kBins);
  for (int b = 0; b < kBins; ++b) cfgs[b] = make_cfg(c, spec[b].group, spec[b].slots, k_eff, n_cols_b);
  BinCfg &cfgL = cfgs[1];
  // packed word: key bits must leave room for the largest possible count (= users of the item)
  int key_bits = 1;
  while (((1LL << key_bits) - 1) <= (long long)n_cols_b) ++key_bits;  // keys <= 2^kb - 2
  int count_bits = 32 - key_bits;
  if (count_bits < 1 || (long long)max_marg_a >= (1LL << count_bits))
    return set_error(CCO_E_UNSUPPORTED,
                     "an item with %d users and %d columns does not fit the packed 32-bit accumulator word "
                     "(key %d bits + count %d bits)", max_marg_a, n_cols_b, key_bits, count_bits);
  // thresholds on w, descending: bin b takes rows with h_thr[b-1] >= w > h_thr[b]; a hashed table also needs w <= cap
  std::vector<uint32_t> h_thr(kBins);
  for (int b = 0; b < kBins; ++b) {
    const BinCfg &f = cfgs[std::min(b + 1, kBins - 1)];   // h_thr[b] = upper limit of bin b+1
    uint32_t lim = b + 1 < kBins ? spec[b + 1].max_w : 0u;
    if (b + 1 < kBins && !f.dense) lim = std::min<uint32_t>(lim, (uint32_t)f.cap);
    h_thr[b] = lim;
    if (b > 0) h_thr[b] = std::min(h_thr[b], h_thr[b - 1]);
  }
  int32_t *d_bounds;
  CKR(ar.alloc(&d_bounds, kBins + 3));
  BinThresholds bt;
  memset(&bt, 0, sizeof bt);
  for (int b = 0; b < kBins; ++b) bt.t[b] = h_thr[b];
  k_bin_bounds<<<1, 32, 0, s>>>(n_my, sorted_work, kBins, bt, d_bounds);
  c->launches++;
  // per-column constants of B' for the fused LLR
  ColTerm *col_terms;
  CKR(ar.alloc(&col_terms, std::max<int32_t>(n_cols_b, 1)));
  if (n_cols_b > 0) {
    k_col_terms<<<grid_for(n_cols_b, 256, c->sm_count, 4), 256, 0, s>>>(n_cols_b, B.marg, n_users, flags, col_terms);
    c->launches++;
  }
  // 3. outputs ----------------------------------------------------------------------------------------
  int32_t stride = emit_all ? n_cols_b : std::min<int32_t>(prm.top_k, n_cols_b);
  if (stride < 1) stride = 1;
  int32_t *o_col, *o_cnt, *o_len;
  double *o_llr = nullptr;
  unsigned long long *d_distinct;
  int *d_err;
  size_t cells = (size_t)std::max(n_items_a, 1) * stride;
  CKR(ar.alloc(&o_col, cells));
  CKR(ar.alloc(&o_cnt, cells));
  if (!emit_all) CKR(ar.alloc(&o_llr, cells));
  CKR(ar.alloc(&o_len, n_items_a + 1));
  CKR(ar.alloc(&d_distinct, 2));
  CKR(ar.alloc(&d_err, 1));
  CK(cudaMemsetAsync(o_len, 0, sizeof(int32_t) * ((size_t)n_items_a + 1), s));
  CK(cudaMemsetAsync(d_distinct, 0, 16, s));
  CK(cudaMemsetAsync(d_err, 0, 4, s));
  RowArgs a;
  memset(&a, 0, sizeof a);
  a.at_ptr = at_ptr;
  a.at_users = at_users;
  a.b_ptr = B.rp;
  a.b_col = B.col;
  a.marg_a = marg_a;
  a.marg_b = B.marg;
  a.col_terms = col_terms;
  a.rows_sorted = rows_sorted;
  a.row_work = row_work;
  a.bin_bounds = d_bounds;
  a.n_cols_b = n_cols_b;
  a.n_users = n_users;
  a.self = self ? 1 : 0;
  a.top_k = k_eff;
  a.has_min_llr = prm.has_min_llr;
  a.min_llr = prm.min_llr;
  a.flags = flags;
  a.count_bits = count_bits;
  a.out_stride = stride;
  a.out_col = o_col;
  a.out_llr = o_llr;
  a.out_cnt = o_cnt;
  a.out_len = o_len;
  a.stat_distinct = d_distinct;
  a.stat_evaluated = d_distinct + 1;
  a.err_flag = d_err;
  a.emit_all = emit_all ? 1 : 0;
  CK(cudaEventRecord(c->ev[4], s));
  if (n_my > 0) {
    // the bins touch disjoint rows: run them concurrently (tails of one bin overlap the bulk of another)
    CK(cudaEventRecord(c->bin_ev[8], s));
    for (int b = 0; b < kBins; ++b) {
      if (b == 0 && cfgL.dense) continue;                 // dense L takes every large row in bin 1
      RowArgs ab = a;
      ab.bin = b;
      ab.slots = cfgs[b].slots;
      ab.cap = cfgs[b].cap;
      ab.tsize_x16 = 32;
      ab.cbuf = cfgs[b].cbuf;
      ab.caux = cfgs[b].caux;
      ab.keep_max = cfgs[b].keep_max;
      ab.final_max = cfgs[b].final_max;
      ab.group_smem_bytes = (int32_t)cfgs[b].region;
      CK(cudaStreamWaitEvent(c->bin_stream[b], c->bin_ev[8], 0));
      CKR(launch_rows(c, ab, cfgs[b], c->bin_stream[b]));
      CK(cudaEventRecord(c->bin_ev[b], c->bin_stream[b]));
      CK(cudaStreamWaitEvent(s, c->bin_ev[b], 0));
    }
  }
  CK(cudaEventRecord(c->ev[5], s));
  // 4. pack + copy back ---------------------------------------------------------------------------------
  long long *len64, *out_ptr;
  CKR(ar.alloc(&len64, n_my + 1));
  CKR(ar.alloc(&out_ptr, n_my + 1));
  CK(cudaMemsetAsync(len64 + n_my, 0, 8, s));
  if (n_my > 0) {
    k_len_to_i64<<<grid_for(n_my, 256, c->sm_count), 256, 0, s>>>(row_lo, n_my, o_len, len64);
    c->launches++;
  }
  CKR(exclusive_sum_i64(c, ar, len64, out_ptr, (long long)n_my + 1));
  long long total = 0;
  unsigned long long h_distinct[2] = {0, 0};
  int h_err = 0;
  CKR(mail_fetch(c, &total, out_ptr + n_my, 8));
  CKR(mail_fetch(c, h_distinct, d_distinct, 16));
  CKR(mail_fetch(c, &h_err, d_err, 4));
  CKR(mail_wait(c));
  io->products = hp2[1] - hp2[0];
  if (h_err) return set_error(CCO_E_CUDA, "internal: shared-memory hash table overflow");
  io->nnz = total;
  io->distinct = (int64_t)h_distinct[0];
  io->evaluated = (int64_t)h_distinct[1];
  int32_t *p_col, *p_cnt;
  double *p_llr = nullptr;
  CKR(ar.alloc(&p_col, std::max<long long>(total, 1)));
  CKR(ar.alloc(&p_cnt, std::max<long long>(total, 1)));
  if (!emit_all) CKR(ar.alloc(&p_llr, std::max<long long>(total, 1)));
  if (n_my > 0 && total > 0) {
    k_compact_rows<<<grid_for((long long)n_my * 32, 256, c->sm_count), 256, 0, s>>>(row_lo, n_my, stride, out_ptr, o_len,
                                                                                 o_col, o_llr, o_cnt, p_col, p_llr, p_cnt);
    c->launches++;
  }
  rm->row_begin = row_lo;
  rm->row_end = row_hi;
  rm->n_cols = n_cols_b;
  rm->row_ptr = (int64_t *)c->pinned_get(sizeof(int64_t) * ((size_t)n_my + 1));
  rm->col = (int32_t *)c->pinned_get(sizeof(int32_t) * (size_t)std::max<long long>(total, 1));
  rm->cnt = (int32_t *)c->pinned_get(sizeof(int32_t) * (size_t)std::max<long long>(total, 1));
  rm->llr = (double *)c->pinned_get(sizeof(double) * (size_t)std::max<long long>(total, 1));
  if (!rm->row_ptr || !rm->col || !rm->cnt || !rm->llr) return set_error(CCO_E_OOM, "pinned host allocation failed");
  // device -> host on the copy stream so the transfer overlaps the next indicator's kernels; train_dataset joins the
  // copy stream before it returns (and before the arena frees the packed buffers)
  cudaStream_t cs = c->copy_stream;
  CK(cudaEventRecord(c->copy_ev[0], s));
  CK(cudaStreamWaitEvent(cs, c->copy_ev[0], 0));
  CK(cudaMemcpyAsync(rm->row_ptr, out_ptr, sizeof(int64_t) * ((size_t)n_my + 1), cudaMemcpyDeviceToHost, cs));
  if (total > 0 && !(flags & CCO_FLAG_RESULT_ON_DEVICE)) {
    CK(cudaMemcpyAsync(rm->col, p_col, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, cs));
    CK(cudaMemcpyAsync(rm->cnt, p_cnt, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, cs));
    if (!emit_all) CK(cudaMemcpyAsync(rm->llr, p_llr, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, cs));
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]));
  if (ms_rows) *ms_rows = ms;
  // the strided buffers are dead once k_compact_rows has been enqueued (stream order); the packed ones stay until the
  // copy stream is joined
  for (void *p : {(void *)o_col, (void *)o_cnt, (void *)o_llr})
    if (p) ar.release(p);
  return CCO_OK;
}

static int validate_host(int32_t n_mats, const cco_csr_t *mats, const cco_indicator_params_t *params) {
  if (n_mats < 1 || !mats || !params) return set_error(CCO_E_INVALID_ARG, "need at least the primary matrix and its params");
  for (int i = 0; i < n_mats; ++i) {
    const cco_csr_t &m = mats[i];
    if (!m.row_ptr) return set_error(CCO_E_INVALID_ARG, "matrix %d: null row_ptr", i);
    if (m.n_rows < 0 || m.n_rows >= 0x7fffffffLL) return set_error(CCO_E_INVALID_ARG, "matrix %d: n_rows out of range", i);
    if (m.n_cols < 0 || m.n_cols >= 0x7ffffffe) return set_error(CCO_E_INVALID_ARG, "matrix %d: n_cols out of range", i);
    if (m.n_rows != mats[0].n_rows)
      return set_error(CCO_E_SHAPE_MISMATCH, "matrix %d has %lld rows, the primary has %lld: all event types share the user dictionary",
                       i, (long long)m.n_rows, (long long)mats[0].n_rows);
    if (m.row_ptr[0] != 0) return set_error(CCO_E_INVALID_ARG, "matrix %d: row_ptr[0] != 0", i);
    long long nnz = m.row_ptr[m.n_rows];
    if (nnz < 0 || nnz >= 0xffffffffLL) return set_error(CCO_E_UNSUPPORTED, "matrix %d: nnz %lld outside [0, 2^32)", i, nnz);
    if (nnz > 0 && !m.col_idx) return set_error(CCO_E_INVALID_ARG, "matrix %d: null col_idx", i);
    if (params[i].max_interactions < 1) return set_error(CCO_E_INVALID_ARG, "matrix %d: max_interactions must be >= 1", i);
    if (params[i].top_k < 1) return set_error(CCO_E_INVALID_ARG, "matrix %d: top_k must be >= 1", i);
    if (params[i].top_k > CCO_MAX_TOP_K)
      return set_error(CCO_E_UNSUPPORTED, "matrix %d: top_k %d > CCO_MAX_TOP_K (%d)", i, params[i].top_k, CCO_MAX_TOP_K);
    if (params[i].has_min_llr && params[i].min_llr != params[i].min_llr)
      return set_error(CCO_E_INVALID_ARG, "matrix %d: min_llr is NaN", i);
  }
  return CCO_OK;
}

// host CSR -> device (the dataset owns its buffers; they live until cco_dataset_free)
static void dataset_release(cco_dataset *d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  for (auto p : d->rp)
    if (p) cudaFreeAsync(p, d->ctx->stream);
  for (auto p : d->col)
    if (p) cudaFreeAsync(p, d->ctx->stream);
  for (auto e : d->ready)
    if (e) cudaEventDestroy(e);
  delete d;
}

static int dataset_upload(cco_ctx *c, int32_t n_mats, const cco_csr_t *mats, uint32_t flags, cco_dataset **out, bool async = false) {
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  cco_dataset *d = new cco_dataset();
  d->ctx = c;
  d->n_mats = n_mats;
  d->n_users = mats[0].n_rows;
  d->rp.assign(n_mats, nullptr);
  d->col.assign(n_mats, nullptr);
  d->n_cols.assign(n_mats, 0);
  d->nnz.assign(n_mats, 0);
  d->ready.assign(n_mats, nullptr);
  struct G {
    cco_dataset *d;
    bool ok = false;
    ~G() {
      if (!ok) dataset_release(d);
    }
  } g{d};
  // allocations are ordered on the main stream; the copies run on the copy stream (H2D engine) so that the caller
  // of the async form can start preparing matrix i while matrix i+1 is still in flight
  cudaStream_t cs = c->copy_stream;
  for (int i = 0; i < n_mats; ++i) {
    const cco_csr_t &m = mats[i];
    d->n_cols[i] = m.n_cols;
    d->nnz[i] = m.row_ptr[m.n_rows];
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, sizeof(int64_t) * ((size_t)m.n_rows + 1), s);
    if (e != cudaSuccess) return set_error(CCO_E_OOM, "cudaMallocAsync row_ptr: %s", cudaGetErrorString(e));
    d->rp[i] = (long long *)p;
    e = cudaMallocAsync(&p, sizeof(int32_t) * (size_t)std::max<long long>(d->nnz[i], 4), s);
    if (e != cudaSuccess) return set_error(CCO_E_OOM, "cudaMallocAsync col_idx: %s", cudaGetErrorString(e));
    d->col[i] = (int32_t *)p;
    CK(cudaEventCreateWithFlags(&d->ready[i], cudaEventDisableTiming));
  }
  CK(cudaEventRecord(c->copy_ev[0], s));
  CK(cudaStreamWaitEvent(cs, c->copy_ev[0], 0));
  CK(cudaEventRecord(c->ev[6], cs));
  for (int i = 0; i < n_mats; ++i) {
    const cco_csr_t &m = mats[i];
    CK(cudaMemcpyAsync(d->rp[i], m.row_ptr, sizeof(int64_t) * ((size_t)m.n_rows + 1), cudaMemcpyHostToDevice, cs));
    if (d->nnz[i] > 0)
      CK(cudaMemcpyAsync(d->col[i], m.col_idx, sizeof(int32_t) * (size_t)d->nnz[i], cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(d->ready[i], cs));
  }
  CK(cudaEventRecord(c->ev[7], cs));
  if (async && (flags & CCO_FLAG_ASSUME_CANONICAL)) {
    d->h2d_pending = true;
    g.ok = true;
    *out = d;
    return CCO_OK;
  }
  for (int i = 0; i < n_mats; ++i) CK(cudaStreamWaitEvent(s, d->ready[i], 0));
  // check + (if needed) canonicalise in place
  if (!(flags & CCO_FLAG_ASSUME_CANONICAL)) {
    Arena ar(s);
    int *d_flags;
    CKR(ar.alloc(&d_flags, 2 * n_mats));
    CK(cudaMemsetAsync(d_flags, 0, sizeof(int) * 2 * n_mats, s));
    for (int i = 0; i < n_mats; ++i) {
      k_check_rows<<<grid_for(d->n_users * kSG, 256, c->sm_count), 256, 0, s>>>(d->n_users, (int32_t)d->n_cols[i], d->rp[i],
                                                                             d->col[i], d_flags + 2 * i);
      c->launches++;
    }
    std::vector<int> h(2 * n_mats);
    CK(cudaMemcpyAsync(h.data(), d_flags, sizeof(int) * 2 * n_mats, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    for (int i = 0; i < n_mats; ++i)
      if (h[2 * i]) return set_error(CCO_E_INVALID_ARG, "matrix %d: row_ptr not monotone or column index out of [0, n_cols)", i);
    for (int i = 0; i < n_mats; ++i)
      if (h[2 * i + 1]) {
        DevRaw r;
        r.n_rows = d->n_users;
        r.n_cols = (int32_t)d->n_cols[i];
        r.nnz = d->nnz[i];
        r.rp = d->rp[i];
        r.col = d->col[i];
        CKR(canonicalize_device(c, ar, r));
        d->nnz[i] = r.nnz;
      }
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(&d->ms_h2d, c->ev[6], c->ev[7]));
  g.ok = true;
  *out = d;
  return CCO_OK;
}

static int train_dataset(cco_ctx *c, const cco_dataset *ds, const cco_indicator_params_t *params, int32_t seed, uint32_t flags,
                         cco_result **out) {
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const int n_mats = ds->n_mats;
  c->mail_pending.clear();
  c->mail_used = 0;
  Arena ar(s);
  struct CopyJoin {  // destroyed before `ar`: no packed buffer is freed while the copy stream still reads it
    cco_ctx *c;
    ~CopyJoin() { cudaStreamSynchronize(c->copy_stream); }
  } copy_join{c};
  cco_result *res = new cco_result();
  res->ctx = c;
  res->mats.resize(n_mats);
  memset(&res->stats, 0, sizeof res->stats);
  struct Guard {
    cco_result *r;
    bool ok = false;
    ~Guard() {
      if (!ok) cco_result_free(r);
    }
  } guard{res};
  cco_stats_t &st = res->stats;
  st.n_mats = n_mats;
  st.n_users = ds->n_users;
  st.ms_h2d = ds->ms_h2d;
  const bool h2d_pending = ds->h2d_pending;
  const long long n_users = ds->n_users;
  std::vector<DevRaw> raw(n_mats);
  for (int i = 0; i < n_mats; ++i) {
    raw[i].n_rows = n_users;
    raw[i].n_cols = (int32_t)ds->n_cols[i];
    raw[i].nnz = ds->nnz[i];
    raw[i].rp = ds->rp[i];
    raw[i].col = ds->col[i];
    st.nnz_in_total += raw[i].nnz;
  }
  CK(cudaEventRecord(c->ev[1], s));
  // raw column counts: this rank histograms its user slice; ONE allreduce sums all matrices' counts
  long long total_cols = 0;
  std::vector<long long> col_off(n_mats + 1, 0);
  for (int i = 0; i < n_mats; ++i) {
    col_off[i] = total_cols;
    total_cols += raw[i].n_cols;
  }
  col_off[n_mats] = total_cols;
  constexpr int kHistCopies = 16;
  const long long copy_stride = std::max<long long>(total_cols, 1);
  int32_t *raw_counts;
  CKR(ar.alloc(&raw_counts, (size_t)copy_stride * kHistCopies));
  CK(cudaMemsetAsync(raw_counts, 0, sizeof(int32_t) * (size_t)copy_stride * kHistCopies, s));
  const long long u_lo = n_users * c->rank / c->world, u_hi = n_users * (c->rank + 1) / c->world;
  for (int i = 0; i < n_mats; ++i) {
    CK(cudaStreamWaitEvent(s, ds->ready[i], 0));  // matrix i has landed (async upload: later ones may still be in flight)
    if (raw[i].nnz == 0 || u_hi == u_lo) continue;
    k_col_histogram<<<grid_for(raw[i].nnz / c->world + 1, 256, c->sm_count), 256, 0, s>>>(u_lo, u_hi, raw[i].rp, raw[i].col,
                                                                                       raw_counts + col_off[i], kHistCopies, copy_stride);
    c->launches++;
  }
  if (total_cols > 0) {
    k_sum_copies<<<grid_for(total_cols, 256, c->sm_count), 256, 0, s>>>(total_cols, kHistCopies, copy_stride, raw_counts);
    c->launches++;
  }
  if (c->world > 1) {
    int r = g_nccl.AllReduce(raw_counts, raw_counts, (size_t)total_cols, kNcclInt32, kNcclSum, c->comm, s);
    if (r != 0) return set_error(CCO_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
  }
  // sampleDownAndBinarize every matrix
  std::vector<DevMat> dm(n_mats);
  if (c->world > 1) {
    CKR(downsample_sharded_all(c, ar, raw, raw_counts, col_off, params, seed, flags, dm));
  } else {
    for (int i = 0; i < n_mats; ++i)
      CKR(downsample_device(c, ar, raw[i], raw_counts + col_off[i], params[i].max_interactions, seed, flags, &dm[i]));
  }
  // `drmA.t`
  const int32_t n_items_a = dm[0].n_cols;
  uint32_t *at_ptr, *cursor;
  int32_t *at_users, *d_max;
  CKR(ar.alloc(&at_ptr, n_items_a + 1));
  CKR(ar.alloc(&cursor, n_items_a + 1));
  CKR(ar.alloc(&d_max, 1));
  CKR(ar.alloc(&at_users, std::max<long long>(raw[0].nnz, 1)));
  CK(cudaMemsetAsync(d_max, 0, 4, s));
  {
    uint32_t *marg_pad;
    CKR(ar.alloc(&marg_pad, n_items_a + 1));
    CK(cudaMemcpyAsync(marg_pad, dm[0].marg, sizeof(int32_t) * (size_t)n_items_a, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemsetAsync(marg_pad + n_items_a, 0, 4, s));
    CKR(exclusive_sum_u32(c, ar, marg_pad, at_ptr, (long long)n_items_a + 1));
    ar.release(marg_pad);
  }
  CK(cudaMemcpyAsync(cursor, at_ptr, sizeof(uint32_t) * ((size_t)n_items_a + 1), cudaMemcpyDeviceToDevice, s));
  k_transpose_scatter<<<grid_for(n_users * kSG, 256, c->sm_count), 256, 0, s>>>(n_users, dm[0].rp, dm[0].col, cursor, at_users);
  c->launches++;
  if (n_items_a > 0) {
    k_max_i32<<<grid_for(n_items_a, 256, c->sm_count, 2), 256, 0, s>>>(n_items_a, dm[0].marg, d_max);
    c->launches++;
  }
  int32_t max_marg_a = 0;
  std::vector<uint32_t> h_nnz(n_mats);
  CKR(mail_fetch(c, &max_marg_a, d_max, 4));
  for (int i = 0; i < n_mats; ++i) CKR(mail_fetch(c, &h_nnz[i], dm[i].rp + n_users, 4));
  CK(cudaEventRecord(c->ev[2], s));
  CKR(mail_wait(c));
  for (int i = 0; i < n_mats && i < 16; ++i) st.nnz_downsampled[i] = h_nnz[i];

  for (int i = 0; i < n_mats; ++i) {
    IndicatorOut io;
    float ms_rows = 0;
    CKR(run_indicator(c, ar, at_ptr, at_users, n_items_a, dm[0].marg, max_marg_a, dm[i], n_users, i == 0, params[i], flags,
                      false, c->rank, c->world, &res->mats[i], &io, &ms_rows));
    if (i < 16) {
      st.products[i] = io.products;
      st.distinct_cells[i] = io.distinct;
      st.llr_evaluated[i] = io.evaluated;
      st.out_nnz[i] = io.nnz;
      st.ms_indicator[i] = ms_rows;
    }
  }
  CK(cudaEventRecord(c->copy_ev[1], c->copy_stream));
  CK(cudaStreamWaitEvent(s, c->copy_ev[1], 0));
  CK(cudaEventRecord(c->ev[3], s));
  CK(cudaStreamSynchronize(s));
  CK(cudaStreamSynchronize(c->copy_stream));
  if (h2d_pending) CK(cudaEventElapsedTime(&st.ms_h2d, c->ev[6], c->ev[7]));
  CK(cudaEventElapsedTime(&st.ms_prepare, c->ev[1], c->ev[2]));
  CK(cudaEventElapsedTime(&st.ms_cooccurrence, c->ev[2], c->ev[3]));
  CK(cudaEventElapsedTime(&st.ms_total, c->ev[1], c->ev[3]));
  if (!h2d_pending) st.ms_total += st.ms_h2d;  // async upload overlaps the prepare stage: already inside the bracket
  st.n_kernel_launches = c->launches;
  guard.ok = true;
  *out = res;
  return CCO_OK;
}

static int train_impl(cco_ctx *c, int32_t n_mats, const cco_csr_t *mats, const cco_indicator_params_t *params, int32_t seed,
                      uint32_t flags, cco_result **out) {
  CKR(validate_host(n_mats, mats, params));
  c->launches = 0;
  cco_dataset *ds = nullptr;
  CKR(dataset_upload(c, n_mats, mats, flags, &ds, /*async=*/true));
  int rc = train_dataset(c, ds, params, seed, flags, out);
  cudaStreamSynchronize(c->copy_stream);  // the caller's host buffers are free again when cco_train returns
  dataset_release(ds);
  return rc;
}

}  // namespace cco

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int cco_abi_version(void) { return CCO_ABI_VERSION; }
const char *cco_last_error(void) { return g_err; }
const char *cco_status_string(int s) {
  switch (s) {
    case CCO_OK: return "ok";
    case CCO_E_INVALID_ARG: return "invalid argument";
    case CCO_E_CUDA: return "CUDA error";
    case CCO_E_NCCL: return "NCCL error";
    case CCO_E_OOM: return "out of memory";
    case CCO_E_SHAPE_MISMATCH: return "shape mismatch";
    case CCO_E_UNSUPPORTED: return "unsupported";
  }
  return "unknown";
}

int cco_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return set_error(CCO_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

int cco_nccl_unique_id(unsigned char out[128]) {
  if (!out) return set_error(CCO_E_INVALID_ARG, "null output");
  CKR(load_nccl());
  ncclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return set_error(CCO_E_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memcpy(out, id.internal, 128);
  return CCO_OK;
}

int cco_create(const cco_config_t *cfg, cco_ctx_t **out) {
  if (!cfg || !out) return set_error(CCO_E_INVALID_ARG, "null argument");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return set_error(CCO_E_INVALID_ARG, "bad rank/world_size %d/%d", cfg->rank, cfg->world_size);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_error(CCO_E_CUDA, "no CUDA device (%s): this library has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (cfg->device < 0 || cfg->device >= n) return set_error(CCO_E_INVALID_ARG, "device %d not in [0,%d)", cfg->device, n);
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, cfg->device));
  if (p.major != 10)
    return set_error(CCO_E_CUDA, "device %d is sm_%d%d; this build contains sm_100a code only", cfg->device, p.major, p.minor);
  CK(cudaSetDevice(cfg->device));
  if (cfg->world_size > 1 && !cfg->nccl_unique_id) return set_error(CCO_E_INVALID_ARG, "world_size > 1 needs nccl_unique_id");
  cco_ctx *c = new cco_ctx();
  c->device = cfg->device;
  c->rank = cfg->rank;
  c->world = cfg->world_size;
  c->sm_count = p.multiProcessorCount;
  c->smem_optin = p.sharedMemPerBlockOptin;
  // every failure below releases what was created so far (cco_destroy tolerates a half-built context)
  auto init = [&]() -> int {
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto &ev : c->ev) CK(cudaEventCreate(&ev));
    for (auto &ev : c->tev) CK(cudaEventCreate(&ev));
    for (auto &ev : c->copy_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &st : c->bin_stream) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &ev : c->bin_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaHostAlloc((void **)&c->mail_h, kMailBytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->mail_d, c->mail_h, 0));
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
    uint64_t thr = UINT64_MAX;  // keep freed blocks: steady-state trains allocate nothing
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    if (c->world > 1) {
      CKR(load_nccl());
      ncclUniqueId id;
      memcpy(id.internal, cfg->nccl_unique_id, 128);
      int rc = g_nccl.CommInitRank(&c->comm, c->world, id, c->rank);
      if (rc != 0) return set_error(CCO_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
    }
    return CCO_OK;
  };
  const int st = init();
  if (st != CCO_OK) {
    char keep[sizeof g_err];
    memcpy(keep, g_err, sizeof keep);   // cco_destroy must not clobber the message
    cco_destroy(c);
    memcpy(g_err, keep, sizeof keep);
    return st;
  }
  *out = c;
  return CCO_OK;
}

int cco_destroy(cco_ctx_t *c) {
  if (!c) return CCO_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  for (auto &b : c->pinned) cudaFreeHost(b.p);
  if (c->mail_h) cudaFreeHost(c->mail_h);
  for (auto &ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : c->tev)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : c->copy_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto &st : c->bin_stream)
    if (st) cudaStreamDestroy(st);
  for (auto &ev : c->bin_ev)
    if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
  return CCO_OK;
}

int cco_host_alloc(cco_ctx_t *c, size_t bytes, void **out) {
  if (!c || !out) return set_error(CCO_E_INVALID_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  void *p = c->pinned_get(bytes);
  if (!p) return set_error(CCO_E_OOM, "cudaHostAlloc(%zu) failed", bytes);
  *out = p;
  return CCO_OK;
}
int cco_host_free(cco_ctx_t *c, void *p) {
  if (!c) return set_error(CCO_E_INVALID_ARG, "null context");
  if (p) c->pinned_put(p);
  return CCO_OK;
}

int cco_train(cco_ctx_t *ctx, int32_t n_mats, const cco_csr_t *mats, const cco_indicator_params_t *params, int32_t seed,
              uint32_t flags, cco_result_t **out) {
  if (!ctx || !out) return set_error(CCO_E_INVALID_ARG, "null argument");
  *out = nullptr;
  return train_impl(ctx, n_mats, mats, params, seed, flags, out);
}

int cco_cooccurrences_idss(cco_ctx_t *ctx, int32_t n_mats, const cco_csr_t *mats, int32_t seed,
                           int32_t max_interesting_items_per_thing, int32_t max_num_interactions, uint32_t flags,
                           cco_result_t **out) {
  if (n_mats < 1) return set_error(CCO_E_INVALID_ARG, "need at least the primary matrix");
  std::vector<cco_indicator_params_t> p(n_mats);
  for (auto &q : p) {
    q.max_interactions = max_num_interactions;
    q.top_k = max_interesting_items_per_thing;
    q.has_min_llr = 0;
    q.min_llr = 0.0;
  }
  return cco_train(ctx, n_mats, mats, p.data(), seed, flags, out);
}

int cco_dataset_shape(const cco_dataset_t *ds, int32_t i, int64_t *n_rows, int32_t *n_cols, int64_t *nnz) {
  if (!ds || i < 0 || i >= ds->n_mats) return set_error(CCO_E_INVALID_ARG, "bad dataset/index");
  if (n_rows) *n_rows = ds->n_users;
  if (n_cols) *n_cols = (int32_t)ds->n_cols[i];
  if (nnz) *nnz = ds->nnz[i];
  return CCO_OK;
}

int cco_dataset_download(const cco_dataset_t *ds, int32_t i, int64_t **row_ptr, int32_t **col_idx) {
  if (!ds || !row_ptr || !col_idx || i < 0 || i >= ds->n_mats) return set_error(CCO_E_INVALID_ARG, "bad argument");
  cco_ctx *c = ds->ctx;
  CK(cudaSetDevice(c->device));
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * ((size_t)ds->n_users + 1));
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)std::max<long long>(ds->nnz[i], 1));
  if (!rp || !ci) return set_error(CCO_E_OOM, "malloc failed");
  CK(cudaStreamSynchronize(c->copy_stream));
  CK(cudaMemcpyAsync(rp, ds->rp[i], sizeof(int64_t) * ((size_t)ds->n_users + 1), cudaMemcpyDeviceToHost, c->stream));
  if (ds->nnz[i] > 0)
    CK(cudaMemcpyAsync(ci, ds->col[i], sizeof(int32_t) * (size_t)ds->nnz[i], cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *row_ptr = rp;
  *col_idx = ci;
  return CCO_OK;
}

// Preparator.prepare on the device: histogram + scans for the dictionaries, one radix sort + unique
// per event type for the binary CSR.
int cco_ingest(cco_ctx_t *c, int32_t n_types, const cco_events_t *ev, int64_t n_users_raw, int32_t min_events_per_user,
               int32_t *user_map, int32_t *const *item_maps, cco_dataset_t **out) {
  if (!c || !ev || !user_map || !item_maps || !out || n_types < 1) return set_error(CCO_E_INVALID_ARG, "bad argument");
  if (n_users_raw < 0 || n_users_raw >= 0x7fffffffLL) return set_error(CCO_E_INVALID_ARG, "n_users_raw out of range");
  for (int t = 0; t < n_types; ++t) {
    if (ev[t].n_events < 0 || ev[t].n_events >= 0xffffffffLL || ev[t].n_items_raw < 0)
      return set_error(CCO_E_INVALID_ARG, "type %d: bad event count / item space", t);
    if (ev[t].n_events > 0 && (!ev[t].user || !ev[t].item)) return set_error(CCO_E_INVALID_ARG, "type %d: null event arrays", t);
    if (!item_maps[t] && ev[t].n_items_raw > 0) return set_error(CCO_E_INVALID_ARG, "type %d: null item_map", t);
    // ids are range-checked on the host: they index device arrays
    for (int64_t i = 0; i < ev[t].n_events; ++i)
      if (ev[t].user[i] < 0 || ev[t].user[i] >= n_users_raw || ev[t].item[i] < 0 || ev[t].item[i] >= ev[t].n_items_raw)
        return set_error(CCO_E_INVALID_ARG, "type %d: user or item id out of range at event %lld", t, (long long)i);
  }
  *out = nullptr;
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  c->mail_pending.clear();
  c->mail_used = 0;
  Arena ar(s);
  cco_dataset *d = new cco_dataset();
  d->ctx = c;
  d->n_mats = n_types;
  d->rp.assign(n_types, nullptr);
  d->col.assign(n_types, nullptr);
  d->n_cols.assign(n_types, 0);
  d->nnz.assign(n_types, 0);
  d->ready.assign(n_types, nullptr);
  struct G {
    cco_dataset *d;
    bool ok = false;
    ~G() {
      if (!ok) dataset_release(d);
    }
  } g{d};
  const long long nu = std::max<long long>(n_users_raw, 1);
  // events to the device
  std::vector<long long *> d_user(n_types, nullptr);
  std::vector<int32_t *> d_item(n_types, nullptr);
  for (int t = 0; t < n_types; ++t) {
    CKR(ar.alloc(&d_user[t], std::max<long long>(ev[t].n_events, 1)));
    CKR(ar.alloc(&d_item[t], std::max<long long>(ev[t].n_events, 1)));
    if (ev[t].n_events > 0) {
      CK(cudaMemcpyAsync(d_user[t], ev[t].user, sizeof(int64_t) * (size_t)ev[t].n_events, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(d_item[t], ev[t].item, sizeof(int32_t) * (size_t)ev[t].n_events, cudaMemcpyHostToDevice, s));
    }
  }
  // user dictionary from the primary events
  int32_t *cnt, *d_user_map;
  uint32_t *uflag, *upos;
  CKR(ar.alloc(&cnt, nu));
  CKR(ar.alloc(&uflag, nu + 1));
  CKR(ar.alloc(&upos, nu + 1));
  CKR(ar.alloc(&d_user_map, nu));
  CK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)nu, s));
  CK(cudaMemsetAsync(uflag, 0, sizeof(uint32_t) * ((size_t)nu + 1), s));
  if (ev[0].n_events > 0)
    k_ingest_count_users<<<grid_for(ev[0].n_events, 256, c->sm_count), 256, 0, s>>>(ev[0].n_events, d_user[0], cnt);
  const int32_t need = min_events_per_user > 1 ? min_events_per_user : 1;
  if (n_users_raw > 0)
    k_ingest_user_flags<<<grid_for(n_users_raw, 256, c->sm_count), 256, 0, s>>>(n_users_raw, cnt, need, uflag);
  CKR(exclusive_sum_u32(c, ar, uflag, upos, nu + 1));
  if (n_users_raw > 0)
    k_ingest_make_map<<<grid_for(n_users_raw, 256, c->sm_count), 256, 0, s>>>(n_users_raw, uflag, upos, d_user_map);
  c->launches += 3;
  uint32_t n_users = 0;
  CKR(mail_fetch(c, &n_users, upos + n_users_raw, 4));
  if (n_users_raw > 0)
    CK(cudaMemcpyAsync(user_map, d_user_map, sizeof(int32_t) * (size_t)n_users_raw, cudaMemcpyDeviceToHost, s));
  CKR(mail_wait(c));
  d->n_users = n_users;
  for (int t = 0; t < n_types; ++t) {
    const long long ne = ev[t].n_events, ni = std::max<int32_t>(ev[t].n_items_raw, 1);
    uint32_t *iflag, *ipos;
    int32_t *d_item_map;
    CKR(ar.alloc(&iflag, ni + 1));
    CKR(ar.alloc(&ipos, ni + 1));
    CKR(ar.alloc(&d_item_map, ni));
    CK(cudaMemsetAsync(iflag, 0, sizeof(uint32_t) * ((size_t)ni + 1), s));
    if (ne > 0)
      k_ingest_item_flags<<<grid_for(ne, 256, c->sm_count), 256, 0, s>>>(ne, d_user[t], d_item[t], d_user_map, iflag);
    CKR(exclusive_sum_u32(c, ar, iflag, ipos, ni + 1));
    k_ingest_make_map<<<grid_for(ni, 256, c->sm_count), 256, 0, s>>>(ev[t].n_items_raw, iflag, ipos, d_item_map);
    uint32_t n_items = 0;
    CKR(mail_fetch(c, &n_items, ipos + ev[t].n_items_raw, 4));
    if (ev[t].n_items_raw > 0)
      CK(cudaMemcpyAsync(item_maps[t], d_item_map, sizeof(int32_t) * (size_t)ev[t].n_items_raw, cudaMemcpyDeviceToHost, s));
    // sort surviving (user, item) keys, drop duplicates, rebuild row_ptr
    unsigned long long *k0, *k1, *d_kept;
    CKR(ar.alloc(&k0, std::max<long long>(ne, 1)));
    CKR(ar.alloc(&k1, std::max<long long>(ne, 1)));
    CKR(ar.alloc(&d_kept, 1));
    CK(cudaMemsetAsync(d_kept, 0, 8, s));
    if (ne > 0)
      k_ingest_keys<<<grid_for(ne, 256, c->sm_count), 256, 0, s>>>(ne, d_user[t], d_item[t], d_user_map, d_item_map, k0, d_kept);
    c->launches += 3;
    unsigned long long kept = 0;
    CKR(mail_fetch(c, &kept, d_kept, 8));
    CKR(mail_wait(c));
    d->n_cols[t] = n_items;
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, sizeof(int64_t) * ((size_t)n_users + 1), s);
    if (e != cudaSuccess) return set_error(CCO_E_OOM, "cudaMallocAsync row_ptr: %s", cudaGetErrorString(e));
    d->rp[t] = (long long *)p;
    e = cudaMallocAsync(&p, sizeof(int32_t) * (size_t)std::max<unsigned long long>(kept, 4), s);
    if (e != cudaSuccess) return set_error(CCO_E_OOM, "cudaMallocAsync col_idx: %s", cudaGetErrorString(e));
    d->col[t] = (int32_t *)p;
    CK(cudaEventCreateWithFlags(&d->ready[t], cudaEventDisableTiming));
    long long n_unique = 0;
    if (kept > 0) {
      cub::DoubleBuffer<unsigned long long> db(k0, k1);
      size_t tb = 0;
      CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, (long long)ne, 0, 64, s));
      void *tmp;
      CKR(ar.alloc((char **)&tmp, tb));
      CK(cub::DeviceRadixSort::SortKeys(tmp, tb, db, (long long)ne, 0, 64, s));
      ar.release(tmp);
      unsigned long long *sorted = db.Current(), *other = db.Alternate();
      uint32_t *flag, *pos;
      CKR(ar.alloc(&flag, kept + 1));
      CKR(ar.alloc(&pos, kept + 1));
      CK(cudaMemsetAsync(flag + kept, 0, 4, s));
      k_unique_flags<<<grid_for((long long)kept, 256, c->sm_count), 256, 0, s>>>((long long)kept, sorted, flag);
      CKR(exclusive_sum_u32(c, ar, flag, pos, (long long)kept + 1));
      uint32_t nuq = 0;
      CKR(mail_fetch(c, &nuq, pos + kept, 4));
      k_unique_scatter<<<grid_for((long long)kept, 256, c->sm_count), 256, 0, s>>>((long long)kept, sorted, flag, pos, other, d->col[t]);
      CKR(mail_wait(c));
      n_unique = nuq;
      k_rowptr_from_keys<<<grid_for((long long)n_users + 1, 256, c->sm_count), 256, 0, s>>>((long long)n_users, n_unique, other, d->rp[t]);
      c->launches += 3;
      ar.release(flag);
      ar.release(pos);
    } else {
      CK(cudaMemsetAsync(d->rp[t], 0, sizeof(int64_t) * ((size_t)n_users + 1), s));
    }
    d->nnz[t] = n_unique;
    CK(cudaEventRecord(d->ready[t], s));
    ar.release(k0);
    ar.release(k1);
    ar.release(iflag);
    ar.release(ipos);
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  g.ok = true;
  *out = d;
  return CCO_OK;
}

int cco_partition_rows(const int64_t *work_prefix, int32_t n_items, int32_t world_size, int32_t *bounds) {
  if (!work_prefix || !bounds || n_items < 0 || world_size < 1) return set_error(CCO_E_INVALID_ARG, "bad argument");
  // weight of row i = its products + 1 (so rows without work are spread too); contiguous ranges of equal weight
  const long long total = (long long)work_prefix[n_items] + n_items;
  for (int r = 0; r <= world_size; ++r) {
    if (r == 0) { bounds[r] = 0; continue; }
    if (r == world_size) { bounds[r] = n_items; continue; }
    const long long target = (long long)((__int128)total * r / world_size);
    int lo = 0, hi = n_items;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((long long)work_prefix[mid] + mid < target) lo = mid + 1; else hi = mid;
    }
    bounds[r] = lo;
  }
  return CCO_OK;
}

int cco_dataset_upload(cco_ctx_t *ctx, int32_t n_mats, const cco_csr_t *mats, uint32_t flags, cco_dataset_t **out) {
  if (!ctx || !out) return set_error(CCO_E_INVALID_ARG, "null argument");
  *out = nullptr;
  std::vector<cco_indicator_params_t> p(std::max(n_mats, 1), cco_indicator_params_t{1, 1, 0, 0.0});
  CKR(validate_host(n_mats, mats, p.data()));
  return dataset_upload(ctx, n_mats, mats, flags, out);
}
int cco_dataset_free(cco_dataset_t *ds) {
  dataset_release(ds);
  return CCO_OK;
}
int cco_train_dataset(cco_ctx_t *ctx, const cco_dataset_t *ds, const cco_indicator_params_t *params, int32_t seed,
                      uint32_t flags, cco_result_t **out) {
  if (!ctx || !ds || !params || !out) return set_error(CCO_E_INVALID_ARG, "null argument");
  if (ds->ctx != ctx) return set_error(CCO_E_INVALID_ARG, "dataset belongs to another context");
  *out = nullptr;
  for (int i = 0; i < ds->n_mats; ++i) {
    if (params[i].max_interactions < 1) return set_error(CCO_E_INVALID_ARG, "matrix %d: max_interactions must be >= 1", i);
    if (params[i].top_k < 1) return set_error(CCO_E_INVALID_ARG, "matrix %d: top_k must be >= 1", i);
    if (params[i].top_k > CCO_MAX_TOP_K)
      return set_error(CCO_E_UNSUPPORTED, "matrix %d: top_k %d > CCO_MAX_TOP_K (%d)", i, params[i].top_k, CCO_MAX_TOP_K);
  }
  ctx->launches = 0;
  return train_dataset(ctx, ds, params, seed, flags, out);
}
int cco_timer_start(cco_ctx_t *ctx) {
  if (!ctx) return set_error(CCO_E_INVALID_ARG, "null context");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->tev[0], ctx->stream));
  return CCO_OK;
}
int cco_timer_stop(cco_ctx_t *ctx, float *ms) {
  if (!ctx || !ms) return set_error(CCO_E_INVALID_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->tev[1], ctx->stream));
  CK(cudaEventSynchronize(ctx->tev[1]));
  CK(cudaEventElapsedTime(ms, ctx->tev[0], ctx->tev[1]));
  return CCO_OK;
}

int cco_result_num_matrices(const cco_result_t *r) { return r ? (int)r->mats.size() : set_error(CCO_E_INVALID_ARG, "null result"); }

int cco_result_row_range(const cco_result_t *r, int32_t i, int64_t *row_begin, int64_t *row_end) {
  if (!r || i < 0 || i >= (int)r->mats.size()) return set_error(CCO_E_INVALID_ARG, "bad result/index");
  if (row_begin) *row_begin = r->mats[i].row_begin;
  if (row_end) *row_end = r->mats[i].row_end;
  return CCO_OK;
}

int cco_result_matrix(const cco_result_t *r, int32_t i, int64_t *n_rows, int32_t *n_cols, const int64_t **row_ptr,
                      const int32_t **col_idx, const double **llr, const int32_t **count) {
  if (!r || i < 0 || i >= (int)r->mats.size()) return set_error(CCO_E_INVALID_ARG, "bad result/index");
  const ResultMat &m = r->mats[i];
  if (n_rows) *n_rows = m.row_end - m.row_begin;
  if (n_cols) *n_cols = m.n_cols;
  if (row_ptr) *row_ptr = m.row_ptr;
  if (col_idx) *col_idx = m.col;
  if (llr) *llr = m.llr;
  if (count) *count = m.cnt;
  return CCO_OK;
}

int cco_result_stats(const cco_result_t *r, cco_stats_t *out) {
  if (!r || !out) return set_error(CCO_E_INVALID_ARG, "null argument");
  *out = r->stats;
  return CCO_OK;
}

int cco_result_free(cco_result_t *r) {
  if (!r) return CCO_OK;
  for (auto &m : r->mats) {
    for (void *p : {(void *)m.row_ptr, (void *)m.col, (void *)m.llr, (void *)m.cnt})
      if (p) r->ctx->pinned_put(p);
  }
  delete r;
  return CCO_OK;
}

void cco_free(void *p) { free(p); }

// ---- debug / parity entries ------------------------------------------------------------------------
int cco_debug_llr(cco_ctx_t *c, int64_t n, const int64_t *k11, const int64_t *k12, const int64_t *k21, const int64_t *k22,
                  uint32_t flags, double *out) {
  if (!c || n < 0 || (n > 0 && (!k11 || !k12 || !k21 || !k22 || !out))) return set_error(CCO_E_INVALID_ARG, "bad argument");
  if (n == 0) return CCO_OK;
  for (int64_t i = 0; i < n; ++i)
    if (k11[i] < 0 || k12[i] < 0 || k21[i] < 0 || k22[i] < 0)
      return set_error(CCO_E_INVALID_ARG, "negative count at %lld (Preconditions.checkArgument in LogLikelihood)", (long long)i);
  CK(cudaSetDevice(c->device));
  Arena ar(c->stream);
  long long *d[4];
  double *dout;
  const int64_t *h[4] = {k11, k12, k21, k22};
  for (int j = 0; j < 4; ++j) {
    CKR(ar.alloc(&d[j], n));
    CK(cudaMemcpyAsync(d[j], h[j], sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  CKR(ar.alloc(&dout, n));
  k_debug_llr<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, d[0], d[1], d[2], d[3], flags, dout);
  CK(cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  return CCO_OK;
}

int cco_debug_downsample(cco_ctx_t *c, const cco_csr_t *m, int32_t max_interactions, int32_t seed, uint32_t flags,
                         int64_t **row_ptr, int32_t **col_idx, int32_t *raw_col_counts, int32_t *new_col_counts) {
  if (!c || !m || !row_ptr || !col_idx) return set_error(CCO_E_INVALID_ARG, "null argument");
  cco_indicator_params_t prm = {max_interactions, 1, 0, 0.0};
  CKR(validate_host(1, m, &prm));
  CK(cudaSetDevice(c->device));
  cco_dataset *ds = nullptr;
  CKR(dataset_upload(c, 1, m, flags, &ds));
  struct DG { cco_dataset *d; ~DG() { dataset_release(d); } } dg{ds};
  Arena ar(c->stream);
  std::vector<DevRaw> raw(1);
  raw[0].n_rows = ds->n_users; raw[0].n_cols = (int32_t)ds->n_cols[0]; raw[0].nnz = ds->nnz[0];
  raw[0].rp = ds->rp[0]; raw[0].col = ds->col[0];
  int32_t *counts;
  CKR(ar.alloc(&counts, std::max<int32_t>(m->n_cols, 1)));
  CK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)std::max<int32_t>(m->n_cols, 1), c->stream));
  if (raw[0].nnz > 0 && m->n_rows > 0)
    k_col_histogram<<<grid_for(raw[0].nnz, 256, c->sm_count), 256, 0, c->stream>>>(0, m->n_rows, raw[0].rp, raw[0].col, counts, 1, 0);
  DevMat dm;
  CKR(downsample_device(c, ar, raw[0], counts, max_interactions, seed, flags, &dm));
  std::vector<uint32_t> rp32((size_t)m->n_rows + 1);
  CK(cudaMemcpyAsync(rp32.data(), dm.rp, sizeof(uint32_t) * rp32.size(), cudaMemcpyDeviceToHost, c->stream));
  if (raw_col_counts && m->n_cols > 0)
    CK(cudaMemcpyAsync(raw_col_counts, counts, sizeof(int32_t) * (size_t)m->n_cols, cudaMemcpyDeviceToHost, c->stream));
  if (new_col_counts && m->n_cols > 0)
    CK(cudaMemcpyAsync(new_col_counts, dm.marg, sizeof(int32_t) * (size_t)m->n_cols, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  size_t nnz = rp32[m->n_rows];
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * rp32.size());
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(nnz, 1));
  if (!rp || !ci) return set_error(CCO_E_OOM, "malloc failed");
  for (size_t i = 0; i < rp32.size(); ++i) rp[i] = rp32[i];
  if (nnz) CK(cudaMemcpy(ci, dm.col, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost));
  *row_ptr = rp;
  *col_idx = ci;
  return CCO_OK;
}

int cco_debug_cooccurrence(cco_ctx_t *c, const cco_csr_t *a, const cco_csr_t *b, int64_t **row_ptr, int32_t **col_idx,
                           int32_t **count) {
  if (!c || !a || !b || !row_ptr || !col_idx || !count) return set_error(CCO_E_INVALID_ARG, "null argument");
  cco_csr_t two[2] = {*a, *b};
  cco_indicator_params_t prm[2] = {{0x7fffffff, 1, 0, 0.0}, {0x7fffffff, 1, 0, 0.0}};
  CKR(validate_host(2, two, prm));
  CK(cudaSetDevice(c->device));
  c->mail_pending.clear();
  c->mail_used = 0;
  cudaStream_t s = c->stream;
  cco_dataset *ds = nullptr;
  CKR(dataset_upload(c, 2, two, 0, &ds));
  struct DG { cco_dataset *d; ~DG() { dataset_release(d); } } dg{ds};
  Arena ar(s);
  std::vector<DevRaw> raw(2);
  for (int i = 0; i < 2; ++i) {
    raw[i].n_rows = ds->n_users; raw[i].n_cols = (int32_t)ds->n_cols[i]; raw[i].nnz = ds->nnz[i];
    raw[i].rp = ds->rp[i]; raw[i].col = ds->col[i];
  }
  // identity "downsample" (m = INT_MAX) gives the device CSR + marginals
  std::vector<DevMat> dm(2);
  for (int i = 0; i < 2; ++i) {
    int32_t *counts;
    CKR(ar.alloc(&counts, std::max<int32_t>(raw[i].n_cols, 1)));
    CK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)std::max<int32_t>(raw[i].n_cols, 1), s));
    if (raw[i].nnz > 0 && raw[i].n_rows > 0)
      k_col_histogram<<<grid_for(raw[i].nnz, 256, c->sm_count), 256, 0, s>>>(0, raw[i].n_rows, raw[i].rp, raw[i].col, counts, 1, 0);
    CKR(downsample_device(c, ar, raw[i], counts, 0x7fffffff, 0, 0, &dm[i]));
  }
  const int32_t n_items_a = dm[0].n_cols;
  uint32_t *at_ptr, *cursor, *marg_pad;
  int32_t *at_users, *d_max;
  CKR(ar.alloc(&at_ptr, n_items_a + 1));
  CKR(ar.alloc(&cursor, n_items_a + 1));
  CKR(ar.alloc(&marg_pad, n_items_a + 1));
  CKR(ar.alloc(&d_max, 1));
  CKR(ar.alloc(&at_users, std::max<long long>(raw[0].nnz, 1)));
  CK(cudaMemsetAsync(d_max, 0, 4, s));
  CK(cudaMemcpyAsync(marg_pad, dm[0].marg, sizeof(int32_t) * (size_t)n_items_a, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(marg_pad + n_items_a, 0, 4, s));
  CKR(exclusive_sum_u32(c, ar, marg_pad, at_ptr, (long long)n_items_a + 1));
  CK(cudaMemcpyAsync(cursor, at_ptr, sizeof(uint32_t) * ((size_t)n_items_a + 1), cudaMemcpyDeviceToDevice, s));
  k_transpose_scatter<<<grid_for(a->n_rows * kSG, 256, c->sm_count), 256, 0, s>>>(a->n_rows, dm[0].rp, dm[0].col, cursor, at_users);
  if (n_items_a > 0) k_max_i32<<<grid_for(n_items_a, 256, c->sm_count, 2), 256, 0, s>>>(n_items_a, dm[0].marg, d_max);
  int32_t max_marg_a = 0;
  CK(cudaMemcpyAsync(&max_marg_a, d_max, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  ResultMat rm;
  IndicatorOut io;
  cco_indicator_params_t p1 = {0x7fffffff, 1, 0, 0.0};
  int rc = run_indicator(c, ar, at_ptr, at_users, n_items_a, dm[0].marg, max_marg_a, dm[1], a->n_rows, false, p1, 0, true, 0, 1,
                         &rm, &io, nullptr);
  cudaStreamSynchronize(c->copy_stream);
  auto put = [&]() {
    for (void *p : {(void *)rm.row_ptr, (void *)rm.col, (void *)rm.llr, (void *)rm.cnt})
      if (p) c->pinned_put(p);
  };
  if (rc != CCO_OK) {
    put();
    return rc;
  }
  size_t nnz = (size_t)rm.row_ptr[n_items_a];
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_items_a + 1));
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(nnz, 1));
  int32_t *cn = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(nnz, 1));
  if (!rp || !ci || !cn) {
    put();
    return set_error(CCO_E_OOM, "malloc failed");
  }
  memcpy(rp, rm.row_ptr, sizeof(int64_t) * ((size_t)n_items_a + 1));
  // cells of a row come back in table order: sort each row by column for the caller
  std::vector<std::pair<int32_t, int32_t>> tmp;
  for (int32_t r = 0; r < n_items_a; ++r) {
    size_t lo = (size_t)rp[r], hi = (size_t)rp[r + 1];
    tmp.resize(hi - lo);
    for (size_t q = lo; q < hi; ++q) tmp[q - lo] = {rm.col[q], rm.cnt[q]};
    std::sort(tmp.begin(), tmp.end());
    for (size_t q = lo; q < hi; ++q) {
      ci[q] = tmp[q - lo].first;
      cn[q] = tmp[q - lo].second;
    }
  }
  put();
  *row_ptr = rp;
  *col_idx = ci;
  *count = cn;
  return CCO_OK;
}

}  // extern "C"
