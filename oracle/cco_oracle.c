/*
 * cco_oracle.c -- CPU restatement of the CCO train path (see cco_oracle.h for the
 * "test infrastructure only" notice, the reference anchors and the parity-pinning status).
 *
 * Each function cites what it restates.  "[Mahout]" = Apache Mahout 0.13.0 source, which is a
 * pinned third-party dependency of the reference (build.sbt:15,34-38) and is not in the
 * reference tree; those parts follow SURVEY.md Appendix A.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include "cco_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
#define FAIL(...)                                 \
  do {                                            \
    snprintf(g_err, sizeof g_err, __VA_ARGS__);   \
    return -1;                                    \
  } while (0)

void orc_free(void *p) { free(p); }
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * LogLikelihood  [Mahout mahout-math org.apache.mahout.math.stats.LogLikelihood; SURVEY.md A.3]
 *   xLogX(x)              = x == 0 ? 0.0 : x * Math.log(x)
 *   entropy(a,b)          = xLogX(a+b) - xLogX(a) - xLogX(b)                 (left to right)
 *   entropy(a,b,c,d)      = xLogX(a+b+c+d) - xLogX(a) - xLogX(b) - xLogX(c) - xLogX(d)
 *   entropy(long...)      = xLogX(sum) - (sum of xLogX in argument order)      (varargs form)
 *   logLikelihoodRatio    = rowEntropy + columnEntropy < matrixEntropy ? 0.0
 *                           : 2.0 * (rowEntropy + columnEntropy - matrixEntropy)
 * The two entropy forms differ only in the last bits; which one 0.13.0 calls from
 * logLikelihoodRatio cannot be checked here (no Mahout source), so both exist and the
 * left-to-right specialisation is the default (DESIGN.md "LLR evaluation order").
 * ------------------------------------------------------------------------------------------- */
double orc_xlogx(int64_t x) { return x == 0 ? 0.0 : (double)x * log((double)x); }

static double entropy2(int64_t a, int64_t b, int flags) {
  if (flags & ORC_FLAG_ENTROPY_VARARGS) {
    double r = 0.0;
    r += orc_xlogx(a);
    r += orc_xlogx(b);
    return orc_xlogx(a + b) - r;
  }
  return orc_xlogx(a + b) - orc_xlogx(a) - orc_xlogx(b);
}

static double entropy4(int64_t a, int64_t b, int64_t c, int64_t d, int flags) {
  if (flags & ORC_FLAG_ENTROPY_VARARGS) {
    double r = 0.0;
    r += orc_xlogx(a);
    r += orc_xlogx(b);
    r += orc_xlogx(c);
    r += orc_xlogx(d);
    return orc_xlogx(a + b + c + d) - r;
  }
  return orc_xlogx(a + b + c + d) - orc_xlogx(a) - orc_xlogx(b) - orc_xlogx(c) - orc_xlogx(d);
}

double orc_llr(int64_t k11, int64_t k12, int64_t k21, int64_t k22, int flags) {
  /* Preconditions.checkArgument(k >= 0) -> IllegalArgumentException in Mahout; NaN here */
  if (k11 < 0 || k12 < 0 || k21 < 0 || k22 < 0) return NAN;
  double row_entropy = entropy2(k11 + k12, k21 + k22, flags);
  double column_entropy = entropy2(k11 + k21, k12 + k22, flags);
  double matrix_entropy = entropy4(k11, k12, k21, k22, flags);
  if (row_entropy + column_entropy < matrix_entropy) return 0.0; /* round off error */
  return 2.0 * (row_entropy + column_entropy - matrix_entropy);
}

/* SimilarityAnalysis.logLikelihoodRatio(numInteractionsWithA, ..WithB, ..WithAandB, numInteractions)
 * [Mahout; SURVEY.md 8a-H5] */
static inline double llr_from_marginals(int64_t with_a, int64_t with_b, int64_t with_ab, int64_t n,
                                        int flags) {
  int64_t k11 = with_ab;
  int64_t k12 = with_a - with_ab;
  int64_t k21 = with_b - with_ab;
  int64_t k22 = n - with_a - with_b + with_ab;
  return orc_llr(k11, k12, k21, k22, flags);
}

/* ---------------------------------------------------------------------------------------------
 * Shared-memory parallel helpers of the CPU arm (OpenMP).  They only change HOW FAST the restated
 * algorithm runs on the host cores (bench.py --impl reference), never what it computes.
 * ------------------------------------------------------------------------------------------- */
static int g_threads = 0; /* set by orc_train / orc_set_threads; 0 = OpenMP default */
void orc_set_threads(int n) {
  g_threads = n > 0 ? n : 0;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n); /* overrides OMP_NUM_THREADS (torchrun exports 1 to its children) */
#endif
}
static int team_size(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* counts[j] = number of entries of idx[0..n) equal to j; per-thread private histograms (at most 32 of them, so the
 * scratch stays <= 32 * n_cols * 4 bytes) folded column-parallel: no atomics on Zipf-hot columns */
static int col_histogram(const int32_t *idx, int64_t n, int32_t n_cols, int32_t *counts) {
  const int32_t nc = n_cols > 0 ? n_cols : 1;
  int T = team_size();
  if (T > 32) T = 32;
  if (n < (1 << 16) || T < 2) {
    memset(counts, 0, sizeof(int32_t) * (size_t)nc);
    for (int64_t i = 0; i < n; ++i) counts[idx[i]]++;
    return 0;
  }
  int32_t *priv = (int32_t *)calloc((size_t)T * (size_t)nc, sizeof(int32_t));
  if (!priv) FAIL("out of memory");
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int t = 0, nt = 1;
#endif
    int32_t *mine = priv + (size_t)t * (size_t)nc;
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (int64_t i = lo; i < hi; ++i) mine[idx[i]]++;
#pragma omp barrier
#pragma omp for schedule(static)
    for (int32_t j = 0; j < nc; ++j) {
      int32_t acc = 0;
      for (int k = 0; k < nt; ++k) acc += priv[(size_t)k * (size_t)nc + j];
      counts[j] = acc;
    }
  }
  free(priv);
  return 0;
}

/* in-place inclusive prefix sum of x[1..n] (x[0] stays): two-pass block scan */
static void prefix_sum_i64(int64_t *x, int64_t n) {
  int T = team_size();
  if (n < (1 << 16) || T < 2) {
    for (int64_t r = 0; r < n; ++r) x[r + 1] += x[r];
    return;
  }
  if (T > 64) T = 64;
  int64_t block_sum[65];
  memset(block_sum, 0, sizeof block_sum);
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int t = 0, nt = 1;
#endif
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    int64_t s = 0;
    for (int64_t r = lo; r < hi; ++r) s += x[r + 1];
    block_sum[t + 1] = s;
#pragma omp barrier
#pragma omp single
    for (int k = 0; k < nt; ++k) block_sum[k + 1] += block_sum[k];
    int64_t run = x[0] + block_sum[t];
    for (int64_t r = lo; r < hi; ++r) {
      run += x[r + 1];
      x[r + 1] = run;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Deterministic counter-based sampler (this repo's definition, shared bit-for-bit with the CUDA
 * path -- include/cco_b200.h "Sampler").  Mahout seeds java.util.Random per Spark block
 * (MurmurHash(keys(0), seed)) and draws one nextDouble() per non-zero in hash-iteration order,
 * which no independent implementation can reproduce; parity with Mahout is exact
 * only when downsampling is the identity (every row and column count <= m).
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}
uint64_t orc_hash64(int32_t seed, int64_t u, int32_t j) {
  uint64_t x = mix64(((uint64_t)(uint32_t)seed << 32) | (uint64_t)(uint32_t)u);
  return mix64(x + (uint64_t)(uint32_t)j * 0x9e3779b97f4a7c15ULL);
}
double orc_u01(uint64_t h) { return (double)(h >> 11) * 0x1.0p-53; }

/* ---------------------------------------------------------------------------------------------
 * Input canonicalisation: Preparator.scala:201-208 builds each row with
 * RandomAccessSparseVector.setQuick(col, 1.0) -> duplicates collapse, order is irrelevant.
 * ------------------------------------------------------------------------------------------- */
static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

static int validate(const orc_csr_t *m) {
  if (!m || !m->row_ptr) FAIL("null matrix");
  if (m->n_rows < 0 || m->n_rows > 0x7fffffff) FAIL("n_rows out of range (Mahout row keys are Int)");
  if (m->n_cols < 0) FAIL("n_cols < 0");
  if (m->row_ptr[0] != 0) FAIL("row_ptr[0] != 0");
  int64_t bad_row = -1, bad_idx = -1;
#pragma omp parallel for schedule(static, 65536)
  for (int64_t r = 0; r < m->n_rows; ++r)
    if (m->row_ptr[r + 1] < m->row_ptr[r]) {
#pragma omp critical
      bad_row = r;
    }
  if (bad_row >= 0) FAIL("row_ptr not monotone at row %lld", (long long)bad_row);
  int64_t nnz = m->row_ptr[m->n_rows];
  if (nnz > 0 && !m->col_idx) FAIL("null col_idx");
#pragma omp parallel for schedule(static, 65536)
  for (int64_t i = 0; i < nnz; ++i)
    if (m->col_idx[i] < 0 || m->col_idx[i] >= m->n_cols) {
#pragma omp critical
      bad_idx = i;
    }
  if (bad_idx >= 0) FAIL("col_idx out of range at %lld", (long long)bad_idx);
  return 0;
}

int orc_canonicalize(const orc_csr_t *in, int64_t **row_ptr, int32_t **col_idx) {
  if (validate(in)) return -1;
  int64_t nnz = in->row_ptr[in->n_rows];
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(in->n_rows + 1));
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!rp || !ci) FAIL("out of memory");
  /* fast path: every row already strictly ascending (what Preparator-built matrices look like after a sort) */
  int all_sorted = 1;
#pragma omp parallel for schedule(static, 4096) reduction(&& : all_sorted)
  for (int64_t r = 0; r < in->n_rows; ++r)
    for (int64_t i = in->row_ptr[r] + 1; i < in->row_ptr[r + 1]; ++i)
      if (in->col_idx[i] <= in->col_idx[i - 1]) all_sorted = 0;
  if (all_sorted) {
    memcpy(rp, in->row_ptr, sizeof(int64_t) * (size_t)(in->n_rows + 1));
    if (nnz > 0) memcpy(ci, in->col_idx, sizeof(int32_t) * (size_t)nnz);
    *row_ptr = rp;
    *col_idx = ci;
    return 0;
  }
  int64_t w = 0;
  rp[0] = 0;
  for (int64_t r = 0; r < in->n_rows; ++r) {
    int64_t s = in->row_ptr[r], e = in->row_ptr[r + 1], start = w;
    memcpy(ci + w, in->col_idx + s, sizeof(int32_t) * (size_t)(e - s));
    int sorted = 1;
    for (int64_t i = s + 1; i < e; ++i)
      if (in->col_idx[i] <= in->col_idx[i - 1]) { sorted = 0; break; }
    if (!sorted) {
      qsort(ci + start, (size_t)(e - s), sizeof(int32_t), cmp_i32);
      int64_t u = start;
      for (int64_t i = start; i < start + (e - s); ++i)
        if (i == start || ci[i] != ci[u - 1]) ci[u++] = ci[i];
      w = u;
    } else {
      w = start + (e - s);
    }
    rp[r + 1] = w;
  }
  *row_ptr = rp;
  *col_idx = ci;
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * sampleDownAndBinarize(drmM, seed, maxNumInteractions)  [Mahout SimilarityAnalysis; SURVEY.md A.1]
 *   numInteractions = drmI.numNonZeroElementsPerColumn            (raw column counts)
 *   perRowSampleRate   = min(m, d_r) / d_r     (Int/Int in Scala -> ORC_FLAG_ROWRATE_INTDIV)
 *   perThingSampleRate = min(m, c_j) / c_j     (c_j Double -> real division)
 *   keep (r, j) iff random.nextDouble() <= min(perRowSampleRate, perThingSampleRate); value 1
 * ------------------------------------------------------------------------------------------- */
static inline int keep_entry(const orc_csr_t *in, const int32_t *cc, int64_t r, int64_t d, double row_rate, int64_t i,
                             int32_t m, int32_t seed) {
  int32_t j = in->col_idx[i];
  double c = (double)cc[j];
  double col_rate = (c < (double)m ? c : (double)m) / c;
  double rate = row_rate < col_rate ? row_rate : col_rate;
  (void)d;
  return orc_u01(orc_hash64(seed, r, j)) <= rate;
}

int orc_downsample(const orc_csr_t *in, int32_t m, int32_t seed, int flags, int64_t **row_ptr,
                   int32_t **col_idx, int32_t *raw_col_counts, int32_t *new_col_counts) {
  if (m < 1) FAIL("max_interactions must be >= 1");
  int64_t nnz = in->row_ptr[in->n_rows];
  int32_t *cc = (int32_t *)calloc((size_t)(in->n_cols > 0 ? in->n_cols : 1), sizeof(int32_t));
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(in->n_rows + 1));
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!cc || !rp || !ci) FAIL("out of memory");
  if (col_histogram(in->col_idx, nnz, in->n_cols, cc)) return -1;
  if (raw_col_counts) memcpy(raw_col_counts, cc, sizeof(int32_t) * (size_t)in->n_cols);
  /* pass 1 (parallel over rows): kept entries per row; the keep decision is a pure function of (seed, row, col) */
  rp[0] = 0;
#pragma omp parallel for schedule(static, 4096)
  for (int64_t r = 0; r < in->n_rows; ++r) {
    int64_t s = in->row_ptr[r], e = in->row_ptr[r + 1], d = e - s, kept = 0;
    double row_rate = 1.0;
    if (d > 0) {
      int64_t md = d < m ? d : m;
      row_rate = (flags & ORC_FLAG_ROWRATE_INTDIV) ? (double)(md / d) : (double)md / (double)d;
    }
    for (int64_t i = s; i < e; ++i) kept += keep_entry(in, cc, r, d, row_rate, i, m, seed);
    rp[r + 1] = kept;
  }
  prefix_sum_i64(rp, in->n_rows);
  /* pass 2 (parallel over rows): ordered write */
#pragma omp parallel for schedule(static, 4096)
  for (int64_t r = 0; r < in->n_rows; ++r) {
    int64_t s = in->row_ptr[r], e = in->row_ptr[r + 1], d = e - s, w = rp[r];
    double row_rate = 1.0;
    if (d > 0) {
      int64_t md = d < m ? d : m;
      row_rate = (flags & ORC_FLAG_ROWRATE_INTDIV) ? (double)(md / d) : (double)md / (double)d;
    }
    for (int64_t i = s; i < e; ++i)
      if (keep_entry(in, cc, r, d, row_rate, i, m, seed)) ci[w++] = in->col_idx[i];
  }
  if (new_col_counts && col_histogram(ci, rp[in->n_rows], in->n_cols, new_col_counts)) return -1;
  free(cc);
  *row_ptr = rp;
  *col_idx = ci;
  return 0;
}

/* item-major view of a canonical CSR: users of each item (what `drmA.t` provides).  The scatter runs on every host
 * thread with an atomic per-item cursor, so the order of users inside an item is arbitrary -- the integer
 * co-occurrence counts do not depend on it. */
static int transpose(const int64_t *rp, const int32_t *ci, int64_t n_rows, int32_t n_cols,
                     int64_t **t_ptr, int32_t **t_idx) {
  int64_t nnz = rp[n_rows];
  int64_t *tp = (int64_t *)calloc((size_t)n_cols + 2, sizeof(int64_t));
  int32_t *ti = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_cols > 0 ? n_cols : 1));
  int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_cols > 0 ? n_cols : 1));
  if (!tp || !ti || !cnt || !cur) FAIL("out of memory");
  if (col_histogram(ci, nnz, n_cols, cnt)) return -1;
  for (int32_t j = 0; j < n_cols; ++j) tp[j + 1] = cnt[j];
  prefix_sum_i64(tp, n_cols);
  memcpy(cur, tp, sizeof(int64_t) * (size_t)n_cols);
#pragma omp parallel for schedule(static, 4096)
  for (int64_t r = 0; r < n_rows; ++r)
    for (int64_t i = rp[r]; i < rp[r + 1]; ++i) {
      int64_t pos = __atomic_fetch_add(&cur[ci[i]], 1, __ATOMIC_RELAXED);
      ti[pos] = (int32_t)r;
    }
  free(cnt);
  free(cur);
  *t_ptr = tp; /* tp[j]..tp[j+1] delimits item j */
  *t_idx = ti;
  return 0;
}

/* `drmA.t %*% drmB`  [Mahout sparkbindings AtB; SURVEY.md 8a-H4]: integer co-occurrence counts */
int orc_cooccurrence(const orc_csr_t *a, const orc_csr_t *b, int64_t **row_ptr, int32_t **col_idx,
                     int32_t **count) {
  if (validate(a) || validate(b)) return -1;
  if (a->n_rows != b->n_rows) FAIL("row cardinality mismatch");
  int64_t *tp;
  int32_t *ti;
  if (transpose(a->row_ptr, a->col_idx, a->n_rows, a->n_cols, &tp, &ti)) return -1;
  int32_t *acc = (int32_t *)calloc((size_t)(b->n_cols > 0 ? b->n_cols : 1), sizeof(int32_t));
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * ((size_t)a->n_cols + 1));
  size_t cap = 1024, w = 0;
  int32_t *ci = (int32_t *)malloc(cap * sizeof(int32_t)), *cn = (int32_t *)malloc(cap * sizeof(int32_t));
  if (!acc || !rp || !ci || !cn) FAIL("out of memory");
  rp[0] = 0;
  for (int32_t item = 0; item < a->n_cols; ++item) {
    for (int64_t p = tp[item]; p < tp[item + 1]; ++p) {
      int64_t u = ti[p];
      for (int64_t q = b->row_ptr[u]; q < b->row_ptr[u + 1]; ++q) acc[b->col_idx[q]]++;
    }
    for (int32_t j = 0; j < b->n_cols; ++j)
      if (acc[j]) {
        if (w == cap) {
          cap *= 2;
          ci = (int32_t *)realloc(ci, cap * sizeof(int32_t));
          cn = (int32_t *)realloc(cn, cap * sizeof(int32_t));
          if (!ci || !cn) FAIL("out of memory");
        }
        ci[w] = j;
        cn[w] = acc[j];
        acc[j] = 0;
        ++w;
      }
    rp[item + 1] = (int64_t)w;
  }
  free(acc);
  free(tp);
  free(ti);
  *row_ptr = rp;
  *col_idx = ci;
  *count = cn;
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * computeSimilarities(drm = A'^T B', numUsers, k, rowMarg, colMarg, crossCooccurrence, minLLROpt)
 * [Mahout SimilarityAnalysis; SURVEY.md A.2, 8a-H5/H6]
 *   per row: skip the diagonal iff !crossCooccurrence; llr = logLikelihoodRatio(...);
 *   keep if minLLR.isEmpty || llr >= minLLR; bounded priority queue of size k (strict '>' at
 *   the cut); writing 0.0 into the sparse result stores nothing -> LLR == 0 cells vanish.
 * Tie policy of this repo (DESIGN.md "Ties"): total order (llr desc, col asc).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  double llr;
  int32_t col;
  int32_t cnt;
} cand_t;

/* returns 1 if x ranks strictly better than y under (llr desc, col asc) */
static inline int better(const cand_t *x, const cand_t *y) {
  return x->llr > y->llr || (x->llr == y->llr && x->col < y->col);
}
static int cmp_cand(const void *a, const void *b) {
  const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
  return better(x, y) ? -1 : (better(y, x) ? 1 : 0);
}
/* min-heap on `better` (root = worst kept) */
static void heap_sift_down(cand_t *h, int n, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, w = i;
    if (l < n && better(&h[w], &h[l])) w = l;
    if (r < n && better(&h[w], &h[r])) w = r;
    if (w == i) return;
    cand_t t = h[i];
    h[i] = h[w];
    h[w] = t;
    i = w;
  }
}
static void heap_sift_up(cand_t *h, int i) {
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!better(&h[p], &h[i])) return;
    cand_t t = h[i];
    h[i] = h[p];
    h[p] = t;
    i = p;
  }
}

static int similarity(const int64_t *at_ptr, const int32_t *at_idx, int32_t n_items_a,
                      const int64_t *b_rp, const int32_t *b_ci, int32_t n_items_b, int64_t n_users,
                      const int32_t *marg_a, const int32_t *marg_b, int self, const orc_params_t *prm,
                      int flags, int n_threads, orc_result_t *out) {
  if (prm->top_k < 1) FAIL("top_k must be >= 1");
  int32_t stride = prm->top_k < n_items_b ? prm->top_k : n_items_b;
  if (stride < 1) stride = 1;
  cand_t *kept = (cand_t *)malloc(sizeof(cand_t) * (size_t)n_items_a * (size_t)stride);
  int32_t *len = (int32_t *)calloc((size_t)n_items_a + 1, sizeof(int32_t));
  if (!kept || !len) FAIL("out of memory");
  int64_t products = 0, distinct = 0;
  int oom = 0;
#ifdef _OPENMP
  if (n_threads < 1) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads) reduction(+ : products, distinct)
  {
    int32_t *acc = (int32_t *)calloc((size_t)(n_items_b > 0 ? n_items_b : 1), sizeof(int32_t));
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_items_b > 0 ? n_items_b : 1));
    cand_t *heap = (cand_t *)malloc(sizeof(cand_t) * (size_t)stride);
    if (!acc || !touched || !heap) {
#pragma omp atomic write
      oom = 1;
    }
#pragma omp barrier
    if (!oom) {
#pragma omp for schedule(dynamic, 64)
      for (int32_t item = 0; item < n_items_a; ++item) {
        int32_t nt = 0;
        for (int64_t p = at_ptr[item]; p < at_ptr[item + 1]; ++p) {
          int64_t u = at_idx[p];
          int64_t s = b_rp[u], e = b_rp[u + 1];
          products += e - s;
          for (int64_t q = s; q < e; ++q) {
            int32_t j = b_ci[q];
            if (acc[j]++ == 0) touched[nt++] = j;
          }
        }
        distinct += nt;
        int hn = 0;
        for (int32_t t = 0; t < nt; ++t) {
          int32_t j = touched[t];
          int32_t k11 = acc[j];
          acc[j] = 0;
          if (self && j == item) continue; /* crossCooccurrence || thingB != thingA */
          cand_t c;
          c.llr = llr_from_marginals(marg_a[item], marg_b[j], k11, n_users, flags);
          c.col = j;
          c.cnt = k11;
          if (prm->has_min_llr && !(c.llr >= prm->min_llr)) continue;
          if (!(c.llr > 0.0)) continue; /* llrBlock(row, col) = 0.0 stores nothing */
          if (hn < stride) {
            heap[hn] = c;
            heap_sift_up(heap, hn);
            ++hn;
          } else if (better(&c, &heap[0])) {
            heap[0] = c;
            heap_sift_down(heap, hn, 0);
          }
        }
        qsort(heap, (size_t)hn, sizeof(cand_t), cmp_cand);
        memcpy(kept + (size_t)item * stride, heap, sizeof(cand_t) * (size_t)hn);
        len[item] = hn;
      }
    }
    free(acc);
    free(touched);
    free(heap);
  }
  if (oom) {
    free(kept);
    free(len);
    FAIL("out of memory");
  }
  out->n_rows = n_items_a;
  out->n_cols = n_items_b;
  out->row_ptr = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_items_a + 1));
  if (!out->row_ptr) FAIL("out of memory");
  out->row_ptr[0] = 0;
  for (int32_t i = 0; i < n_items_a; ++i) out->row_ptr[i + 1] = len[i];
  prefix_sum_i64(out->row_ptr, n_items_a);
  const int64_t total = out->row_ptr[n_items_a];
  out->col_idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
  out->llr = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
  out->count = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
  if (!out->col_idx || !out->llr || !out->count) FAIL("out of memory");
#pragma omp parallel for schedule(static, 1024)
  for (int32_t i = 0; i < n_items_a; ++i) {
    int64_t w = out->row_ptr[i];
    for (int32_t t = 0; t < len[i]; ++t, ++w) {
      const cand_t *c = &kept[(size_t)i * stride + t];
      out->col_idx[w] = c->col;
      out->llr[w] = c->llr;
      out->count[w] = c->cnt;
    }
  }
  out->products = products;
  out->distinct_cells = distinct;
  free(kept);
  free(len);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * crossOccurrenceDownsampled(datasets, randomSeed) / cooccurrencesIDSs(...)
 * [Mahout SimilarityAnalysis; called from URAlgorithm.scala:323-329, 343-346]
 *   A' = sampleDownAndBinarize(A, seed, m_0); N = A'.nrow; colA = nnzPerColumn(A')
 *   out[0] = computeSimilarities(A'^T A', N, k_0, colA, colA, self, minLLR_0)
 *   out[i] = computeSimilarities(A'^T B'_i, N, k_i, colA, colB_i, cross, minLLR_i),
 *            B'_i = sampleDownAndBinarize(B_i, seed, m_i)
 * ------------------------------------------------------------------------------------------- */
int orc_train(int n_mats, const orc_csr_t *mats, const orc_params_t *params, int32_t seed, int flags,
              int n_threads, orc_result_t *results) {
  if (n_mats < 1) FAIL("need at least the primary matrix");
  orc_set_threads(n_threads); /* every parallel region of this call, not only the similarity loop */
  memset(results, 0, sizeof(orc_result_t) * (size_t)n_mats);
  for (int i = 0; i < n_mats; ++i) {
    if (validate(&mats[i])) return -1;
    if (mats[i].n_rows != mats[0].n_rows) FAIL("matrix %d: row cardinality differs from the primary", i);
    if (params[i].max_interactions < 1) FAIL("matrix %d: max_interactions must be >= 1", i);
    if (params[i].top_k < 1) FAIL("matrix %d: top_k must be >= 1", i);
  }
  const int64_t n_users = mats[0].n_rows;
  int64_t *a_rp = NULL, *at_ptr = NULL;
  int32_t *a_ci = NULL, *at_idx = NULL, *marg_a = NULL;
  int rc = -1;
  {
    int64_t *c_rp;
    int32_t *c_ci;
    if (orc_canonicalize(&mats[0], &c_rp, &c_ci)) return -1;
    orc_csr_t canon = {mats[0].n_rows, mats[0].n_cols, c_rp, c_ci};
    marg_a = (int32_t *)malloc(sizeof(int32_t) * (size_t)(mats[0].n_cols > 0 ? mats[0].n_cols : 1));
    int r = orc_downsample(&canon, params[0].max_interactions, seed, flags, &a_rp, &a_ci, NULL, marg_a);
    free(c_rp);
    free(c_ci);
    if (r) goto done;
    if (transpose(a_rp, a_ci, n_users, mats[0].n_cols, &at_ptr, &at_idx)) goto done;
  }
  for (int i = 0; i < n_mats; ++i) {
    int64_t *b_rp = a_rp;
    int32_t *b_ci = a_ci, *marg_b = marg_a;
    if (i > 0) {
      int64_t *c_rp;
      int32_t *c_ci;
      if (orc_canonicalize(&mats[i], &c_rp, &c_ci)) goto done;
      orc_csr_t canon = {mats[i].n_rows, mats[i].n_cols, c_rp, c_ci};
      marg_b = (int32_t *)malloc(sizeof(int32_t) * (size_t)(mats[i].n_cols > 0 ? mats[i].n_cols : 1));
      int r = orc_downsample(&canon, params[i].max_interactions, seed, flags, &b_rp, &b_ci, NULL, marg_b);
      free(c_rp);
      free(c_ci);
      if (r) {
        free(marg_b);
        goto done;
      }
    }
    int r = similarity(at_ptr, at_idx, mats[0].n_cols, b_rp, b_ci, mats[i].n_cols, n_users, marg_a, marg_b,
                       i == 0, &params[i], flags, n_threads, &results[i]);
    results[i].nnz_a = a_rp[n_users];
    results[i].nnz_b = b_rp[n_users];
    if (i > 0) {
      free(b_rp);
      free(b_ci);
      free(marg_b);
    }
    if (r) goto done;
  }
  rc = 0;
done:
  free(a_rp);
  free(a_ci);
  free(at_ptr);
  free(at_idx);
  free(marg_a);
  if (rc)
    for (int i = 0; i < n_mats; ++i) orc_free_result(&results[i]);
  return rc;
}

/* ---------------------------------------------------------------------------------------------
 * orc_ingest: Preparator.prepare (Preparator.scala:44-87) with IndexedDatasetSpark.apply(elements, minEventsPerUser)
 * (:102-158, user filter counting duplicate events :129-132) and apply(elements, existingRowIDs) (:160-214, events
 * of unknown users dropped :175-178, item ids from the surviving events :184, setQuick dedup :201-208).
 * ------------------------------------------------------------------------------------------- */
void orc_free_ingested(orc_ingested_t *r) {
  if (!r) return;
  free(r->row_ptr);
  free(r->col_idx);
  free(r->item_map);
  memset(r, 0, sizeof *r);
}

int orc_ingest(int n_types, const orc_events_t *ev, int64_t n_users_raw, int32_t min_events_per_user, int32_t *user_map,
               orc_ingested_t *out) {
  if (n_types < 1 || !ev || !user_map || !out) FAIL("bad argument");
  if (n_users_raw < 0 || n_users_raw > 0x7fffffff) FAIL("n_users_raw out of range");
  memset(out, 0, sizeof(orc_ingested_t) * (size_t)n_types);
  for (int t = 0; t < n_types; ++t)
    for (int64_t i = 0; i < ev[t].n_events; ++i) {
      if (ev[t].user[i] < 0 || ev[t].user[i] >= n_users_raw) FAIL("type %d: user id out of range at %lld", t, (long long)i);
      if (ev[t].item[i] < 0 || ev[t].item[i] >= ev[t].n_items_raw) FAIL("type %d: item id out of range at %lld", t, (long long)i);
    }
  /* user dictionary from the primary events (duplicates count) */
  int64_t *cnt = (int64_t *)calloc((size_t)(n_users_raw > 0 ? n_users_raw : 1), sizeof(int64_t));
  if (!cnt) FAIL("out of memory");
  for (int64_t i = 0; i < ev[0].n_events; ++i) cnt[ev[0].user[i]]++;
  const int64_t need = min_events_per_user > 1 ? min_events_per_user : 1;
  int32_t n_users = 0;
  for (int64_t u = 0; u < n_users_raw; ++u) user_map[u] = cnt[u] >= need ? n_users++ : -1;
  free(cnt);
  for (int t = 0; t < n_types; ++t) {
    orc_ingested_t *o = &out[t];
    const int32_t ni = ev[t].n_items_raw;
    o->item_map = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ni > 0 ? ni : 1));
    o->row_ptr = (int64_t *)calloc((size_t)n_users + 1, sizeof(int64_t));
    if (!o->item_map || !o->row_ptr) FAIL("out of memory");
    /* item dictionary: items with a surviving event */
    for (int32_t j = 0; j < ni; ++j) o->item_map[j] = -1;
    for (int64_t i = 0; i < ev[t].n_events; ++i)
      if (user_map[ev[t].user[i]] >= 0) o->item_map[ev[t].item[i]] = 0;
    int32_t n_items = 0;
    for (int32_t j = 0; j < ni; ++j)
      if (o->item_map[j] == 0) o->item_map[j] = n_items++;
    o->n_rows = n_users;
    o->n_cols = n_items;
    /* bucket the surviving events by user, then sort + dedup each row */
    for (int64_t i = 0; i < ev[t].n_events; ++i) {
      int32_t r = user_map[ev[t].user[i]];
      if (r >= 0) o->row_ptr[r + 1]++;
    }
    for (int32_t r = 0; r < n_users; ++r) o->row_ptr[r + 1] += o->row_ptr[r];
    int64_t kept = o->row_ptr[n_users];
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(kept > 0 ? kept : 1));
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_users + 1));
    o->col_idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(kept > 0 ? kept : 1));
    if (!tmp || !cur || !o->col_idx) FAIL("out of memory");
    memcpy(cur, o->row_ptr, sizeof(int64_t) * ((size_t)n_users + 1));
    for (int64_t i = 0; i < ev[t].n_events; ++i) {
      int32_t r = user_map[ev[t].user[i]];
      if (r >= 0) tmp[cur[r]++] = o->item_map[ev[t].item[i]];
    }
    int64_t w = 0;
    for (int32_t r = 0; r < n_users; ++r) {
      int64_t s = o->row_ptr[r], e = o->row_ptr[r + 1];
      qsort(tmp + s, (size_t)(e - s), sizeof(int32_t), cmp_i32);
      o->row_ptr[r] = w;
      for (int64_t i = s; i < e; ++i)
        if (i == s || tmp[i] != tmp[i - 1]) o->col_idx[w++] = tmp[i];
    }
    o->row_ptr[n_users] = w;
    free(tmp);
    free(cur);
  }
  return 0;
}

void orc_free_result(orc_result_t *r) {
  if (!r) return;
  free(r->row_ptr);
  free(r->col_idx);
  free(r->llr);
  free(r->count);
  memset(r, 0, sizeof *r);
}
