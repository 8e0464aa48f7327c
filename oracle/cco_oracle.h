/*
 * cco_oracle.h -- CPU restatement of the Correlated Cross-Occurrence (CCO) train path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker / CPU baseline.
 *
 * What it restates: Apache Mahout 0.13.0 `SimilarityAnalysis.crossOccurrenceDownsampled`
 * / `cooccurrencesIDSs` (math-scala) + `LogLikelihood.logLikelihoodRatio` (mahout-math),
 * the third-party code behind the two call sites
 *   /root/reference/src/main/scala/URAlgorithm.scala:323-329  (cooccurrencesIDSs)
 *   /root/reference/src/main/scala/URAlgorithm.scala:343-346  (crossOccurrenceDownsampled)
 * Mahout is pinned at 0.13.0 in /root/reference/build.sbt:15,34-38 and is NOT vendored in
 * the reference tree, so the algorithm is restated from its published source (SURVEY.md
 * Appendix A) and anchored on the reference's call sites, fixtures and expected files.
 *
 * PARITY PINNING STATUS: "parity unpinned" numerically -- the reference has no test that
 * asserts a co-occurrence count, an LLR value or an indicator list, and there is no JVM in
 * this image to run Mahout.  What IS pinned (tests/test_oracle_golden.py):
 *   - six LLR known-answer values from Mahout's LogLikelihoodTest;
 *   - indicator MEMBERSHIP for data/sample-handmade-item-set-data.txt implied by
 *     data/integration-test-item-set-expected.txt:16-40;
 *   - the zero/non-zero score pattern of data/integration-test-expected.txt:16-54
 *     (needs N=3 after minEventsPerUser, LLR==0 cells absent).
 */
#ifndef CCO_ORACLE_H
#define CCO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* binary user x item matrix, CSR, values implicit 1 (Preparator.scala:201-208: setQuick(idx, 1.0)) */
typedef struct {
  int64_t n_rows;         /* U: size of the shared user dictionary (Preparator.scala:213) */
  int32_t n_cols;         /* I: size of this event type's item dictionary */
  const int64_t *row_ptr; /* [n_rows+1] */
  const int32_t *col_idx; /* [nnz]; any order inside a row, duplicates collapse */
} orc_csr_t;

/* DownsamplableCrossOccurrenceDataset(iD, maxElementsPerRow, maxInterestingElements, minLLROpt)
 * built at URAlgorithm.scala:336-340 */
typedef struct {
  int32_t max_interactions; /* m: maxItemsPerUser / maxEventsPerEventType, default 500 */
  int32_t top_k;            /* k: maxCorrelatorsPerItem / maxCorrelatorsPerEventType, default 50 */
  int32_t has_min_llr;      /* Option[Double] */
  double min_llr;
} orc_params_t;

enum {
  ORC_FLAG_ROWRATE_INTDIV = 1, /* literal Mahout Int/Int row sample rate */
  ORC_FLAG_ENTROPY_VARARGS = 2 /* entropy(...) = xLogX(sum) - (sum of xLogX) instead of left-to-right */
};

typedef struct {
  int64_t n_rows; /* = I_A (primary items) */
  int32_t n_cols; /* = I_B of this event type */
  int64_t *row_ptr;
  int32_t *col_idx; /* sorted per row by (llr desc, col asc) */
  double *llr;
  int32_t *count; /* k11 of each kept cell */
  int64_t products;       /* P(A',B') = sum_u degA'(u) * degB'(u) */
  int64_t distinct_cells; /* nnz(A'^T B') (diagonal included for A^T A) */
  int64_t nnz_a;          /* nnz(A') after downsampling */
  int64_t nnz_b;          /* nnz(B') after downsampling */
} orc_result_t;

double orc_xlogx(int64_t x);
double orc_llr(int64_t k11, int64_t k12, int64_t k21, int64_t k22, int flags);
uint64_t orc_hash64(int32_t seed, int64_t u, int32_t j);
double orc_u01(uint64_t h);

/* sort + dedup each row; returns 0 or <0 on malformed input. Outputs malloc'd (free with orc_free). */
int orc_canonicalize(const orc_csr_t *in, int64_t **row_ptr, int32_t **col_idx);

/* Mahout sampleDownAndBinarize with the repo's deterministic counter-based sampler.
 * `in` must be canonical.  raw_col_counts / new_col_counts may be NULL. */
int orc_downsample(const orc_csr_t *in, int32_t m, int32_t seed, int flags, int64_t **row_ptr,
                   int32_t **col_idx, int32_t *raw_col_counts, int32_t *new_col_counts);

/* full integer co-occurrence matrix C = A^T B (canonical inputs), CSR over items of A,
 * columns ascending; diagonal kept.  For tests only (O(nnz(C)) memory). */
int orc_cooccurrence(const orc_csr_t *a, const orc_csr_t *b, int64_t **row_ptr, int32_t **col_idx,
                     int32_t **count);

/* the whole path: downsample every matrix, A'^T A' and A'^T B'_i, LLR, top-k.
 * results: caller-provided array of n_mats entries, filled; free each with orc_free_result. */
int orc_train(int n_mats, const orc_csr_t *mats, const orc_params_t *params, int32_t seed, int flags,
              int n_threads, orc_result_t *results);

/* ---- next row: the ingest that sits right before the boundary --------------------------------------
 * Preparator.prepare + IndexedDatasetSpark.apply (Preparator.scala:44-87, 100-216) on integer-tokenised events.
 * Type 0 is the primary event.  The user dictionary is fixed by the primary events (users with at least
 * min_events_per_user primary events, duplicates counted; min 0/1 = every user with a primary event); every type is
 * restricted to those users; each type's item dictionary holds only items that still have an event; duplicates collapse.
 * Dictionary order here: ascending raw id (Mahout's is the arbitrary order of distinct().collect()). */
typedef struct {
  int64_t n_events;
  const int64_t *user; /* raw user id in [0, n_users_raw) */
  const int32_t *item; /* raw item id in [0, n_items_raw) */
  int32_t n_items_raw;
} orc_events_t;

typedef struct {
  int64_t n_rows;    /* = size of the user dictionary, equal for every type */
  int32_t n_cols;    /* = size of this type's item dictionary */
  int64_t *row_ptr;  /* binary CSR, columns ascending */
  int32_t *col_idx;
  int32_t *item_map; /* [n_items_raw]: new column id or -1 */
} orc_ingested_t;

/* user_map: caller-provided [n_users_raw], receives the new row id or -1.  out: caller-provided [n_types]. */
int orc_ingest(int n_types, const orc_events_t *ev, int64_t n_users_raw, int32_t min_events_per_user, int32_t *user_map,
               orc_ingested_t *out);
void orc_free_ingested(orc_ingested_t *r);

void orc_free_result(orc_result_t *r);
void orc_free(void *p);
int orc_max_threads(void);
/* thread count of every OpenMP region of later calls on this thread (0 = OpenMP default); orc_train calls it with its
 * n_threads argument, so OMP_NUM_THREADS=1 exported by a launcher (torchrun) does not serialise the CPU arm */
void orc_set_threads(int n);
const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
