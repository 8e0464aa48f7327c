/* Plain-C consumer of include/cco_b200.h: proves the boundary is a C ABI (no C++/CUDA/torch types) and that a box
 * without a usable B200 gets a loud CCO_E_CUDA, never a CPU fallback.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "cco_b200.h"

int main(void) {
  if (cco_abi_version() != CCO_ABI_VERSION) return 10;
  if (strcmp(cco_status_string(CCO_E_SHAPE_MISMATCH), "shape mismatch") != 0) return 11;
  int32_t bounds[3];
  int64_t prefix[5] = {0, 10, 10, 30, 40};
  if (cco_partition_rows(prefix, 4, 2, bounds) != CCO_OK || bounds[0] != 0 || bounds[2] != 4) return 12;
  cco_config_t cfg = {0, 0, 1, 0, NULL};
  cco_ctx_t *ctx = NULL;
  int st = cco_create(&cfg, &ctx);
  if (st == CCO_OK) { /* a B200 is present: run one tiny train through the C ABI */
    int64_t rp[4] = {0, 2, 3, 5};
    int32_t ci[5] = {0, 1, 1, 0, 2};
    cco_csr_t m = {3, 3, rp, ci};
    cco_indicator_params_t p = {500, 50, 0, 0.0};
    cco_result_t *res = NULL;
    st = cco_train(ctx, 1, &m, &p, 1, 0, &res);
    if (st != CCO_OK) { printf("train failed: %s\n", cco_last_error()); return 13; }
    int64_t n_rows; int32_t n_cols; const int64_t *orp; const int32_t *oci; const double *llr; const int32_t *cnt;
    if (cco_result_matrix(res, 0, &n_rows, &n_cols, &orp, &oci, &llr, &cnt) != CCO_OK || n_rows != 3 || n_cols != 3) return 14;
    cco_result_free(res);
    cco_destroy(ctx);
    printf("gpu ok\n");
    return 0;
  }
  if (st != CCO_E_CUDA) return 15;
  if (strstr(cco_last_error(), "no CPU fallback") == NULL && strstr(cco_last_error(), "sm_100a") == NULL) return 16;
  printf("no gpu: %s\n", cco_last_error());
  return 0;
}
