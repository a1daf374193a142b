/* Plain-C client of the C ABI (include/deepmod_b200.h): proves the header is C (not only C++) and shows the call
 * sequence a non-Python host makes.  Without a GPU dm_create fails with DM_ERR_CUDA and says so -- there is no CPU path.
 *
 *   gcc -std=c99 -Wall -Werror -Iinclude examples/c_abi_probe.c -Ldeepmod_b200 -ldeepmod_b200 -Wl,-rpath,$PWD/deepmod_b200 -o /tmp/probe
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "deepmod_b200.h"

int main(void) {
  printf("dm_version %d\n", dm_version());
  dm_ctx* ctx = NULL;
  if (dm_create(&ctx, 0, NULL, DM_FP32) != DM_ERR_ARG) return 2;          /* bad argument: a status code, never a crash */
  /* the 14 tensors in the reference layout (here: zeros) */
  static float k0[107 * 400], k12[200 * 400], b[400], cw[200 * 2], cb[2];
  dm_weights w;
  memset(&w, 0, sizeof(w));
  for (int d = 0; d < 2; ++d)
    for (int l = 0; l < 3; ++l) { w.kernel[d][l] = l == 0 ? k0 : k12; w.bias[d][l] = b; }
  w.cls_w = cw; w.cls_b = cb;
  int rc = dm_create(&ctx, 0, &w, DM_F16);
  if (rc != DM_OK) {
    printf("dm_create: status %d: %s\n", rc, dm_last_error(NULL));
    return rc == DM_ERR_CUDA ? 0 : 3;                                      /* expected on a box without a B200 */
  }
  /* with a device: one empty detect call, the accumulator of a 1 kb contig, the exchange step of a 1-rank job */
  const int64_t len = 1000;
  dm_batch batch;
  memset(&batch, 0, sizeof(batch));
  int64_t rows = -1;
  uint64_t cov = 1, mod = 1, n = 1, chk = 1;
  rc = dm_set_genome(ctx, 1, &len, 'C');
  if (rc == DM_OK) rc = dm_detect_batch(ctx, &batch, NULL, NULL, NULL);
  if (rc == DM_OK) rc = dm_reduce_comm(ctx, NULL, 0, 1);
  if (rc == DM_OK) rc = dm_hist_totals(ctx, &cov, &mod, &n, &chk);
  if (rc == DM_OK) rc = dm_hist_nonzero(ctx, 0, 1, 0, NULL, NULL, NULL, &rows);
  printf("status %d, rows %lld, totals %llu %llu %llu %llu, %lld kernel launches\n", rc, (long long)rows, (unsigned long long)cov,
         (unsigned long long)mod, (unsigned long long)n, (unsigned long long)chk, (long long)dm_launch_count(ctx));
  dm_destroy(ctx);
  return rc == DM_OK && rows == 0 && cov == 0 ? 0 : 4;
}
