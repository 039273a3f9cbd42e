/* Replays, in plain C against the link-level symbols, what the UNMODIFIED Jalla library issues for
 *     solveLinearSystem a b          (/root/reference/Numeric/Jalla/Matrix.hs:513-521)
 * when `a` and `b` live in a GPU-resident CMatrix: matrixCopy b and matrixCopy a are loops of one strided cblas_dcopy per
 * row (Matrix.hs:855-871 through CVector.hs:284-288), then LAPACKE_dgesv in ColumnMajor with a host ipiv
 * (allocaArray, Matrix.hs:502).  All matrix pointers are jb_malloc'd DEVICE memory.  Exit 0 = the solution satisfies
 * |A x - b| <= 1e-10, 77 = no GPU (the library must fail loudly then), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "jalla_b200.h"

int main(void) {
  const int n = 200, nrhs = 3;
  if (jb_init(-1) != 0) { printf("no GPU: %s\n", jb_last_error_string()); return 77; }
  double *A = malloc(sizeof(double) * n * n), *B = malloc(sizeof(double) * n * nrhs), *X = malloc(sizeof(double) * n * nrhs);
  unsigned long long s = 88172645463325252ull;
  for (int i = 0; i < n * n + n * nrhs; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double v = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    if (i < n * n) A[i] = v; else B[i - n * n] = v;
  }
  double *dA = jb_malloc(sizeof(double) * n * n, JB_MEM_DEVICE), *dB = jb_malloc(sizeof(double) * n * nrhs, JB_MEM_DEVICE);
  double *dA2 = jb_malloc(sizeof(double) * n * n, JB_MEM_DEVICE), *dX = jb_malloc(sizeof(double) * n * nrhs, JB_MEM_DEVICE);
  if (!dA || !dB || !dA2 || !dX) { printf("jb_malloc: %s\n", jb_last_error_string()); return 2; }
  if (jb_memcpy_h2d(dA, A, sizeof(double) * n * n) || jb_memcpy_h2d(dB, B, sizeof(double) * n * nrhs)) return 3;
  for (int i = 0; i < n; i++) cblas_dcopy(nrhs, dB + i, n, dX + i, n);      /* matrixCopy b NoTrans */
  for (int i = 0; i < n; i++) cblas_dcopy(n, dA + i, n, dA2 + i, n);        /* matrixCopy a NoTrans */
  if (jb_last_error()) { printf("cblas_dcopy: %s\n", jb_last_error_string()); return 4; }
  int *ipiv = malloc(sizeof(int) * n);
  int info = LAPACKE_dgesv(102, n, nrhs, dA2, n, ipiv, dX, n);
  if (info != 0) { printf("LAPACKE_dgesv info %d: %s\n", info, jb_last_error_string()); return 5; }
  if (jb_memcpy_d2h(X, dX, sizeof(double) * n * nrhs)) return 6;
  double worst = 0;
  for (int c = 0; c < nrhs; c++)
    for (int i = 0; i < n; i++) {
      double r = -B[i + c * n];
      for (int j = 0; j < n; j++) r += A[i + j * n] * X[j + c * n];
      if (fabs(r) > worst) worst = fabs(r);
    }
  jb_free(dA); jb_free(dB); jb_free(dA2); jb_free(dX);
  if (!(worst <= 1e-10)) { printf("residual %g\n", worst); return 7; }
  printf("solve sequence ok: residual %.2e, %llu kernel launches\n", worst, (unsigned long long)jb_kernel_launches());
  return 0;
}
