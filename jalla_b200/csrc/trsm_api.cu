// cblas_?trsm entry points (BlasOps.trsm, /root/reference/Numeric/Jalla/Foreign/BlasOps.hs:57; raw imports
// Foreign/BLAS.chs:150-178 region) — a level-3 sibling of GEMM that the blocked factorisations and the
// distributed Cholesky use; exported because SURVEY.md §8(f) lists it as the next row behind the same class.
#include "jb_common.cuh"

namespace jb {
template <typename T>
static void trsm_entry(const char *name, int order, int side, int uplo, int trans, int diag, int m, int n, T alpha,
                       const T *A, int lda, T *B, int ldb) {
  int bad = 0;
  if (order != 101 && order != 102) bad = 1;
  else if (side != 141 && side != 142) bad = 2;
  else if (uplo != 121 && uplo != 122) bad = 3;
  else if (trans < 111 || trans > 113) bad = 4;
  else if (diag != 131 && diag != 132) bad = 5;
  else if (m < 0) bad = 6;
  else if (n < 0) bad = 7;
  if (!bad && order == 101) {   // row-major B is the column-major B^T: flip side and triangle, swap m and n
    side = (side == 141) ? 142 : 141;
    uplo = (uplo == 121) ? 122 : 121;
    int t = m; m = n; n = t;
  }
  if (!bad) {
    const int ka = (side == 141) ? m : n;
    if (lda < (ka > 1 ? ka : 1)) bad = 10;
    else if (ldb < (m > 1 ? m : 1)) bad = 12;
  }
  if (bad) { set_error(JB_ERR_ARG, "%s: parameter %d had an illegal value", name, bad); return; }
  if (m == 0 || n == 0) return;
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  const int ka = (side == 141) ? m : n;
  Staged sa, sb;
  if (sa.open(A, ka, ka, lda, sizeof(T), true, s)) return;
  if (sb.open(B, m, n, ldb, sizeof(T), true, s)) { sa.close(false, s); return; }
  int rc = trsm_device<T>(side, uplo, trans, diag, m, n, alpha, (const T *)sa.dev, sa.ld, (T *)sb.dev, sb.ld, s);
  bool host = sa.staged || sb.staged;
  sa.close(false, s);
  sb.close(rc == 0, s);
  if (host) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) set_error(JB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  }
}
template <typename T> static int scalar_from(const void *p, T *out) {
  if (!p) { set_error(JB_ERR_ARG, "null scalar pointer"); return 1; }
  if (classify(p) == Residency::Device) {
    if (ensure_device()) return 1;
    if (cudaMemcpy(out, p, sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error(JB_ERR_CUDA, "scalar read failed"); return 1; }
  } else memcpy(out, p, sizeof(T));
  return 0;
}
}  // namespace jb

using namespace jb;
extern "C" {
void jb_cblas_strsm(int order, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *a, int lda, float *b, int ldb) {
  trsm_entry<float>("cblas_strsm", order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb);
}
void jb_cblas_dtrsm(int order, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *a, int lda, double *b, int ldb) {
  trsm_entry<double>("cblas_dtrsm", order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb);
}
void jb_cblas_ctrsm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) {
  cfloat al; if (scalar_from(alpha, &al)) return;
  trsm_entry<cfloat>("cblas_ctrsm", order, side, uplo, trans, diag, m, n, al, (const cfloat *)a, lda, (cfloat *)b, ldb);
}
void jb_cblas_ztrsm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) {
  cdouble al; if (scalar_from(alpha, &al)) return;
  trsm_entry<cdouble>("cblas_ztrsm", order, side, uplo, trans, diag, m, n, al, (const cdouble *)a, lda, (cdouble *)b, ldb);
}
void cblas_strsm(int order, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *a, int lda, float *b, int ldb) {
  jb_cblas_strsm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb);
}
void cblas_dtrsm(int order, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *a, int lda, double *b, int ldb) {
  jb_cblas_dtrsm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb);
}
void cblas_ctrsm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) {
  jb_cblas_ctrsm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb);
}
void cblas_ztrsm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) {
  jb_cblas_ztrsm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb);
}
}
