// cblas_?syrk / cblas_?herk entry points (BlasOps.syrk / herk, /root/reference/Numeric/Jalla/Foreign/BlasOps.hs:54,73):
// C <- alpha*op(A)*op(A)^T + beta*C on one triangle.  They are the trailing update of the blocked Cholesky, exported
// because SURVEY.md §8(f) rank 3 lists them as the next level-3 rows behind the same class.  Implemented on the GEMM
// kernels' triangle mode (only tiles touching the `uplo` triangle are computed, only its elements written).
#include "jb_common.cuh"

namespace jb {
// herm: 0 = syrk (A A^T), 1 = herk (A A^H, real alpha/beta, imaginary part of the diagonal forced to zero)
template <typename T>
__global__ void zero_diag_imag_kernel(int n, T *C, int ldc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { T v = C[(size_t)i + (size_t)i * ldc]; C[(size_t)i + (size_t)i * ldc] = Num<T>::make(Num<T>::re(v), 0); }
}

template <typename T>
static void syrk_entry(const char *name, int herm, int order, int uplo, int trans, int n, int k, T alpha, const T *A, int lda,
                       T beta, T *C, int ldc) {
  int bad = 0;
  if (order != 101 && order != 102) bad = 1;
  else if (uplo != 121 && uplo != 122) bad = 2;
  else if (trans < 111 || trans > 113) bad = 3;
  else if (n < 0) bad = 4;
  else if (k < 0) bad = 5;
  if (!bad && order == 101) {   // row-major: the stored triangle flips and op flips (C^T = C for the referenced part)
    uplo = (uplo == 121) ? 122 : 121;
    trans = (trans == 111) ? (herm ? 113 : 112) : 111;
  }
  if (!bad) {
    const int rows_a = (trans == 111) ? n : k;
    if (lda < (rows_a > 1 ? rows_a : 1)) bad = 8;
    else if (ldc < (n > 1 ? n : 1)) bad = 11;
  }
  if (bad) { set_error(JB_ERR_ARG, "%s: parameter %d had an illegal value", name, bad); return; }
  if (n == 0) return;
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  const long ra = (trans == 111) ? n : k, ca = (trans == 111) ? k : n;
  Staged sa, sc;
  if (sa.open(A, ra, k > 0 ? ca : 0, lda, sizeof(T), true, s)) return;
  if (sc.open(C, n, n, ldc, sizeof(T), true, s)) { sa.close(false, s); return; }
  // op(A) op(A)^T (or ^H): second operand is the same matrix with the opposite transpose flag
  const int ta = (trans == 111) ? 111 : (herm ? 113 : 112);
  const int tb = (trans == 111) ? (herm ? 113 : 112) : 111;
  int rc = gemm_device<T>(ta, tb, n, n, k, alpha, (const T *)sa.dev, sa.ld, (const T *)sa.dev, sa.ld, beta, (T *)sc.dev, sc.ld, s,
                          uplo == 122 ? 'L' : 'U');
  if (!rc && herm && Num<T>::cplx) {
    zero_diag_imag_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, (T *)sc.dev, sc.ld);
    count_launch();
  }
  bool host = sa.staged || sc.staged;
  sa.close(false, s);
  sc.close(rc == 0, s);
  if (host) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) set_error(JB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  }
}
template <typename T> static int get_scalar(const void *p, T *out) {
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
void jb_cblas_ssyrk(int order, int uplo, int trans, int n, int k, float alpha, const float *a, int lda, float beta, float *c, int ldc) {
  syrk_entry<float>("cblas_ssyrk", 0, order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc);
}
void jb_cblas_dsyrk(int order, int uplo, int trans, int n, int k, double alpha, const double *a, int lda, double beta, double *c, int ldc) {
  syrk_entry<double>("cblas_dsyrk", 0, order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc);
}
void jb_cblas_csyrk(int order, int uplo, int trans, int n, int k, const void *alpha, const void *a, int lda, const void *beta, void *c, int ldc) {
  cfloat al, be; if (get_scalar(alpha, &al) || get_scalar(beta, &be)) return;
  syrk_entry<cfloat>("cblas_csyrk", 0, order, uplo, trans, n, k, al, (const cfloat *)a, lda, be, (cfloat *)c, ldc);
}
void jb_cblas_zsyrk(int order, int uplo, int trans, int n, int k, const void *alpha, const void *a, int lda, const void *beta, void *c, int ldc) {
  cdouble al, be; if (get_scalar(alpha, &al) || get_scalar(beta, &be)) return;
  syrk_entry<cdouble>("cblas_zsyrk", 0, order, uplo, trans, n, k, al, (const cdouble *)a, lda, be, (cdouble *)c, ldc);
}
void jb_cblas_cherk(int order, int uplo, int trans, int n, int k, float alpha, const void *a, int lda, float beta, void *c, int ldc) {
  syrk_entry<cfloat>("cblas_cherk", 1, order, uplo, trans, n, k, cfloat{alpha, 0.f}, (const cfloat *)a, lda, cfloat{beta, 0.f}, (cfloat *)c, ldc);
}
void jb_cblas_zherk(int order, int uplo, int trans, int n, int k, double alpha, const void *a, int lda, double beta, void *c, int ldc) {
  syrk_entry<cdouble>("cblas_zherk", 1, order, uplo, trans, n, k, cdouble{alpha, 0.0}, (const cdouble *)a, lda, cdouble{beta, 0.0}, (cdouble *)c, ldc);
}
void cblas_ssyrk(int order, int uplo, int trans, int n, int k, float alpha, const float *a, int lda, float beta, float *c, int ldc) { jb_cblas_ssyrk(order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
void cblas_dsyrk(int order, int uplo, int trans, int n, int k, double alpha, const double *a, int lda, double beta, double *c, int ldc) { jb_cblas_dsyrk(order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
void cblas_csyrk(int order, int uplo, int trans, int n, int k, const void *alpha, const void *a, int lda, const void *beta, void *c, int ldc) { jb_cblas_csyrk(order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
void cblas_zsyrk(int order, int uplo, int trans, int n, int k, const void *alpha, const void *a, int lda, const void *beta, void *c, int ldc) { jb_cblas_zsyrk(order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
void cblas_cherk(int order, int uplo, int trans, int n, int k, float alpha, const void *a, int lda, float beta, void *c, int ldc) { jb_cblas_cherk(order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
void cblas_zherk(int order, int uplo, int trans, int n, int k, double alpha, const void *a, int lda, double beta, void *c, int ldc) { jb_cblas_zherk(order, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
}
