// The remaining level-3 siblings of GEMM behind Jalla's BlasOps class and the helpers SURVEY.md §8(f) marks "next":
//   cblas_?trmm, cblas_?symm, cblas_{c,z}hemm   (BlasOps.trmm / symm, BlasOpsComplex.hemm: Foreign/BlasOps.hs:53-57,72;
//                                                raw imports Foreign/BLAS.chs:150-176)
//   LAPACKE_?potri                               (Foreign/LAPACKE.chs:737-740, lapacke.h:2853-2863)
//   jb_?laset / jb_?set_diag / jb_?mult_diag     (device forms of fill / setDiag / idMatrix, Matrix.hs:525-526,798-848,
//                                                1023-1036, and of matrixMultDiag's scaleColumn loop, Matrix.hs:660-665,1007-1009)
// The level-3 routines run on the tiled GEMM: a triangular or symmetric operand is expanded into a full scratch matrix
// (O(n^2) extra traffic against O(n^2 m) flops), so they inherit the DMMA / FFMA kernels' throughput.
#include "jb_common.cuh"

namespace jb {

// W (n x n, ld n) <- the full matrix described by one triangle of A.
// mode 0: triangular (other triangle zero, unit diagonal if `unit`), 1: symmetric, 2: Hermitian.
template <typename T>
__global__ void expand_kernel(int mode, int lower, int unit, int n, const T *__restrict__ A, int lda, T *__restrict__ W, int ldw) {
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const bool stored = lower ? (i >= j) : (i <= j);
      T v;
      if (i == j) {
        v = (mode == 0 && unit) ? Num<T>::one() : A[(size_t)i + (size_t)j * lda];
        if (mode == 2) v = Num<T>::make(Num<T>::re(v), 0);
      } else if (stored) v = A[(size_t)i + (size_t)j * lda];
      else if (mode == 0) v = Num<T>::zero();
      else { v = A[(size_t)j + (size_t)i * lda]; if (mode == 2) v = Num<T>::conj(v); }
      W[(size_t)i + (size_t)j * ldw] = v;
    }
}
template <typename T>
__global__ void laset_kernel(int m, int n, T off, T diag, T *__restrict__ A, int lda) {
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
      A[(size_t)i + (size_t)j * lda] = (i == j) ? diag : off;
}
// diagonal d (0 main, > 0 above, < 0 below) <- values[0..count) (or the scalar when values == nullptr)
template <typename T>
__global__ void set_diag_kernel(int m, int n, int d, const T *__restrict__ values, T scalar, int count, T *__restrict__ A, int lda) {
  const int i0 = d < 0 ? -d : 0, j0 = d > 0 ? d : 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x)
    if (i0 + t < m && j0 + t < n) A[(size_t)(i0 + t) + (size_t)(j0 + t) * lda] = values ? values[t] : scalar;
}
// out (r x c) <- op(A) * diag(d): column j of op(A) scaled by d[j]; trans 111 copy, 112 transpose, 113 conjugate transpose
template <typename T>
__global__ void mult_diag_kernel(int trans, int r, int c, const T *__restrict__ A, int lda, const T *__restrict__ d, T *__restrict__ out, int ldo) {
  for (int j = blockIdx.y; j < c; j += gridDim.y) {
    const T s = d[j];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r; i += gridDim.x * blockDim.x) {
      T v = (trans == 111) ? A[(size_t)i + (size_t)j * lda] : A[(size_t)j + (size_t)i * lda];
      if (trans == 113) v = Num<T>::conj(v);
      out[(size_t)i + (size_t)j * ldo] = Num<T>::mul(v, s);
    }
  }
}
template <typename T>
__global__ void copy_tri_kernel(int lower, int n, const T *__restrict__ W, int ldw, T *__restrict__ A, int lda) {
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      if (lower ? (i >= j) : (i <= j)) A[(size_t)i + (size_t)j * lda] = W[(size_t)i + (size_t)j * ldw];
}
template <typename T>
__global__ void first_zero_diag_kernel(int n, const T *__restrict__ A, int lda, int *info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && Num<T>::is_zero(A[(size_t)i + (size_t)i * lda])) atomicMin(info, i + 1);
}
static inline dim3 grid2(int m, int n) { return dim3((unsigned)std::min((m + 255) / 256, 1024), (unsigned)std::min(std::max(n, 1), 4096)); }
static inline int max1(int x) { return x > 1 ? x : 1; }

template <typename T> static int scalar_of(const void *p, T *out) {
  if (!p) { set_error(JB_ERR_ARG, "null scalar pointer"); return 1; }
  if (classify(p) == Residency::Device) {
    if (ensure_device()) return 1;
    if (cudaMemcpy(out, p, sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error(JB_ERR_CUDA, "scalar read failed"); return 1; }
  } else memcpy(out, p, sizeof(T));
  return 0;
}
static void finish_host(const char *name, bool host, cudaStream_t s) {
  if (!host) return;
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) set_error(JB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
}

// B <- alpha op(A) B (side 141) or alpha B op(A) (side 142), A triangular
template <typename T>
static void trmm_entry(const char *name, int order, int side, int uplo, int trans, int diag, int m, int n, T alpha, const T *A, int lda, T *B, int ldb) {
  int bad = 0;
  if (order != 101 && order != 102) bad = 1;
  else if (side != 141 && side != 142) bad = 2;
  else if (uplo != 121 && uplo != 122) bad = 3;
  else if (trans < 111 || trans > 113) bad = 4;
  else if (diag != 131 && diag != 132) bad = 5;
  else if (m < 0) bad = 6;
  else if (n < 0) bad = 7;
  if (!bad && order == 101) { side = (side == 141) ? 142 : 141; uplo = (uplo == 121) ? 122 : 121; int t = m; m = n; n = t; }
  const int ka = (side == 141) ? m : n;
  if (!bad) { if (lda < max1(ka)) bad = 10; else if (ldb < max1(m)) bad = 12; }
  if (bad) { set_error(JB_ERR_ARG, "%s: parameter %d had an illegal value", name, bad); return; }
  if (m == 0 || n == 0) return;
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  Staged sa, sb;
  if (sa.open(A, ka, ka, lda, sizeof(T), true, s)) return;
  if (sb.open(B, m, n, ldb, sizeof(T), true, s)) { sa.close(false, s); return; }
  T *W = nullptr, *B0 = nullptr;
  const int ldw = ka + (ka & 1), ld0 = m + (m & 1);
  int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)ldw * ka, s);
  if (!rc) rc = ws_alloc((void **)&B0, sizeof(T) * (size_t)ld0 * n, s);
  if (!rc) {
    expand_kernel<T><<<grid2(ka, ka), 256, 0, s>>>(0, uplo == 122, diag == 132, ka, (const T *)sa.dev, sa.ld, W, ldw);
    count_launch();
    rc = lacpy_device<T>(111, m, n, (const T *)sb.dev, sb.ld, B0, ld0, s);
    if (!rc) {
      if (side == 141) rc = gemm_device<T>(trans, 111, m, n, m, alpha, W, ldw, B0, ld0, Num<T>::zero(), (T *)sb.dev, sb.ld, s);
      else rc = gemm_device<T>(111, trans, m, n, n, alpha, B0, ld0, W, ldw, Num<T>::zero(), (T *)sb.dev, sb.ld, s);
    }
  }
  ws_free(W, s); ws_free(B0, s);
  const bool host = sa.staged || sb.staged;
  sa.close(false, s);
  sb.close(rc == 0, s);
  finish_host(name, host, s);
}

// C <- alpha A B + beta C (side 141) or alpha B A + beta C (side 142), A symmetric (herm = 0) or Hermitian (herm = 1)
template <typename T>
static void symm_entry(const char *name, int herm, int order, int side, int uplo, int m, int n, T alpha, const T *A, int lda, const T *B, int ldb,
                       T beta, T *C, int ldc) {
  int bad = 0;
  if (order != 101 && order != 102) bad = 1;
  else if (side != 141 && side != 142) bad = 2;
  else if (uplo != 121 && uplo != 122) bad = 3;
  else if (m < 0) bad = 4;
  else if (n < 0) bad = 5;
  if (!bad && order == 101) { side = (side == 141) ? 142 : 141; uplo = (uplo == 121) ? 122 : 121; int t = m; m = n; n = t; }
  const int ka = (side == 141) ? m : n;
  if (!bad) { if (lda < max1(ka)) bad = 8; else if (ldb < max1(m)) bad = 10; else if (ldc < max1(m)) bad = 13; }
  if (bad) { set_error(JB_ERR_ARG, "%s: parameter %d had an illegal value", name, bad); return; }
  if (m == 0 || n == 0) return;
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  Staged sa, sb, sc;
  const bool need_c = !(Num<T>::re(beta) == 0 && Num<T>::im(beta) == 0);
  if (sa.open(A, ka, ka, lda, sizeof(T), true, s)) return;
  if (sb.open(B, m, n, ldb, sizeof(T), true, s)) { sa.close(false, s); return; }
  if (sc.open(C, m, n, ldc, sizeof(T), need_c, s)) { sa.close(false, s); sb.close(false, s); return; }
  T *W = nullptr;
  const int ldw = ka + (ka & 1);
  int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)ldw * ka, s);
  if (!rc) {
    expand_kernel<T><<<grid2(ka, ka), 256, 0, s>>>(herm ? 2 : 1, uplo == 122, 0, ka, (const T *)sa.dev, sa.ld, W, ldw);
    count_launch();
    // (row-major input: the column-major view of the stored triangle completes to A^T, which is exactly the operand the
    //  transposed product C^T = alpha B^T A^T + beta C^T needs — no further transposition)
    const int tw = 111;
    if (side == 141) rc = gemm_device<T>(tw, 111, m, n, m, alpha, W, ldw, (const T *)sb.dev, sb.ld, beta, (T *)sc.dev, sc.ld, s);
    else rc = gemm_device<T>(111, tw, m, n, n, alpha, (const T *)sb.dev, sb.ld, W, ldw, beta, (T *)sc.dev, sc.ld, s);
  }
  ws_free(W, s);
  const bool host = sa.staged || sb.staged || sc.staged;
  sa.close(false, s); sb.close(false, s);
  sc.close(rc == 0, s);
  finish_host(name, host, s);
}

// inverse of an SPD / HPD matrix from its Cholesky factor, uplo triangle only (the other triangle is not referenced)
template <typename T>
static int api_potri(int order, char uplo, int n, T *a, int lda) {
  if (order != 101 && order != 102) return -1;
  if (!(uplo == 'L' || uplo == 'l' || uplo == 'U' || uplo == 'u')) return -2;
  if (n < 0) return -3;
  if (lda < max1(n)) return -5;
  if (n == 0) return 0;
  if (int rc = ensure_device()) return rc;
  cudaStream_t s = tls().stream;
  // row-major triangle == column-major opposite triangle of the conjugate (see api_potrf): same bytes, flipped uplo
  char u = uplo;
  if (order == 101) u = (uplo == 'L' || uplo == 'l') ? 'U' : 'L';
  const int lower = (u == 'L' || u == 'l');
  Staged st;
  int *d_misc = nullptr;
  T *W = nullptr;
  int rc = st.open(a, n, n, lda, sizeof(T), true, s);
  if (rc) return rc;
  int h[2] = {0x7fffffff, 0};
  rc = ws_alloc((void **)&d_misc, 2 * sizeof(int), s);
  if (!rc && cudaMemcpyAsync(d_misc, h, sizeof(h), cudaMemcpyHostToDevice, s) != cudaSuccess) rc = JB_ERR_CUDA;
  if (!rc) {
    nancheck_device<T>(lower ? 'L' : 'U', n, n, (const T *)st.dev, st.ld, d_misc + 1, s);
    first_zero_diag_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, (const T *)st.dev, st.ld, d_misc);
    count_launch();
    if (cudaMemcpyAsync(h, d_misc, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) rc = JB_ERR_CUDA;
  }
  int info = 0;
  if (!rc) {
    if (h[1]) info = -4;
    else if (h[0] != 0x7fffffff) info = h[0];       // ?trtri: the factor is singular
    else {
      const int ldw = n + (n & 1);
      rc = ws_alloc((void **)&W, sizeof(T) * (size_t)ldw * n, s);
      if (!rc) {
        laset_kernel<T><<<grid2(n, n), 256, 0, s>>>(n, n, Num<T>::zero(), Num<T>::one(), W, ldw);
        count_launch();
        rc = potrs_device<T>(u, n, n, (const T *)st.dev, st.ld, W, ldw, s);
        if (!rc) { copy_tri_kernel<T><<<grid2(n, n), 256, 0, s>>>(lower, n, W, ldw, (T *)st.dev, st.ld); count_launch(); }
      }
    }
  }
  ws_free(W, s); ws_free(d_misc, s);
  int r2 = st.close(rc == 0 && info == 0, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) rc = JB_ERR_CUDA;
  if (rc == JB_ERR_CUDA) set_error(JB_ERR_CUDA, "LAPACKE_?potri failed on the device");
  return rc ? rc : (info ? info : r2);
}

template <typename T> static int laset_entry(int m, int n, T off, T diag, T *a, int lda) {
  if (int rc = ensure_device()) return rc;
  if (m < 0 || n < 0 || lda < max1(m)) { set_error(JB_ERR_ARG, "jb_?laset: bad argument"); return JB_ERR_ARG; }
  if (m == 0 || n == 0) return 0;
  laset_kernel<T><<<grid2(m, n), 256, 0, tls().stream>>>(m, n, off, diag, a, lda);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 0;
}
// values: `count` elements in host or device memory, or NULL for `count` copies of *scalar
template <typename T> static int set_diag_entry(int m, int n, int d, const T *values, T scalar, int count, T *a, int lda) {
  if (int rc = ensure_device()) return rc;
  if (m < 0 || n < 0 || lda < max1(m) || count < 0) { set_error(JB_ERR_ARG, "jb_?set_diag: bad argument"); return JB_ERR_ARG; }
  const int len = d >= 0 ? std::min(m, n - d) : std::min(m + d, n);
  if (count > len) count = len > 0 ? len : 0;
  if (count == 0) return 0;
  cudaStream_t s = tls().stream;
  T *dv = nullptr; bool staged = false;
  if (values && classify(values) != Residency::Device) {
    if (int rc = ws_alloc((void **)&dv, sizeof(T) * (size_t)count, s)) return rc;
    JB_CUDA(cudaMemcpyAsync(dv, values, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaStreamSynchronize(s));          // the caller's list is a temporary (withArray)
    staged = true;
  } else dv = const_cast<T *>(values);
  set_diag_kernel<T><<<(count + 255) / 256, 256, 0, s>>>(m, n, d, dv, scalar, count, a, lda);
  count_launch();
  if (staged) ws_free(dv, s);
  JB_CUDA(cudaGetLastError());
  return 0;
}
template <typename T> static int mult_diag_entry(int trans, int r, int c, const T *a, int lda, const T *d, T *out, int ldo) {
  if (int rc = ensure_device()) return rc;
  if (trans < 111 || trans > 113 || r < 0 || c < 0 || lda < max1(trans == 111 ? r : c) || ldo < max1(r)) { set_error(JB_ERR_ARG, "jb_?mult_diag: bad argument"); return JB_ERR_ARG; }
  if (r == 0 || c == 0) return 0;
  cudaStream_t s = tls().stream;
  T *dv = nullptr; bool staged = false;
  if (classify(d) != Residency::Device) {
    if (int rc = ws_alloc((void **)&dv, sizeof(T) * (size_t)c, s)) return rc;
    JB_CUDA(cudaMemcpyAsync(dv, d, sizeof(T) * (size_t)c, cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaStreamSynchronize(s));
    staged = true;
  } else dv = const_cast<T *>(d);
  mult_diag_kernel<T><<<grid2(r, c), 256, 0, s>>>(trans, r, c, a, lda, dv, out, ldo);
  count_launch();
  if (staged) ws_free(dv, s);
  JB_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace jb

using namespace jb;
extern "C" {
void jb_cblas_strmm(int order, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *a, int lda, float *b, int ldb) {
  trmm_entry<float>("cblas_strmm", order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb); }
void jb_cblas_dtrmm(int order, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *a, int lda, double *b, int ldb) {
  trmm_entry<double>("cblas_dtrmm", order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb); }
void jb_cblas_ctrmm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) {
  cfloat al; if (scalar_of(alpha, &al)) return;
  trmm_entry<cfloat>("cblas_ctrmm", order, side, uplo, trans, diag, m, n, al, (const cfloat *)a, lda, (cfloat *)b, ldb); }
void jb_cblas_ztrmm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) {
  cdouble al; if (scalar_of(alpha, &al)) return;
  trmm_entry<cdouble>("cblas_ztrmm", order, side, uplo, trans, diag, m, n, al, (const cdouble *)a, lda, (cdouble *)b, ldb); }
void cblas_strmm(int order, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *a, int lda, float *b, int ldb) { jb_cblas_strmm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb); }
void cblas_dtrmm(int order, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *a, int lda, double *b, int ldb) { jb_cblas_dtrmm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb); }
void cblas_ctrmm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) { jb_cblas_ctrmm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb); }
void cblas_ztrmm(int order, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *a, int lda, void *b, int ldb) { jb_cblas_ztrmm(order, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb); }

void jb_cblas_ssymm(int order, int side, int uplo, int m, int n, float alpha, const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc) {
  symm_entry<float>("cblas_ssymm", 0, order, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc); }
void jb_cblas_dsymm(int order, int side, int uplo, int m, int n, double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc) {
  symm_entry<double>("cblas_dsymm", 0, order, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc); }
#define CPLX_SYMM(name, T, herm)                                                                                              \
  void jb_##name(int order, int side, int uplo, int m, int n, const void *alpha, const void *a, int lda, const void *b, int ldb, const void *beta, void *c, int ldc) { \
    T al, be; if (scalar_of(alpha, &al) || scalar_of(beta, &be)) return;                                                      \
    symm_entry<T>(#name, herm, order, side, uplo, m, n, al, (const T *)a, lda, (const T *)b, ldb, be, (T *)c, ldc); }         \
  void name(int order, int side, int uplo, int m, int n, const void *alpha, const void *a, int lda, const void *b, int ldb, const void *beta, void *c, int ldc) { \
    jb_##name(order, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc); }
CPLX_SYMM(cblas_csymm, cfloat, 0)
CPLX_SYMM(cblas_zsymm, cdouble, 0)
CPLX_SYMM(cblas_chemm, cfloat, 1)
CPLX_SYMM(cblas_zhemm, cdouble, 1)
void cblas_ssymm(int order, int side, int uplo, int m, int n, float alpha, const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc) { jb_cblas_ssymm(order, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc); }
void cblas_dsymm(int order, int side, int uplo, int m, int n, double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc) { jb_cblas_dsymm(order, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc); }

#define POTRI_API(t, T, CT)                                                                                                   \
  int jb_LAPACKE_##t##potri(int order, char uplo, int n, CT *a, int lda) { return api_potri<T>(order, uplo, n, (T *)a, lda); } \
  int LAPACKE_##t##potri(int order, char uplo, int n, CT *a, int lda) { return jb_LAPACKE_##t##potri(order, uplo, n, a, lda); }
POTRI_API(s, float, float)
POTRI_API(d, double, double)
POTRI_API(c, cfloat, void)
POTRI_API(z, cdouble, void)

int jb_slaset(int m, int n, float off, float diag, float *a, int lda) { return laset_entry<float>(m, n, off, diag, a, lda); }
int jb_dlaset(int m, int n, double off, double diag, double *a, int lda) { return laset_entry<double>(m, n, off, diag, a, lda); }
int jb_claset(int m, int n, const void *off, const void *diag, void *a, int lda) {
  cfloat o, d; if (scalar_of(off, &o) || scalar_of(diag, &d)) return JB_ERR_ARG; return laset_entry<cfloat>(m, n, o, d, (cfloat *)a, lda); }
int jb_zlaset(int m, int n, const void *off, const void *diag, void *a, int lda) {
  cdouble o, d; if (scalar_of(off, &o) || scalar_of(diag, &d)) return JB_ERR_ARG; return laset_entry<cdouble>(m, n, o, d, (cdouble *)a, lda); }
int jb_sset_diag(int m, int n, int d, const float *values, int count, float *a, int lda) { return set_diag_entry<float>(m, n, d, values, 0.f, count, a, lda); }
int jb_dset_diag(int m, int n, int d, const double *values, int count, double *a, int lda) { return set_diag_entry<double>(m, n, d, values, 0.0, count, a, lda); }
int jb_cset_diag(int m, int n, int d, const void *values, int count, void *a, int lda) { return set_diag_entry<cfloat>(m, n, d, (const cfloat *)values, Num<cfloat>::zero(), count, (cfloat *)a, lda); }
int jb_zset_diag(int m, int n, int d, const void *values, int count, void *a, int lda) { return set_diag_entry<cdouble>(m, n, d, (const cdouble *)values, Num<cdouble>::zero(), count, (cdouble *)a, lda); }
int jb_smult_diag(int trans, int r, int c, const float *a, int lda, const float *d, float *out, int ldo) { return mult_diag_entry<float>(trans, r, c, a, lda, d, out, ldo); }
int jb_dmult_diag(int trans, int r, int c, const double *a, int lda, const double *d, double *out, int ldo) { return mult_diag_entry<double>(trans, r, c, a, lda, d, out, ldo); }
int jb_cmult_diag(int trans, int r, int c, const void *a, int lda, const void *d, void *out, int ldo) { return mult_diag_entry<cfloat>(trans, r, c, (const cfloat *)a, lda, (const cfloat *)d, (cfloat *)out, ldo); }
int jb_zmult_diag(int trans, int r, int c, const void *a, int lda, const void *d, void *out, int ldo) { return mult_diag_entry<cdouble>(trans, r, c, (const cdouble *)a, lda, (const cdouble *)d, (cdouble *)out, ldo); }
}
