// cblas_?gemm entry points — the symbols Jalla imports at
// /root/reference/Numeric/Jalla/Foreign/BLAS.chs:149 (s), :155 (d), :161 (c), :167 (z)
// and reaches from unsafeMatrixMult (Numeric/Jalla/Matrix.hs:477-490) through the
// BlasOps dictionaries (Foreign/BlasOps.hs:52,108,132,156,180).
#include "jb_common.cuh"

namespace jb {

// Large all-host GEMM: stream column panels of B / C through the GPU so the PCIe copies
// overlap the kernel (the e2e path bench.py times).  Declared here, defined below.
template <typename T>
static int gemm_host_pipelined(int ta, int tb, int m, int n, int k, T alpha, const T *A, int lda, const T *B, int ldb,
                               T beta, T *C, int ldc, cudaStream_t s);

template <typename T>
static void gemm_entry(const char *name, int order, int ta, int tb, int m, int n, int k, T alpha, const T *A, int lda,
                       const T *B, int ldb, T beta, T *C, int ldc) {
  // argument screening in cblas_xerbla order; we record instead of aborting
  int bad = 0;
  if (order != 101 && order != 102) bad = 1;
  else if (ta < 111 || ta > 113) bad = 2;
  else if (tb < 111 || tb > 113) bad = 3;
  else if (m < 0) bad = 4;
  else if (n < 0) bad = 5;
  else if (k < 0) bad = 6;
  if (!bad && order == 101) {   // row-major C = op(A)op(B)  <=>  column-major C^T = op(B)^T op(A)^T
    const T *tp = A; A = B; B = tp;
    int t = ta; ta = tb; tb = t;
    t = lda; lda = ldb; ldb = t;
    t = m; m = n; n = t;
  }
  if (!bad) {
    int rows_a = (ta == 111) ? m : k, rows_b = (tb == 111) ? k : n;
    if (lda < (rows_a > 1 ? rows_a : 1)) bad = (order == 102) ? 9 : 11;
    else if (ldb < (rows_b > 1 ? rows_b : 1)) bad = (order == 102) ? 11 : 9;
    else if (ldc < (m > 1 ? m : 1)) bad = 14;
  }
  if (bad) { set_error(JB_ERR_ARG, "%s: parameter %d had an illegal value", name, bad); return; }
  if (m == 0 || n == 0) return;
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;

  const bool beta0 = Num<T>::is_zero(beta);
  long ra = (ta == 111) ? m : k, ca = (ta == 111) ? k : m;
  long rb = (tb == 111) ? k : n, cb = (tb == 111) ? n : k;
  bool hostA = classify(A) == Residency::Host, hostB = classify(B) == Residency::Host, hostC = classify(C) == Residency::Host;
  if (hostA && hostB && hostC && (double)m * n * k >= 5e10 && k > 0) {
    gemm_host_pipelined<T>(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s);
    return;
  }
  Staged sa, sb, sc;
  bool need_ab = k > 0 && !Num<T>::is_zero(alpha);
  if (sa.open(A, ra, need_ab ? ca : 0, lda, sizeof(T), true, s)) return;
  if (sb.open(B, rb, need_ab ? cb : 0, ldb, sizeof(T), true, s)) { sa.close(false, s); return; }
  if (sc.open(C, m, n, ldc, sizeof(T), !beta0, s)) { sa.close(false, s); sb.close(false, s); return; }
  int rc = gemm_device<T>(ta, tb, m, n, need_ab ? k : 0, alpha, (const T *)sa.dev, sa.ld, (const T *)sb.dev, sb.ld, beta,
                          (T *)sc.dev, sc.ld, s, 0);
  bool any_host = sa.staged || sb.staged || sc.staged;
  sa.close(false, s);
  sb.close(false, s);
  sc.close(rc == 0, s);
  if (any_host) {   // host-visible output must be complete on return
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) set_error(JB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  }
}

template <typename T>
static int gemm_host_pipelined(int ta, int tb, int m, int n, int k, T alpha, const T *A, int lda, const T *B, int ldb,
                               T beta, T *C, int ldc, cudaStream_t s) {
  // A resident once; B and C move in column panels of op(B)/C on two copy streams,
  // double buffered, while the compute stream multiplies the previous panel.
  const bool beta0 = Num<T>::is_zero(beta);
  long ra = (ta == 111) ? m : k, ca = (ta == 111) ? k : m;
  const int NP = 1024;                       // panel width (columns of C)
  int npanels = (n + NP - 1) / NP;
  cudaStream_t sin, sout;
  JB_CUDA(cudaStreamCreateWithFlags(&sin, cudaStreamNonBlocking));
  JB_CUDA(cudaStreamCreateWithFlags(&sout, cudaStreamNonBlocking));
  T *dA = nullptr, *dB[2] = {nullptr, nullptr}, *dC[2] = {nullptr, nullptr};
  int ldda = (int)(ra + (ra & 1)), lddc = m + (m & 1);
  // B panel storage: NoTrans -> k x NP (ld k), Trans -> NP x k (ld NP)
  int lddb = (tb == 111) ? (k + (k & 1)) : NP;
  size_t bpanel = (tb == 111) ? (size_t)lddb * NP : (size_t)lddb * k;
  cudaEvent_t ev_in[2], ev_done[2], ev_out[2], ev_start;
  int rc = 0;
  cudaError_t e = cudaSuccess;
#define PL_CUDA(x) do { e = (x); if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "gemm (host pipeline): %s", cudaGetErrorString(e)); rc = JB_ERR_CUDA; goto done; } } while (0)
  for (int i = 0; i < 2; i++) { ev_in[i] = ev_done[i] = ev_out[i] = nullptr; }
  ev_start = nullptr;
  PL_CUDA(cudaMalloc(&dA, (size_t)ldda * ca * sizeof(T)));
  for (int i = 0; i < 2; i++) {
    PL_CUDA(cudaMalloc(&dB[i], bpanel * sizeof(T)));
    PL_CUDA(cudaMalloc(&dC[i], (size_t)lddc * NP * sizeof(T)));
    PL_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
  }
  PL_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
  PL_CUDA(cudaEventRecord(ev_start, s));
  PL_CUDA(cudaStreamWaitEvent(sin, ev_start, 0));
  PL_CUDA(cudaStreamWaitEvent(sout, ev_start, 0));
  PL_CUDA(cudaMemcpy2DAsync(dA, (size_t)ldda * sizeof(T), A, (size_t)lda * sizeof(T), (size_t)ra * sizeof(T), (size_t)ca,
                            cudaMemcpyHostToDevice, sin));
  for (int p = 0; p < npanels; p++) {
    int b = p & 1, j0 = p * NP, nb = (n - j0 < NP) ? (n - j0) : NP;
    if (p >= 2) PL_CUDA(cudaStreamWaitEvent(sin, ev_done[b], 0));    // buffer b's previous multiply finished
    if (tb == 111)
      PL_CUDA(cudaMemcpy2DAsync(dB[b], (size_t)lddb * sizeof(T), B + (size_t)j0 * ldb, (size_t)ldb * sizeof(T),
                                (size_t)k * sizeof(T), (size_t)nb, cudaMemcpyHostToDevice, sin));
    else
      PL_CUDA(cudaMemcpy2DAsync(dB[b], (size_t)lddb * sizeof(T), B + j0, (size_t)ldb * sizeof(T), (size_t)nb * sizeof(T),
                                (size_t)k, cudaMemcpyHostToDevice, sin));
    if (!beta0) {
      if (p >= 2) PL_CUDA(cudaStreamWaitEvent(sin, ev_out[b], 0));
      PL_CUDA(cudaMemcpy2DAsync(dC[b], (size_t)lddc * sizeof(T), C + (size_t)j0 * ldc, (size_t)ldc * sizeof(T),
                                (size_t)m * sizeof(T), (size_t)nb, cudaMemcpyHostToDevice, sin));
    }
    PL_CUDA(cudaEventRecord(ev_in[b], sin));
    PL_CUDA(cudaStreamWaitEvent(s, ev_in[b], 0));
    if (p >= 2) PL_CUDA(cudaStreamWaitEvent(s, ev_out[b], 0));       // C buffer b drained to the host
    rc = gemm_device<T>(ta, tb, m, nb, k, alpha, dA, ldda, dB[b], lddb, beta, dC[b], lddc, s, 0);
    if (rc) goto done;
    PL_CUDA(cudaEventRecord(ev_done[b], s));
    PL_CUDA(cudaStreamWaitEvent(sout, ev_done[b], 0));
    PL_CUDA(cudaMemcpy2DAsync(C + (size_t)j0 * ldc, (size_t)ldc * sizeof(T), dC[b], (size_t)lddc * sizeof(T),
                              (size_t)m * sizeof(T), (size_t)nb, cudaMemcpyDeviceToHost, sout));
    PL_CUDA(cudaEventRecord(ev_out[b], sout));
  }
  PL_CUDA(cudaStreamSynchronize(sout));
  PL_CUDA(cudaStreamSynchronize(s));
done:
#undef PL_CUDA
  cudaStreamSynchronize(sin); cudaStreamSynchronize(sout); cudaStreamSynchronize(s);
  for (int i = 0; i < 2; i++) {
    if (dB[i]) cudaFree(dB[i]);
    if (dC[i]) cudaFree(dC[i]);
    if (ev_in[i]) cudaEventDestroy(ev_in[i]);
    if (ev_done[i]) cudaEventDestroy(ev_done[i]);
    if (ev_out[i]) cudaEventDestroy(ev_out[i]);
  }
  if (ev_start) cudaEventDestroy(ev_start);
  if (dA) cudaFree(dA);
  cudaStreamDestroy(sin); cudaStreamDestroy(sout);
  return rc;
}

// complex alpha/beta arrive as pointers to call-scoped temporaries that may live on the
// host (BlasOps.hs:156,180) or on the device; read them once on entry (H10).
template <typename T>
static int read_scalar(const void *p, T *out) {
  if (!p) { set_error(JB_ERR_ARG, "null scalar pointer"); return JB_ERR_ARG; }
  if (classify(p) == Residency::Device) {
    if (ensure_device()) return JB_ERR_CUDA;
    JB_CUDA(cudaMemcpyAsync(out, p, sizeof(T), cudaMemcpyDeviceToHost, tls().stream));
    JB_CUDA(cudaStreamSynchronize(tls().stream));
  } else {
    memcpy(out, p, sizeof(T));
  }
  return 0;
}

}  // namespace jb

using namespace jb;
extern "C" {

void jb_cblas_sgemm(int order, int ta, int tb, int m, int n, int k, float alpha, const float *a, int lda, const float *b,
                    int ldb, float beta, float *c, int ldc) {
  gemm_entry<float>("cblas_sgemm", order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void jb_cblas_dgemm(int order, int ta, int tb, int m, int n, int k, double alpha, const double *a, int lda,
                    const double *b, int ldb, double beta, double *c, int ldc) {
  gemm_entry<double>("cblas_dgemm", order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void jb_cblas_cgemm(int order, int ta, int tb, int m, int n, int k, const void *alpha, const void *a, int lda,
                    const void *b, int ldb, const void *beta, void *c, int ldc) {
  cfloat al, be;
  if (read_scalar(alpha, &al) || read_scalar(beta, &be)) return;
  gemm_entry<cfloat>("cblas_cgemm", order, ta, tb, m, n, k, al, (const cfloat *)a, lda, (const cfloat *)b, ldb, be,
                     (cfloat *)c, ldc);
}
void jb_cblas_zgemm(int order, int ta, int tb, int m, int n, int k, const void *alpha, const void *a, int lda,
                    const void *b, int ldb, const void *beta, void *c, int ldc) {
  cdouble al, be;
  if (read_scalar(alpha, &al) || read_scalar(beta, &be)) return;
  gemm_entry<cdouble>("cblas_zgemm", order, ta, tb, m, n, k, al, (const cdouble *)a, lda, (const cdouble *)b, ldb, be,
                      (cdouble *)c, ldc);
}

// Unprefixed aliases: the literal symbols Jalla's `foreign import ccall unsafe` names.
void cblas_sgemm(int order, int ta, int tb, int m, int n, int k, float alpha, const float *a, int lda, const float *b,
                 int ldb, float beta, float *c, int ldc) {
  jb_cblas_sgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cblas_dgemm(int order, int ta, int tb, int m, int n, int k, double alpha, const double *a, int lda, const double *b,
                 int ldb, double beta, double *c, int ldc) {
  jb_cblas_dgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cblas_cgemm(int order, int ta, int tb, int m, int n, int k, const void *alpha, const void *a, int lda, const void *b,
                 int ldb, const void *beta, void *c, int ldc) {
  jb_cblas_cgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cblas_zgemm(int order, int ta, int tb, int m, int n, int k, const void *alpha, const void *a, int lda, const void *b,
                 int ldb, const void *beta, void *c, int ldc) {
  jb_cblas_zgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
}
