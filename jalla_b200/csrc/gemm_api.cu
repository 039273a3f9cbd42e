// cblas_?gemm entry points — the symbols Jalla imports at
// /root/reference/Numeric/Jalla/Foreign/BLAS.chs:149 (s), :155 (d), :161 (c), :167 (z)
// and reaches from unsafeMatrixMult (Numeric/Jalla/Matrix.hs:477-490) through the
// BlasOps dictionaries (Foreign/BlasOps.hs:52,108,132,156,180).
#include "jb_common.cuh"

namespace jb {

// Large all-host GEMM: stream row blocks of A and column panels of B / C through the GPU so the PCIe
// copies overlap the kernel (the e2e path bench.py times).  Declared here, defined below.
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
  // C is cut into R row blocks x P column panels.  Issue order (copies on the H2D stream, products on the call's stream,
  // results on the D2H stream): A block 0 and B panel 0, tile (0,0); then panel j, tile (0,j) for the other panels; then
  // block i, tiles (i, *).  A tile waits only for its own operands and its result leaves while the next one multiplies;
  // PCIe moves the operands (2 x N^2) faster than the GEMM consumes them, so after the first tile the copies hide under
  // the kernel.  Copies and products are issued interleaved because a copy from PAGEABLE memory blocks the caller while
  // the host threads fill a pinned slot (HostStage) — the device must already hold the work that overlaps it.
  const bool beta0 = Num<T>::is_zero(beta);
  const int R = (m >= 8192) ? 2 : 1;
  const int MB = ((m + R - 1) / R + 127) / 128 * 128;          // rows of op(A)/C per block
  const int NP = 2048;                                         // columns of C per panel
  const int nblocks = (m + MB - 1) / MB, npanels = (n + NP - 1) / NP;
  const long ra = (ta == 111) ? m : k, ca = (ta == 111) ? k : m;
  const long rb = (tb == 111) ? k : n, cb = (tb == 111) ? n : k;
  const int ldda = (int)(ra + (ra & 1)), lddb = (int)(rb + (rb & 1)), lddc = m + (m & 1);
  const size_t E = sizeof(T);
  cudaStream_t sin = nullptr, sout = nullptr;
  T *dA = nullptr, *dB = nullptr, *dC = nullptr;
  cudaEvent_t ev_start = nullptr, ev_tile = nullptr, ev_in = nullptr;
  HostStage stage;
  int rc = 0;
  cudaError_t e = cudaSuccess;
#define PL_CUDA(x) do { e = (x); if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "gemm (host pipeline): %s", cudaGetErrorString(e)); rc = JB_ERR_CUDA; goto done; } } while (0)
#define PL_RC(x) do { rc = (x); if (rc) goto done; } while (0)
  PL_CUDA(cudaStreamCreateWithFlags(&sin, cudaStreamNonBlocking));
  PL_CUDA(cudaStreamCreateWithFlags(&sout, cudaStreamNonBlocking));
  if ((rc = ws_alloc((void **)&dA, (size_t)ldda * ca * E, s))) goto done;
  if ((rc = ws_alloc((void **)&dB, (size_t)lddb * cb * E, s))) goto done;
  if ((rc = ws_alloc((void **)&dC, (size_t)lddc * n * E, s))) goto done;
  PL_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
  PL_CUDA(cudaEventCreateWithFlags(&ev_tile, cudaEventDisableTiming));
  PL_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
  PL_CUDA(cudaEventRecord(ev_start, s));                       // buffers come from s's pool: order the side streams after it
  PL_CUDA(cudaStreamWaitEvent(sin, ev_start, 0));
  PL_CUDA(cudaStreamWaitEvent(sout, ev_start, 0));
  {
    // rows [r0, r0 + mr) of op(A), inner indices [k0, k0 + kc)
    auto copy_a = [&](int r0, int mr, int k0, int kc) -> int {
      if (ta == 111) return stage.h2d(dA + r0 + (size_t)k0 * ldda, (size_t)ldda * E, A + r0 + (size_t)k0 * lda, (size_t)lda * E, (size_t)mr * E, (size_t)kc, sin);
      return stage.h2d(dA + k0 + (size_t)r0 * ldda, (size_t)ldda * E, A + k0 + (size_t)r0 * lda, (size_t)lda * E, (size_t)kc * E, (size_t)mr, sin);
    };
    // columns [c0, c0 + nc) of op(B), inner indices [k0, k0 + kc)
    auto copy_b = [&](int c0, int nc, int k0, int kc) -> int {
      if (tb == 111) return stage.h2d(dB + k0 + (size_t)c0 * lddb, (size_t)lddb * E, B + k0 + (size_t)c0 * ldb, (size_t)ldb * E, (size_t)kc * E, (size_t)nc, sin);
      return stage.h2d(dB + c0 + (size_t)k0 * lddb, (size_t)lddb * E, B + c0 + (size_t)k0 * ldb, (size_t)ldb * E, (size_t)nc * E, (size_t)kc, sin);
    };
    auto landed = [&]() -> cudaError_t {                       // the call's stream waits for everything queued on sin so far
      cudaError_t ce = cudaEventRecord(ev_in, sin);
      return ce == cudaSuccess ? cudaStreamWaitEvent(s, ev_in, 0) : ce;
    };
    // The first tile would wait for ALL of A block 0 and B panel 0 (1.25 GiB at N = 16384: 24 ms of a 270 ms call).  Its
    // operands therefore arrive in NKC slices along K, and the tile is accumulated slice by slice (beta, then 1) as they
    // land: the first product starts after 1/NKC of that wait.
    const int NKC = (k >= 8192 && option("gemm.e2e_kchunks") != 0) ? 4 : 1;
    const int KC = ((k + NKC - 1) / NKC + 15) / 16 * 16;
    for (int i = 0; i < nblocks; i++) {
      const int r0 = i * MB, mr = (m - r0 < MB) ? (m - r0) : MB;
      if (!beta0) PL_RC(stage.h2d(dC + r0, (size_t)lddc * E, C + r0, (size_t)ldc * E, (size_t)mr * E, (size_t)n, sin));
      if (i > 0) { PL_RC(copy_a(r0, mr, 0, k)); PL_CUDA(landed()); }
      for (int j = 0; j < npanels; j++) {
        const int c0 = j * NP, nc = (n - c0 < NP) ? (n - c0) : NP;
        const T *pa = (ta == 111) ? dA + r0 : dA + (size_t)r0 * ldda;
        const T *pb = (tb == 111) ? dB + (size_t)c0 * lddb : dB + c0;
        T *pc = dC + r0 + (size_t)c0 * lddc;
        if (i == 0 && j == 0) {
          for (int c = 0; c < NKC; c++) {
            const int k0 = c * KC, kc = (k - k0 < KC) ? (k - k0) : KC;
            if (kc <= 0) break;
            PL_RC(copy_a(r0, mr, k0, kc));
            PL_RC(copy_b(c0, nc, k0, kc));
            PL_CUDA(landed());
            const T *pak = (ta == 111) ? pa + (size_t)k0 * ldda : pa + k0;
            const T *pbk = (tb == 111) ? pb + k0 : pb + (size_t)k0 * lddb;
            PL_RC(gemm_device<T>(ta, tb, mr, nc, kc, alpha, pak, ldda, pbk, lddb, c == 0 ? beta : Num<T>::one(), pc, lddc, s, 0));
          }
        } else {
          if (i == 0) { PL_RC(copy_b(c0, nc, 0, k)); PL_CUDA(landed()); }
          PL_RC(gemm_device<T>(ta, tb, mr, nc, k, alpha, pa, ldda, pb, lddb, beta, pc, lddc, s, 0));
        }
        PL_CUDA(cudaEventRecord(ev_tile, s));
        PL_CUDA(cudaStreamWaitEvent(sout, ev_tile, 0));
        PL_RC(stage.d2h(C + r0 + (size_t)c0 * ldc, (size_t)ldc * E, pc, (size_t)lddc * E, (size_t)mr * E, (size_t)nc, sout));
      }
    }
  }
  PL_RC(stage.finish());
  PL_CUDA(cudaStreamSynchronize(sout));
  PL_CUDA(cudaStreamSynchronize(s));
done:
#undef PL_CUDA
#undef PL_RC
  stage.finish();
  if (sin) cudaStreamSynchronize(sin);
  if (sout) cudaStreamSynchronize(sout);
  cudaStreamSynchronize(s);
  if (ev_start) cudaEventDestroy(ev_start);
  if (ev_tile) cudaEventDestroy(ev_tile);
  if (ev_in) cudaEventDestroy(ev_in);
  ws_free(dC, s); ws_free(dB, s); ws_free(dA, s);
  if (sin) cudaStreamDestroy(sin);
  if (sout) cudaStreamDestroy(sout);
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
