// LAPACKE_?{gesv,getrf,getrs,getri,potrf,potrs,posv} entry points: the symbols Jalla
// imports at /root/reference/Numeric/Jalla/Foreign/LAPACKE.chs:551-554,691-694,697-700,
// 703-706,733-736,741-744,601-604 and calls from Numeric/Jalla/Matrix.hs:503,766,769.
// Contract restated from lapacke/include/lapacke.h (prototypes :962-975,1099-1134,
// 2758-2771,2844-2875; order constants :143-144; error codes :146-147).
#include "jb_common.cuh"

namespace jb {

static inline int max1(int x) { return x > 1 ? x : 1; }
static inline bool is_trans_char(char c) { return c == 'N' || c == 'n' || c == 'T' || c == 't' || c == 'C' || c == 'c'; }
static inline bool is_uplo_char(char c) { return c == 'L' || c == 'l' || c == 'U' || c == 'u'; }

// One LAPACKE call = one scope: owns staging of every operand, the device copies of
// ipiv/info, and the final synchronisation (host-visible outputs must be complete on
// return: SURVEY.md §8b "Threading").
template <typename T>
struct Call {
  cudaStream_t s;
  int *d_misc = nullptr;   // [0] info, [1] nan flag A, [2] nan flag B
  int rc = 0;
  explicit Call() : s(tls().stream) {}
  int begin() {
    if ((rc = ensure_device())) return rc;
    if ((rc = ws_alloc((void **)&d_misc, 4 * sizeof(int), s))) return rc;
    cudaError_t e = cudaMemsetAsync(d_misc, 0, 4 * sizeof(int), s);
    if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); return rc = JB_ERR_CUDA; }
    return 0;
  }
  // Matrix operand in caller's order: returns a column-major device view.  For order=101
  // the row-major r x c matrix is the column-major c x r matrix with the same ld; it is
  // transposed on the device into scratch (and back on close when `out`).
  struct Mat {
    Staged st;           // raw storage on device (possibly staged from host)
    T *cm = nullptr;     // column-major r x c view
    int ld = 0;
    bool transposed = false;
    int r = 0, c = 0;
  };
  int open(Mat &M, int order, const T *p, int r, int c, int ld, bool in) {
    M.r = r; M.c = c;
    long sr = (order == 102) ? r : c, sc = (order == 102) ? c : r;
    if ((rc = M.st.open(p, sr, sc, ld, sizeof(T), in, s))) return rc;
    if (order == 102) { M.cm = (T *)M.st.dev; M.ld = M.st.ld; return 0; }
    M.transposed = true;
    M.ld = max1(r) + (r & 1);
    if ((rc = ws_alloc((void **)&M.cm, sizeof(T) * (size_t)M.ld * max1(c), s))) return rc;
    if (in) rc = lacpy_device<T>(112, c, r, (const T *)M.st.dev, M.st.ld, M.cm, M.ld, s);
    return rc;
  }
  int close(Mat &M, bool out) {
    int r2 = 0;
    if (M.transposed) {
      if (out && rc == 0) r2 = lacpy_device<T>(112, M.r, M.c, M.cm, M.ld, (T *)M.st.dev, M.st.ld, s);
      ws_free(M.cm, s);
      M.transposed = false;
    }
    int r3 = M.st.close(out && rc == 0, s);
    return r2 ? r2 : r3;
  }
  // ipiv may be host (Jalla: allocaArray, Matrix.hs:502,765) or device memory
  struct Piv { int *dev = nullptr; int *host = nullptr; int n = 0; bool staged = false; };
  int open_piv(Piv &P, const int *ipiv, int n, bool in) {
    P.n = n;
    if (n <= 0 || classify(ipiv) == Residency::Device) { P.dev = const_cast<int *>(ipiv); return 0; }
    P.staged = true; P.host = const_cast<int *>(ipiv);
    if ((rc = ws_alloc((void **)&P.dev, sizeof(int) * (size_t)n, s))) return rc;
    if (in) {
      cudaError_t e = cudaMemcpyAsync(P.dev, ipiv, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "ipiv H2D: %s", cudaGetErrorString(e)); return rc = JB_ERR_CUDA; }
    }
    return 0;
  }
  void close_piv(Piv &P, bool out) {
    if (!P.staged) return;
    if (out && rc == 0) cudaMemcpyAsync(P.host, P.dev, sizeof(int) * (size_t)P.n, cudaMemcpyDeviceToHost, s);
    ws_free(P.dev, s);
    P.staged = false;
  }
  int nan_flags(int *fa, int *fb) {
    int h[4];
    cudaError_t e = cudaMemcpyAsync(h, d_misc, sizeof(h), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "nan check: %s", cudaGetErrorString(e)); return rc = JB_ERR_CUDA; }
    *fa = h[1]; if (fb) *fb = h[2];
    return 0;
  }
  // final sync; returns LAPACK info (or a JB_ERR_* code)
  int finish() {
    int info = 0;
    if (d_misc) {
      cudaError_t e = cudaMemcpyAsync(&info, d_misc, sizeof(int), cudaMemcpyDeviceToHost, s);
      ws_free(d_misc, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "LAPACKE call failed on the device: %s", cudaGetErrorString(e)); cudaGetLastError(); return JB_ERR_CUDA; }
    }
    return rc ? rc : info;
  }
};

template <typename T>
static int api_getrf(int order, int m, int n, T *a, int lda, int *ipiv) {
  if (order != 101 && order != 102) return -1;
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (lda < max1(order == 102 ? m : n)) return -5;
  if (m == 0 || n == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  typename Call<T>::Mat A; typename Call<T>::Piv P;
  const int mn = m < n ? m : n;
  if (c.open(A, order, a, m, n, lda, true) == 0 && c.open_piv(P, ipiv, mn, false) == 0) {
    nancheck_device<T>(0, m, n, A.cm, A.ld, c.d_misc + 1, c.s);
    int fa = 0;
    if (c.nan_flags(&fa, nullptr) == 0) {
      if (fa) { c.close_piv(P, false); c.close(A, false); c.finish(); return -4; }
      c.rc = getrf_device<T>(m, n, A.cm, A.ld, P.dev, c.d_misc, c.s);
    }
  }
  c.close_piv(P, true);
  int r2 = c.close(A, true);
  int info = c.finish();
  return info ? info : r2;
}

template <typename T>
static int api_getrs(int order, char trans, int n, int nrhs, const T *a, int lda, const int *ipiv, T *b, int ldb) {
  if (order != 101 && order != 102) return -1;
  if (!is_trans_char(trans)) return -2;
  if (n < 0) return -3;
  if (nrhs < 0) return -4;
  if (lda < max1(n)) return -6;
  if (ldb < max1(order == 102 ? n : nrhs)) return -9;
  if (n == 0 || nrhs == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  typename Call<T>::Mat A, B; typename Call<T>::Piv P;
  int early = 0;
  if (c.open(A, order, a, n, n, lda, true) == 0 && c.open(B, order, b, n, nrhs, ldb, true) == 0 &&
      c.open_piv(P, ipiv, n, true) == 0) {
    nancheck_device<T>(0, n, n, A.cm, A.ld, c.d_misc + 1, c.s);
    nancheck_device<T>(0, n, nrhs, B.cm, B.ld, c.d_misc + 2, c.s);
    int fa = 0, fb = 0;
    if (c.nan_flags(&fa, &fb) == 0) {
      if (fa) early = -5; else if (fb) early = -8;
      else c.rc = getrs_device<T>(trans, n, nrhs, A.cm, A.ld, P.dev, B.cm, B.ld, c.s);
    }
  }
  c.close_piv(P, false);
  c.close(A, false);
  int r2 = c.close(B, early == 0);
  int info = c.finish();
  if (early) return early;
  return info ? info : r2;
}

template <typename T>
static int api_gesv(int order, int n, int nrhs, T *a, int lda, int *ipiv, T *b, int ldb) {
  if (order != 101 && order != 102) return -1;
  if (n < 0) return -2;
  if (nrhs < 0) return -3;
  if (lda < max1(n)) return -5;
  if (ldb < max1(order == 102 ? n : nrhs)) return -8;
  if (n == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  typename Call<T>::Mat A, B; typename Call<T>::Piv P;
  int early = 0;
  if (c.open(A, order, a, n, n, lda, true) == 0 && c.open(B, order, b, n, nrhs, ldb, true) == 0 &&
      c.open_piv(P, ipiv, n, false) == 0) {
    nancheck_device<T>(0, n, n, A.cm, A.ld, c.d_misc + 1, c.s);
    nancheck_device<T>(0, n, nrhs, B.cm, B.ld, c.d_misc + 2, c.s);
    int fa = 0, fb = 0;
    if (c.nan_flags(&fa, &fb) == 0) {
      if (fa) early = -4; else if (fb) early = -7;
      else if (nrhs >= 1 && nrhs <= 8 && n >= 128 && n <= 2048 && option("gesv.augment") != 0) {
        // Few right-hand sides, launch-bound sizes: factorise the augmented matrix [A | B].  The extra columns take the
        // panel interchanges and every block-row solve / rank-32 update with the trailing matrix (one more column tile in
        // launches that exist anyway) and leave getrf as Y = inv(L) P B — the forward half of ?getrs and the replay of
        // the interchanges on B cost nothing.  Two device copies of A pay for it (dgesv 512, one right-hand side: the
        // forward half of the fused solve was ~0.1 ms of 0.93).
        const int ldw = n + (n & 1);
        T *W = nullptr;
        if ((c.rc = ws_alloc((void **)&W, sizeof(T) * (size_t)ldw * (size_t)(n + nrhs), c.s)) == 0) {
          T *Y = W + (size_t)n * ldw;
          c.rc = lacpy_device<T>(111, n, n, A.cm, A.ld, W, ldw, c.s);
          if (c.rc == 0) c.rc = lacpy_device<T>(111, n, nrhs, B.cm, B.ld, Y, ldw, c.s);
          if (c.rc == 0) c.rc = getrf_device<T>(n, n + nrhs, W, ldw, P.dev, c.d_misc, c.s);
          if (c.rc == 0) c.rc = lacpy_device<T>(111, n, n, W, ldw, A.cm, A.ld, c.s);
          if (c.rc == 0) {
            int info = 0;       // ?gesv solves only when the factorisation succeeded; B stays as it was otherwise
            cudaMemcpyAsync(&info, c.d_misc, sizeof(int), cudaMemcpyDeviceToHost, c.s);
            cudaStreamSynchronize(c.s);
            if (info == 0) {
              c.rc = getrs_upper_device<T>(n, nrhs, W, ldw, Y, ldw, c.s);
              if (c.rc == 0) c.rc = lacpy_device<T>(111, n, nrhs, Y, ldw, B.cm, B.ld, c.s);
            }
          }
          ws_free(W, c.s);
        }
      } else {
        c.rc = getrf_device<T>(n, n, A.cm, A.ld, P.dev, c.d_misc, c.s);
        if (c.rc == 0 && nrhs > 0) {
          // ?gesv solves only when the factorisation succeeded (info == 0)
          int info = 0;
          cudaMemcpyAsync(&info, c.d_misc, sizeof(int), cudaMemcpyDeviceToHost, c.s);
          cudaStreamSynchronize(c.s);
          if (info == 0) c.rc = getrs_device<T>('N', n, nrhs, A.cm, A.ld, P.dev, B.cm, B.ld, c.s);
        }
      }
    }
  }
  c.close_piv(P, early == 0);
  int r1 = c.close(A, early == 0);
  int r2 = c.close(B, early == 0);
  int info = c.finish();
  if (early) return early;
  return info ? info : (r1 ? r1 : r2);
}

template <typename T>
static int api_getri(int order, int n, T *a, int lda, const int *ipiv) {
  if (order != 101 && order != 102) return -1;
  if (n < 0) return -2;
  if (lda < max1(n)) return -4;
  if (n == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  typename Call<T>::Mat A; typename Call<T>::Piv P;
  int early = 0;
  if (c.open(A, order, a, n, n, lda, true) == 0 && c.open_piv(P, ipiv, n, true) == 0) {
    nancheck_device<T>(0, n, n, A.cm, A.ld, c.d_misc + 1, c.s);
    int fa = 0;
    if (c.nan_flags(&fa, nullptr) == 0) {
      if (fa) early = -3;
      else c.rc = getri_device<T>(n, A.cm, A.ld, P.dev, c.d_misc, c.s);
    }
  }
  c.close_piv(P, false);
  int r1 = c.close(A, early == 0);
  int info = c.finish();
  if (early) return early;
  return info ? info : r1;
}

template <typename T>
static int api_potrf(int order, char uplo, int n, T *a, int lda) {
  if (order != 101 && order != 102) return -1;
  if (!is_uplo_char(uplo)) return -2;
  if (n < 0) return -3;
  if (lda < max1(n)) return -5;
  if (n == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  // Row-major uplo triangle == column-major opposite triangle of A^T = conj(A) for a
  // Hermitian matrix: factor in place without any transpose.  L L^H = A  <=>  (conj L)(conj L)^H
  // = conj(A); the stored row-major lower triangle of L is the column-major upper
  // triangle U = L^T with U^H U = conj(L) L^T = conj(A) — i.e. exactly potrf('U') on the same bytes.
  char u = uplo;
  if (order == 101) u = (uplo == 'L' || uplo == 'l') ? 'U' : 'L';
  Staged st;
  int early = 0;
  if ((c.rc = st.open(a, n, n, lda, sizeof(T), true, c.s)) == 0) {
    nancheck_device<T>((u == 'L' || u == 'l') ? 'L' : 'U', n, n, (const T *)st.dev, st.ld, c.d_misc + 1, c.s);
    int fa = 0;
    if (c.nan_flags(&fa, nullptr) == 0) {
      if (fa) early = -4;
      else c.rc = potrf_device<T>(u, n, (T *)st.dev, st.ld, c.d_misc, c.s);
    }
  }
  int r1 = st.close(early == 0 && c.rc == 0, c.s);
  int info = c.finish();
  if (early) return early;
  return info ? info : r1;
}

template <typename T>
static int api_potrs(int order, char uplo, int n, int nrhs, const T *a, int lda, T *b, int ldb) {
  if (order != 101 && order != 102) return -1;
  if (!is_uplo_char(uplo)) return -2;
  if (n < 0) return -3;
  if (nrhs < 0) return -4;
  if (lda < max1(n)) return -6;
  if (ldb < max1(order == 102 ? n : nrhs)) return -8;
  if (n == 0 || nrhs == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  typename Call<T>::Mat A, B;
  int early = 0;
  if (c.open(A, order, a, n, n, lda, true) == 0 && c.open(B, order, b, n, nrhs, ldb, true) == 0) {
    nancheck_device<T>((uplo == 'L' || uplo == 'l') ? 'L' : 'U', n, n, A.cm, A.ld, c.d_misc + 1, c.s);
    nancheck_device<T>(0, n, nrhs, B.cm, B.ld, c.d_misc + 2, c.s);
    int fa = 0, fb = 0;
    if (c.nan_flags(&fa, &fb) == 0) {
      if (fa) early = -5; else if (fb) early = -7;
      else c.rc = potrs_device<T>(uplo, n, nrhs, A.cm, A.ld, B.cm, B.ld, c.s);
    }
  }
  c.close(A, false);
  int r2 = c.close(B, early == 0);
  int info = c.finish();
  if (early) return early;
  return info ? info : r2;
}

template <typename T>
static int api_posv(int order, char uplo, int n, int nrhs, T *a, int lda, T *b, int ldb) {
  if (order != 101 && order != 102) return -1;
  if (!is_uplo_char(uplo)) return -2;
  if (n < 0) return -3;
  if (nrhs < 0) return -4;
  if (lda < max1(n)) return -6;
  if (ldb < max1(order == 102 ? n : nrhs)) return -8;
  if (n == 0) return 0;
  Call<T> c;
  if (c.begin()) return c.rc;
  typename Call<T>::Mat A, B;
  int early = 0;
  if (c.open(A, order, a, n, n, lda, true) == 0 && c.open(B, order, b, n, nrhs, ldb, true) == 0) {
    nancheck_device<T>((uplo == 'L' || uplo == 'l') ? 'L' : 'U', n, n, A.cm, A.ld, c.d_misc + 1, c.s);
    nancheck_device<T>(0, n, nrhs, B.cm, B.ld, c.d_misc + 2, c.s);
    int fa = 0, fb = 0;
    if (c.nan_flags(&fa, &fb) == 0) {
      if (fa) early = -5; else if (fb) early = -7;
      else {
        c.rc = potrf_device<T>(uplo, n, A.cm, A.ld, c.d_misc, c.s);
        if (c.rc == 0 && nrhs > 0) {
          int info = 0;
          cudaMemcpyAsync(&info, c.d_misc, sizeof(int), cudaMemcpyDeviceToHost, c.s);
          cudaStreamSynchronize(c.s);
          if (info == 0) c.rc = potrs_device<T>(uplo, n, nrhs, A.cm, A.ld, B.cm, B.ld, c.s);
        }
      }
    }
  }
  int r1 = c.close(A, early == 0);
  int r2 = c.close(B, early == 0);
  int info = c.finish();
  if (early) return early;
  return info ? info : (r1 ? r1 : r2);
}

}  // namespace jb

using namespace jb;
extern "C" {
#define LAPACKE_APIS(t, T, CT)                                                                                     \
  int jb_LAPACKE_##t##gesv(int order, int n, int nrhs, CT *a, int lda, int *ipiv, CT *b, int ldb) {               \
    return api_gesv<T>(order, n, nrhs, (T *)a, lda, ipiv, (T *)b, ldb); }                                           \
  int jb_LAPACKE_##t##getrf(int order, int m, int n, CT *a, int lda, int *ipiv) {                                  \
    return api_getrf<T>(order, m, n, (T *)a, lda, ipiv); }                                                          \
  int jb_LAPACKE_##t##getrs(int order, char trans, int n, int nrhs, const CT *a, int lda, const int *ipiv, CT *b, int ldb) { \
    return api_getrs<T>(order, trans, n, nrhs, (const T *)a, lda, ipiv, (T *)b, ldb); }                            \
  int jb_LAPACKE_##t##getri(int order, int n, CT *a, int lda, const int *ipiv) {                                   \
    return api_getri<T>(order, n, (T *)a, lda, ipiv); }                                                             \
  int jb_LAPACKE_##t##potrf(int order, char uplo, int n, CT *a, int lda) {                                         \
    return api_potrf<T>(order, uplo, n, (T *)a, lda); }                                                             \
  int jb_LAPACKE_##t##potrs(int order, char uplo, int n, int nrhs, const CT *a, int lda, CT *b, int ldb) {         \
    return api_potrs<T>(order, uplo, n, nrhs, (const T *)a, lda, (T *)b, ldb); }                                   \
  int jb_LAPACKE_##t##posv(int order, char uplo, int n, int nrhs, CT *a, int lda, CT *b, int ldb) {                \
    return api_posv<T>(order, uplo, n, nrhs, (T *)a, lda, (T *)b, ldb); }                                           \
  int LAPACKE_##t##gesv(int order, int n, int nrhs, CT *a, int lda, int *ipiv, CT *b, int ldb) {                  \
    return jb_LAPACKE_##t##gesv(order, n, nrhs, a, lda, ipiv, b, ldb); }                                            \
  int LAPACKE_##t##getrf(int order, int m, int n, CT *a, int lda, int *ipiv) {                                     \
    return jb_LAPACKE_##t##getrf(order, m, n, a, lda, ipiv); }                                                      \
  int LAPACKE_##t##getrs(int order, char trans, int n, int nrhs, const CT *a, int lda, const int *ipiv, CT *b, int ldb) { \
    return jb_LAPACKE_##t##getrs(order, trans, n, nrhs, a, lda, ipiv, b, ldb); }                                    \
  int LAPACKE_##t##getri(int order, int n, CT *a, int lda, const int *ipiv) {                                      \
    return jb_LAPACKE_##t##getri(order, n, a, lda, ipiv); }                                                         \
  int LAPACKE_##t##potrf(int order, char uplo, int n, CT *a, int lda) {                                            \
    return jb_LAPACKE_##t##potrf(order, uplo, n, a, lda); }                                                         \
  int LAPACKE_##t##potrs(int order, char uplo, int n, int nrhs, const CT *a, int lda, CT *b, int ldb) {            \
    return jb_LAPACKE_##t##potrs(order, uplo, n, nrhs, a, lda, b, ldb); }                                           \
  int LAPACKE_##t##posv(int order, char uplo, int n, int nrhs, CT *a, int lda, CT *b, int ldb) {                   \
    return jb_LAPACKE_##t##posv(order, uplo, n, nrhs, a, lda, b, ldb); }

LAPACKE_APIS(s, float, float)
LAPACKE_APIS(d, double, double)
LAPACKE_APIS(c, cfloat, void)
LAPACKE_APIS(z, cdouble, void)
}
