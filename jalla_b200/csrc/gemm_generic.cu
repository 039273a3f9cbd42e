// Generic SIMT GEMM for every element type / transpose combination / alignment.
// It is the path for shapes the tuned kernels decline (odd leading dimensions, tiny
// problems, complex float) and the trailing-update workhorse of the blocked
// factorisations until a tuned kernel takes the shape.  Computes
//   C <- alpha*op(A)*op(B) + beta*C      (column-major; cblas_?gemm semantics,
//   /root/reference/Numeric/Jalla/Matrix.hs:477-490 is the caller being served)
// beta == 0 never reads C (Jalla passes uninitialised memory: Matrix.hs:458-459).
#include "jb_common.cuh"

namespace jb {

template <typename T, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_generic_kernel(int ta, int tb, int m, int n, int k, T alpha, const T *__restrict__ A, int lda,
                    const T *__restrict__ B, int ldb, T beta, T *__restrict__ C, int ldc, int tri) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ T As[BK][BM + 1];
  __shared__ T Bs[BK][BN + 1];
  const int tiles_m = (m + BM - 1) / BM;
  const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
  const int row0 = tm * BM, col0 = tn * BN;
  if (tri == 'L' && row0 + BM - 1 < col0) return;
  if (tri == 'U' && col0 + BN - 1 < row0) return;
  const int tid = threadIdx.x;
  const int tx = tid % (BM / TM), ty = tid / (BM / TM);
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = Num<T>::zero();

  for (int k0 = 0; k0 < k; k0 += BK) {
    // op(A) tile: BM x BK
    for (int e = tid; e < BM * BK; e += NT) {
      int i, p;
      if (ta == 111) { i = e % BM; p = e / BM; } else { p = e % BK; i = e / BK; }
      int gi = row0 + i, gp = k0 + p;
      T v = Num<T>::zero();
      if (gi < m && gp < k) {
        v = (ta == 111) ? A[(size_t)gi + (size_t)gp * lda] : A[(size_t)gp + (size_t)gi * lda];
        if (ta == 113) v = Num<T>::conj(v);
      }
      As[p][i] = v;
    }
    for (int e = tid; e < BN * BK; e += NT) {
      int j, p;
      if (tb == 111) { p = e % BK; j = e / BK; } else { j = e % BN; p = e / BN; }
      int gj = col0 + j, gp = k0 + p;
      T v = Num<T>::zero();
      if (gj < n && gp < k) {
        v = (tb == 111) ? B[(size_t)gp + (size_t)gj * ldb] : B[(size_t)gj + (size_t)gp * ldb];
        if (tb == 113) v = Num<T>::conj(v);
      }
      Bs[p][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < BK; p++) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[p][tx + i * (BM / TM)];
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = Bs[p][ty + j * (BN / TN)];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = Num<T>::fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool beta0 = Num<T>::is_zero(beta);
#pragma unroll
  for (int j = 0; j < TN; j++) {
    int gj = col0 + ty + j * (BN / TN);
    if (gj >= n) continue;
#pragma unroll
    for (int i = 0; i < TM; i++) {
      int gi = row0 + tx + i * (BM / TM);
      if (gi >= m) continue;
      if (tri == 'L' && gi < gj) continue;
      if (tri == 'U' && gi > gj) continue;
      T *c = C + (size_t)gi + (size_t)gj * ldc;
      T r = Num<T>::mul(alpha, acc[i][j]);
      if (!beta0) r = Num<T>::add(r, Num<T>::mul(beta, *c));
      *c = r;
    }
  }
}

template <typename T>
static int launch_generic(int ta, int tb, int m, int n, int k, T alpha, const T *A, int lda, const T *B, int ldb,
                          T beta, T *C, int ldc, cudaStream_t s, int tri) {
  // skinny or mid-sized shapes (nrhs=1 solves, small panels, the reference's 512^3 case): smaller tiles so that
  // the grid still covers the 148 SMs
  if (m <= 64 || n <= 64 || (long)((m + 63) / 64) * ((n + 63) / 64) < 120) {
    constexpr int BM = 32, BN = 32, BK = 16;
    long tiles = (long)((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    gemm_generic_kernel<T, BM, BN, BK, 2, 2><<<(unsigned)tiles, 256, 0, s>>>(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  } else {
    constexpr int BM = 64, BN = 64, BK = 16;
    long tiles = (long)((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    gemm_generic_kernel<T, BM, BN, BK, 4, 4><<<(unsigned)tiles, 256, 0, s>>>(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  }
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T> struct FastPath {
  static int try_it(int, int, int, int, int, T, const T *, int, const T *, int, T, T *, int, cudaStream_t, int) { return 0; }
};
template <> struct FastPath<double> {
  static int try_it(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
                    double beta, double *C, int ldc, cudaStream_t s, int tri) {
    return dgemm_dmma_try(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri);
  }
};
template <> struct FastPath<float> {
  static int try_it(int ta, int tb, int m, int n, int k, float alpha, const float *A, int lda, const float *B, int ldb,
                    float beta, float *C, int ldc, cudaStream_t s, int tri) {
    if (int took = sgemm_tma_try(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri)) return took;
    return sgemm_simt_try(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri);
  }
};
template <> struct FastPath<cfloat> {
  static int try_it(int ta, int tb, int m, int n, int k, cfloat alpha, const cfloat *A, int lda, const cfloat *B, int ldb,
                    cfloat beta, cfloat *C, int ldc, cudaStream_t s, int tri) {
    return cgemm_tma_try(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri);
  }
};
template <> struct FastPath<cdouble> {
  static int try_it(int ta, int tb, int m, int n, int k, cdouble alpha, const cdouble *A, int lda, const cdouble *B, int ldb,
                    cdouble beta, cdouble *C, int ldc, cudaStream_t s, int tri) {
    return zgemm_dmma_try(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri);
  }
};

template <typename T>
int gemm_device(int ta, int tb, int m, int n, int k, T alpha, const T *A, int lda, const T *B, int ldb, T beta, T *C,
                int ldc, cudaStream_t s, int tri) {
  if (m <= 0 || n <= 0) return 0;
  if (Num<T>::is_zero(alpha)) k = 0;          // BLAS: A and B are not referenced when alpha == 0
  if (k <= 0 && !Num<T>::is_zero(beta) && Num<T>::re(beta) == 1 && Num<T>::im(beta) == 0) return 0;
  if (!Num<T>::cplx) { if (ta == 113) ta = 112; if (tb == 113) tb = 112; }
  if (k > 0 && option("gemm.force_generic") <= 0) {
    int took = FastPath<T>::try_it(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri);
    if (took < 0) return took;
    if (took) return 0;
  }
  return launch_generic<T>(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s, tri);
}
template int gemm_device<float>(int, int, int, int, int, float, const float *, int, const float *, int, float, float *, int, cudaStream_t, int);
template int gemm_device<double>(int, int, int, int, int, double, const double *, int, const double *, int, double, double *, int, cudaStream_t, int);
template int gemm_device<cfloat>(int, int, int, int, int, cfloat, const cfloat *, int, const cfloat *, int, cfloat, cfloat *, int, cudaStream_t, int);
template int gemm_device<cdouble>(int, int, int, int, int, cdouble, const cdouble *, int, const cdouble *, int, cdouble, cdouble *, int, cudaStream_t, int);

}  // namespace jb
