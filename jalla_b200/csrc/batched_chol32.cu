// Batched 32x32 Double / Float Cholesky factorise / solve, lower storage — the SPD sibling of the batched LU (the
// batched form of /root/reference/Numeric/Jalla/Foreign/LAPACKE.chs:733-736 potrf and :741-744 potrs).
//
// The generic kernels (batched_small.cu: potrf32_kernel / potrs32_kernel, all types, n <= 32, both triangles) branch per
// element on lane == j / lane >= j inside the update and read the factor both by rows and by columns (the column reads are
// 256-byte strided): 14.6 ms / 8.7 ms per 10^6 systems = 0.09 / 0.08 of the HBM roofline.  Here, for the headline shape:
//   potrf: one warp per system, lane = row i of L, the row in registers.  Step k: d = a_kk by shuffle, s = sqrt(d),
//     r = 1/s, l_i = a_ik * r (the ?potf2 sequence of oracle_impl.inc).  Unlike LU nothing has to be selected: EVERY lane
//     publishes its own l_i with one full-warp 8-byte store (double buffered by step parity, one __syncwarp per step), then
//     a_ij = fma(-l_i, l_j, a_ij) for j > k with broadcast 128-bit loads of l — unpredicated: entries above the diagonal are
//     garbage nobody stores, and the diagonal's fma(-l_j, l_j, a_jj) IS ?potf2's real accumulation.
//     A non-positive or NaN pivot only records info (first failing column); the system's memory is then left untouched.
//   potrs: rows of L by coalesced column loads, the transpose (rows of L^T) through a padded shared-memory tile, forward
//     and backward substitution with the oracle's x_k = y_k * (1/l_kk), x_i = fma(-l_ik, x_k, x_i); up to four right-hand
//     sides per pass.
// Results are bit-identical to the generic kernels and the oracle (tests/test_batched.py).  The kernels are templates over
// the real type: Float runs the same sequence in single precision (one 128-bit broadcast load carries four entries).
#include "batched_common.cuh"

namespace jb {

__device__ __forceinline__ double2 lds2d(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

template <typename R> __device__ __forceinline__ R mul_rn(R a, R b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename R> __device__ __forceinline__ R fma_rn(R a, R b, R c);
template <> __device__ __forceinline__ double fma_rn<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <> __device__ __forceinline__ float fma_rn<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <typename R> __device__ __forceinline__ R sqrt_rn(R a);
template <> __device__ __forceinline__ double sqrt_rn<double>(double a) { return __dsqrt_rn(a); }
template <> __device__ __forceinline__ float sqrt_rn<float>(float a) { return __fsqrt_rn(a); }
template <typename R> __device__ __forceinline__ R rcp_rn(R a);
template <> __device__ __forceinline__ double rcp_rn<double>(double a) { return __ddiv_rn(1.0, a); }
template <> __device__ __forceinline__ float rcp_rn<float>(float a) { return __fdiv_rn(1.0f, a); }

template <typename R, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
potrf32r_kernel(R *__restrict__ Ag, long long stride_a, int *__restrict__ info_g, int batch) {
  constexpr int V = 16 / sizeof(R);                     // entries per 128-bit broadcast load
  __shared__ __align__(16) R sl_all[WARPS][2][N32 + V];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  R (*sl)[N32 + V] = sl_all[warp];
  for (long long sys = (long long)blockIdx.x * WARPS + warp; sys < batch; sys += (long long)gridDim.x * WARPS) {
    R *const A = Ag + sys * stride_a;
    R a[N32];
#pragma unroll
    for (int j = 0; j < N32; j++) a[j] = lane >= j ? __ldcs(A + lane + N32 * j) : (R)0;
    int info = 0;
#pragma unroll
    for (int k = 0; k < N32; k++) {
      const R d = __shfl_sync(0xffffffffu, a[k], k);
      if (!(d > (R)0) && info == 0) info = k + 1;              // warp-uniform; what follows for this system is never stored
      const R sq = sqrt_rn<R>(d), rinv = rcp_rn<R>(sq);
      const R l = mul_rn<R>(a[k], rinv);
      a[k] = lane == k ? sq : l;
      R *const buf = sl[k & 1];
      buf[lane] = l;
      __syncwarp();
#pragma unroll
      for (int j = (k + 1) & ~(V - 1); j < N32; j += V) {
        if constexpr (V == 2) {
          const double2 u = lds2d(reinterpret_cast<const double *>(&buf[j]));
          if (j > k) a[j] = fma_rn<R>(-l, u.x, a[j]);
          a[j + 1] = fma_rn<R>(-l, u.y, a[j + 1]);
        } else {
          const float4 u = *reinterpret_cast<const float4 *>(&buf[j]);
          if (j > k) a[j] = fma_rn<R>(-l, u.x, a[j]);
          if (j + 1 > k) a[j + 1] = fma_rn<R>(-l, u.y, a[j + 1]);
          if (j + 2 > k) a[j + 2] = fma_rn<R>(-l, u.z, a[j + 2]);
          a[j + 3] = fma_rn<R>(-l, u.w, a[j + 3]);
        }
      }
    }
    if (info == 0) {
#pragma unroll
      for (int j = 0; j < N32; j++)
        if (lane >= j) __stcs(A + lane + N32 * j, a[j]);
    }
    if (lane == 0 && info_g) info_g[sys] = info;
    __syncwarp();
  }
}

constexpr int CH_LD = N32 + 1;        // padded tile: row stores and column loads are both conflict-free
template <typename R, int WARPS, int MINB, int NR>
__global__ void __launch_bounds__(32 * WARPS, MINB)
potrs32r_kernel(const R *__restrict__ Ag, long long stride_a, R *__restrict__ Bg, int ldb, long long stride_b, int nrhs, int batch) {
  __shared__ R st_all[WARPS][N32 * CH_LD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  R *const st = st_all[warp];
  for (long long sys = (long long)blockIdx.x * WARPS + warp; sys < batch; sys += (long long)gridDim.x * WARPS) {
    const R *const A = Ag + sys * stride_a;
    R r[N32];                    // r[j] = L[lane][j]; after the forward sweep r[j] = L[j][lane]
#pragma unroll
    for (int j = 0; j < N32; j++) r[j] = lane >= j ? __ldcs(A + lane + N32 * j) : (R)0;
    R diag = (R)1;
#pragma unroll
    for (int j = 0; j < N32; j++) if (j == lane) diag = r[j];
    const R dinv = rcp_rn<R>(diag);
    R x[NR];
    R *const B = Bg + sys * stride_b;
#pragma unroll
    for (int c = 0; c < NR; c++) x[c] = c < nrhs ? B[lane + (size_t)c * ldb] : (R)0;
    // L y = b
#pragma unroll
    for (int k = 0; k < N32; k++) {
#pragma unroll
      for (int c = 0; c < NR; c++) {
        if (lane == k) x[c] = mul_rn<R>(x[c], dinv);
        const R xk = __shfl_sync(0xffffffffu, x[c], k);
        if (lane > k) x[c] = fma_rn<R>(-r[k], xk, x[c]);
      }
    }
    // rows of L^T
#pragma unroll
    for (int j = 0; j < N32; j++) st[lane * CH_LD + j] = r[j];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < N32; j++) r[j] = st[j * CH_LD + lane];
    __syncwarp();
    // L^T x = y
#pragma unroll
    for (int k = N32 - 1; k >= 0; k--) {
#pragma unroll
      for (int c = 0; c < NR; c++) {
        if (lane == k) x[c] = mul_rn<R>(x[c], dinv);
        const R xk = __shfl_sync(0xffffffffu, x[c], k);
        if (lane < k) x[c] = fma_rn<R>(-r[k], xk, x[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NR; c++)
      if (c < nrhs) B[lane + (size_t)c * ldb] = x[c];
  }
}

// Return 1 when the call was taken (Double or Float, n == 32, lda == 32, lower storage), 0 when it is not this kernel's shape.
template <typename R>
static int potrf32_try_t(char uplo, int n, R *a, int lda, long long stride_a, int *info, int batch, cudaStream_t s) {
  if (!(uplo == 'L' || uplo == 'l') || n != N32 || lda != N32 || option("chol32.generic") > 0) return 0;
  const long cap = (long)num_sms() * 16;
  potrf32r_kernel<R, 1, 16><<<(int)(batch < cap ? batch : cap), 32, 0, s>>>(a, stride_a, info, batch);
  count_launch();
  return 1;
}
template <typename R>
static int potrs32_try_t(char uplo, int n, int nrhs, const R *a, int lda, long long stride_a, R *b, int ldb, long long stride_b, int batch, cudaStream_t s) {
  if (!(uplo == 'L' || uplo == 'l') || n != N32 || lda != N32 || nrhs > 4 || option("chol32.generic") > 0) return 0;
  const long ctas = ((long)batch + 3) / 4, cap = (long)num_sms() * 4;
  if (nrhs == 1) potrs32r_kernel<R, 4, 3, 1><<<(int)(ctas < cap ? ctas : cap), 128, 0, s>>>(a, stride_a, b, ldb, stride_b, nrhs, batch);
  else if (nrhs == 2) potrs32r_kernel<R, 4, 3, 2><<<(int)(ctas < cap ? ctas : cap), 128, 0, s>>>(a, stride_a, b, ldb, stride_b, nrhs, batch);
  else potrs32r_kernel<R, 4, 3, 4><<<(int)(ctas < cap ? ctas : cap), 128, 0, s>>>(a, stride_a, b, ldb, stride_b, nrhs, batch);
  count_launch();
  return 1;
}
int potrf32d_try(char uplo, int n, double *a, int lda, long long stride_a, int *info, int batch, cudaStream_t s) {
  return potrf32_try_t<double>(uplo, n, a, lda, stride_a, info, batch, s);
}
int potrf32s_try(char uplo, int n, float *a, int lda, long long stride_a, int *info, int batch, cudaStream_t s) {
  return potrf32_try_t<float>(uplo, n, a, lda, stride_a, info, batch, s);
}
int potrs32d_try(char uplo, int n, int nrhs, const double *a, int lda, long long stride_a, double *b, int ldb, long long stride_b, int batch, cudaStream_t s) {
  return potrs32_try_t<double>(uplo, n, nrhs, a, lda, stride_a, b, ldb, stride_b, batch, s);
}
int potrs32s_try(char uplo, int n, int nrhs, const float *a, int lda, long long stride_a, float *b, int ldb, long long stride_b, int batch, cudaStream_t s) {
  return potrs32_try_t<float>(uplo, n, nrhs, a, lda, stride_a, b, ldb, stride_b, batch, s);
}

}  // namespace jb
