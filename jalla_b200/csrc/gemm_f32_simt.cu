// FP32 GEMM with full-precision FMA on the CUDA cores (no TF32, no tensor pipe) for sm_100a — the kernel
// behind cblas_sgemm (/root/reference/Numeric/Jalla/Foreign/BLAS.chs:149, reached from Matrix.hs:482).
//
// 128x128 C tile, BK = 8, 256 threads, 8x8 register micro-tile per thread (two 4-wide groups per dimension so
// every shared-memory fragment read is a conflict-free 128-bit load), shared memory double buffered with the
// next tile's global loads held in registers across the FFMA block.  Operands are brought to one shared
// layout As[k][m] / Bs[k][n]: the operand whose contiguous dimension is m (or n) is copied with 128-bit
// loads/stores, the one whose contiguous dimension is k is transposed on the way in (the reason this kernel
// stages through registers instead of TMA: TMA cannot transpose 4-byte elements).  All four NoTrans/Trans
// combinations, arbitrary m, n, k; requires 16-byte aligned bases and leading dimensions that are multiples of 4
// (everything else goes to gemm_generic.cu).  Bound: FP32 FMA pipe, measured peak 72.35 TFLOP/s (tools/peaks.cu).
#include "jb_common.cuh"

namespace jb {
namespace sgemm {

constexpr int BM = 128, BN = 128, BK = 8, THREADS = 256;    // BK = 16 measured slower (47.7 vs 48.8 TFLOP/s)
constexpr int CH = BK / 8;                   // 4-element chunks per thread per operand per k tile
constexpr int PAD = 4;                       // row pitch 132 floats: transposing stores hit distinct banks
constexpr int PITCH = BM + PAD;

// Loads one 4-element chunk of an operand tile into registers.
//  CONTIG_MN = true : the chunk is 4 consecutive m (or n) at one k   -> one 128-bit load
//  CONTIG_MN = false: the chunk is 4 consecutive k at one m (or n)   -> one 128-bit load, transposed on store
template <bool CONTIG_MN>
__device__ __forceinline__ float4 load_chunk(const float *__restrict__ P, int ld, int idx0, int k0, int idx_max, int k_max,
                                             int tid, int c) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (CONTIG_MN) {
    const int i4 = (tid & 31) * 4, kk = (tid >> 5) + 8 * c;  // 32 chunks along idx x 8 k per pass
    const int gi = idx0 + i4, gk = k0 + kk;
    if (gk < k_max) {
      const float *p = P + (size_t)gk * ld + gi;
      if (gi + 3 < idx_max) v = *reinterpret_cast<const float4 *>(p);
      else {
        if (gi < idx_max) v.x = p[0];
        if (gi + 1 < idx_max) v.y = p[1];
        if (gi + 2 < idx_max) v.z = p[2];
      }
    }
  } else {
    const int ii = tid >> 1, k4 = (tid & 1) * 4 + 8 * c;     // 128 idx x 2 chunks along k per pass
    const int gi = idx0 + ii, gk = k0 + k4;
    if (gi < idx_max) {
      const float *p = P + (size_t)gi * ld + gk;
      if (gk + 3 < k_max) v = *reinterpret_cast<const float4 *>(p);
      else {
        if (gk < k_max) v.x = p[0];
        if (gk + 1 < k_max) v.y = p[1];
        if (gk + 2 < k_max) v.z = p[2];
      }
    }
  }
  return v;
}
template <bool CONTIG_MN>
__device__ __forceinline__ void store_chunk(float (*S)[PITCH], float4 v, int tid, int c) {
  if (CONTIG_MN) {
    const int i4 = (tid & 31) * 4, kk = (tid >> 5) + 8 * c;
    *reinterpret_cast<float4 *>(&S[kk][i4]) = v;
  } else {
    const int ii = tid >> 1, k4 = (tid & 1) * 4 + 8 * c;
    S[k4 + 0][ii] = v.x; S[k4 + 1][ii] = v.y; S[k4 + 2][ii] = v.z; S[k4 + 3][ii] = v.w;
  }
}

// A_MN: op(A) has m contiguous in memory (NoTrans, column-major).  B_MN: op(B) has n contiguous (Trans).
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 2)
sgemm_simt_kernel(int m, int n, int k, float alpha, const float *__restrict__ A, int lda, const float *__restrict__ B,
                  int ldb, float beta, float *__restrict__ C, int ldc, int tri, int tiles_m) {
  __shared__ __align__(16) float As[2][BK][PITCH];
  __shared__ __align__(16) float Bs[2][BK][PITCH];
  const int tid = threadIdx.x;
  // band rasterisation (16 tile rows) for L2 reuse
  const int tiles_n = (n + BN - 1) / BN;
  const int per_band = 16 * tiles_n;
  const int band = blockIdx.x / per_band, in_band = blockIdx.x % per_band;
  const int first_m = band * 16, gsz = min(tiles_m - first_m, 16);
  const int tm = first_m + in_band % gsz, tn = in_band / gsz;
  const int row0 = tm * BM, col0 = tn * BN;
  if (tri == 'L' && row0 + BM - 1 < col0) return;
  if (tri == 'U' && col0 + BN - 1 < row0) return;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  const int KT = (k + BK - 1) / BK;
  float4 ra[CH], rb[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) {
    ra[c] = load_chunk<A_MN>(A, lda, row0, 0, m, k, tid, c);
    rb[c] = load_chunk<B_MN>(B, ldb, col0, 0, n, k, tid, c);
  }
#pragma unroll
  for (int c = 0; c < CH; c++) { store_chunk<A_MN>(As[0], ra[c], tid, c); store_chunk<B_MN>(Bs[0], rb[c], tid, c); }
  __syncthreads();
  for (int kt = 0; kt < KT; kt++) {
    const int cur = kt & 1;
    if (kt + 1 < KT) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        ra[c] = load_chunk<A_MN>(A, lda, row0, (kt + 1) * BK, m, k, tid, c);
        rb[c] = load_chunk<B_MN>(B, ldb, col0, (kt + 1) * BK, n, k, tid, c);
      }
    }
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][kk][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][kk][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
#pragma unroll
      for (int c = 0; c < CH; c++) { store_chunk<A_MN>(As[cur ^ 1], ra[c], tid, c); store_chunk<B_MN>(Bs[cur ^ 1], rb[c], tid, c); }
      __syncthreads();
    }
  }

  // epilogue: C = alpha*acc + beta*C; rows of a thread come in two groups of four consecutive elements
  const bool beta0 = (beta == 0.f);
  const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && tri == 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = col0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
    if (col >= n) continue;
    float *ccol = C + (size_t)col * ldc;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r0 = row0 + h * 64 + tx * 4;
      if (r0 >= m) continue;
      float v[4] = {alpha * acc[h * 4 + 0][j], alpha * acc[h * 4 + 1][j], alpha * acc[h * 4 + 2][j], alpha * acc[h * 4 + 3][j]};
      if (vec_ok && r0 + 3 < m) {
        if (!beta0) {
          const float4 c0 = *reinterpret_cast<const float4 *>(ccol + r0);
          v[0] = fmaf(beta, c0.x, v[0]); v[1] = fmaf(beta, c0.y, v[1]); v[2] = fmaf(beta, c0.z, v[2]); v[3] = fmaf(beta, c0.w, v[3]);
        }
        *reinterpret_cast<float4 *>(ccol + r0) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int row = r0 + e;
          if (row >= m) continue;
          if (tri == 'L' && row < col) continue;
          if (tri == 'U' && row > col) continue;
          float r = v[e];
          if (!beta0) r = fmaf(beta, ccol[row], r);
          ccol[row] = r;
        }
      }
    }
  }
}

template <bool A_MN, bool B_MN>
static int launch(int m, int n, int k, float alpha, const float *A, int lda, const float *B, int ldb, float beta, float *C,
                  int ldc, int tri, cudaStream_t s) {
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  sgemm_simt_kernel<A_MN, B_MN><<<tiles_m * tiles_n, THREADS, 0, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, tiles_m);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 1;
}
}  // namespace sgemm

int sgemm_simt_try(int ta, int tb, int m, int n, int k, float alpha, const float *A, int lda, const float *B, int ldb,
                   float beta, float *C, int ldc, cudaStream_t s, int tri) {
  using namespace sgemm;
  if (m < 64 || n < 64 || k < 1) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 3) || (ldb & 3)) return 0;
  const bool a_mn = (ta == 111), b_mn = (tb != 111);
  if (a_mn && b_mn) return launch<true, true>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, s);
  if (a_mn && !b_mn) return launch<true, false>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, s);
  if (!a_mn && b_mn) return launch<false, true>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, s);
  return launch<false, false>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, s);
}
}  // namespace jb
