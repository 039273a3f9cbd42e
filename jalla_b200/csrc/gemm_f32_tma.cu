// FP32 GEMM with full-precision FMA on the CUDA cores (no TF32, no tensor pipe), operands staged by TMA — the kernel
// behind cblas_sgemm (/root/reference/Numeric/Jalla/Foreign/BLAS.chs:149, reached from Matrix.hs:482) for shapes the
// TMA descriptors can describe (16-byte aligned bases, leading dimensions that are multiples of 4).
//
// Round 1's kernel (gemm_f32_simt.cu, kept for everything else) staged both operands through registers because the
// operand whose contiguous dimension is k has to be transposed on its way into shared memory; it issued 79 % of its slots
// for a 69 % busy FMA pipe.  Here nothing is transposed: an operand arrives in shared memory in its own major and the
// FRAGMENT loads adapt —
//   MN-major tile [32 k][128 idx] (no swizzle): a thread reads 4 consecutive rows/columns at one k     (LDS.128)
//   K-major  tile [128 idx][32 k] (128-byte swizzle): a thread reads 4 consecutive k of one row/column (LDS.128)
// and the thread's 8 rows (columns) are 2 x 4 consecutive ones in the first case, {t, t+16, ..., t+112} in the second, so
// that every warp-wide load touches each bank group at most twice (the minimum for 256 bytes).  Either way a 4-k block
// costs 16 vector loads for 256 FFMAs.  One elected thread per CTA issues the TMA loads into a 3-stage mbarrier ring,
// two CTAs of 256 threads share an SM (persistent tile loop, band rasterisation for L2 reuse).  Large products whose A is
// MN-major (NoTrans — what Jalla's ## passes) run a 256 x 128 tile instead: 16 x 8 per thread, one CTA per SM with up to
// 255 registers, 4 stages — 6 operand loads per 64 FFMA2 instead of 4 per 32.
// Bound: FP32 FMA pipe, measured peak 72.35 TFLOP/s (tools/peaks.cu).
#include "dmma_common.cuh"

namespace jb {
namespace sgemm_tma {
using dmma::smem_u32; using dmma::mbar_init; using dmma::mbar_expect_tx; using dmma::mbar_arrive; using dmma::mbar_wait; using dmma::tma_load_2d;

constexpr int BK = 32, THREADS = 256, WARPS = THREADS / 32;
constexpr int GROUP_M = 16;
// Tile shapes: a thread owns TM rows x TN columns of the 16*TM x 16*TN C tile.
//   <8, 8>   128 x 128, 3 stages, two CTAs per SM (128 registers)                       — every operand layout
//   <16, 8>  256 x 128, 4 stages, one CTA per SM (up to 255 registers), A MN-major only  — half the operand loads per FFMA2
template <int TM, int TN> struct Shape {
  static constexpr int BM = 16 * TM, BN = 16 * TN;
  static constexpr int STAGES = (TM == 8 && TN == 8) ? 3 : 4, MINB = (TM == 8 && TN == 8) ? 2 : 1;
  static constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 64;
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// Packed FP32 FMA (FFMA2 on sm_100a): d.lo = a.lo*b.lo + c.lo, d.hi = a.hi*b.hi + c.hi, each an IEEE fma.rn.f32 — the
// same arithmetic as two FFMAs in half the issue slots, and its 64-bit operands read BOTH register banks at once, so
// the even/odd bank conflicts that hold scalar FFMA streams at ~2/3 of the pipe (ncu: 20 % dispatch stalls) cannot occur.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// index (row of A / column of B inside the tile) of a thread's e-th element, e = 0..7; t = its position 0..15
template <bool MN> __device__ __forceinline__ int elem_idx(int t, int e) { return MN ? (e >> 2) * 64 + t * 4 + (e & 3) : t + 16 * e; }

template <bool A_MN, bool B_MN, int TM, int TN>
__global__ void __launch_bounds__(THREADS, Shape<TM, TN>::MINB)
sgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int m, int n, int k, float alpha,
                 float beta, float *__restrict__ C, int ldc, int tri) {
  using SH = Shape<TM, TN>;
  constexpr int BM = SH::BM, BN = SH::BN, STAGES = SH::STAGES, STAGE_BYTES = SH::STAGE_BYTES, TILE_BYTES = SH::A_BYTES;
  static_assert(TM % 4 == 0 && TN % 4 == 0 && (A_MN || TM == 8) && (B_MN || TN == 8), "K-major operands are read 8 rows per thread");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + STAGES * STAGE_BYTES, bar_empty = bar_full + STAGES * 8;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN, ntiles = tiles_m * tiles_n;
  const int KT = (k + BK - 1) / BK;
  if (tid == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(bar_full + i * 8, 1); mbar_init(bar_empty + i * 8, WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  __syncthreads();

  auto coords = [&](int tile, int &tm, int &tn) {
    const int per_band = GROUP_M * tiles_n, band = tile / per_band, in_band = tile % per_band;
    const int first_m = band * GROUP_M, gsz = min(tiles_m - first_m, GROUP_M);
    tm = first_m + in_band % gsz; tn = in_band / gsz;
  };
  auto skipped = [&](int tm, int tn) {
    if (tri == 'L') return tm * BM + BM - 1 < tn * BN;
    if (tri == 'U') return tn * BN + BN - 1 < tm * BM;
    return false;
  };
  // the loads of this CTA as one sequence over (tile, kt); `issued` counts how many have been handed to TMA
  int ld_tile = blockIdx.x, ld_kt = 0, ld_stage = 0; uint32_t ld_phase = 0;
  auto advance_tile = [&]() { while (ld_tile < ntiles) { int tm, tn; coords(ld_tile, tm, tn); if (!skipped(tm, tn)) break; ld_tile += gridDim.x; } };
  auto issue_one = [&]() {          // thread 0 only
    if (ld_tile >= ntiles) return;
    int tm, tn; coords(ld_tile, tm, tn);
    mbar_wait(bar_empty + ld_stage * 8, ld_phase ^ 1);
    const uint32_t full = bar_full + ld_stage * 8, sa = base + ld_stage * STAGE_BYTES, sb = sa + TILE_BYTES;
    mbar_expect_tx(full, STAGE_BYTES);
    if (A_MN) tma_load_2d(sa, &mapA, tm * BM, ld_kt * BK, full); else tma_load_2d(sa, &mapA, ld_kt * BK, tm * BM, full);
    if (B_MN) tma_load_2d(sb, &mapB, tn * BN, ld_kt * BK, full); else tma_load_2d(sb, &mapB, ld_kt * BK, tn * BN, full);
    if (++ld_stage == STAGES) { ld_stage = 0; ld_phase ^= 1; }
    if (++ld_kt == KT) { ld_kt = 0; ld_tile += gridDim.x; advance_tile(); }
  };
  if (tid == 0) {
    advance_tile();
    for (int i = 0; i < STAGES - 1; i++) issue_one();
  }

  int stage = 0; uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int tm, tn; coords(tile, tm, tn);
    if (skipped(tm, tn)) continue;
    // accumulators as packed pairs: along the rows (acc2[ip][j] = rows 2ip, 2ip+1 of column j) unless only B is MN-major
    constexpr bool PACK_I = A_MN || !B_MN;
    u64 acc2[TM * TN / 2];
#pragma unroll
    for (int i = 0; i < TM * TN / 2; i++) acc2[i] = 0ull;

    for (int kt = 0; kt < KT; kt++) {
      if (tid == 0) issue_one();                       // keeps STAGES-1 loads in flight
      mbar_wait(bar_full + stage * 8, phase);
      const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + TILE_BYTES;
#pragma unroll
      for (int kb = 0; kb < BK / 4; kb++) {
        // K-major operands: 4 consecutive k of each of the thread's 8 rows / columns, held for the whole 4-k block
        float4 aq[8], bq[8];
        if (!A_MN) {
#pragma unroll
          for (int e = 0; e < 8; e++) { const int r = tx + 16 * e; aq[e] = lds128(sa + r * 128 + (((kb ^ r) & 7) << 4)); }
        }
        if (!B_MN) {
#pragma unroll
          for (int e = 0; e < 8; e++) { const int c = ty + 16 * e; bq[e] = lds128(sb + c * 128 + (((kb ^ c) & 7) << 4)); }
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          float a[TM], b[TN];
          if (A_MN) {
            const uint32_t row = sa + ((kb * 4 + kk) * BM + tx * 4) * 4;
#pragma unroll
            for (int g = 0; g < TM / 4; g++) {
              const float4 v = lds128(row + g * 256);
              a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) a[e] = kk == 0 ? aq[e].x : (kk == 1 ? aq[e].y : (kk == 2 ? aq[e].z : aq[e].w));
          }
          if (B_MN) {
            const uint32_t row = sb + ((kb * 4 + kk) * BN + ty * 4) * 4;
#pragma unroll
            for (int g = 0; g < TN / 4; g++) {
              const float4 v = lds128(row + g * 256);
              b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) b[e] = kk == 0 ? bq[e].x : (kk == 1 ? bq[e].y : (kk == 2 ? bq[e].z : bq[e].w));
          }
          if (PACK_I) {
            u64 ap[TM / 2], bd[TN];
#pragma unroll
            for (int p = 0; p < TM / 2; p++) ap[p] = pack2(a[2 * p], a[2 * p + 1]);
#pragma unroll
            for (int j = 0; j < TN; j++) bd[j] = pack2(b[j], b[j]);
#pragma unroll
            for (int p = 0; p < TM / 2; p++)
#pragma unroll
              for (int j = 0; j < TN; j++) acc2[p * TN + j] = ffma2(ap[p], bd[j], acc2[p * TN + j]);
          } else {
            u64 ad[TM], bp[TN / 2];
#pragma unroll
            for (int i = 0; i < TM; i++) ad[i] = pack2(a[i], a[i]);
#pragma unroll
            for (int q = 0; q < TN / 2; q++) bp[q] = pack2(b[2 * q], b[2 * q + 1]);
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
              for (int q = 0; q < TN / 2; q++) acc2[i * (TN / 2) + q] = ffma2(ad[i], bp[q], acc2[i * (TN / 2) + q]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + stage * 8);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // epilogue: C = alpha*acc + beta*C, one column of the micro-tile at a time
    const int row0 = tm * BM, col0 = tn * BN;
    const bool beta0 = (beta == 0.f);
    const bool vec_ok = A_MN && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && tri == 0;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      float acc[TM];
#pragma unroll
      for (int i = 0; i < TM; i++) {
        float lo, hi;
        if (PACK_I) { unpack2(acc2[(i >> 1) * TN + j], lo, hi); acc[i] = (i & 1) ? hi : lo; }
        else { unpack2(acc2[i * (TN / 2) + (j >> 1)], lo, hi); acc[i] = (j & 1) ? hi : lo; }
      }
      const int col = col0 + elem_idx<B_MN>(ty, j);
      if (col >= n) continue;
      float *ccol = C + (size_t)col * ldc;
      if (vec_ok) {
#pragma unroll
        for (int hh = 0; hh < TM / 4; hh++) {
          const int r0 = row0 + hh * 64 + tx * 4;
          if (r0 >= m) continue;
          float v[4] = {alpha * acc[hh * 4 + 0], alpha * acc[hh * 4 + 1], alpha * acc[hh * 4 + 2], alpha * acc[hh * 4 + 3]};
          if (r0 + 3 < m) {
            if (!beta0) {
              const float4 c0 = *reinterpret_cast<const float4 *>(ccol + r0);
              v[0] = fmaf(beta, c0.x, v[0]); v[1] = fmaf(beta, c0.y, v[1]); v[2] = fmaf(beta, c0.z, v[2]); v[3] = fmaf(beta, c0.w, v[3]);
            }
            *reinterpret_cast<float4 *>(ccol + r0) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (r0 + e < m) ccol[r0 + e] = beta0 ? v[e] : fmaf(beta, ccol[r0 + e], v[e]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; i++) {
          const int row = row0 + elem_idx<A_MN>(tx, i);
          if (row >= m) continue;
          if (tri == 'L' && row < col) continue;
          if (tri == 'U' && row > col) continue;
          const float v = alpha * acc[i];
          ccol[row] = beta0 ? v : fmaf(beta, ccol[row], v);
        }
      }
    }
  }
}

// dim0 = contiguous dimension.  MN-major tile: box [128 idx][32 k], no swizzle; K-major tile: box [32 k][128 idx], 128-byte swizzle.
static int make_map_f32(CUtensorMap *map, const float *ptr, long dim0, long dim1, int ld, int box0, int box1, bool swizzle) {
  dmma::EncodeTiledFn enc = dmma::get_encode();
  if (!enc) { set_error(JB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable"); return JB_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(JB_ERR_CUDA, "cuTensorMapEncodeTiled (f32) failed (%d)", (int)r); return JB_ERR_CUDA; }
  return 0;
}

template <bool A_MN, bool B_MN, int TM, int TN>
static int launch(const CUtensorMap &ma, const CUtensorMap &mb, int m, int n, int k, float alpha, float beta, float *C, int ldc, int tri, cudaStream_t s) {
  using SH = Shape<TM, TN>;
  JB_ONCE_PER_DEVICE(JB_CUDA(cudaFuncSetAttribute(sgemm_tma_kernel<A_MN, B_MN, TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH::SMEM_BYTES)));
  const int ntiles = ((m + SH::BM - 1) / SH::BM) * ((n + SH::BN - 1) / SH::BN);
  int grid = std::min(ntiles, SH::MINB * num_sms());
  const int cap = option("gemm.max_ctas");
  if (cap > 0 && SH::MINB * cap < grid) grid = SH::MINB * cap;
  sgemm_tma_kernel<A_MN, B_MN, TM, TN><<<grid, THREADS, SH::SMEM_BYTES, s>>>(ma, mb, m, n, k, alpha, beta, C, ldc, tri);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 1;
}
}  // namespace sgemm_tma

int sgemm_tma_try(int ta, int tb, int m, int n, int k, float alpha, const float *A, int lda, const float *B, int ldb, float beta, float *C,
                  int ldc, cudaStream_t s, int tri) {
  using namespace sgemm_tma;
  if (m < 128 || n < 128 || k < 32) return 0;
  if ((long)((m + 127) / 128) * ((n + 127) / 128) < 2 * 148) return 0;       // too few tiles to fill two CTAs per SM
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 3) || (ldb & 3)) return 0;
  if (option("sgemm.simt") > 0) return 0;
  const bool a_mn = (ta == 111), b_mn = (tb != 111);
  // the 256 x 128 tile (one CTA per SM) from 8 tiles per SM on, when A is MN-major (op(A) = A: what Jalla's ## passes)
  const int want = option("sgemm.tile");       // 128 / 256 force a shape (tests, A/B timing); unset = by size
  const bool big = a_mn && m >= 256 && want != 128 && (want == 256 || (long)((m + 255) / 256) * ((n + 127) / 128) >= 8 * 148);
  const int bm = big ? 256 : 128;
  CUtensorMap ma, mb;
  int rc;
  // op(A) is m x k: NoTrans = stored m x k, m contiguous (MN-major); Trans = stored k x m, k contiguous (K-major)
  if (a_mn) rc = make_map_f32(&ma, A, m, k, lda, bm, BK, false); else rc = make_map_f32(&ma, A, k, m, lda, BK, 128, true);
  if (rc) return rc;
  // op(B) is k x n: NoTrans = stored k x n, k contiguous (K-major); Trans = stored n x k, n contiguous (MN-major)
  if (b_mn) rc = make_map_f32(&mb, B, n, k, ldb, 128, BK, false); else rc = make_map_f32(&mb, B, k, n, ldb, BK, 128, true);
  if (rc) return rc;
  if (big && b_mn) return launch<true, true, 16, 8>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (big) return launch<true, false, 16, 8>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (a_mn && b_mn) return launch<true, true, 8, 8>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (a_mn && !b_mn) return launch<true, false, 8, 8>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (!a_mn && b_mn) return launch<false, true, 8, 8>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  return launch<false, false, 8, 8>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
}
}  // namespace jb
