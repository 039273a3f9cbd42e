// Complex-float GEMM on the CUDA cores, four real multiplies per complex MAC ("4M", never 3M: SURVEY.md §7.3 H6), operands
// staged by TMA — the kernel behind cblas_cgemm (/root/reference/Numeric/Jalla/Foreign/BLAS.chs:161, reached from
// Matrix.hs:482 via BlasOps.hs:156-166) for 16-byte aligned operands with even leading dimensions.
//
// Same skeleton as the FP32 kernel (gemm_f32_tma.cu): persistent 128 x 128 (complex) tile loop with band rasterisation, one
// elected thread feeding a 5-stage TMA -> mbarrier ring (BK = 16 complex), 256 threads with an 8 x 8 micro-tile each, one
// CTA per SM.  Interleaved (re, im) storage is what makes the packed FP32 FMA fit: an accumulator IS a register pair
//     (cr, ci) += (br, bi) * ar          fma.rn.f32x2 with a broadcast scalar (SASS FFMA2 ... R.F32)
//     (cr, ci) += (-bi, br) * ai
// — exactly fmaf(-ai, bi, fmaf(ar, br, cr)) and fmaf(ai, br, fmaf(ar, bi, ci)), the sequence of Num<cfloat>::fma — so a
// complex MAC is two issue slots, the (br, bi) pair comes straight out of shared memory and only (-bi, br) costs two ALU
// operations per B element and k (16 per 128 FFMA2).  Conjugation is a sign on ai (op(A)) or on bi (op(B)).
// Nothing is transposed on the way in; the fragment loads adapt to the operand's own major as in the FP32 kernel:
//   MN-major tile [16 k][128 idx] (no swizzle):       one LDS.128 = two consecutive rows/columns at one k
//   K-major  tile [128 idx][16 k] (128-byte swizzle): one LDS.128 = two consecutive k of one row/column
// Bound: FP32 FMA pipe (8 real flops per complex MAC = 4 FMAs = 2 FFMA2), measured peak 72.35 TFLOP/s (tools/peaks.cu).
#include "dmma_common.cuh"

namespace jb {
namespace cgemm_tma {
using dmma::smem_u32; using dmma::mbar_init; using dmma::mbar_expect_tx; using dmma::mbar_arrive; using dmma::mbar_wait; using dmma::tma_load_2d;

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 5, THREADS = 256, WARPS = THREADS / 32;
constexpr int TILE_BYTES = BM * BK * 8, STAGE_BYTES = 2 * TILE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 64;
constexpr int GROUP_M = 16;

typedef unsigned long long u64;
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fneg(float x) { return __int_as_float(__float_as_int(x) ^ (int)0x80000000); }   // LOP3: ALU pipe, not the FMA pipe

// index (row of op(A) / column of op(B) inside the tile) of a thread's e-th element, e = 0..7; t = its position 0..15
template <bool MN> __device__ __forceinline__ int elem_idx(int t, int e) { return MN ? (e >> 1) * 32 + t * 2 + (e & 1) : t + 16 * e; }

// CA / CB: conjugate op(A) / op(B)
template <bool A_MN, bool B_MN, bool CA, bool CB>
__global__ void __launch_bounds__(THREADS, 1)
cgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int m, int n, int k, cfloat alpha,
                 cfloat beta, cfloat *__restrict__ C, int ldc, int tri) {
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
  // the loads of this CTA as one sequence over (tile, kt)
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
    u64 acc[8][8];                  // acc[i][j] = (re, im) of C(row i, column j) of the micro-tile
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = 0ull;

    for (int kt = 0; kt < KT; kt++) {
      if (tid == 0) issue_one();                       // keeps STAGES-1 loads in flight
      mbar_wait(bar_full + stage * 8, phase);
      const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + TILE_BYTES;
#pragma unroll
      for (int kp = 0; kp < BK / 2; kp++) {
        // K-major operands: two consecutive k of each of the thread's 8 rows / columns
        float4 aq[8], bq[8];
        if (!A_MN) {
#pragma unroll
          for (int e = 0; e < 8; e++) { const int r = tx + 16 * e; aq[e] = lds128(sa + r * 128 + (((kp ^ r) & 7) << 4)); }
        }
        if (!B_MN) {
#pragma unroll
          for (int e = 0; e < 8; e++) { const int c = ty + 16 * e; bq[e] = lds128(sb + c * 128 + (((kp ^ c) & 7) << 4)); }
        }
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          float ar[8], ai[8], br[8], bi[8];
          if (A_MN) {
            const uint32_t row = sa + ((kp * 2 + kk) * BM + tx * 2) * 8;
#pragma unroll
            for (int g = 0; g < 4; g++) {
              const float4 v = lds128(row + g * 256);
              ar[2 * g] = v.x; ai[2 * g] = v.y; ar[2 * g + 1] = v.z; ai[2 * g + 1] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) { ar[e] = kk == 0 ? aq[e].x : aq[e].z; ai[e] = kk == 0 ? aq[e].y : aq[e].w; }
          }
          if (B_MN) {
            const uint32_t row = sb + ((kp * 2 + kk) * BN + ty * 2) * 8;
#pragma unroll
            for (int g = 0; g < 4; g++) {
              const float4 v = lds128(row + g * 256);
              br[2 * g] = v.x; bi[2 * g] = v.y; br[2 * g + 1] = v.z; bi[2 * g + 1] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) { br[e] = kk == 0 ? bq[e].x : bq[e].z; bi[e] = kk == 0 ? bq[e].y : bq[e].w; }
          }
          u64 p1[8], p2[8], sr[8], si[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            p1[j] = CB ? pack2(br[j], fneg(bi[j])) : pack2(br[j], bi[j]);            // op(B) element (re, im)
            p2[j] = CB ? pack2(bi[j], br[j]) : pack2(fneg(bi[j]), br[j]);            // i * op(B) element
          }
#pragma unroll
          for (int i = 0; i < 8; i++) {
            sr[i] = pack2(ar[i], ar[i]);
            const float s = CA ? fneg(ai[i]) : ai[i];
            si[i] = pack2(s, s);
          }
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = ffma2(p1[j], sr[i], acc[i][j]);
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = ffma2(p2[j], si[i], acc[i][j]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + stage * 8);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // epilogue: C = alpha*acc + beta*C (complex), one column of the micro-tile at a time
    const int row0 = tm * BM, col0 = tn * BN;
    const bool beta0 = (beta.x == 0.f && beta.y == 0.f);
    const bool vec_ok = A_MN && ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && tri == 0;
    auto axpby = [&](u64 a, cfloat c0) {
      float re, im; unpack2(a, re, im);
      cfloat v = {fmaf(-alpha.y, im, alpha.x * re), fmaf(alpha.y, re, alpha.x * im)};
      if (!beta0) { v.x += fmaf(-beta.y, c0.y, beta.x * c0.x); v.y += fmaf(beta.y, c0.x, beta.x * c0.y); }
      return v;
    };
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int col = col0 + elem_idx<B_MN>(ty, j);
      if (col >= n) continue;
      cfloat *ccol = C + (size_t)col * ldc;
      if (vec_ok) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const int r0 = row0 + g * 32 + tx * 2;
          if (r0 >= m) continue;
          if (r0 + 1 < m) {
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!beta0) c0 = *reinterpret_cast<const float4 *>(ccol + r0);
            const cfloat v0 = axpby(acc[2 * g][j], cfloat{c0.x, c0.y}), v1 = axpby(acc[2 * g + 1][j], cfloat{c0.z, c0.w});
            *reinterpret_cast<float4 *>(ccol + r0) = make_float4(v0.x, v0.y, v1.x, v1.y);
          } else {
            ccol[r0] = axpby(acc[2 * g][j], beta0 ? cfloat{0.f, 0.f} : ccol[r0]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int row = row0 + elem_idx<A_MN>(tx, i);
          if (row >= m) continue;
          if (tri == 'L' && row < col) continue;
          if (tri == 'U' && row > col) continue;
          ccol[row] = axpby(acc[i][j], beta0 ? cfloat{0.f, 0.f} : ccol[row]);
        }
      }
    }
  }
}

// Complex elements travel as opaque 8-byte words.  dim0 = contiguous dimension.  MN-major tile: box [128 idx][16 k], no
// swizzle; K-major tile: box [16 k][128 idx], 128-byte swizzle.
static int make_map_c32(CUtensorMap *map, const cfloat *ptr, long dim0, long dim1, int ld, int box0, int box1, bool swizzle) {
  dmma::EncodeTiledFn enc = dmma::get_encode();
  if (!enc) { set_error(JB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable"); return JB_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(cfloat)};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<cfloat *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(JB_ERR_CUDA, "cuTensorMapEncodeTiled (c32) failed (%d)", (int)r); return JB_ERR_CUDA; }
  return 0;
}

template <bool A_MN, bool B_MN, bool CA, bool CB>
static int launch(const CUtensorMap &ma, const CUtensorMap &mb, int m, int n, int k, cfloat alpha, cfloat beta, cfloat *C, int ldc, int tri, cudaStream_t s) {
  JB_ONCE_PER_DEVICE(JB_CUDA(cudaFuncSetAttribute(cgemm_tma_kernel<A_MN, B_MN, CA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  const int ntiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  int grid = std::min(ntiles, num_sms());
  const int cap = option("gemm.max_ctas");
  if (cap > 0 && cap < grid) grid = cap;
  cgemm_tma_kernel<A_MN, B_MN, CA, CB><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, m, n, k, alpha, beta, C, ldc, tri);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 1;
}
template <bool A_MN, bool B_MN>
static int launch_conj(bool ca, bool cb, const CUtensorMap &ma, const CUtensorMap &mb, int m, int n, int k, cfloat alpha, cfloat beta, cfloat *C, int ldc,
                       int tri, cudaStream_t s) {
  // a conjugated operand is always a transposed one (trans 113): op(A) conjugated => A is K-major, op(B) conjugated => B is MN-major
  constexpr bool CAN_A = !A_MN, CAN_B = B_MN;
  if constexpr (CAN_A && CAN_B) {
    if (ca && cb) return launch<A_MN, B_MN, true, true>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  }
  if constexpr (CAN_A) {
    if (ca && !cb) return launch<A_MN, B_MN, true, false>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  }
  if constexpr (CAN_B) {
    if (cb && !ca) return launch<A_MN, B_MN, false, true>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  }
  return launch<A_MN, B_MN, false, false>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
}
}  // namespace cgemm_tma

int cgemm_tma_try(int ta, int tb, int m, int n, int k, cfloat alpha, const cfloat *A, int lda, const cfloat *B, int ldb, cfloat beta, cfloat *C,
                  int ldc, cudaStream_t s, int tri) {
  using namespace cgemm_tma;
  if (m < 128 || n < 128 || k < 16) return 0;
  if ((long)((m + BM - 1) / BM) * ((n + BN - 1) / BN) < 120 && option("gemm.force_dmma") <= 0) return 0;   // too few tiles: finer-grained generic kernel
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return 0;                       // TMA: 16-byte aligned base and pitch
  if (option("cgemm.generic") > 0) return 0;
  const bool a_mn = (ta == 111), b_mn = (tb != 111);
  CUtensorMap ma, mb;
  int rc;
  // op(A) is m x k: NoTrans = stored m x k, m contiguous (MN-major); (Conj)Trans = stored k x m, k contiguous (K-major)
  if (a_mn) rc = make_map_c32(&ma, A, m, k, lda, BM, BK, false); else rc = make_map_c32(&ma, A, k, m, lda, BK, BM, true);
  if (rc) return rc;
  // op(B) is k x n: NoTrans = stored k x n, k contiguous (K-major); (Conj)Trans = stored n x k, n contiguous (MN-major)
  if (b_mn) rc = make_map_c32(&mb, B, n, k, ldb, BN, BK, false); else rc = make_map_c32(&mb, B, k, n, ldb, BK, BN, true);
  if (rc) return rc;
  const bool ca = (ta == 113), cb = (tb == 113);
  if (a_mn && b_mn) return launch_conj<true, true>(ca, cb, ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (a_mn && !b_mn) return launch_conj<true, false>(ca, cb, ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (!a_mn && b_mn) return launch_conj<false, true>(ca, cb, ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  return launch_conj<false, false>(ca, cb, ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
}
}  // namespace jb
