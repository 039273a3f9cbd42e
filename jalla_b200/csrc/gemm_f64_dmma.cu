// FP64 GEMM on the DMMA tensor pipe for sm_100a — the kernel behind cblas_dgemm
// (/root/reference/Numeric/Jalla/Foreign/BLAS.chs:155, reached from Matrix.hs:482) and the
// trailing updates of the blocked LU / Cholesky.
//
// There is no FP64 kind of tcgen05.mma (kinds are f16/tf32/f8f6f4/i8/mxf*), so FP64 tensor
// work is the warp-level mma.sync.m8n8k4.f64, which sm_100a executes natively as
// DMMA.8x8x4 (16 issue cycles per SMSP; measured 37.07 TFLOP/s chip peak, tools/peaks.cu).
//
// Structure: persistent CTAs (one per SM), 128x128 C tile, BK=16, a TMA producer warp
// feeding a 5-stage shared-memory ring through mbarriers (its warpgroup releases registers
// with setmaxnreg), and 8 consumer warps (2x4),
// each owning a 64x32 sub-tile as 8x4 DMMA accumulators (128 registers/thread).
// Operands arrive with 128-byte TMA swizzle; fragment rows/columns are *permuted* per lane
// (rho / sigma below) so that every 64-bit fragment load is bank-conflict free for both
// operand majors — the permutation is undone for free when C is written.  All four
// NoTrans/Trans combinations are served without materialising a transpose: an operand is
// either K-major (one 16x128 box per stage) or MN-major (eight 16x16 boxes per stage).
#include "dmma_common.cuh"

namespace jb {
namespace dmma {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 5;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;     // 2 consumer warpgroups + 1 producer warpgroup
constexpr int TILE_BYTES = BM * BK * 8;                 // 16 KiB per operand per stage
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int GROUP_M = 16;                             // tile rasterisation band (L2 reuse)

// Per-lane fragment geometry.  g = lane>>2 selects the row (A side) / column (B side) of an
// 8-wide group, t = lane&3 the k index inside a k4 step.
//  K-major tile  [idx][k]: 128-byte rows, element (r,kap) at r*128 + (((kap>>1)^(r&7))<<4) + (kap&1)*8
//  MN-major tile [box][k][idx16]: box b at b*2048, element (i,kap) at kap*128 + (((i>>1)^(kap&7))<<4) + (i&1)*8
template <bool KM> __device__ __forceinline__ int frag_idx(int j, int g) {
  if (KM) return 8 * j + 2 * (g & 3) + (g >> 2);
  return 16 * (j >> 1) + ((g >> 1) & 1) * 8 + (g & 1) + 2 * (2 * (j & 1) + (g >> 2));
}
template <bool KM> __device__ __forceinline__ uint32_t frag_off(int j, int s, int g, int t) {
  const int kap = 4 * s + t;
  if (KM) {
    const int rho = 2 * (g & 3) + (g >> 2);
    return (uint32_t)((8 * j + rho) * 128 + ((((kap >> 1) ^ rho) & 7) << 4) + (kap & 1) * 8);
  }
  const int i15 = ((g >> 1) & 1) * 8 + (g & 1) + 2 * (2 * (j & 1) + (g >> 2));
  return (uint32_t)((j >> 1) * 2048 + kap * 128 + ((((i15 >> 1) ^ (kap & 7)) & 7) << 4) + (i15 & 1) * 8);
}

static_assert(BM == BN, "TileWalker's triangle test assumes square tiles");

template <bool A_KM, bool B_KM>
__global__ void __launch_bounds__(THREADS, 1)
dgemm_dmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int m, int n, int k,
                  double alpha, double beta, double *__restrict__ C, int ldc, int tri, int balanced) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // 128B swizzle atoms need 1024B alignment
  const uint32_t bar_full = base + STAGES * STAGE_BYTES;            // STAGES x 8 bytes
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const int KT = (k + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(bar_full + i * 8, 1); mbar_init(bar_empty + i * 8, CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp >= CONSUMER_WARPS) {
    // ------------------------------------------------------------- TMA producer (one lane)
    // registers are allocated per warpgroup: the producer group hands its share to the consumers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == CONSUMER_WARPS && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
      int stage = 0; uint32_t phase = 0;
      // beta != 0: every CTA of a wave reaches its epilogue at the same time and the 148 x 128 KB burst of C reads
      // ran at ~3 TB/s (13 us per tile, 28 % of a K = 256 update).  The producer lane spreads an L2 prefetch of the
      // tile's C columns over the main loop instead, so the epilogue reads hit L2.
      const bool pf = beta != 0.0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && ((ldc & 1) == 0);
      const int pf_step = (BN + KT - 1) / KT;
      TileWalker<GROUP_M> walk(tiles_m, tiles_n, tri, balanced);
      int tm, tn;
      for (int tile = blockIdx.x, st; (st = walk.at(tile, tm, tn)) != 0; tile += gridDim.x) {
        if (st == 2) continue;
        int pf_col = 0;
        const int pf_cols = min(BN, n - tn * BN);
        const uint32_t pf_bytes = (uint32_t)(min(BM, m - tm * BM) * 8) & ~15u;
        const double *pf_base = C + (size_t)tn * BN * ldc + (size_t)tm * BM;
        for (int kt = 0; kt < KT; kt++) {
          mbar_wait(bar_empty + stage * 8, phase ^ 1);
          const uint32_t full = bar_full + stage * 8;
          mbar_expect_tx(full, STAGE_BYTES);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + TILE_BYTES;
          if (A_KM) tma_load_2d(sa, &mapA, kt * BK, tm * BM, full);
          else {
#pragma unroll
            for (int b = 0; b < BM / 16; b++) tma_load_2d(sa + b * 2048, &mapA, tm * BM + b * 16, kt * BK, full);
          }
          if (B_KM) tma_load_2d(sb, &mapB, kt * BK, tn * BN, full);
          else {
#pragma unroll
            for (int b = 0; b < BN / 16; b++) tma_load_2d(sb + b * 2048, &mapB, tn * BN + b * 16, kt * BK, full);
          }
          if (pf && pf_bytes)
            for (int q = 0; q < pf_step && pf_col < pf_cols; q++, pf_col++) prefetch_l2_bulk(pf_base + (size_t)pf_col * ldc, pf_bytes);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // --------------------------------------------------------------- DMMA consumers
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;              // 2 x 4 warps, warp tile 64 x 32
  // byte offsets of this lane's fragments inside a stage, per k4 step
  uint32_t offA[4], offB[4];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    offA[s] = frag_off<A_KM>(wm * 8, s, g, t);
    offB[s] = TILE_BYTES + frag_off<B_KM>(wn * 4, s, g, t);
  }
  // group-to-group strides are lane independent: K-major 1024 B per 8 rows, MN-major 2048 B per
  // 16 rows with the odd group of a box 32 B further along the row (sigma_1 = sigma_0 + 4 elements,
  // which flips chunk bit 1: +/-32 bytes depending on the lane's swizzle — handled via xor below)
  int stage = 0; uint32_t phase = 0;
  const bool beta0 = (beta == 0.0);

  TileWalker<GROUP_M> walk(tiles_m, tiles_n, tri, balanced);
  int tm, tn;
  for (int tile = blockIdx.x, st; (st = walk.at(tile, tm, tn)) != 0; tile += gridDim.x) {
    if (st == 2) continue;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    double fa[2][8], fb[2][4];
    auto load_frags = [&](int buf, uint32_t sbase, int s) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint32_t o;
        if (A_KM) o = offA[s] + i * 1024;
        else o = (offA[s] ^ ((i & 1) << 5)) + (i >> 1) * 2048;
        fa[buf][i] = lds64(sbase + o);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint32_t o;
        if (B_KM) o = offB[s] + j * 1024;
        else o = (offB[s] ^ ((j & 1) << 5)) + (j >> 1) * 2048;
        fb[buf][j] = lds64(sbase + o);
      }
    };

    mbar_wait(bar_full + stage * 8, phase);
    load_frags(0, base + stage * STAGE_BYTES, 0);
    for (int kt = 0; kt < KT; kt++) {
      const uint32_t sbase = base + stage * STAGE_BYTES;
      int nstage = stage + 1; uint32_t nphase = phase;
      if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int cur = s & 1, nxt = cur ^ 1;
        if (s < 3) load_frags(nxt, sbase, s + 1);
        else if (kt + 1 < KT) {
          mbar_wait(bar_full + nstage * 8, nphase);
          load_frags(nxt, base + nstage * STAGE_BYTES, 0);
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], fa[cur][i], fb[cur][j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + stage * 8);
      stage = nstage; phase = nphase;
    }

    // ------------------------------------------------------------- epilogue: C = alpha*acc + beta*C
    const int row0 = tm * BM + wm * 64, col0 = tn * BN + wn * 32;
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int col = col0 + frag_idx<B_KM>(j, 2 * t + h) ;
        if (col >= n) continue;
        double *ccol = C + (size_t)col * ldc;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int row = row0 + frag_idx<A_KM>(i, g);
          if (row >= m) continue;
          if (tri == 'L' && row < col) continue;
          if (tri == 'U' && row > col) continue;
          double r = alpha * acc[i][j][h];
          if (!beta0) r = fma(beta, ccol[row], r);
          ccol[row] = r;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ host side
// Tensor map of a column-major matrix `rows x cols` (ld) seen as dim0 = rows (contiguous), dim1 = cols.
static int make_map(CUtensorMap *map, const double *ptr, long rows, long cols, int ld, int box0, int box1) {
  return make_map_f64(map, ptr, rows, cols, (size_t)ld * sizeof(double), box0, box1);
}

template <bool A_KM, bool B_KM>
static int launch(const CUtensorMap &ma, const CUtensorMap &mb, int m, int n, int k, double alpha, double beta, double *C,
                  int ldc, int tri, cudaStream_t s) {
  JB_ONCE_PER_DEVICE(JB_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<A_KM, B_KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  const int ntiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  int grid = ntiles < num_sms() ? ntiles : num_sms();
  // "gemm.max_ctas": a CTA of this persistent kernel owns a whole SM (registers + shared memory), so callers that
  // overlap it with other kernels — the NCCL panel broadcasts of the sharded GEMM — leave a few SMs free
  const int cap = option("gemm.max_ctas");
  if (cap > 0 && cap < grid) grid = cap;
  const int tcap = tls().gemm_cta_cap;          // set by the look-ahead factorisations around their trailing update
  if (tcap > 0 && tcap < grid) grid = tcap;
  dgemm_dmma_kernel<A_KM, B_KM><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, option("gemm.tri_balance") != 0);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace dmma

int dgemm_dmma_try(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
                   double beta, double *C, int ldc, cudaStream_t s, int tri) {
  using namespace dmma;
  if (m < 64 || n < 64 || k < 1) return 0;
  // a 128x128 tile keeps one SM busy for 2*128*128*k flop: below ~64 tiles the generic kernel's smaller tiles finish
  // sooner because they spread over all SMs (512^3: 74 us here vs ~20 us there)
  if ((long)((m + BM - 1) / BM) * ((n + BN - 1) / BN) < 64 && option("gemm.force_dmma") <= 0) return 0;
  // TMA needs 16-byte aligned bases and row pitches
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return 0;
  const bool a_km = (ta != 111), b_km = (tb == 111);
  CUtensorMap ma, mb;
  int rc;
  // op(A) is m x k.  Trans: stored k x m (k contiguous) -> K-major; NoTrans: stored m x k -> MN-major.
  if (a_km) rc = make_map(&ma, A, k, m, lda, BK, BM); else rc = make_map(&ma, A, m, k, lda, 16, BK);
  if (rc) return rc;
  if (b_km) rc = make_map(&mb, B, k, n, ldb, BK, BN); else rc = make_map(&mb, B, n, k, ldb, 16, BK);
  if (rc) return rc;
  if (a_km && b_km) return launch<true, true>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (a_km && !b_km) return launch<true, false>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  if (!a_km && b_km) return launch<false, true>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
  return launch<false, false>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, s);
}

}  // namespace jb
