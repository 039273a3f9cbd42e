// Complex-double GEMM on the FP64 DMMA tensor pipe (sm_100a) — the kernel behind cblas_zgemm
// (/root/reference/Numeric/Jalla/Foreign/BLAS.chs:167, reached from Matrix.hs:482 via BlasOps.hs:180).
//
// Four-real-multiply scheme ("4M"; a 3M scheme fails the reference's own prop_matrixMultDiag pins,
// SURVEY.md §4): with interleaved (re,im) storage one 128-bit fragment load yields (ar,ai) of a
// complex element, and each complex MAC block is four DMMA.8x8x4:
//     Cr += ar*br ; Cr += (-ai)*bi ; Ci += ar*bi ; Ci += ai*br
// Same skeleton as the real kernel (gemm_f64_dmma.cu): persistent CTAs, TMA producer warpgroup,
// mbarrier ring, 8 consumer warps; here the C tile is 64x128 complex, BK = 8 complex (128-byte rows),
// warp tile 32x32 complex = 4x4 DMMA tiles x {Cr,Ci} = 128 accumulator registers.  Fragment rows are
// permuted (rho) so the 128-bit loads are bank-conflict free under the 128B TMA swizzle for both
// operand majors; NoTrans / Trans / ConjTrans are all served in place (conjugation = sign of the
// imaginary fragment).  Complex data is always 16-byte aligned, so every leading dimension qualifies.
#include "dmma_common.cuh"

namespace jb {
namespace zdmma {
using namespace dmma;

constexpr int BM = 64, BN = 128, BK = 8;        // complex elements
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
constexpr int A_BYTES = BM * BK * 16, B_BYTES = BN * BK * 16;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 24 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int GROUP_M = 16;

__device__ __forceinline__ void lds128(uint32_t addr, double &x, double &y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ int rho_of(int g) { return (g >> 1) | ((g & 1) << 2); }
// byte offset of complex element (group j, lane row rho, kappa) inside an operand tile
template <bool KM> __device__ __forceinline__ uint32_t frag_off(int j, int kap, int rho) {
  if (KM) return (uint32_t)((8 * j + rho) * 128 + ((kap ^ rho) << 4));
  return (uint32_t)(j * 1024 + kap * 128 + ((rho ^ kap) << 4));
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int &tm, int &tn) {
  const int per_band = GROUP_M * tiles_n;
  const int band = tile / per_band, in_band = tile % per_band;
  const int first_m = band * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  tm = first_m + in_band % gsz;
  tn = in_band / gsz;
}
__device__ __forceinline__ bool tile_skipped(int tri, int tm, int tn) {
  if (tri == 'L') return tm * BM + BM - 1 < tn * BN;
  if (tri == 'U') return tn * BN + BN - 1 < tm * BM;
  return false;
}

template <bool A_KM, bool B_KM>
__global__ void __launch_bounds__(THREADS, 1)
zgemm_dmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int m, int n, int k,
                  cdouble alpha, cdouble beta, cdouble *__restrict__ C, int ldc, int tri, int conj_a, int conj_b) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + STAGES * STAGE_BYTES;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const int ntiles = tiles_m * tiles_n;
  const int KT = (k + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(bar_full + i * 8, 1); mbar_init(bar_empty + i * 8, CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp >= CONSUMER_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == CONSUMER_WARPS && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
      int stage = 0; uint32_t phase = 0;
      // beta != 0: L2 prefetch of the tile's C columns spread over the main loop (see gemm_f64_dmma.cu)
      const bool pf = !(beta.x == 0.0 && beta.y == 0.0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
      const int pf_step = (BN + KT - 1) / KT;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int tm, tn; tile_coords(tile, tiles_m, tiles_n, tm, tn);
        if (tile_skipped(tri, tm, tn)) continue;
        int pf_col = 0;
        const int pf_cols = min(BN, n - tn * BN);
        const uint32_t pf_bytes = (uint32_t)(min(BM, m - tm * BM) * 16);
        const cdouble *pf_base = C + (size_t)tn * BN * ldc + (size_t)tm * BM;
        for (int kt = 0; kt < KT; kt++) {
          mbar_wait(bar_empty + stage * 8, phase ^ 1);
          const uint32_t full = bar_full + stage * 8;
          mbar_expect_tx(full, STAGE_BYTES);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          // coordinates are in FP64 words: a complex index c is word 2c
          if (A_KM) tma_load_2d(sa, &mapA, 2 * kt * BK, tm * BM, full);
          else {
#pragma unroll
            for (int b = 0; b < BM / 8; b++) tma_load_2d(sa + b * 1024, &mapA, 2 * (tm * BM + b * 8), kt * BK, full);
          }
          if (B_KM) tma_load_2d(sb, &mapB, 2 * kt * BK, tn * BN, full);
          else {
#pragma unroll
            for (int b = 0; b < BN / 8; b++) tma_load_2d(sb + b * 1024, &mapB, 2 * (tn * BN + b * 8), kt * BK, full);
          }
          if (pf)
            for (int q = 0; q < pf_step && pf_col < pf_cols; q++, pf_col++) prefetch_l2_bulk(pf_base + (size_t)pf_col * ldc, pf_bytes);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;              // 2 x 4 warps, warp tile 32 x 32 complex
  const int rho = rho_of(g);
  uint32_t offA[2], offB[2];
#pragma unroll
  for (int s = 0; s < 2; s++) {
    offA[s] = frag_off<A_KM>(wm * 4, 4 * s + t, rho);
    offB[s] = A_BYTES + frag_off<B_KM>(wn * 4, 4 * s + t, rho);
  }
  const double sgn_a = conj_a ? -1.0 : 1.0, sgn_b = conj_b ? -1.0 : 1.0;
  int stage = 0; uint32_t phase = 0;
  const bool beta0 = (beta.x == 0.0 && beta.y == 0.0);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int tm, tn; tile_coords(tile, tiles_m, tiles_n, tm, tn);
    if (tile_skipped(tri, tm, tn)) continue;
    double cr[4][4][2], ci[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { cr[i][j][0] = cr[i][j][1] = 0.0; ci[i][j][0] = ci[i][j][1] = 0.0; }

    double ar[2][4], ai[2][4], br[2][4], bi[2][4];
    auto load_frags = [&](int buf, uint32_t sbase, int s) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        lds128(sbase + offA[s] + i * 1024, ar[buf][i], ai[buf][i]);
        ai[buf][i] *= sgn_a;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        lds128(sbase + offB[s] + j * 1024, br[buf][j], bi[buf][j]);
        bi[buf][j] *= sgn_b;
      }
    };

    mbar_wait(bar_full + stage * 8, phase);
    load_frags(0, base + stage * STAGE_BYTES, 0);
    for (int kt = 0; kt < KT; kt++) {
      const uint32_t sbase = base + stage * STAGE_BYTES;
      int nstage = stage + 1; uint32_t nphase = phase;
      if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const int cur = s & 1, nxt = cur ^ 1;
        if (s < 1) load_frags(nxt, sbase, s + 1);
        else if (kt + 1 < KT) {
          mbar_wait(bar_full + nstage * 8, nphase);
          load_frags(nxt, base + nstage * STAGE_BYTES, 0);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const double nai = -ai[cur][i];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            dmma884(cr[i][j][0], cr[i][j][1], ar[cur][i], br[cur][j]);
            dmma884(ci[i][j][0], ci[i][j][1], ar[cur][i], bi[cur][j]);
            dmma884(cr[i][j][0], cr[i][j][1], nai, bi[cur][j]);
            dmma884(ci[i][j][0], ci[i][j][1], ai[cur][i], br[cur][j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + stage * 8);
      stage = nstage; phase = nphase;
    }

    // epilogue: C = alpha*(cr + i ci) + beta*C, one 16-byte store per element
    const int row0 = tm * BM + wm * 32, col0 = tn * BN + wn * 32;
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int col = col0 + 8 * j + rho_of(2 * t + h);
        if (col >= n) continue;
        cdouble *ccol = C + (size_t)col * ldc;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int row = row0 + 8 * i + rho;
          if (row >= m) continue;
          if (tri == 'L' && row < col) continue;
          if (tri == 'U' && row > col) continue;
          const double xr = cr[i][j][h], xi = ci[i][j][h];
          cdouble r;
          r.x = alpha.x * xr - alpha.y * xi;
          r.y = alpha.x * xi + alpha.y * xr;
          if (!beta0) {
            const cdouble c0 = ccol[row];
            r.x += beta.x * c0.x - beta.y * c0.y;
            r.y += beta.x * c0.y + beta.y * c0.x;
          }
          *reinterpret_cast<double2 *>(&ccol[row]) = make_double2(r.x, r.y);
        }
      }
    }
  }
}

template <bool A_KM, bool B_KM>
static int launch(const CUtensorMap &ma, const CUtensorMap &mb, int m, int n, int k, cdouble alpha, cdouble beta,
                  cdouble *C, int ldc, int tri, int ca, int cb, cudaStream_t s) {
  JB_ONCE_PER_DEVICE(JB_CUDA(cudaFuncSetAttribute(zgemm_dmma_kernel<A_KM, B_KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  const int ntiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  int grid = ntiles < num_sms() ? ntiles : num_sms();
  // "gemm.max_ctas": a CTA of this persistent kernel owns a whole SM (registers + shared memory), so callers that
  // overlap it with other kernels — the NCCL panel broadcasts of the sharded GEMM — leave a few SMs free
  const int cap = option("gemm.max_ctas");
  if (cap > 0 && cap < grid) grid = cap;
  const int tcap = tls().gemm_cta_cap;          // set by the look-ahead factorisations around their trailing update
  if (tcap > 0 && tcap < grid) grid = tcap;
  zgemm_dmma_kernel<A_KM, B_KM><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, ca, cb);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace zdmma

int zgemm_dmma_try(int ta, int tb, int m, int n, int k, cdouble alpha, const cdouble *A, int lda, const cdouble *B,
                   int ldb, cdouble beta, cdouble *C, int ldc, cudaStream_t s, int tri) {
  using namespace zdmma;
  if (m < 32 || n < 32 || k < 1) return 0;
  if ((long)((m + BM - 1) / BM) * ((n + BN - 1) / BN) < 48 && option("gemm.force_dmma") <= 0) return 0;   // see gemm_f64_dmma.cu
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15)) return 0;
  const bool a_km = (ta != 111), b_km = (tb == 111);
  const int ca = (ta == 113), cb = (tb == 113);
  CUtensorMap ma, mb;
  int rc;
  // tensor maps over FP64 words: a complex column-major matrix (rows x cols, ld) is 2*rows words x cols, pitch 16*ld
  if (a_km) rc = dmma::make_map_f64(&ma, A, 2L * k, m, (size_t)lda * 16, 16, BM);
  else rc = dmma::make_map_f64(&ma, A, 2L * m, k, (size_t)lda * 16, 16, BK);
  if (rc) return rc;
  if (b_km) rc = dmma::make_map_f64(&mb, B, 2L * k, n, (size_t)ldb * 16, 16, BN);
  else rc = dmma::make_map_f64(&mb, B, 2L * n, k, (size_t)ldb * 16, 16, BK);
  if (rc) return rc;
  if (a_km && b_km) return launch<true, true>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, ca, cb, s);
  if (a_km && !b_km) return launch<true, false>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, ca, cb, s);
  if (!a_km && b_km) return launch<false, true>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, ca, cb, s);
  return launch<false, false>(ma, mb, m, n, k, alpha, beta, C, ldc, tri, ca, cb, s);
}

}  // namespace jb
