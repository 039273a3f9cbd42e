// Blocked right-looking LU / Cholesky and the triangular solves behind
// LAPACKE_?{getrf,getrs,getri,gesv,potrf,potrs,posv} — the routines Jalla reaches at
// /root/reference/Numeric/Jalla/Matrix.hs:503 (gesv), :766 (getrf), :769 (getri) and binds
// raw at Foreign/LAPACKE.chs:703-706,733-736,741-744,601-604.
// Structure (reference-LAPACK's, re-cut for the GPU): a narrow panel is factorised by one
// CTA that keeps the panel in shared memory, row interchanges are applied by a swap
// kernel, the block row is solved against the unit-lower diagonal block, and the trailing
// matrix is updated by this library's own GEMM (gemm_device: DMMA for double).
#include "jb_common.cuh"
#include <vector>
#include <atomic>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// Inside the look-ahead loops an error must not `return`: the epilogue joins the second stream into the caller's and
// frees the scratch.  Sets rc and leaves the loop instead.
#define JB_CUDA_BREAK(call)                                                                    \
  {                                                                                            \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      jb::set_error(JB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      rc = JB_ERR_CUDA;                                                                        \
      break;                                                                                   \
    }                                                                                          \
  }

namespace jb {

constexpr int NB = 32;           // panel width of the blocked algorithms
constexpr int PANEL_THREADS = 1024;

// ------------------------------------------------------------------ LU panel (getf2)
// Factorises the pm x jb panel P (column-major, ld) in place with partial pivoting.
// Pivot = first row of maximal |re|+|im| (i?amax semantics, lowest index wins ties).
// ipiv (1-based, global row numbers = local + row_off) and the first zero pivot
// (info, 1-based global column) are written to device memory.
template <typename T>
__global__ void __launch_bounds__(PANEL_THREADS)
lu_panel_kernel(int pm, int jb, T *__restrict__ Pg, int ldg, int *__restrict__ ipiv, int *info, int row_off,
                int use_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using R = typename Num<T>::R;
  __shared__ R red_val[32];
  __shared__ int red_idx[32];
  __shared__ T s_row[NB];
  T *P = Pg;
  int ld = ldg;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (use_smem) {
    P = reinterpret_cast<T *>(smem_raw);
    ld = pm | 1;   // odd leading dimension: row-wise accesses (swaps) spread over banks
    for (int c = 0; c < jb; c++)
      for (int r = tid; r < pm; r += nt) P[r + (size_t)c * ld] = Pg[r + (size_t)c * ldg];
    __syncthreads();
  }
  // Two barriers per column: (B1) after the per-warp pivot candidates are posted — every warp then reduces them
  // redundantly, so no second round trip — and (B2) after the interchange + pivot-row stash.  The rank-1 update
  // needs no barrier of its own: a thread's contribution to the next column's search only involves rows it has
  // just updated itself, and the next interchange happens behind the next B1.
  const int lane = tid & 31, nw = (nt + 31) >> 5;
  for (int c = 0; c < jb && c < pm; c++) {
    // --- pivot search over rows c..pm-1 of column c
    R best = (R)-1;
    int bi = 0x7fffffff;
    for (int r = c + tid; r < pm; r += nt) {
      R v = Num<T>::abs1(P[r + (size_t)c * ld]);
      if (v > best) { best = v; bi = r; }
    }
    // NaNs never win (matches i?amax on OpenBLAS where comparisons with NaN are false)
    for (int off = 16; off > 0; off >>= 1) {
      R ov = __shfl_xor_sync(0xffffffffu, best, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { red_val[tid >> 5] = best; red_idx[tid >> 5] = bi; }
    __syncthreads();                                                       // B1
    best = lane < nw ? red_val[lane] : (R)-1;
    bi = lane < nw ? red_idx[lane] : 0x7fffffff;
    for (int off = 16; off > 0; off >>= 1) {
      R ov = __shfl_xor_sync(0xffffffffu, best, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = c;       // column of NaNs: keep the diagonal like i?amax returning 1
    const int p = bi;
    if (tid == 0) ipiv[c] = p + row_off + 1;
    // --- swap rows c and p across the panel, stash the pivot row
    if (tid < jb) {
      T a = P[c + (size_t)tid * ld], b = P[p + (size_t)tid * ld];
      if (p != c) { P[c + (size_t)tid * ld] = b; P[p + (size_t)tid * ld] = a; }
      s_row[tid] = (p != c) ? b : a;
    }
    __syncthreads();                                                       // B2
    const T piv = s_row[c];
    const bool zero = Num<T>::is_zero(piv);
    if (zero) {
      if (tid == 0 && *info == 0) *info = c + row_off + 1;
    }
    const T rinv = zero ? Num<T>::zero() : Num<T>::recip(piv);
    // --- scale the column and rank-1 update of the remaining panel columns
    for (int r = c + 1 + tid; r < pm; r += nt) {
      T l = P[r + (size_t)c * ld];
      if (!zero) { l = Num<T>::cmul(l, rinv); P[r + (size_t)c * ld] = l; }
      for (int cc = c + 1; cc < jb; cc++)
        P[r + (size_t)cc * ld] = Num<T>::msub(P[r + (size_t)cc * ld], l, s_row[cc]);
    }
  }
  __syncthreads();
  if (use_smem) {
    for (int c = 0; c < jb; c++)
      for (int r = tid; r < pm; r += nt) Pg[r + (size_t)c * ldg] = P[r + (size_t)c * ld];
  }
}

// ------------------------------------------------------------------ LU panel, many CTAs (tall panels)
// Same algorithm and arithmetic as lu_panel_kernel, but the pm rows are split over a cooperative grid: each
// CTA keeps its slab of the panel in shared memory, and per column there is exactly ONE grid-wide barrier.
// Before the barrier every CTA publishes its best pivot candidate together with that row's panel entries (and
// the owner of the diagonal row publishes that row), so after the barrier every CTA can pick the global pivot
// (max magnitude, lowest row wins ties), fetch the pivot row, perform its share of the interchange and update
// its own rows without further communication.  Scratch is double buffered by column parity.
struct PanelScratch {
  // all arrays have a leading [2] parity dimension
  double *cand_val;   // [2][G]   magnitude of the CTA's candidate (-1: none)
  int *cand_row;      // [2][G]   its panel row index
  void *cand_data;    // [2][G][NB] that row's panel entries
  void *diag_data;    // [2][NB]  entries of panel row c (the row being swapped down)
};

// scratch written by other CTAs is read through the L2 (never a possibly stale L1 line)
template <typename T> __device__ __forceinline__ T ld_cg(const T *p) {
  if constexpr (sizeof(T) == 4) { float v = __ldcg(reinterpret_cast<const float *>(p)); return *reinterpret_cast<T *>(&v); }
  else if constexpr (sizeof(T) == 8) { double v = __ldcg(reinterpret_cast<const double *>(p)); return *reinterpret_cast<T *>(&v); }
  else { double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); return *reinterpret_cast<T *>(&v); }
}

template <typename T>
__global__ void __launch_bounds__(256)
lu_panel_coop_kernel(int pm, int jb, T *__restrict__ Pg, int ldg, int *__restrict__ ipiv, int *info, int row_off,
                     int rows_per_cta, PanelScratch sc) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using R = typename Num<T>::R;
  T *S = reinterpret_cast<T *>(smem_raw);                 // [jb][lds] column-major slab
  __shared__ R red_val[8];
  __shared__ int red_idx[8];
  __shared__ T s_u[NB];
  __shared__ int s_piv, s_win;
  const int tid = threadIdx.x, nt = blockDim.x, G = gridDim.x, cta = blockIdx.x;
  const int r_lo = cta * rows_per_cta, r_hi = min(pm, r_lo + rows_per_cta), nr = max(r_hi - r_lo, 0);
  const int lds = rows_per_cta | 1;
  for (int c = 0; c < jb; c++)
    for (int r = tid; r < nr; r += nt) S[r + (size_t)c * lds] = Pg[(r_lo + r) + (size_t)c * ldg];
  __syncthreads();
  T *cand_data = reinterpret_cast<T *>(sc.cand_data);
  T *diag_data = reinterpret_cast<T *>(sc.diag_data);
  const int ncols = min(jb, pm);
  for (int c = 0; c < ncols; c++) {
    const int par = c & 1;
    // ---- local candidate among my rows >= c
    R best = (R)-1; int bi = 0x7fffffff;
    for (int r = max(c - r_lo, 0) + tid; r < nr; r += nt) {
      R v = Num<T>::abs1(S[r + (size_t)c * lds]);
      if (v > best) { best = v; bi = r_lo + r; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      R ov = __shfl_down_sync(0xffffffffu, best, off);
      int oi = __shfl_down_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red_val[tid >> 5] = best; red_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      const int nw = nt >> 5;
      best = tid < nw ? red_val[tid] : (R)-1;
      bi = tid < nw ? red_idx[tid] : 0x7fffffff;
      for (int off = 16; off > 0; off >>= 1) {
        R ov = __shfl_down_sync(0xffffffffu, best, off);
        int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (tid == 0) {
        // a slab that holds row c but only NaNs/none better keeps row c as its candidate (i?amax semantics)
        if (bi == 0x7fffffff && c >= r_lo && c < r_hi) { bi = c; best = (R)-0.5; }
        s_piv = bi;
        sc.cand_val[par * G + cta] = (double)best;
        sc.cand_row[par * G + cta] = bi;
      }
    }
    __syncthreads();
    const int my_cand = s_piv;
    if (tid < jb) {
      if (my_cand != 0x7fffffff) cand_data[((size_t)par * G + cta) * NB + tid] = S[(my_cand - r_lo) + (size_t)tid * lds];
      if (c >= r_lo && c < r_hi) diag_data[par * NB + tid] = S[(c - r_lo) + (size_t)tid * lds];
    }
    __threadfence();
    grid.sync();
    // ---- global pivot: every CTA reduces the G candidates identically
    if (tid < 32) {
      double bv = -1.0; int br = 0x7fffffff, bw = 0;
      for (int g = tid; g < G; g += 32) {
        double v = __ldcg(&sc.cand_val[par * G + g]); int rr = __ldcg(&sc.cand_row[par * G + g]);
        if (rr != 0x7fffffff && (v > bv || (v == bv && rr < br))) { bv = v; br = rr; bw = g; }
      }
      for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, bv, off);
        int orr = __shfl_down_sync(0xffffffffu, br, off), ow = __shfl_down_sync(0xffffffffu, bw, off);
        if (orr != 0x7fffffff && (ov > bv || (ov == bv && orr < br))) { bv = ov; br = orr; bw = ow; }
      }
      if (tid == 0) { s_piv = br; s_win = bw; }
    }
    __syncthreads();
    const int p = s_piv, win = s_win;
    if (tid < jb) s_u[tid] = ld_cg<T>(&cand_data[((size_t)par * G + win) * NB + tid]);
    if (cta == 0 && tid == 0) ipiv[c] = p + row_off + 1;
    __syncthreads();
    // ---- my share of the interchange (row c <- pivot row, row p <- old row c)
    if (p != c && tid < jb) {
      if (p >= r_lo && p < r_hi) S[(p - r_lo) + (size_t)tid * lds] = ld_cg<T>(&diag_data[par * NB + tid]);
      if (c >= r_lo && c < r_hi) S[(c - r_lo) + (size_t)tid * lds] = s_u[tid];
    }
    __syncthreads();
    const T piv = s_u[c];
    const bool zero = Num<T>::is_zero(piv);
    if (zero && cta == 0 && tid == 0 && *info == 0) *info = c + row_off + 1;
    const T rinv = zero ? Num<T>::zero() : Num<T>::recip(piv);
    for (int r = max(c + 1 - r_lo, 0) + tid; r < nr; r += nt) {
      T l = S[r + (size_t)c * lds];
      if (!zero) { l = Num<T>::cmul(l, rinv); S[r + (size_t)c * lds] = l; }
      for (int cc = c + 1; cc < jb; cc++) S[r + (size_t)cc * lds] = Num<T>::msub(S[r + (size_t)cc * lds], l, s_u[cc]);
    }
    __syncthreads();
  }
  for (int c = 0; c < jb; c++)
    for (int r = tid; r < nr; r += nt) Pg[(r_lo + r) + (size_t)c * ldg] = S[r + (size_t)c * lds];
}

// ------------------------------------------------------------------ LU panel, rows in registers
// The panel kernels above keep the panel in shared memory and pay an LDS + STS per element per column for the
// rank-1 update (measured 2.3-4.4 us per column).  Here a THREAD owns a panel row.  The not-yet-eliminated part of
// its row lives in registers, ROTATED by one position per column step — a[j] is panel column c+j at step c — so the
// column loop stays rolled (one small loop body in the instruction cache; a fully unrolled variant was 370 KB of
// SASS and instruction-fetch bound) while every register index is still static: the update writes
// a[j] = a[j+1] - l*u[c+1+j].  Multipliers, final once computed, are parked in shared memory L[c][thread] and
// written back at the end.  The pivot search per column is three redux.sync on an order-preserving integer key per
// warp plus one shared-memory hop across the CTA's warps.
//
// With more rows than one CTA holds, a cooperative grid of G CTAs exchanges, per column, a 3-word record (key, row)
// and the candidate row's entries through global scratch written as FLAGGED WORDS: every 8-byte word carries 4 bytes
// of payload and a 4-byte tag (launch sequence * 64 + column), stored and polled with single 8-byte relaxed
// accesses, which are single-copy atomic.  A reader therefore needs no fence, no counter and no grid barrier: it
// polls the G records until their tags match, reduces them, then polls the winner's row.  Two L2 round trips per
// column.  Buffers alternate with column parity: a CTA can run at most one column ahead of the slowest reader,
// because publishing column c+1 needs every CTA's record for column c.  The launch stays cooperative so that all
// G CTAs are co-resident while they poll.
struct PanelLLScratch {
  unsigned long long *rec;    // [2][G][4]        key high word, key low word, row, unused
  unsigned long long *cand;   // [2][G][NB * W]   candidate row of each CTA, W = sizeof(T) / 4 words per entry
  unsigned long long *diag;   // [2][NB * W]      panel row c before the interchange
  long long *prof;            // optional [8] per-phase cycle sums of CTA 0 (debug option getrf.panel_prof), or null
};
#ifndef JB_PANEL_PROF
#define JB_PANEL_PROF 0
#endif
#define JB_PROF(i) do { if (!JB_PANEL_PROF) break; if (sc.prof && cta == 0 && tid == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long *)&sc.prof[i], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)

__device__ __forceinline__ void ll_store(unsigned long long *p, unsigned data, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | data;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// warp-wide (max key, then lowest row) — rows with key 0 are only kept when flagged valid by the caller
__device__ __forceinline__ void warp_best(unsigned long long key, int row, bool valid, unsigned long long &bk, int &br) {
  constexpr int NONE = 0x7fffffff;
  const unsigned hi = __reduce_max_sync(0xffffffffu, valid ? (unsigned)(key >> 32) : 0u);
  const unsigned lo = __reduce_max_sync(0xffffffffu, (valid && (unsigned)(key >> 32) == hi) ? (unsigned)key : 0u);
  bk = ((unsigned long long)hi << 32) | lo;
  br = __reduce_min_sync(0xffffffffu, (valid && key == bk) ? row : NONE);
}

constexpr int PANEL_REC_PER_LANE = 5;     // the record poll covers G <= 160 CTAs
constexpr int PANEL_MULTI_THREADS = 256;  // rows per CTA of the multi-CTA instantiation (registers to spare)
// Second multi-CTA instantiation: 512 rows per CTA for at most 32 CTAs (one record per lane, 128 registers).  Half the
// CTAs for the same panel: a panel of 8k-16k rows then needs <= 32 co-resident CTAs, few enough to run one block ahead
// of the trailing update (getrf_device).  Not for complex double (the multipliers would need 256 KB of shared memory).
constexpr int PANEL_WIDE_THREADS = 512, PANEL_WIDE_MAX_CTAS = 32;

// A panel row as it is passed around: `rot` is the owner's register image (rot[j] = panel column c+j at step c, so
// every access is a static, 16-byte aligned pair), `lpart` its multipliers in absolute columns 0..c-1.
template <typename T> struct alignas(16) PanelRow { T rot[NB]; T lpart[NB]; };
template <typename T> struct alignas(2 * sizeof(T)) PanelPair { T x, y; };

// a[j] <- a[j+1] - l * u[j+1] for the W registers that still hold columns (a[0] is done by the caller); a[W-1] <- 0
template <typename T, int W>
__device__ __forceinline__ void panel_rotate(T (&a)[NB], const T l, const PanelPair<T> *up) {
#pragma unroll
  for (int jj = 1; jj < W / 2; jj++) {
    const PanelPair<T> q = up[jj];
    a[2 * jj - 1] = Num<T>::msub(a[2 * jj], l, q.x);
    a[2 * jj] = Num<T>::msub(a[2 * jj + 1], l, q.y);
  }
  a[W - 1] = Num<T>::zero();
}

template <typename T, int MAXT, bool MULTI, int RPL = PANEL_REC_PER_LANE>
__global__ void __launch_bounds__(MAXT, 1)
lu_panel_reg_kernel(int pm, int jb, T *__restrict__ Pg, int ldg, int *__restrict__ ipiv, int *info, int row_off,
                    PanelLLScratch sc, unsigned tag0) {
  constexpr int W = sizeof(T) / 4;
  constexpr int PW = 2 * NB * W;                        // words per row packet
  constexpr int NONE = 0x7fffffff;
  using Pair = PanelPair<T>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *Lm = reinterpret_cast<T *>(smem_raw);              // [NB][nt] multipliers of the rows of this CTA
  __shared__ unsigned long long s_key[MAXT / 32];
  __shared__ int s_krow[MAXT / 32];
  __shared__ PanelRow<T> s_cand, s_diag, s_u, s_d;
  __shared__ int s_p;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int G = gridDim.x, cta = blockIdx.x, base = cta * nt;
  const int row = base + tid;                           // panel row owned by this thread
  const bool live = row < pm;
  const int ncols = min(jb, pm);
  T a[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) a[j] = (live && j < jb) ? Pg[row + (size_t)j * ldg] : Num<T>::zero();

  long long t_prev = clock64();
  for (int c = 0; c < ncols; c++) {
    // ---- candidate: first row of maximal |re|+|im| among rows >= c; NaNs never win (key 0)
    unsigned long long key = 0;
    if (live && row >= c) {
      const double v = (double)Num<T>::abs1(a[0]);
      if (v == v) key = (unsigned long long)__double_as_longlong(v) + 1ull;
    }
    unsigned long long kb;
    int br;
    warp_best(key, row, key != 0, kb, br);
    if (lane == 0) { s_key[warp] = kb; s_krow[warp] = br; }
    __syncthreads();
    JB_PROF(0);
    {
      const bool has = lane < nw;
      const unsigned long long k = has ? s_key[lane] : 0ull;
      const int r = has ? s_krow[lane] : NONE;
      warp_best(k, r, has && r != NONE, kb, br);
    }
    const bool own_c = c >= base && c < base + nt;
    if (br == NONE && own_c) { br = c; kb = 0; }        // nothing comparable: the diagonal stands (i?amax returns 1)
    const unsigned tag = tag0 + c;
    const int par = c & 1;
    if (MULTI && warp == 0 && lane < 3) {               // the record leaves first; the row follows while it travels
      const unsigned v = lane == 0 ? (unsigned)(kb >> 32) : lane == 1 ? (unsigned)kb : (unsigned)br;
      ll_store(sc.rec + ((size_t)par * G + cta) * 4 + lane, v, tag);
    }
    if (row == br) {
#pragma unroll
      for (int j = 0; j < NB; j += 2) *reinterpret_cast<Pair *>(&s_cand.rot[j]) = Pair{a[j], a[j + 1]};
    }
    if (row == c && (MULTI || br != c)) {
#pragma unroll
      for (int j = 0; j < NB; j += 2) *reinterpret_cast<Pair *>(&s_diag.rot[j]) = Pair{a[j], a[j + 1]};
    }
    if (warp == (1 % nw) && lane < c && br != NONE) s_cand.lpart[lane] = Lm[(size_t)lane * nt + (br - base)];
    if (warp == (2 % nw) && lane < c && own_c) s_diag.lpart[lane] = Lm[(size_t)lane * nt + (c - base)];
    __syncthreads();
    JB_PROF(1);
    int p = br;
    const PanelRow<T> *u = &s_cand, *d = &s_diag;
    if constexpr (MULTI) {
      if (warp == 0) {
        if (br != NONE) {
          unsigned long long *dst = sc.cand + ((size_t)par * G + cta) * PW;
          const unsigned *src = reinterpret_cast<const unsigned *>(&s_cand);
#pragma unroll
          for (int w = 0; w < 2 * W; w++) ll_store(dst + lane + 32 * w, src[lane + 32 * w], tag);
        }
        if (own_c) {
          unsigned long long *dst = sc.diag + (size_t)par * PW;
          const unsigned *src = reinterpret_cast<const unsigned *>(&s_diag);
#pragma unroll
          for (int w = 0; w < 2 * W; w++) ll_store(dst + lane + 32 * w, src[lane + 32 * w], tag);
        }
        JB_PROF(2);
        // every CTA reduces the G records identically; a lane's records are polled together
        unsigned long long gk = 0;
        int gr = NONE;
        {
          bool ok;
          unsigned long long w0[RPL], w1[RPL], w2[RPL];
          do {
#pragma unroll
            for (int i = 0; i < RPL; i++) {
              const int g = lane + 32 * i;
              if (g < G) {
                const unsigned long long *rg = sc.rec + ((size_t)par * G + g) * 4;
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0[i]) : "l"(rg) : "memory");
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1[i]) : "l"(rg + 1) : "memory");
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w2[i]) : "l"(rg + 2) : "memory");
              }
            }
            ok = true;
#pragma unroll
            for (int i = 0; i < RPL; i++)
              if (lane + 32 * i < G)
                ok = ok && (unsigned)(w0[i] >> 32) == tag && (unsigned)(w1[i] >> 32) == tag && (unsigned)(w2[i] >> 32) == tag;
          } while (!ok);
#pragma unroll
          for (int i = 0; i < RPL; i++)
            if (lane + 32 * i < G) {
              const unsigned long long k = ((unsigned long long)(unsigned)w0[i] << 32) | (unsigned)w1[i];
              const int r = (int)(unsigned)w2[i];
              if (r != NONE && (gr == NONE || k > gk || (k == gk && r < gr))) { gk = k; gr = r; }
            }
        }
        unsigned long long gbest;
        warp_best(gk, gr, gr != NONE, gbest, p);
        JB_PROF(3);
        const int win = p / nt;
        const bool need_d = p != c && p >= base && p < base + nt;
        {
          // the winner's row and, where this CTA holds row p, old row c: one round trip
          unsigned uw[2 * W], dw[2 * W];
          const unsigned long long *srcu = sc.cand + ((size_t)par * G + win) * PW;
          const unsigned long long *srcd = sc.diag + (size_t)par * PW;
          bool ok;
          do {
            unsigned long long xu[2 * W], xd[2 * W];
#pragma unroll
            for (int w = 0; w < 2 * W; w++)
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(xu[w]) : "l"(srcu + lane + 32 * w) : "memory");
            if (need_d) {
#pragma unroll
              for (int w = 0; w < 2 * W; w++)
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(xd[w]) : "l"(srcd + lane + 32 * w) : "memory");
            }
            ok = true;
#pragma unroll
            for (int w = 0; w < 2 * W; w++) { ok = ok && (unsigned)(xu[w] >> 32) == tag; uw[w] = (unsigned)xu[w]; }
            if (need_d) {
#pragma unroll
              for (int w = 0; w < 2 * W; w++) { ok = ok && (unsigned)(xd[w] >> 32) == tag; dw[w] = (unsigned)xd[w]; }
            }
          } while (!ok);
          unsigned *du = reinterpret_cast<unsigned *>(&s_u), *dd = reinterpret_cast<unsigned *>(&s_d);
#pragma unroll
          for (int w = 0; w < 2 * W; w++) du[lane + 32 * w] = uw[w];
          if (need_d) {
#pragma unroll
            for (int w = 0; w < 2 * W; w++) dd[lane + 32 * w] = dw[w];
          }
        }
        if (lane == 0) s_p = p;
        JB_PROF(4);
      }
      __syncthreads();
      JB_PROF(5);
      p = s_p;
      u = &s_u;
      d = &s_d;
    }
    if (cta == 0 && tid == 0) ipiv[c] = p + row_off + 1;
    // ---- interchange: panel row c is final from here on (multipliers left of the diagonal, U right of it), so
    // it goes straight to memory from the broadcast copy; the thread that held the pivot row takes what row c held
    if (cta == 0 && warp == 0 && lane < jb) Pg[c + (size_t)lane * ldg] = lane < c ? u->lpart[lane] : u->rot[lane - c];
    if (p != c && p >= base && p < base + nt) {
      if (row == p) {
#pragma unroll
        for (int j = 0; j < NB; j += 2) {
          const Pair q = *reinterpret_cast<const Pair *>(&d->rot[j]);
          a[j] = q.x; a[j + 1] = q.y;
        }
      }
      if (warp == (1 % nw) && lane < c) Lm[(size_t)lane * nt + (p - base)] = d->lpart[lane];
    }
    const Pair *up = reinterpret_cast<const Pair *>(u->rot);
    const Pair u0 = up[0];
    const T piv = u0.x;
    const bool zero = Num<T>::is_zero(piv);
    if (zero && cta == 0 && tid == 0 && *info == 0) *info = c + row_off + 1;
    const T rinv = zero ? Num<T>::zero() : Num<T>::recip(piv);
    // ---- eliminate and rotate: register j takes panel column c+1+j.  Rows <= c are dead from here on and
    // compute garbage that nothing reads (their keys are 0, their Lm slots are never consulted).
    T l = a[0];
    if (!zero) l = Num<T>::cmul(l, rinv);
    Lm[(size_t)c * nt + tid] = l;
    a[0] = Num<T>::msub(a[1], l, u0.y);
    // registers >= NB - c hold no column any more (zeros): the update covers only the width that is left, in four
    // compile-time widths so that every register index stays static (warp-uniform branch on c)
    if (c < 8) panel_rotate<T, NB>(a, l, up);
    else if (c < 16) panel_rotate<T, NB - 8>(a, l, up);
    else if (c < 24) panel_rotate<T, NB - 16>(a, l, up);
    else panel_rotate<T, NB - 24>(a, l, up);
    JB_PROF(6);
  }
  __syncthreads();
  if (live && row >= ncols) {
    for (int j = 0; j < ncols; j++) Pg[row + (size_t)j * ldg] = Lm[(size_t)j * nt + tid];
  }
}

// Applies interchanges ipiv[k0..k1) (1-based rows of the whole matrix; A points at its row `row_base`) to columns
// [c0,c1) of A, forward.
//
// A thread per column walking the chain pays 2 dependent, uncoalesced global round trips per interchange (measured
// 43 us per launch inside dgetrf 4096 — a third of the factorisation).  Here each chunk of <= 32 interchanges is
// first composed into ONE permutation of the <= 64 rows it touches: slots 0..31 are rows q0..q0+31, slots 32..63 the
// pivot rows outside that range (duplicates merged with match.any); one lane replays the chain on slot indices.
// Then a WARP owns a column: lane i fetches the entries that end up in slot i and slot 32+i (independent loads, the
// in-block half contiguous across the warp) and stores them, four columns in flight per warp.
constexpr int LASWP_THREADS = 256, LASWP_MAXCHUNK = 16, LASWP_COLS_PER_WARP = 4;
template <typename T>
__global__ void __launch_bounds__(LASWP_THREADS)
laswp_kernel(T *__restrict__ A, int lda, int c0, int c1, const int *__restrict__ ipiv, int k0, int k1, int row_base,
             int tri) {
  __shared__ int s_rows[LASWP_MAXCHUNK][64], s_cur[LASWP_MAXCHUNK][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = LASWP_THREADS / 32;
  const int nchunk = (k1 - k0 + 31) / 32;
  for (int q = warp; q < nchunk; q += nwarp) {
    const int q0 = k0 + 32 * q, np = min(32, k1 - q0);
    int p = -1 - lane, slot = lane;                  // inactive lanes: unique dummy rows, identity slots
    if (lane < np) {
      p = ipiv[q0 + lane] - 1 - row_base;
      slot = (p >= q0 && p < q0 + 32) ? p - q0 : -1;
    }
    const unsigned same = __match_any_sync(0xffffffffu, p);
    if (slot < 0) slot = 32 + (__ffs(same) - 1);     // first interchange that named this outside row owns its slot
    s_rows[q][lane] = q0 + lane;
    s_rows[q][32 + lane] = p;
    s_cur[q][lane] = lane;
    s_cur[q][32 + lane] = 32 + lane;
    __syncwarp();
    // replay the chain on slot indices: lanes exchange through shuffles, lane 0 applies
    for (int k = 0; k < np; k++) {
      const int b = __shfl_sync(0xffffffffu, slot, k);
      if (lane == 0) { const int t = s_cur[q][k]; s_cur[q][k] = s_cur[q][b]; s_cur[q][b] = t; }
    }
    __syncwarp();
  }
  __syncthreads();
  const int col_first = c0 + (blockIdx.x * nwarp + warp) * LASWP_COLS_PER_WARP;
  if (col_first >= c1) return;
  // tri: column c only takes the interchanges from the end of its own 32-column block on (the deferred left-hand
  // interchanges of a panel sweep; k0 and c0 are multiples of 32 there, so the start is uniform across the warp)
  const int q_first = tri ? max(0, ((col_first / NB + 1) * NB - k0) / 32) : 0;
  for (int q = q_first; q < nchunk; q++) {
    // slot i finally holds what slot s_cur[i] held before the chunk
    const int src_lo = s_cur[q][lane], src_hi = s_cur[q][32 + lane];
    const bool mv_lo = src_lo != lane, mv_hi = src_hi != 32 + lane;
    const int rs_lo = s_rows[q][src_lo], rs_hi = s_rows[q][src_hi];
    const int rd_lo = s_rows[q][lane], rd_hi = s_rows[q][32 + lane];
    T v_lo[LASWP_COLS_PER_WARP], v_hi[LASWP_COLS_PER_WARP];
#pragma unroll
    for (int j = 0; j < LASWP_COLS_PER_WARP; j++) {
      const int c = col_first + j;
      if (c < c1) {
        const T *col = A + (size_t)c * lda;
        if (mv_lo) v_lo[j] = col[rs_lo];
        if (mv_hi) v_hi[j] = col[rs_hi];
      }
    }
    __syncwarp();                                    // every lane has read before any lane overwrites
#pragma unroll
    for (int j = 0; j < LASWP_COLS_PER_WARP; j++) {
      const int c = col_first + j;
      if (c < c1) {
        T *col = A + (size_t)c * lda;
        if (mv_lo) col[rd_lo] = v_lo[j];
        if (mv_hi) col[rd_hi] = v_hi[j];
      }
    }
    __syncwarp();                                    // the next chunk reads what this one wrote
  }
}

template <typename T>
static int laswp_launch(T *A, int lda, int c0, int c1, const int *d_ipiv, int k0, int k1, int row_base, cudaStream_t s,
                        int tri = 0) {
  if (c1 <= c0 || k1 <= k0) return 0;
  const int cols_per_cta = (LASWP_THREADS / 32) * LASWP_COLS_PER_WARP;
  for (int q0 = k0; q0 < k1; q0 += 32 * LASWP_MAXCHUNK) {
    const int q1 = (k1 - q0 < 32 * LASWP_MAXCHUNK) ? k1 : q0 + 32 * LASWP_MAXCHUNK;
    // tri: columns whose block ends at or after q1 have nothing to take from this launch
    const int cend = tri ? min(c1, ((q1 - 1) / NB) * NB) : c1;
    if (cend <= c0) continue;
    laswp_kernel<T><<<(cend - c0 + cols_per_cta - 1) / cols_per_cta, LASWP_THREADS, 0, s>>>(A, lda, c0, cend, d_ipiv, q0, q1, row_base, tri);
    count_launch();
  }
  return 0;
}

template <typename T> __device__ __forceinline__ T shfl_T(T v, int src) {
  if constexpr (Num<T>::cplx) { v.x = __shfl_sync(0xffffffffu, v.x, src); v.y = __shfl_sync(0xffffffffu, v.y, src); return v; }
  else return __shfl_sync(0xffffffffu, v, src);
}
// Interchanges of one panel applied to the columns right of it AND the unit-lower solve of the panel's block row
// against the diagonal block, in one launch (they touch the same columns; separately they were two launch-bound
// kernels per panel step).  Same permutation scheme as laswp_kernel for the <= 32 interchanges k0..k1; the warp that
// owns a column then has the block row in its lanes (lane = row k0 + lane) and solves L11 x = b by shuffles, four
// columns interleaved.  L11 = A[k0.., k0..] (unit diagonal, strictly lower part used).
template <typename T>
__global__ void __launch_bounds__(LASWP_THREADS)
laswp_trsm_kernel(T *__restrict__ A, int lda, int c0, int c1, const int *__restrict__ ipiv, int k0, int k1, int row_base) {
  __shared__ int s_rows[64], s_cur[64];
  __shared__ T Ms[NB][NB + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = LASWP_THREADS / 32;
  const int np = k1 - k0;
  for (int e = threadIdx.x; e < NB * NB; e += LASWP_THREADS) {
    const int r = e % NB, c = e / NB;
    Ms[r][c] = (r < np && c < r) ? A[(k0 + r) + (size_t)(k0 + c) * lda] : Num<T>::zero();
  }
  if (warp == 0) {
    int p = -1 - lane, slot = lane;
    if (lane < np) {
      p = ipiv[k0 + lane] - 1 - row_base;
      slot = (p >= k0 && p < k0 + 32) ? p - k0 : -1;
    }
    const unsigned same = __match_any_sync(0xffffffffu, p);
    if (slot < 0) slot = 32 + (__ffs(same) - 1);
    s_rows[lane] = k0 + lane;
    s_rows[32 + lane] = p;
    s_cur[lane] = lane;
    s_cur[32 + lane] = 32 + lane;
    __syncwarp();
    for (int k = 0; k < np; k++) {
      const int b = __shfl_sync(0xffffffffu, slot, k);
      if (lane == 0) { const int t = s_cur[k]; s_cur[k] = s_cur[b]; s_cur[b] = t; }
    }
  }
  __syncthreads();
  const int col_first = c0 + (blockIdx.x * nwarp + warp) * LASWP_COLS_PER_WARP;
  if (col_first >= c1) return;
  const int src_lo = s_cur[lane], src_hi = s_cur[32 + lane];
  const bool mv_hi = src_hi != 32 + lane;
  const int rs_lo = s_rows[src_lo], rs_hi = s_rows[src_hi];
  const int rd_hi = s_rows[32 + lane];
  const bool in_rows = lane < np;                       // slots >= np of a ragged last panel are other rows: untouched
  T x[LASWP_COLS_PER_WARP], v_hi[LASWP_COLS_PER_WARP];
#pragma unroll
  for (int j = 0; j < LASWP_COLS_PER_WARP; j++) {
    const int c = col_first + j;
    x[j] = Num<T>::zero();
    if (c < c1) {
      const T *col = A + (size_t)c * lda;
      if (in_rows || src_lo != lane) x[j] = col[rs_lo];   // what ends up in row k0 + lane
      if (mv_hi) v_hi[j] = col[rs_hi];
    }
  }
  __syncwarp();                                         // every lane has read before any lane overwrites
  // unit-lower forward substitution, lane = row; rows >= np carry values that no lane below them uses (Ms is zero there)
#pragma unroll
  for (int i = 0; i < NB; i++) {
    if (i < np) {
#pragma unroll
      for (int j = 0; j < LASWP_COLS_PER_WARP; j++) {
        const T xi = shfl_T<T>(x[j], i);
        if (lane > i && in_rows) x[j] = Num<T>::msub(x[j], Ms[lane][i], xi);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < LASWP_COLS_PER_WARP; j++) {
    const int c = col_first + j;
    if (c < c1) {
      T *col = A + (size_t)c * lda;
      if (in_rows || src_lo != lane) col[k0 + lane] = x[j];
      if (mv_hi) col[rd_hi] = v_hi[j];
    }
  }
}

// ------------------------------------------------------------------ small triangular solves
// Effective triangular block M (jb x jb, jb <= 32) = op(Ablk) with op applied on load.
// mode bits: 1 = effective-lower (forward), 0 = effective-upper (backward); unit = implicit 1 diagonal.
template <typename T>
__device__ inline void load_tri_block(T (*Ms)[NB + 1], const T *__restrict__ Ablk, int lda, int jb, int trans) {
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    int r = e % NB, c = e / NB;
    T v = Num<T>::zero();
    if (r < jb && c < jb) {
      v = (trans == 111) ? Ablk[r + (size_t)c * lda] : Ablk[c + (size_t)r * lda];
      if (trans == 113) v = Num<T>::conj(v);
    }
    Ms[r][c] = v;
  }
}

// Left side: solve M X = B for X (overwrites B: jb x nrhs, column-major).  One warp per
// right-hand side, lane = row; the solved component is broadcast by shuffle.
template <typename T>
__global__ void __launch_bounds__(256)
trsm_small_left_kernel(int lower, int unit, int trans, int jb, int nrhs, const T *__restrict__ Ablk, int lda,
                       T *__restrict__ B, int ldb, const int *skip) {
  if (skip && *skip) return;
  __shared__ T Ms[NB][NB + 1];
  load_tri_block<T>(Ms, Ablk, lda, jb, trans);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = blockIdx.x * nw + warp; c < nrhs; c += gridDim.x * nw) {
    T *b = B + (size_t)c * ldb;
    T x = lane < jb ? b[lane] : Num<T>::zero();
    if (lower) {
      for (int i = 0; i < jb; i++) {
        T xi = x;
        if (!unit && lane == i) xi = Num<T>::cmul(x, Num<T>::recip(Ms[i][i]));
        if constexpr (Num<T>::cplx) {
          xi.x = __shfl_sync(0xffffffffu, xi.x, i); xi.y = __shfl_sync(0xffffffffu, xi.y, i);
        } else {
          xi = __shfl_sync(0xffffffffu, xi, i);
        }
        if (lane == i) x = xi;
        else if (lane > i) x = Num<T>::msub(x, Ms[lane][i], xi);
      }
    } else {
      for (int i = jb - 1; i >= 0; i--) {
        T xi = x;
        if (!unit && lane == i) xi = Num<T>::cmul(x, Num<T>::recip(Ms[i][i]));
        if constexpr (Num<T>::cplx) {
          xi.x = __shfl_sync(0xffffffffu, xi.x, i); xi.y = __shfl_sync(0xffffffffu, xi.y, i);
        } else {
          xi = __shfl_sync(0xffffffffu, xi, i);
        }
        if (lane == i) x = xi;
        else if (lane < i) x = Num<T>::msub(x, Ms[lane][i], xi);
      }
    }
    if (lane < jb) b[lane] = x;
  }
}

// Right side: solve X M = B for X (B is nrows x jb, column-major), M effective-upper
// (= op(Ablk)); one thread per row of B, the row lives in registers.
// Used by Cholesky-lower: L21 <- A21 * L11^{-H}  (M = L11^H, trans = 113).
template <typename T>
__global__ void __launch_bounds__(256)
trsm_small_right_upper_kernel(int trans, int jb, int nrows, const T *__restrict__ Ablk, int lda, T *__restrict__ B,
                              int ldb, const int *skip) {
  if (skip && *skip) return;
  __shared__ T Ms[NB][NB + 1];
  load_tri_block<T>(Ms, Ablk, lda, jb, trans);
  __syncthreads();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  T x[NB];
#pragma unroll
  for (int c = 0; c < NB; c++) x[c] = c < jb ? B[r + (size_t)c * ldb] : Num<T>::zero();
#pragma unroll
  for (int c = 0; c < NB; c++) {
    if (c < jb) {
      T s = x[c];
#pragma unroll
      for (int p = 0; p < c; p++) s = Num<T>::msub(s, x[p], Ms[p][c]);
      x[c] = Num<T>::cmul(s, Num<T>::recip(Ms[c][c]));
    }
  }
#pragma unroll
  for (int c = 0; c < NB; c++)
    if (c < jb) B[r + (size_t)c * ldb] = x[c];
}

// ------------------------------------------------------------------ Cholesky diagonal block (potf2)
// One CTA of 32x32 threads... kept simple: 1024 threads, block in shared memory,
// right-looking; info = first non-positive pivot (1-based global index), block left
// partially factored like ?potf2.  Only the uplo triangle is read or written.
template <typename T>
__global__ void __launch_bounds__(1024)
potf2_block_kernel(int lower, int jb, T *__restrict__ A, int lda, int *info, int off) {
  if (*info) return;
  using R = typename Num<T>::R;
  __shared__ T S[NB][NB + 1];   // S[r][c], lower-effective: for upper we load the conj transpose
  __shared__ int s_fail;
  const int r = threadIdx.x % NB, c = threadIdx.x / NB;
  if (threadIdx.x == 0) s_fail = 0;
  T v = Num<T>::zero();
  if (r < jb && c < jb && r >= c) {
    v = lower ? A[r + (size_t)c * lda] : Num<T>::conj(A[c + (size_t)r * lda]);
  }
  S[r][c] = v;
  __syncthreads();
  for (int j = 0; j < jb; j++) {
    R d = Num<T>::re(S[j][j]);
    if (!(d > (R)0)) {
      if (threadIdx.x == 0) { s_fail = j + 1; }
      __syncthreads();
      break;
    }
    R sq = sqrt(d), rinv = (R)1 / sq;
    __syncthreads();
    if (c == j && r == j) S[j][j] = Num<T>::make(sq, 0);
    if (c == j && r > j && r < jb) S[r][j] = Num<T>::scale(S[r][j], rinv);
    __syncthreads();
    if (c > j && r >= c && r < jb) S[r][c] = Num<T>::msub(S[r][c], S[r][j], Num<T>::conj(S[c][j]));
    __syncthreads();
  }
  if (r < jb && c < jb && r >= c) {
    T o = S[r][c];
    if (lower) A[r + (size_t)c * lda] = o; else A[c + (size_t)r * lda] = Num<T>::conj(o);
  }
  if (threadIdx.x == 0 && s_fail) *info = off + s_fail;
}

// ------------------------------------------------------------------ Cholesky panel, rows in registers
// Factorises a panel of the lower-form matrix — its jb x jb diagonal block AND every row below it — in one launch:
// potf2 on the block and the triangular solve of the rows under it, fused.  Same register scheme as the LU panel
// (a thread owns a row, the part not yet eliminated is rotated through its registers, finished multipliers wait in
// shared memory), but without a pivot search the only thing a column step needs from another thread is row c of
// the diagonal block, scaled: u[j] = conj(L[c+j][c]).  Warp 0 of EVERY CTA carries the <= 32 diagonal-block rows
// redundantly (mirrored to full Hermitian rows on load, only the stored triangle is read), so CTAs never talk to
// each other and the grid needs no co-residency: CTA g owns rows 32 + g*(NT-32) ... and the CTA that finishes last
// writes the diagonal block back.  Row c's owner posts its scaled row before the column's single barrier (double buffered by
// column parity).  Arithmetic per entry is potf2's: s = sqrt(d), l = a * (1/s), a -= l * conj(l').
// `lower == 0` runs the same code on the conjugate-transposed addressing (the panel is then a block row of U).
// info = first non-positive (or NaN) pivot, 1-based + off; the panel is then left partially factored like ?potf2.
constexpr int CHOL_PANEL_THREADS = 256;
template <typename T>
__global__ void __launch_bounds__(CHOL_PANEL_THREADS, 1)
chol_panel_kernel(int lower, int m, int jb, T *__restrict__ P, int lda, int *info, int off, int *done_count) {
  // One read per CTA.  info may also be set by CTA 0 of THIS launch (a non-positive pivot in columns off+1 .. off+jb) while
  // CTAs of a later wave are still starting: such a value must not make them skip — every CTA carries the diagonal block
  // itself, reaches the same failed_at and parks its partially updated rows like the others, so what A holds for
  // info > 0 does not depend on scheduling.  Only a failure of an EARLIER panel (info <= off) skips the launch.
  __shared__ int s_skip;
  if (threadIdx.x == 0) { const int v = *info; s_skip = (v > off && v <= off + jb) ? 0 : v; }
  __syncthreads();
  if (s_skip) {
    // an earlier panel failed: nothing to do, but the completion count (see the end of the kernel) stays consistent
    if (threadIdx.x == 0 && atomicAdd(done_count, 1) == (int)gridDim.x - 1) *done_count = 0;
    return;
  }
  using R = typename Num<T>::R;
  using Pair = PanelPair<T>;
  constexpr int NT = CHOL_PANEL_THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *Lm = reinterpret_cast<T *>(smem_raw);              // [NB][NT] finished multipliers of this CTA's rows
  struct alignas(16) Posted { T rot[NB]; R rinv; int fail; };
  __shared__ Posted s_u[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = warp == 0 ? tid : NB + blockIdx.x * (NT - NB) + (tid - NB);
  const bool live = row < m;
  auto at = [&](int r, int j) -> T * { return lower ? P + r + (size_t)j * lda : P + j + (size_t)r * lda; };
  T a[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) {
    T v = Num<T>::zero();
    if (live && j < jb) {
      // rows of the diagonal block are completed from the stored triangle: (r, j>r) = conj((j, r))
      const bool mirror = row < jb && j > row;
      v = mirror ? *at(j, row) : *at(row, j);
      if (mirror != (lower == 0)) v = Num<T>::conj(v);
    }
    a[j] = v;
  }
  int failed_at = -1;
  for (int c = 0; c < jb; c++) {
    const int par = c & 1;
    if (row == c) {
      const R d = Num<T>::re(a[0]);
      const bool ok = d > (R)0;
      const R sq = sqrt(d), rinv = (R)1 / sq;
      s_u[par].fail = !ok;
      s_u[par].rinv = rinv;
      s_u[par].rot[0] = Num<T>::make(sq, 0);
#pragma unroll
      for (int j = 1; j < NB; j++) s_u[par].rot[j] = Num<T>::scale(a[j], rinv);
    }
    __syncthreads();
    if (s_u[par].fail) { failed_at = c; break; }
    const Pair *up = reinterpret_cast<const Pair *>(s_u[par].rot);
    const Pair u0 = up[0];
    // rows above c are finished and compute garbage nothing reads; row c parks the diagonal entry
    const T l = row == c ? u0.x : Num<T>::scale(a[0], s_u[par].rinv);
    Lm[(size_t)c * NT + tid] = l;
    a[0] = Num<T>::msub(a[1], l, u0.y);
#pragma unroll
    for (int jj = 1; jj < NB / 2; jj++) {
      const Pair q = up[jj];
      a[2 * jj - 1] = Num<T>::msub(a[2 * jj], l, q.x);
      a[2 * jj] = Num<T>::msub(a[2 * jj + 1], l, q.y);
    }
    a[NB - 1] = Num<T>::zero();
  }
  if (failed_at >= 0 && blockIdx.x == 0 && tid == 0) *info = off + failed_at + 1;
  // Every CTA holds the factorised diagonal block (bit-identical), but it must not reach memory while a CTA of a
  // later wave can still load the original: the grid need not be co-resident (large m, or SMs shared with the
  // look-ahead's trailing update).  So the CTA that finishes LAST writes it — by then every CTA has consumed its loads.
  __shared__ int s_last;
  __syncthreads();
  if (tid == 0) {
    s_last = atomicAdd(done_count, 1) == (int)gridDim.x - 1;
    if (s_last) *done_count = 0;                        // ready for the next launch
  }
  __syncthreads();
  const bool writer = live && (warp != 0 || s_last);
  if (!writer) return;
  if (failed_at >= 0) {
    // columns [0, failed_at) are final; park the partially updated rest (a[k] = column failed_at + k) next to them
#pragma unroll
    for (int k = 0; k < NB; k++)
      if (failed_at + k < NB) Lm[(size_t)(failed_at + k) * NT + tid] = a[k];
  }
  const int last = row < jb ? row : jb - 1;             // only the stored triangle of the diagonal block
  for (int j = 0; j <= last; j++) {
    T v = Lm[(size_t)j * NT + tid];
    if (!lower) v = Num<T>::conj(v);
    *at(row, j) = v;
  }
}

// ------------------------------------------------------------------ helpers for getri / getrs
template <typename T>
__global__ void set_identity_kernel(int n, T *W, int ldw) {
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      W[(size_t)i + (size_t)j * ldw] = (i == j) ? Num<T>::one() : Num<T>::zero();
}
template <typename T>
__global__ void diag_zero_kernel(int n, const T *A, int lda, int *info) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (Num<T>::is_zero(A[(size_t)i + (size_t)i * lda])) atomicMin(info, i + 1);
}
// perm[i] = source row that ends up in row i after applying ipiv[0..n) forward to the identity
__global__ void perm_from_ipiv_kernel(int n, const int *__restrict__ ipiv, int *__restrict__ perm, int use_smem) {
  // single CTA; the swap chain is inherently sequential (n dependent swaps of ints), so it
  // runs on one thread against shared memory when the two arrays fit
  extern __shared__ int sm_perm[];
  int *w = use_smem ? sm_perm : perm;
  int *piv = use_smem ? sm_perm + n : nullptr;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { w[i] = i; if (use_smem) piv[i] = ipiv[i]; }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < n; k++) {
      int p = (use_smem ? piv[k] : ipiv[k]) - 1;
      if (p != k && p >= 0 && p < n) { int t = w[k]; w[k] = w[p]; w[p] = t; }
    }
  __syncthreads();
  if (use_smem) for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = w[i];
}
// The same permutation without the n-step serial chain (931 us at n = 16384 — more than the two triangular sweeps of a
// one-column ?getrs together): the interchanges are taken 32 at a time, each chunk composed into ONE permutation of the
// <= 64 positions it touches (the scheme of laswp_kernel: slots 0..31 are positions q0..q0+31, slots 32..63 the pivot
// rows outside that range, one lane replays the chain on slot indices) — chunks are independent there, so 32 warps
// compose 64 chunks at a time — and one warp then applies the composed chunks in order to the index vector in shared
// memory, 64 entries per step.  Single CTA of 1024 threads; w (n ints) + the tables of 64 chunks in shared memory.
constexpr int PERM_THREADS = 1024, PERM_BATCH = 64;
__global__ void __launch_bounds__(PERM_THREADS)
perm_from_ipiv_chunked_kernel(int n, const int *__restrict__ ipiv, int *__restrict__ perm) {
  extern __shared__ int sm_perm[];
  int *const w = sm_perm;                                   // [n]
  int *const t_rows = sm_perm + n;                          // [PERM_BATCH][64]
  int *const t_cur = t_rows + PERM_BATCH * 64;              // [PERM_BATCH][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = PERM_THREADS / 32;
  for (int i = threadIdx.x; i < n; i += PERM_THREADS) w[i] = i;
  const int nchunk = (n + 31) / 32;
  for (int b0 = 0; b0 < nchunk; b0 += PERM_BATCH) {
    const int nb = min(PERM_BATCH, nchunk - b0);
    __syncthreads();                                        // w initialised / the previous batch applied; tables free
    for (int q = warp; q < nb; q += nwarp) {
      const int q0 = 32 * (b0 + q), np = min(32, n - q0);
      int p = -1 - lane, slot = lane;                       // inactive lanes: unique dummy rows, identity slots
      if (lane < np) {
        p = ipiv[q0 + lane] - 1;
        if (p < 0 || p >= n) p = q0 + lane;                 // out-of-range pivot: no interchange (as the serial kernel)
        slot = (p >= q0 && p < q0 + 32) ? p - q0 : -1;
      }
      const unsigned same = __match_any_sync(0xffffffffu, p);
      if (slot < 0) slot = 32 + (__ffs(same) - 1);
      t_rows[q * 64 + lane] = q0 + lane;
      t_rows[q * 64 + 32 + lane] = p;
      t_cur[q * 64 + lane] = lane;
      t_cur[q * 64 + 32 + lane] = 32 + lane;
      __syncwarp();
      for (int k = 0; k < np; k++) {
        const int bsl = __shfl_sync(0xffffffffu, slot, k);
        if (lane == 0) { const int t = t_cur[q * 64 + k]; t_cur[q * 64 + k] = t_cur[q * 64 + bsl]; t_cur[q * 64 + bsl] = t; }
      }
      __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
      for (int q = 0; q < nb; q++) {
        const int src_lo = t_cur[q * 64 + lane], src_hi = t_cur[q * 64 + 32 + lane];
        const bool mv_lo = src_lo != lane, mv_hi = src_hi != 32 + lane;
        int v_lo = 0, v_hi = 0;
        if (mv_lo) v_lo = w[t_rows[q * 64 + src_lo]];
        if (mv_hi) v_hi = w[t_rows[q * 64 + src_hi]];
        __syncwarp();                                       // every lane has read before any lane overwrites
        if (mv_lo) w[t_rows[q * 64 + lane]] = v_lo;
        if (mv_hi) w[t_rows[q * 64 + 32 + lane]] = v_hi;
        __syncwarp();
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += PERM_THREADS) perm[i] = w[i];
}
// out[i,:] = in[perm[i],:]  (inverse=0)   or   out[perm[i],:] = in[i,:]  (inverse=1)
template <typename T>
__global__ void permute_rows_kernel(int inverse, int n, int nrhs, const int *__restrict__ perm, const T *__restrict__ in,
                                    int ldi, T *__restrict__ out, int ldo) {
  for (int j = blockIdx.y; j < nrhs; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      int p = perm[i];
      if (inverse) out[(size_t)p + (size_t)j * ldo] = in[(size_t)i + (size_t)j * ldi];
      else out[(size_t)i + (size_t)j * ldo] = in[(size_t)p + (size_t)j * ldi];
    }
}

// ------------------------------------------------------------------ host-side drivers
template <typename T> struct PanelMaxThreads { static constexpr int value = sizeof(T) == 16 ? 256 : 512; };
struct PanelCtx {
  int variant = 0;            // 0: rows-in-registers panel kernel, 1: the shared-memory panel kernels
  int panel_rows = 0;         // rows (= threads) per CTA of the cooperative register panel; 0 = 256, or 512 where that makes <= 32 CTAs of > 32
  int coop_min_rows = 1024;   // shared-memory kernels: from this many rows on, the multi-CTA one
  PanelScratch sc{};
  bool have_sc = false;
  PanelLLScratch ll{};
  bool have_ll = false;
};
// Scoped thread-local CTA cap for the persistent DMMA GEMMs (restored on every exit path).
struct GemmCapScope {
  int saved;
  explicit GemmCapScope(int cap) : saved(tls().gemm_cta_cap) { if (cap > 0) tls().gemm_cta_cap = cap; }
  ~GemmCapScope() { tls().gemm_cta_cap = saved; }
};
// Second stream + events of a look-ahead factorisation, released on every exit path (destruction of a stream or
// event with work pending is deferred by the runtime until that work has drained).
struct LookaheadStreams {
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_first = nullptr, ev_panel = nullptr;
  int open() {
    int lo = 0, hi = 0;
    JB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    JB_CUDA(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi));
    JB_CUDA(cudaEventCreateWithFlags(&ev_first, cudaEventDisableTiming));
    JB_CUDA(cudaEventCreateWithFlags(&ev_panel, cudaEventDisableTiming));
    return 0;
  }
  ~LookaheadStreams() {
    if (ev_first) cudaEventDestroy(ev_first);
    if (ev_panel) cudaEventDestroy(ev_panel);
    if (s2) cudaStreamDestroy(s2);
  }
};
// Tags of the flagged-word exchange must never repeat on memory a previous launch may have left behind.
static unsigned next_panel_tag() {
  static std::atomic<unsigned> seq{1};
  return seq.fetch_add(1) * 64u + 1u;
}

// Rows per CTA of the cooperative register panel for a panel of pm rows (0: not the cooperative register panel).
template <typename T>
static int panel_rows_for(int pm, const PanelCtx &ctx, int nsm) {
  const bool wide_ok = sizeof(T) < 16 && (pm + PANEL_WIDE_THREADS - 1) / PANEL_WIDE_THREADS <= PANEL_WIDE_MAX_CTAS;
  if (wide_ok && (ctx.panel_rows == PANEL_WIDE_THREADS || (ctx.panel_rows == 0 && pm > PANEL_WIDE_MAX_CTAS * PANEL_MULTI_THREADS)))
    return PANEL_WIDE_THREADS;
  int nt = ctx.panel_rows <= 0 || ctx.panel_rows > PANEL_MULTI_THREADS ? PANEL_MULTI_THREADS : ctx.panel_rows;
  while ((pm + nt - 1) / nt > nsm && nt < PANEL_MULTI_THREADS) nt *= 2;
  return nt;
}

// One level of the right-looking LU with NB = 32 panels on the m x n block at A, whose first row is row
// `row_base` of the caller's matrix (ipiv / info are written in the caller's numbering; d_ipiv is indexed locally).
template <typename T>
static int getrf_nb32(int m, int n, T *A, int lda, int *d_ipiv, int *d_info, int row_base, const PanelCtx &ctx, cudaStream_t s) {
  const int mn = m < n ? m : n;
  const size_t smem_cap = 200 * 1024;
  const int nsm = num_sms();
  PanelScratch sc = ctx.sc;
  const bool have_scratch = ctx.have_sc;
  const int coop_min_rows = ctx.coop_min_rows;
  constexpr int MAXT = PanelMaxThreads<T>::value;
  for (int j0 = 0; j0 < mn; j0 += NB) {
    const int jb = (mn - j0 < NB) ? (mn - j0) : NB;
    const int pm = m - j0;
    // rows-in-registers panel: one CTA up to MAXT rows, a cooperative grid with flagged-word exchange above
    int reg_nt = 0, reg_G = 0;
    if (ctx.variant == 0) {
      if (pm <= MAXT) { reg_nt = ((pm + 31) / 32) * 32; reg_G = 1; }
      else if (ctx.have_ll) {
        reg_nt = panel_rows_for<T>(pm, ctx, nsm);
        reg_G = (pm + reg_nt - 1) / reg_nt;
        if (reg_G > nsm || reg_G < 2) reg_G = 0;      // taller than 148 x 256 rows: the shared-memory kernels
      }
    }
    if (reg_G > 0) {
      T *Pp = A + j0 + (size_t)j0 * lda;
      int *pv = d_ipiv + j0;
      int ldp = lda, roff = row_base + j0, jbb = jb, pmm = pm;
      PanelLLScratch ll = ctx.ll;
      unsigned tag0 = next_panel_tag();
      const size_t shm = (size_t)NB * reg_nt * sizeof(T);
      if (reg_G == 1) {
        lu_panel_reg_kernel<T, MAXT, false><<<1, reg_nt, shm, s>>>(pmm, jbb, Pp, ldp, pv, d_info, roff, ll, tag0);
      } else {
        void *args[] = {&pmm, &jbb, &Pp, &ldp, &pv, &d_info, &roff, &ll, &tag0};
        void *fn = (void *)lu_panel_reg_kernel<T, PANEL_MULTI_THREADS, true>;
        if constexpr (sizeof(T) < 16)
          if (reg_nt == PANEL_WIDE_THREADS) fn = (void *)lu_panel_reg_kernel<T, PANEL_WIDE_THREADS, true, 1>;
        JB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(reg_G), dim3(reg_nt), args, shm, s));
      }
      count_launch();
    } else if (have_scratch && pm >= coop_min_rows) {
      int rows_per_cta = 128;
      int G = (pm + rows_per_cta - 1) / rows_per_cta;
      if (G > nsm) { G = nsm; rows_per_cta = ((pm + G - 1) / G + 31) / 32 * 32; G = (pm + rows_per_cta - 1) / rows_per_cta; }
      size_t shm = (size_t)(rows_per_cta | 1) * jb * sizeof(T);
      T *Pp = A + j0 + (size_t)j0 * lda;
      int *pv = d_ipiv + j0;
      int ldp = lda, roff = row_base + j0, jbb = jb, pmm = pm;
      void *args[] = {&pmm, &jbb, &Pp, &ldp, &pv, &d_info, &roff, &rows_per_cta, &sc};
      JB_CUDA(cudaLaunchCooperativeKernel((void *)lu_panel_coop_kernel<T>, dim3(G), dim3(256), args, shm, s));
      count_launch();
    } else {
      size_t need = (size_t)(pm | 1) * jb * sizeof(T);
      int use_smem = need <= smem_cap;
      int threads = pm >= PANEL_THREADS ? PANEL_THREADS : ((pm + 31) / 32) * 32;
      if (threads < 64) threads = 64;
      lu_panel_kernel<T><<<1, threads, use_smem ? need : 0, s>>>(pm, jb, A + j0 + (size_t)j0 * lda, lda, d_ipiv + j0, d_info,
                                                                row_base + j0, use_smem);
      count_launch();
    }
    // The columns right of the panel take its interchanges and the block-row solve in one launch; the columns left
    // of it (finished L, read by nothing later in this sweep) take all their interchanges in one pass at the end.
    const int nr = n - j0 - jb;
    // (A fused interchange + block-row solve + rank-32 update kernel for small trailing matrices was tried in round 2d:
    // two-warp CTAs of eight columns, L21 streamed from L2 — dgesv 512 went from 0.93 to 1.11 ms, the update loop is a
    // chain of L2 round trips without enough warps to hide them; the two launch-bound kernels below stayed.)
    if (nr > 0) {
      const int cols_per_cta = (LASWP_THREADS / 32) * LASWP_COLS_PER_WARP;
      laswp_trsm_kernel<T><<<(nr + cols_per_cta - 1) / cols_per_cta, LASWP_THREADS, 0, s>>>(A, lda, j0 + jb, n, d_ipiv, j0, j0 + jb,
                                                                                         row_base);
      count_launch();
      const int mr = m - j0 - jb;
      if (mr > 0) {
        int rc = gemm_device<T>(111, 111, mr, nr, jb, Num<T>::neg(Num<T>::one()), A + (j0 + jb) + (size_t)j0 * lda, lda,
                                A + j0 + (size_t)(j0 + jb) * lda, lda, Num<T>::one(),
                                A + (j0 + jb) + (size_t)(j0 + jb) * lda, lda, s, 0);
        if (rc) return rc;
      }
    }
  }
  if (mn > NB)
    if (int rc = laswp_launch<T>(A, lda, 0, ((mn - 1) / NB) * NB, d_ipiv, NB, mn, row_base, s, 1)) return rc;
  return 0;
}

template <typename T>
static int trsm_left_blocked(int stored_lower, int trans, int unit, int n, int nrhs, const T *A, int lda, T *B, int ldb,
                             cudaStream_t s);

// Middle level of the blocked LU: an outer block of n (<= NB2) columns is factorised in sub-blocks of `nbm` columns.  The
// rank-32 updates of getrf_nb32 then stay inside a sub-block, and the rest of the outer block takes ONE rank-nbm update
// per sub-block: the skinny updates, which only stream C (rows x remaining columns, read and written per 32 columns),
// move 2.5x fewer bytes at nbm = 128, NB2 = 512 — they were half of the look-ahead chain's duration
// (profiles/trace_getrf16384_r2c.txt).  An entry's update is now accumulated nbm products at a time instead of 32 (same k
// order, different rounding points), like any change of block size.  Numbering as getrf_nb32: A points at the block, `row_base` is its first row
// in the caller's matrix, d_ipiv is indexed locally and holds caller-numbered rows.
template <typename T>
static int getrf_mid(int m, int n, T *A, int lda, int *d_ipiv, int *d_info, int row_base, const PanelCtx &ctx, int nbm, cudaStream_t s) {
  const int mn = m < n ? m : n;
  if (nbm < 2 * NB || mn <= nbm + NB) return getrf_nb32<T>(m, n, A, lda, d_ipiv, d_info, row_base, ctx, s);
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  for (int j0 = 0; j0 < mn; j0 += nbm) {
    const int jb = (mn - j0 < nbm) ? (mn - j0) : nbm;
    if (int rc = getrf_nb32<T>(m - j0, jb, A + j0 + (size_t)j0 * lda, lda, d_ipiv + j0, d_info, row_base + j0, ctx, s)) return rc;
    const int c0 = j0 + jb;
    if (c0 < n) {                // columns of the block right of the sub-block: interchanges, block-row solve, rank-jb update
      if (int rc = laswp_launch<T>(A, lda, c0, n, d_ipiv, j0, c0, row_base, s)) return rc;
      if (int rc = trsm_left_blocked<T>(1, 111, 1, jb, n - c0, A + j0 + (size_t)j0 * lda, lda, A + j0 + (size_t)c0 * lda, lda, s)) return rc;
      if (m - c0 > 0)
        if (int rc = gemm_device<T>(111, 111, m - c0, n - c0, jb, minus1, A + c0 + (size_t)j0 * lda, lda, A + j0 + (size_t)c0 * lda, lda, one,
                                    A + c0 + (size_t)c0 * lda, lda, s, 0)) return rc;
    }
  }
  // the finished sub-blocks on the left take the later sub-blocks' interchanges in one pass per sub-block boundary
  for (int j0 = nbm; j0 < mn; j0 += nbm) {
    const int jb = (mn - j0 < nbm) ? (mn - j0) : nbm;
    if (int rc = laswp_launch<T>(A, lda, 0, j0, d_ipiv, j0, j0 + jb, row_base, s)) return rc;
  }
  return 0;
}

// Debug timeline of a look-ahead factorisation (options getrf.trace / potrf.trace): timing events at the four points of
// every outer step, printed to stderr after a final synchronisation.  Off: no events are created.
struct StepTrace {
  bool on = false;
  std::vector<cudaEvent_t> ev;      // per step: start (s), first block updated (s), chain end (s2), update end (s)
  std::vector<int> tag;
  void mark(cudaStream_t st, int t) {
    if (!on) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ev.push_back(e); tag.push_back(t);
  }
  void print(const char *what, cudaStream_t s) {
    if (!on || ev.empty()) return;
    cudaStreamSynchronize(s);
    double sum[3] = {0, 0, 0}, crit = 0;
    for (size_t i = 0; i + 3 < ev.size(); i += 4) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, ev[i], ev[i + 1]);
      cudaEventElapsedTime(&b, ev[i + 1], ev[i + 2]);
      cudaEventElapsedTime(&c, ev[i + 1], ev[i + 3]);
      fprintf(stderr, "[%s.trace] J0=%5d first %.3f ms | chain %.3f | rest update %.3f\n", what, tag[i], a, b, c);
      sum[0] += a; sum[1] += b; sum[2] += c; crit += a + (b > c ? b : c);
    }
    fprintf(stderr, "[%s.trace] sums: first %.2f ms, chain %.2f, rest update %.2f, critical path %.2f\n", what, sum[0], sum[1], sum[2], crit);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear(); tag.clear();
  }
};

// Two-level right-looking LU: outer panels of NB2 columns are factorised by the NB=32 routine (tall panels on the
// cooperative multi-CTA kernel); the interchanges, the unit-lower block-row solve and ONE rank-NB2 GEMM update per
// outer step then touch the rest of the matrix — NB2/32 times fewer sweeps over the trailing matrix.
template <typename T>
int getrf_device(int m, int n, T *A, int lda, int *d_ipiv, int *d_info, cudaStream_t s) {
  const int mn = m < n ? m : n;
  if (mn <= 0) return 0;
  const size_t smem_cap = 200 * 1024;
  JB_CUDA(cudaFuncSetAttribute(lu_panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
  const int nsm = num_sms();
  constexpr int MAXT = PanelMaxThreads<T>::value;
  JB_CUDA(cudaFuncSetAttribute(lu_panel_reg_kernel<T, MAXT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)((size_t)NB * MAXT * sizeof(T))));
  JB_CUDA(cudaFuncSetAttribute(lu_panel_reg_kernel<T, PANEL_MULTI_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)((size_t)NB * PANEL_MULTI_THREADS * sizeof(T))));
  if constexpr (sizeof(T) < 16)
    JB_CUDA(cudaFuncSetAttribute(lu_panel_reg_kernel<T, PANEL_WIDE_THREADS, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)NB * PANEL_WIDE_THREADS * sizeof(T))));
  constexpr int W = sizeof(T) / 4;
  PanelCtx ctx;
  ctx.variant = option("getrf.panel_variant") > 0 ? 1 : 0;
  if (option("getrf.panel_rows") > 0) ctx.panel_rows = (option("getrf.panel_rows") + 31) / 32 * 32;      // 512: the wide instantiation wherever it fits
  if (option("getrf.coop_min_rows") > 0) ctx.coop_min_rows = option("getrf.coop_min_rows");
  unsigned char *scratch = nullptr, *scratch_ll = nullptr;
  if (ctx.variant == 0 && m > MAXT) {
    constexpr int PW = 2 * NB * W;        // words per row packet
    const size_t words = 2 * ((size_t)nsm * 4 + (size_t)nsm * PW + (size_t)PW);
    if (int rc = ws_alloc((void **)&scratch_ll, words * sizeof(unsigned long long), s)) return rc;
    JB_CUDA(cudaMemsetAsync(scratch_ll, 0, words * sizeof(unsigned long long), s));
    ctx.ll.rec = (unsigned long long *)scratch_ll;
    ctx.ll.cand = ctx.ll.rec + 2 * (size_t)nsm * 4;
    ctx.ll.diag = ctx.ll.cand + 2 * (size_t)nsm * PW;
    ctx.have_ll = true;
  }
  long long *prof = nullptr;
  if (option("getrf.panel_prof") > 0) {
    if (int rc = ws_alloc((void **)&prof, 8 * sizeof(long long), s)) return rc;
    JB_CUDA(cudaMemsetAsync(prof, 0, 8 * sizeof(long long), s));
    ctx.ll.prof = prof;
  }
  if (m >= ctx.coop_min_rows && (ctx.variant == 1 || m > (long)nsm * PANEL_MULTI_THREADS)) {
    const size_t per = 2 * ((size_t)nsm * (sizeof(double) + sizeof(int) + NB * sizeof(T)) + NB * sizeof(T)) + 256;
    if (int rc = ws_alloc((void **)&scratch, per, s)) { ws_free(scratch_ll, s); return rc; }
    unsigned char *q = scratch;
    ctx.sc.cand_val = (double *)q; q += 2 * (size_t)nsm * sizeof(double);
    ctx.sc.cand_row = (int *)q; q += 2 * (size_t)nsm * sizeof(int);
    q = (unsigned char *)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    ctx.sc.cand_data = q; q += 2 * (size_t)nsm * NB * sizeof(T);
    ctx.sc.diag_data = q;
    ctx.have_sc = true;
    JB_CUDA(cudaFuncSetAttribute(lu_panel_coop_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
  }
  int NB2 = option("getrf.nb2");
  // measured (tools/time_nb2.py): 16384: 212.8 / 185.7 / 178.2 / 175.4 ms for 128 / 256 / 384 / 512; 8192: 46.6 / 43.9 / 43.6 / 43.2
  // with look-ahead (tools/time_nb2_potrf.py): 4096: 13.8 / 14.5 / 14.6 ms for 128 / 256 / 512; 8192: 39.7 / 37.3 / 36.9;
  // 16384: 181.7 / 165.2 / 158.5; 2048 stays single level (5.6 ms against 6.5)
  if (NB2 <= 0) NB2 = (mn >= 8192) ? 512 : (mn >= 4096 ? 128 : mn);
  // sub-block width inside an outer block (getrf_mid); 0 = none.  Measured (profiles/lookahead_r2c.jsonl): no gain for LU —
  // 16384: 145.1 / 143.4 ms at 128 / 256 against 144.4 without, 8192: 36.0 / 34.7 against 33.9 (the chain is the
  // cooperative panel kernel itself, and every sub-block adds an interchange and a solve launch) — off by default.
  int nbm = option("getrf.nbm");
  if (nbm < 0) nbm = 0;
  int rc = 0;
  if (mn <= NB2 + NB) {
    rc = getrf_nb32<T>(m, n, A, lda, d_ipiv, d_info, 0, ctx, s);
  } else {
    const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
    // Look-ahead: the update of outer step J is split into the columns of block J+1 and the rest.  As soon as the
    // first part is done, block J+1 is factorised on a second (higher-priority) stream while the main stream
    // finishes the trailing update — the two touch disjoint columns.  The persistent DMMA GEMM owns every SM it
    // runs on, so during the overlap it is capped to leave the panel kernels their CTAs (thread-local cap).
    const bool lookahead = option("getrf.lookahead") != 0;
    LookaheadStreams la;
    if (lookahead) if (int r = la.open()) { ws_free(scratch, s); ws_free(scratch_ll, s); return r; }
    cudaStream_t s2 = la.s2;
    cudaEvent_t ev_first = la.ev_first, ev_panel = la.ev_panel;
    // right-of-block work of outer step (J0, jb2) on columns [c0, c1): interchanges + U12 solve, then the rank-jb2 update
    auto right_solve = [&](int J0, int jb2, int c0, int c1, cudaStream_t st) -> int {
      if (c1 <= c0) return 0;
      int r = laswp_launch<T>(A, lda, c0, c1, d_ipiv, J0, J0 + jb2, 0, st);
      if (!r) r = trsm_left_blocked<T>(1, 111, 1, jb2, c1 - c0, A + J0 + (size_t)J0 * lda, lda, A + J0 + (size_t)c0 * lda, lda, st);
      return r;
    };
    auto right_gemm = [&](int J0, int jb2, int c0, int c1, cudaStream_t st) -> int {
      const int mr = m - J0 - jb2;
      if (c1 <= c0 || mr <= 0) return 0;
      return gemm_device<T>(111, 111, mr, c1 - c0, jb2, minus1, A + (J0 + jb2) + (size_t)J0 * lda, lda,
                            A + J0 + (size_t)c0 * lda, lda, one, A + (J0 + jb2) + (size_t)c0 * lda, lda, st, 0);
    };
    auto right_update = [&](int J0, int jb2, int c0, int c1, cudaStream_t st) -> int {
      int r = right_solve(J0, jb2, c0, c1, st);
      return r ? r : right_gemm(J0, jb2, c0, c1, st);
    };
    bool factored = false;                      // block at J0 already factorised by the previous step's look-ahead
    StepTrace tr;
    tr.on = lookahead && option("getrf.trace") > 0;
    for (int J0 = 0; J0 < mn && !rc; J0 += NB2) {
      const int jb2 = (mn - J0 < NB2) ? (mn - J0) : NB2;
      tr.mark(s, J0);
      if (!factored)
        rc = getrf_mid<T>(m - J0, jb2, A + J0 + (size_t)J0 * lda, lda, d_ipiv + J0, d_info, J0, ctx, nbm, s);
      factored = false;
      if (rc) break;
      const int c0 = J0 + jb2;
      const int jn = (c0 < mn) ? ((mn - c0 < NB2) ? (mn - c0) : NB2) : 0;      // width of the next block
      const int mr = m - c0;
      const bool ahead = lookahead && jn > 0 && mr > 0 && n - c0 > 0;
      // the interchanges of the finished columns on the left are read by nobody in this sweep: behind the chain's start
      if (J0 > 0 && !ahead) rc = laswp_launch<T>(A, lda, 0, J0, d_ipiv, J0, J0 + jb2, 0, s);
      if (rc) break;
      if (n - c0 <= 0) break;
      if (ahead) {
        rc = right_update(J0, jb2, c0, c0 + jn, s);
        if (rc) break;
        JB_CUDA_BREAK(cudaEventRecord(ev_first, s));
        JB_CUDA_BREAK(cudaStreamWaitEvent(s2, ev_first, 0));
        tr.mark(s, J0);
        if (J0 > 0) rc = laswp_launch<T>(A, lda, 0, J0, d_ipiv, J0, J0 + jb2, 0, s);
        if (rc) break;
        // CTAs the panel phase needs at once: the cooperative panel grid or one CTA, + room for its interchange /
        // small-solve / small-GEMM launches.  The persistent GEMM of the trailing update is capped to leave them free —
        // but only for as many columns as the panel chain is expected to last (model below); the remaining columns run
        // as a second, uncapped launch.  (Capping the whole update costs need/148 of ALL of it: reserving 70 SMs at
        // n = 16384 lost more GEMM time than the panels hid, 175 -> 201 ms.)
        // (The chain's own rank-32 updates stay uncapped: their CTAs run in waves over whatever SMs are free and spread
        // out as soon as the trailing update ends; capping them to the reserved SMs measured 0-1 ms slower.)
        constexpr int MAXT = PanelMaxThreads<T>::value;
        const int prow = panel_rows_for<T>(mr, ctx, nsm);
        int need = (mr > MAXT ? (mr + prow - 1) / prow : 1) + 8;
        int max_need = option("getrf.la_max_need");
        if (max_need <= 0) max_need = 80;
        const int cA = c0 + jn;
        const bool overlap = need <= max_need && nsm - need >= 32;
        int scale = option("getrf.la_chain_pct");
        if (scale <= 0) scale = 115;
        // chain of the next block: jn/32 steps of (32 columns at ~2 us (cooperative) or ~1 us (one CTA), the fused
        // interchange/solve launch, a rank-32 update streaming mr x jn/2 entries)
        const double t_chain = (jn / (double)NB) * (NB * (mr > MAXT ? 2.0 : 1.0) + 15.0 + 0.002 * mr) * scale * 0.01;
        const double rate = (sizeof(T) == 4 || sizeof(T) == 8 && Num<T>::cplx ? 0.36e6 : 0.225e6);    // flop per us and SM on the update
        const double flop_col = (Num<T>::cplx ? 8.0 : 2.0) * (double)(m - c0) * jb2;
        const int chain_sms = need;
        rc = getrf_mid<T>(m - c0, jn, A + c0 + (size_t)c0 * lda, lda, d_ipiv + c0, d_info, c0, ctx, nbm, s2);
        if (rc) break;
        JB_CUDA_BREAK(cudaEventRecord(ev_panel, s2));
        tr.mark(s2, J0);
        if (overlap) {
          long ncap = (long)(t_chain * rate * (nsm - chain_sms) / flop_col);
          ncap = (ncap + 127) / 128 * 128;
          if (ncap > n - cA - 256) ncap = n - cA;           // no sliver launches
          {
            GemmCapScope cap(nsm - chain_sms);
            rc = right_solve(J0, jb2, cA, n, s);
            if (!rc) rc = right_gemm(J0, jb2, cA, cA + (int)ncap, s);
          }
          if (!rc) rc = right_gemm(J0, jb2, cA + (int)ncap, n, s);
        } else {
          rc = right_update(J0, jb2, cA, n, s);
        }
        tr.mark(s, J0);
        JB_CUDA_BREAK(cudaStreamWaitEvent(s, ev_panel, 0));
        factored = true;
      } else {
        rc = right_update(J0, jb2, c0, n, s);
        if (tr.on) { tr.mark(s, J0); tr.mark(s, J0); tr.mark(s, J0); }
      }
    }
    if (lookahead) {              // whatever path left the loop: nothing of the second stream outlives the call's stream order
      cudaEventRecord(ev_panel, s2);
      cudaStreamWaitEvent(s, ev_panel, 0);
    }
    tr.print("getrf", s);
  }
  if (prof) {
    long long h[8];
    cudaMemcpyAsync(h, prof, sizeof(h), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    fprintf(stderr, "[getrf.panel_prof] n=%d cycles of CTA 0 summed over %d columns: search+B1 %lld | pick+copy+B1b %lld | publish %lld | "
            "poll records %lld | winner row %lld | B2 %lld | interchange+update %lld\n", mn, mn, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    ws_free(prof, s);
  }
  ws_free(scratch, s);
  ws_free(scratch_ll, s);
  if (rc) return rc;
  JB_CUDA(cudaGetLastError());
  return 0;
}

// Blocked solve op(Tri) X = B in place; Tri is the n x n triangle stored in A
// (stored_lower selects which triangle of A holds it; trans applies op).
template <typename T>
static int trsm_left_blocked(int stored_lower, int trans, int unit, int n, int nrhs, const T *A, int lda, T *B, int ldb,
                             cudaStream_t s) {
  if (trsm_inv_pays(n, nrhs)) return trsm_inv_device<T>(141, stored_lower, trans, unit, n, nrhs, A, lda, B, ldb, s);
  if (trsv_few_pays(n, nrhs)) return trsv_few_device<T>(stored_lower, trans, unit, n, nrhs, A, lda, B, ldb, s);
  const int eff_lower = (trans == 111) ? stored_lower : !stored_lower;
  const int nblk = (n + NB - 1) / NB;
  int blocks = (nrhs + 7) / 8; if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  for (int bi = 0; bi < nblk; bi++) {
    const int blk = eff_lower ? bi : (nblk - 1 - bi);
    const int j0 = blk * NB, jb = (n - j0 < NB) ? (n - j0) : NB;
    trsm_small_left_kernel<T><<<blocks, 256, 0, s>>>(eff_lower, unit, trans, jb, nrhs, A + j0 + (size_t)j0 * lda, lda,
                                                     B + j0, ldb, nullptr);
    count_launch();
    int rc = 0;
    if (eff_lower) {
      const int rem = n - j0 - jb;
      if (rem > 0) {
        // B[j0+jb:, :] -= op(A)[j0+jb:, j0:j0+jb] * X
        const T *Ab = (trans == 111) ? A + (j0 + jb) + (size_t)j0 * lda : A + j0 + (size_t)(j0 + jb) * lda;
        rc = gemm_device<T>(trans, 111, rem, nrhs, jb, minus1, Ab, lda, B + j0, ldb, one, B + j0 + jb, ldb, s, 0);
      }
    } else if (j0 > 0) {
      // B[0:j0, :] -= op(A)[0:j0, j0:j0+jb] * X
      const T *Ab = (trans == 111) ? A + (size_t)j0 * lda : A + j0;
      rc = gemm_device<T>(trans, 111, j0, nrhs, jb, minus1, Ab, lda, B + j0, ldb, one, B, ldb, s, 0);
    }
    if (rc) return rc;
  }
  JB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int apply_row_perm(int inverse, int n, int nrhs, const int *d_ipiv, T *B, int ldb, cudaStream_t s) {
  int *perm = nullptr;
  T *W = nullptr;
  if (int rc = ws_alloc((void **)&perm, sizeof(int) * (size_t)n, s)) return rc;
  if (int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)n * nrhs, s)) { ws_free(perm, s); return rc; }
  // the swap chain is n dependent steps: keep it in shared memory (opt-in size) up to n = 25600 before falling back
  // to global memory, where every step is two dependent round trips
  const int perm_smem = (size_t)n * 2 * sizeof(int) <= 200 * 1024;
  const size_t chunked_smem = ((size_t)n + 2 * PERM_BATCH * 64) * sizeof(int);
  if (n >= 256 && chunked_smem <= 200 * 1024 && option("getrs.perm_serial") <= 0) {     // n <= 43008
    JB_CUDA(cudaFuncSetAttribute(perm_from_ipiv_chunked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    perm_from_ipiv_chunked_kernel<<<1, PERM_THREADS, chunked_smem, s>>>(n, d_ipiv, perm);
  } else {
    if (perm_smem)
      JB_CUDA(cudaFuncSetAttribute(perm_from_ipiv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    perm_from_ipiv_kernel<<<1, 256, perm_smem ? (size_t)n * 2 * sizeof(int) : 0, s>>>(n, d_ipiv, perm, perm_smem);
  }
  dim3 grid((unsigned)min((n + 255) / 256, 1024), (unsigned)min(nrhs, 4096));
  permute_rows_kernel<T><<<grid, 256, 0, s>>>(inverse, n, nrhs, perm, B, ldb, W, n);
  count_launch(2);
  int rc = lacpy_device<T>(111, n, nrhs, W, n, B, ldb, s);
  ws_free(W, s);
  ws_free(perm, s);
  return rc;
}

// ------------------------------------------------------------------ getrs for a few right-hand sides
// The blocked path above costs ~4 launches per 32 rows whatever nrhs is (dgesv 512 with one right-hand side spent
// 0.48 ms in it).  For a handful of right-hand sides one CTA per column does the whole solve: the column sits in
// shared memory, thread 0 replays the interchanges on it, then 32-row blocks go forward through unit-lower L and back
// through U — warp 0 solves the diagonal block (lane = row, the solved component broadcast by shuffle), all warps
// subtract the block's contribution from the remaining rows with the factor streamed from L2, and the next diagonal
// block is fetched while they do.  Same operations as ?getrs (x_i * (1/u_ii) like the blocked path).
constexpr int GETRS_FUSED_THREADS = 512;
template <typename T>
__global__ void __launch_bounds__(GETRS_FUSED_THREADS)
getrs_fused_kernel(int n, const T *__restrict__ A, int lda, const int *__restrict__ ipiv, T *__restrict__ B, int ldb, int upper_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *x = reinterpret_cast<T *>(smem_raw);
  int *piv = reinterpret_cast<int *>(x + n);
  __shared__ T Ms[NB][NB + 1];
  __shared__ T xb[NB];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  T *b = B + (size_t)blockIdx.x * ldb;
  for (int i = tid; i < n; i += nt) { x[i] = b[i]; piv[i] = upper_only ? i + 1 : ipiv[i]; }
  // diagonal block (r, c) of the factor at block offset j0, zero outside the matrix; two entries per thread
  auto fetch = [&](int j0, T &e0, T &e1) {
    const int r0 = tid & 31, c0 = tid >> 5, c1 = c0 + 16;
    e0 = (j0 + r0 < n && j0 + c0 < n) ? A[(j0 + r0) + (size_t)(j0 + c0) * lda] : Num<T>::zero();
    e1 = (j0 + r0 < n && j0 + c1 < n) ? A[(j0 + r0) + (size_t)(j0 + c1) * lda] : Num<T>::zero();
  };
  auto stash = [&](T e0, T e1) { Ms[tid & 31][tid >> 5] = e0; Ms[tid & 31][(tid >> 5) + 16] = e1; };
  T e0, e1;
  fetch(0, e0, e1);
  __syncthreads();
  if (tid == 0 && !upper_only)
    for (int k = 0; k < n; k++) {
      const int p = piv[k] - 1;
      if (p != k && p >= 0 && p < n) { const T t = x[k]; x[k] = x[p]; x[p] = t; }
    }
  stash(e0, e1);
  __syncthreads();
  const int nblk = (n + NB - 1) / NB;
  // ---- forward: L y = P b, unit diagonal (skipped when the caller already holds y: gesv on the augmented matrix)
  for (int bi = 0; bi < (upper_only ? 0 : nblk); bi++) {
    const int j0 = bi * NB, jb = min(NB, n - j0);
    if (warp == 0) {
      T xl = lane < jb ? x[j0 + lane] : Num<T>::zero();
#pragma unroll
      for (int i = 0; i < NB; i++) {
        const T xi = shfl_T<T>(xl, i);
        if (lane > i) xl = Num<T>::msub(xl, Ms[lane][i], xi);
      }
      if (lane < jb) x[j0 + lane] = xl;
      xb[lane] = lane < jb ? xl : Num<T>::zero();
    }
    if (bi + 1 < nblk) fetch(j0 + NB, e0, e1);
    __syncthreads();
    for (int r = j0 + NB + tid; r < n; r += nt) {
      T acc = x[r];
      const T *Ar = A + r + (size_t)j0 * lda;
      T av[NB];                                   // all 32 entries of the row in flight at once: one L2 round trip
#pragma unroll
      for (int i = 0; i < NB; i++) av[i] = Ar[(size_t)i * lda];
#pragma unroll
      for (int i = 0; i < NB; i++) acc = Num<T>::msub(acc, av[i], xb[i]);
      x[r] = acc;
    }
    if (bi + 1 < nblk) stash(e0, e1);
    __syncthreads();
  }
  // ---- backward: U x = y
  fetch((nblk - 1) * NB, e0, e1);
  stash(e0, e1);
  __syncthreads();
  for (int bi = nblk - 1; bi >= 0; bi--) {
    const int j0 = bi * NB, jb = min(NB, n - j0);
    if (warp == 0) {
      T xl = lane < jb ? x[j0 + lane] : Num<T>::zero();
      const T rinv = lane < jb ? Num<T>::recip(Ms[lane][lane]) : Num<T>::zero();
#pragma unroll
      for (int i = NB - 1; i >= 0; i--) {
        const T xi = shfl_T<T>(Num<T>::cmul(xl, rinv), i);
        if (lane == i) xl = xi;
        else if (lane < i) xl = Num<T>::msub(xl, Ms[lane][i], xi);
      }
      if (lane < jb) x[j0 + lane] = xl;
      xb[lane] = lane < jb ? xl : Num<T>::zero();
    }
    if (bi > 0) fetch(j0 - NB, e0, e1);
    __syncthreads();
    for (int r = tid; r < j0; r += nt) {
      T acc = x[r];
      const T *Ar = A + r + (size_t)j0 * lda;
      if (jb == NB) {
        T av[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) av[i] = Ar[(size_t)i * lda];
#pragma unroll
        for (int i = 0; i < NB; i++) acc = Num<T>::msub(acc, av[i], xb[i]);
      } else {
        for (int i = 0; i < jb; i++) acc = Num<T>::msub(acc, Ar[(size_t)i * lda], xb[i]);
      }
      x[r] = acc;
    }
    if (bi > 0) stash(e0, e1);
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) b[i] = x[i];
}

template <typename T>
int getrs_device(char trans, int n, int nrhs, const T *A, int lda, const int *d_ipiv, T *B, int ldb, cudaStream_t s) {
  if (n <= 0 || nrhs <= 0) return 0;
  int rc;
  if (trans == 'N' || trans == 'n') {
    // options: getrs.fused_max_nrhs (unset: 16, 0: never), getrs.fused_max_n (unset: 4096)
    const int fused_nrhs = option("getrs.fused_max_nrhs") >= 0 ? option("getrs.fused_max_nrhs") : 16;
    // measured (tools/time_getrs_few.py, profiles/getrs_few_r2c.jsonl; ms at n = 1024 / 2048 / 4096, one right-hand side):
    // fused 0.32 / 0.80 / 5.15, bandwidth-bound sweep (trsv_few, <= 8 columns) 0.27 / 0.52 / 0.69, blocked 0.96 / 1.88 / 4.19
    // one right-hand side from order 1024 on: the persistent pipelined sweep (trsv_pipe_kernel, one launch per triangle;
    // tools/time_trsv_pipe.py: dgetrs 1024 / 2048 / 4096 / 16384 x 1 0.259 / 0.393 / 0.690 / 2.86 -> 0.234 / 0.331 / 0.602 / 2.47 ms,
    // at 512 the two block inversions it needs cost more than the fused kernel's whole solve: 0.194 against 0.157)
    const bool pipe1 = nrhs == 1 && n >= 1024 && trsv_few_pays(n, 1) && option("trsv.pipe") != 0;
    const int fused_n = option("getrs.fused_max_n") > 0 ? option("getrs.fused_max_n") : (pipe1 ? 255 : (trsv_few_pays(n, nrhs) ? 768 : 2048));
    if (nrhs <= fused_nrhs && n <= fused_n) {
      const size_t shm = (size_t)n * (sizeof(T) + sizeof(int));
      JB_CUDA(cudaFuncSetAttribute(getrs_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
      getrs_fused_kernel<T><<<nrhs, GETRS_FUSED_THREADS, shm, s>>>(n, A, lda, d_ipiv, B, ldb, 0);
      count_launch();
      JB_CUDA(cudaGetLastError());
      return 0;
    }
    if ((rc = apply_row_perm<T>(0, n, nrhs, d_ipiv, B, ldb, s))) return rc;
    if ((rc = trsm_left_blocked<T>(1, 111, 1, n, nrhs, A, lda, B, ldb, s))) return rc;
    if ((rc = trsm_left_blocked<T>(0, 111, 0, n, nrhs, A, lda, B, ldb, s))) return rc;
  } else {
    int t = (trans == 'C' || trans == 'c') ? 113 : 112;
    if ((rc = trsm_left_blocked<T>(0, t, 0, n, nrhs, A, lda, B, ldb, s))) return rc;   // U^T y = b
    if ((rc = trsm_left_blocked<T>(1, t, 1, n, nrhs, A, lda, B, ldb, s))) return rc;   // L^T z = y
    if ((rc = apply_row_perm<T>(1, n, nrhs, d_ipiv, B, ldb, s))) return rc;            // x = P^T z
  }
  return 0;
}

// U X = Y with the stored upper factor (the second half of ?getrs): what ?gesv still has to do after factorising the
// augmented matrix [A | B], whose extra columns leave getrf as Y = inv(L) P B.
template <typename T>
int getrs_upper_device(int n, int nrhs, const T *A, int lda, T *B, int ldb, cudaStream_t s) {
  if (n <= 0 || nrhs <= 0) return 0;
  const int fused_nrhs = option("getrs.fused_max_nrhs") >= 0 ? option("getrs.fused_max_nrhs") : 16;
  const bool pipe1 = nrhs == 1 && n >= 1024 && trsv_few_pays(n, 1) && option("trsv.pipe") != 0;
  const int fused_n = option("getrs.fused_max_n") > 0 ? option("getrs.fused_max_n") : (pipe1 ? 255 : (trsv_few_pays(n, nrhs) ? 768 : 2048));
  if (nrhs <= fused_nrhs && n <= fused_n) {
    const size_t shm = (size_t)n * (sizeof(T) + sizeof(int));
    JB_CUDA(cudaFuncSetAttribute(getrs_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    getrs_fused_kernel<T><<<nrhs, GETRS_FUSED_THREADS, shm, s>>>(n, A, lda, nullptr, B, ldb, 1);
    count_launch();
    JB_CUDA(cudaGetLastError());
    return 0;
  }
  return trsm_left_blocked<T>(0, 111, 0, n, nrhs, A, lda, B, ldb, s);
}
template int getrs_upper_device<float>(int, int, const float *, int, float *, int, cudaStream_t);
template int getrs_upper_device<double>(int, int, const double *, int, double *, int, cudaStream_t);
template int getrs_upper_device<cfloat>(int, int, const cfloat *, int, cfloat *, int, cudaStream_t);
template int getrs_upper_device<cdouble>(int, int, const cdouble *, int, cdouble *, int, cudaStream_t);

// inv(A) from the LU factors: solves A X = I with the stored factors and overwrites A.
// (?getri forms inv(U) and solves X L = inv(U); same inverse, different operation order —
// the parity criterion for this row is the residual bound, see tests/test_lapacke_parity.py.)
template <typename T>
int getri_device(int n, T *A, int lda, const int *d_ipiv, int *d_info, cudaStream_t s) {
  if (n <= 0) return 0;
  JB_CUDA(cudaMemsetAsync(d_info, 0x7f, sizeof(int), s));   // large sentinel for atomicMin
  diag_zero_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, A, lda, d_info);
  count_launch();
  int h = 0;
  JB_CUDA(cudaMemcpyAsync(&h, d_info, sizeof(int), cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaStreamSynchronize(s));
  if (h != 0x7f7f7f7f) {          // singular: U(i,i) == 0, A left as the LU factors
    JB_CUDA(cudaMemcpyAsync(d_info, &h, sizeof(int), cudaMemcpyHostToDevice, s));
    return 0;
  }
  JB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), s));
  T *W = nullptr;
  if (int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)n * n, s)) return rc;
  dim3 grid((unsigned)min((n + 255) / 256, 1024), (unsigned)min(n, 4096));
  set_identity_kernel<T><<<grid, 256, 0, s>>>(n, W, n);
  count_launch();
  int rc = getrs_device<T>('N', n, n, A, lda, d_ipiv, W, n, s);
  if (!rc) rc = lacpy_device<T>(111, n, n, W, n, A, lda, s);
  ws_free(W, s);
  return rc;
}

// One level of the blocked Cholesky with NB = 32 on the nn x nn block at A (global index offset `off` for info).
template <typename T>
static int potrf_nb32(int lower, int nn, T *A, int lda, int *d_info, int off, cudaStream_t s) {
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  const int tc = Num<T>::cplx ? 113 : 112;
  for (int j0 = 0; j0 < nn; j0 += NB) {
    const int jb = (nn - j0 < NB) ? (nn - j0) : NB;
    T *Ajj = A + j0 + (size_t)j0 * lda;
    potf2_block_kernel<T><<<1, 1024, 0, s>>>(lower, jb, Ajj, lda, d_info, off + j0);
    count_launch();
    const int rem = nn - j0 - jb;
    if (rem <= 0) break;
    int rc;
    if (lower) {
      T *A21 = A + (j0 + jb) + (size_t)j0 * lda;
      trsm_small_right_upper_kernel<T><<<(rem + 255) / 256, 256, 0, s>>>(113, jb, rem, Ajj, lda, A21, lda, d_info);
      count_launch();
      rc = gemm_device<T>(111, tc, rem, rem, jb, minus1, A21, lda, A21, lda, one, A + (j0 + jb) + (size_t)(j0 + jb) * lda, lda, s, 'L');
    } else {
      T *A12 = A + j0 + (size_t)(j0 + jb) * lda;
      int blocks = (rem + 7) / 8; if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
      trsm_small_left_kernel<T><<<blocks, 256, 0, s>>>(1, 0, 113, jb, rem, Ajj, lda, A12, lda, d_info);   // U12 <- U11^{-H} A12
      count_launch();
      rc = gemm_device<T>(tc, 111, rem, rem, jb, minus1, A12, lda, A12, lda, one, A + (j0 + jb) + (size_t)(j0 + jb) * lda, lda, s, 'U');
    }
    if (rc) return rc;
  }
  return 0;
}

// Two-level right-looking Cholesky.  Inside an outer block of NB2 columns every 32-column panel is factorised over
// ALL rows below it by chol_panel_kernel (diagonal block + solve in one launch) and only the remaining columns of
// the outer block are updated (a skinny rank-32 GEMM); the trailing matrix then gets ONE rank-NB2 update per outer
// step through the tuned GEMM in triangle mode (DMMA for double / complex double).
template <typename T>
static int potrf_panels(char uplo, int n, T *A, int lda, int *d_info, int NB2, cudaStream_t s) {
  const int lower = (uplo == 'L' || uplo == 'l');
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  const int tc = Num<T>::cplx ? 113 : 112;
  const size_t shm = (size_t)NB * CHOL_PANEL_THREADS * sizeof(T);
  JB_CUDA(cudaFuncSetAttribute(chol_panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  int *d_count = nullptr;                              // completion counter of the panel launches (self-resetting)
  if (int r = ws_alloc((void **)&d_count, sizeof(int), s)) return r;
  struct CountFree { int *p; cudaStream_t st; ~CountFree() { ws_free(p, st); } } count_free{d_count, s};
  JB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
  // C[r0:, c0:c1] -= L[r0:, k0:k1] * L[c0:c1, k0:k1]^H in lower form (rows >= columns only), or its mirror image
  auto update = [&](int r0, int c0, int c1, int k0, int k1, cudaStream_t st, int r_end = -1) -> int {
    const int mr = (r_end < 0 ? n : r_end) - r0, nc = c1 - c0, kk = k1 - k0;
    if (mr <= 0 || nc <= 0) return 0;
    if (lower)
      return gemm_device<T>(111, tc, mr, nc, kk, minus1, A + r0 + (size_t)k0 * lda, lda, A + c0 + (size_t)k0 * lda, lda, one,
                            A + r0 + (size_t)c0 * lda, lda, st, 'L');
    return gemm_device<T>(tc, 111, nc, mr, kk, minus1, A + k0 + (size_t)c0 * lda, lda, A + k0 + (size_t)r0 * lda, lda, one,
                          A + c0 + (size_t)r0 * lda, lda, st, 'U');
  };
  // the 32-column panels of outer block [J0, J0 + jb2) with the updates that stay inside the block
  // split = true: only the rows of the diagonal block (a few CTAs); the rows below follow as ONE triangular solve in
  // GEMM form (trsm_inv.cu) — the block column below a 512-wide outer block is then 17 launches on the tuned GEMM instead
  // of 16 x (panel over all rows + skinny rank-32 update), and the look-ahead chain needs 8 free SMs instead of 16-64.
  // Middle level (potrf.nbm, default 128): the rank-32 updates stay inside sub-blocks of nbm columns, the rest of the outer
  // block takes one rank-nbm update per sub-block — the skinny updates only stream C, and this moves 2.5x fewer bytes at
  // nbm = 128, NB2 = 512 (they were most of the look-ahead chain: profiles/trace_potrf16384_r2c.txt).  As with any change
  // of block size an entry's update is accumulated in different groups (same k order, different rounding points).
  int nbm = option("potrf.nbm");
  if (nbm < 0) nbm = 128;
  auto factor_block = [&](int J0, int jb2, cudaStream_t st, bool split) -> int {
    const int r_end = split ? J0 + jb2 : n;
    const int sub = (nbm >= 2 * NB && jb2 > nbm + NB) ? nbm : jb2;
    for (int S0 = J0; S0 < J0 + jb2; S0 += sub) {
      const int S1 = (S0 + sub < J0 + jb2) ? S0 + sub : J0 + jb2;
      for (int j0 = S0; j0 < S1; j0 += NB) {
        const int jb = (S1 - j0 < NB) ? (S1 - j0) : NB;
        const int m = r_end - j0;
        const int G = m > NB ? (m - NB + (CHOL_PANEL_THREADS - NB) - 1) / (CHOL_PANEL_THREADS - NB) : 1;
        chol_panel_kernel<T><<<G, CHOL_PANEL_THREADS, shm, st>>>(lower, m, jb, A + j0 + (size_t)j0 * lda, lda, d_info, j0, d_count);
        count_launch();
        if (int rc = update(j0 + jb, j0 + jb, S1, j0, j0 + jb, st, r_end)) return rc;
      }
      if (int rc = update(S1, S1, J0 + jb2, S0, S1, st, r_end)) return rc;
    }
    return 0;
  };
  auto solve_below = [&](int J0, int jb2, cudaStream_t st) -> int {
    const int rem = n - J0 - jb2;
    if (rem <= 0) return 0;
    T *Ajj = A + J0 + (size_t)J0 * lda;
    if (lower) return trsm_inv_device<T>(142, 1, tc, 0, rem, jb2, Ajj, lda, A + (J0 + jb2) + (size_t)J0 * lda, lda, st);   // A21 <- A21 L11^{-H}
    return trsm_inv_device<T>(141, 0, tc, 0, jb2, rem, Ajj, lda, A + J0 + (size_t)(J0 + jb2) * lda, lda, st);              // A12 <- U11^{-H} A12
  };
  const bool split = option("potrf.split") > 0 && NB2 >= 128;      // measured (tools/time_potrf_rsv.py): 16384: 64.4 vs 65.4 ms, 8192: 15.1 vs 13.2 — off by default
  // Look-ahead as in getrf_device: the trailing update is split into the next block's columns and the rest, and the
  // next block is factorised on a second, higher-priority stream while the main stream finishes the rest.  The panel
  // CTAs are independent (no co-residency needed), so a fixed 16 SMs are kept free of the persistent DMMA GEMM.
  const bool lookahead = option("potrf.lookahead") != 0 && n > NB2 + NB;
  LookaheadStreams la;
  if (lookahead) if (int r = la.open()) return r;
  cudaStream_t s2 = la.s2;
  cudaEvent_t ev_first = la.ev_first, ev_panel = la.ev_panel;
  int rc = 0;
  bool factored = false;
  StepTrace tr;
  tr.on = lookahead && option("potrf.trace") > 0;
  for (int J0 = 0; J0 < n && !rc; J0 += NB2) {
    const int jb2 = (n - J0 < NB2) ? (n - J0) : NB2;
    tr.mark(s, J0);
    if (!factored) rc = factor_block(J0, jb2, s, split);
    factored = false;
    if (!rc && split) rc = solve_below(J0, jb2, s);
    if (rc) break;
    const int c0 = J0 + jb2;
    if (c0 >= n) break;
    const int jn = (n - c0 < NB2) ? (n - c0) : NB2;
    if (lookahead && c0 + jn < n) {
      rc = update(c0, c0, c0 + jn, J0, c0, s);
      if (rc) break;
      JB_CUDA_BREAK(cudaEventRecord(ev_first, s));
      JB_CUDA_BREAK(cudaStreamWaitEvent(s2, ev_first, 0));
      tr.mark(s, J0);
      int rsv = option("potrf.reserve_sms");
      const bool timed = option("potrf.timed_cap") != 0;
      const int a0 = c0 + jn;
      const double mt = (double)(n - a0), up_flop = mt * mt * jb2 * (Num<T>::cplx ? 4.0 : 1.0);
      const double per_sm = 0.182e6;                  // flop per us and SM on the triangle update
      const int G = (n - c0 - NB + (CHOL_PANEL_THREADS - NB) - 1) / (CHOL_PANEL_THREADS - NB);
      auto chain_us = [&](int r) { return (jn / (double)NB) * (((G + r - 1) / r) * 21.0 + 35.0); };
      {
        // SMs kept free of the persistent GEMM for the panel chain of the next block.  Unsplit, the chain runs its G
        // independent panel CTAs in waves over the free SMs, so it is as long as the update when few SMs face a tall panel,
        // and the update is longer than it has to be when many SMs wait for a short one: pick the split that balances the
        // two (model: update at ~0.18 TFLOP/s per SM, a panel wave 21 us, 35 us per 32-column step for the skinny update).
        if (rsv < 0 && split) rsv = 8;
        // timed: the cap holds only for the first columns of the update — as many as the chain is expected to last; the
        // rest runs as a second, uncapped launch.  Measured (tools/time_lookahead.py, profiles/lookahead_r2c.jsonl): with
        // the rank-32 chain of round 2b it lost (16384: 65.5 ms against 64.2 — the chain was as long as the update, and
        // an uncapped launch that starts before the chain has ended takes its SMs); with the middle level the chain is
        // shorter than the update and it wins a little (16384: 60.4 against 61.0, 8192: 12.97 / 12.92).
        if (rsv < 0) {
          double best = 1e30;
          for (int r : {24, 32, 48, 64, 96}) {      // below 24 the chain's skinny updates starve (16384: 65.6 ms at 16, 64.6 at 24)
            if (!timed && r > 64) continue;
            const double t_chain = chain_us(r), t_cap = up_flop / (per_sm * (num_sms() - r));
            double t = t_cap > t_chain ? t_cap : t_chain;
            if (timed && t_cap > t_chain) t = t_chain + (up_flop - t_chain * per_sm * (num_sms() - r)) / (per_sm * num_sms());
            if (t < best) { best = t; rsv = r; }
          }
        }
      }
      rc = factor_block(c0, jn, s2, split);      // its skinny updates stay uncapped (waves over the free SMs; capped: 66.3 against 64.2 ms at 16384)
      if (rc) break;
      JB_CUDA_BREAK(cudaEventRecord(ev_panel, s2));
      tr.mark(s2, J0);
      {
        int cs = n;                                      // columns [a0, cs) run capped
        if (timed && rsv > 0) {
          int pct = option("potrf.chain_pct");
          if (pct <= 0) pct = 140;       // 16384: 61.2 / 60.2 / 59.5 ms at 80 / 110 / 140 %
          const double f1 = chain_us(rsv) * pct * 0.01 * per_sm * (num_sms() - rsv) / (jb2 * (Num<T>::cplx ? 4.0 : 1.0));   // (n-a0)^2 - (n-cs)^2
          const double left = mt * mt - f1;
          if (left > 0) {
            cs = n - (int)sqrt(left);
            cs = a0 + (cs - a0 + 127) / 128 * 128;
            if (cs > n - 512) cs = n;
          }
        }
        {
          GemmCapScope cap(rsv > 0 ? num_sms() - rsv : 0);
          rc = update(a0, a0, cs, J0, c0, s);
        }
        if (!rc && cs < n) rc = update(cs, cs, n, J0, c0, s);
      }
      tr.mark(s, J0);
      JB_CUDA_BREAK(cudaStreamWaitEvent(s, ev_panel, 0));
      factored = true;
    } else {
      rc = update(c0, c0, n, J0, c0, s);
      if (tr.on) { tr.mark(s, J0); tr.mark(s, J0); tr.mark(s, J0); }
    }
  }
  if (lookahead) {
    cudaEventRecord(ev_panel, s2);
    cudaStreamWaitEvent(s, ev_panel, 0);
  }
  tr.print("potrf", s);
  if (rc) return rc;
  JB_CUDA(cudaGetLastError());
  return 0;
}

// Two-level right-looking Cholesky: outer panels of NB2 columns are factorised by the NB=32 routine, the
// block column (row) below (right of) them is solved 32 columns at a time, and the trailing matrix gets ONE
// rank-NB2 update per outer step through the tuned GEMM (DMMA for double / complex double) — NB2/32 times
// fewer sweeps over the trailing matrix than the single-level algorithm.
template <typename T>
int potrf_device(char uplo, int n, T *A, int lda, int *d_info, cudaStream_t s) {
  if (n <= 0) return 0;
  const int lower = (uplo == 'L' || uplo == 'l');
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  const int tc = Num<T>::cplx ? 113 : 112;
  int NB2 = option("potrf.nb2");
  // outer block width, measured with look-ahead (tools/time_nb2_potrf.py, ms for 128 / 256 / 512): 2048: 2.23 / 2.18 / 2.35 (single level 2.36);
  // 4096: 4.67 / 4.89 / 5.14; 8192: 15.5 / 13.8 / 14.5; 16384: 81.5 / 68.4 / 64.3
  if (NB2 <= 0) NB2 = (n >= 16384) ? 512 : (n >= 8192 ? 256 : (n >= 2048 ? 128 : n));
  if (option("potrf.variant") <= 0) return potrf_panels<T>(uplo, n, A, lda, d_info, NB2, s);   // 1: separate potf2 / trsm kernels
  if (n <= NB2 + NB) { int rc = potrf_nb32<T>(lower, n, A, lda, d_info, 0, s); if (!rc) JB_CUDA(cudaGetLastError()); return rc; }
  for (int j0 = 0; j0 < n; j0 += NB2) {
    const int jb2 = (n - j0 < NB2) ? (n - j0) : NB2;
    T *Ajj = A + j0 + (size_t)j0 * lda;
    int rc = potrf_nb32<T>(lower, jb2, Ajj, lda, d_info, j0, s);
    if (rc) return rc;
    const int rem = n - j0 - jb2;
    if (rem <= 0) break;
    if (lower) {
      T *A21 = A + (j0 + jb2) + (size_t)j0 * lda;          // rem x jb2 :  A21 <- A21 * L11^{-H}
      for (int c0 = 0; c0 < jb2; c0 += NB) {
        const int cb = (jb2 - c0 < NB) ? (jb2 - c0) : NB;
        trsm_small_right_upper_kernel<T><<<(rem + 255) / 256, 256, 0, s>>>(113, cb, rem, Ajj + c0 + (size_t)c0 * lda, lda,
                                                                          A21 + (size_t)c0 * lda, lda, d_info);
        count_launch();
        const int right = jb2 - c0 - cb;
        if (right > 0) {   // A21[:, c0+cb:] -= X * L11[c0+cb:, c0:c0+cb]^H
          rc = gemm_device<T>(111, tc, rem, right, cb, minus1, A21 + (size_t)c0 * lda, lda, Ajj + (c0 + cb) + (size_t)c0 * lda, lda,
                              one, A21 + (size_t)(c0 + cb) * lda, lda, s, 0);
          if (rc) return rc;
        }
      }
      rc = gemm_device<T>(111, tc, rem, rem, jb2, minus1, A21, lda, A21, lda, one, A + (j0 + jb2) + (size_t)(j0 + jb2) * lda, lda, s, 'L');
    } else {
      T *A12 = A + j0 + (size_t)(j0 + jb2) * lda;          // jb2 x rem :  A12 <- U11^{-H} * A12
      int blocks = (rem + 7) / 8; if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
      for (int r0 = 0; r0 < jb2; r0 += NB) {
        const int rb = (jb2 - r0 < NB) ? (jb2 - r0) : NB;
        trsm_small_left_kernel<T><<<blocks, 256, 0, s>>>(1, 0, 113, rb, rem, Ajj + r0 + (size_t)r0 * lda, lda, A12 + r0, lda, d_info);
        count_launch();
        const int below = jb2 - r0 - rb;
        if (below > 0) {   // A12[r0+rb:, :] -= U11[r0:r0+rb, r0+rb:]^H * X
          rc = gemm_device<T>(tc, 111, below, rem, rb, minus1, Ajj + r0 + (size_t)(r0 + rb) * lda, lda, A12 + r0, lda, one,
                              A12 + r0 + rb, lda, s, 0);
          if (rc) return rc;
        }
      }
      rc = gemm_device<T>(tc, 111, rem, rem, jb2, minus1, A12, lda, A12, lda, one, A + (j0 + jb2) + (size_t)(j0 + jb2) * lda, lda, s, 'U');
    }
    if (rc) return rc;
  }
  JB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int potrs_device(char uplo, int n, int nrhs, const T *A, int lda, T *B, int ldb, cudaStream_t s) {
  if (n <= 0 || nrhs <= 0) return 0;
  const int lower = (uplo == 'L' || uplo == 'l');
  int rc;
  if (lower) {
    if ((rc = trsm_left_blocked<T>(1, 111, 0, n, nrhs, A, lda, B, ldb, s))) return rc;   // L y = b
    if ((rc = trsm_left_blocked<T>(1, 113, 0, n, nrhs, A, lda, B, ldb, s))) return rc;   // L^H x = y
  } else {
    if ((rc = trsm_left_blocked<T>(0, 113, 0, n, nrhs, A, lda, B, ldb, s))) return rc;   // U^H y = b
    if ((rc = trsm_left_blocked<T>(0, 111, 0, n, nrhs, A, lda, B, ldb, s))) return rc;   // U x = y
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// cblas_?trsm on the device (column-major): B <- alpha * op(A)^{-1} B (Left) or alpha * B op(A)^{-1} (Right).
// Level-3 sibling the factorisations need anyway (BlasOps.trsm, /root/reference/Numeric/Jalla/Foreign/BlasOps.hs:57;
// SURVEY.md §8f rank 3).  uplo 121 upper / 122 lower, trans 111/112/113, diag 131 non-unit / 132 unit.
template <typename T>
__global__ void scale_kernel(int m, int n, T alpha, T *B, int ldb) {
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
      B[(size_t)i + (size_t)j * ldb] = Num<T>::mul(alpha, B[(size_t)i + (size_t)j * ldb]);
}

template <typename T>
int trsm_device(int side, int uplo, int trans, int diag, int m, int n, T alpha, const T *A, int lda, T *B, int ldb,
                cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  if (Num<T>::re(alpha) == 0 && Num<T>::im(alpha) == 0) {
    // ?trsm: alpha == 0 sets B := 0 without referencing A (NaN / Inf in B or a zero diagonal must not leak through)
    for (int j = 0; j < n; j++) JB_CUDA(cudaMemsetAsync(B + (size_t)j * ldb, 0, sizeof(T) * (size_t)m, s));
    return 0;
  }
  if (!(Num<T>::re(alpha) == 1 && Num<T>::im(alpha) == 0)) {
    dim3 grid((unsigned)min((m + 255) / 256, 1024), (unsigned)min(n, 4096));
    scale_kernel<T><<<grid, 256, 0, s>>>(m, n, alpha, B, ldb);
    count_launch();
  }
  if (!Num<T>::cplx && trans == 113) trans = 112;
  const int stored_lower = (uplo == 122), unit = (diag == 132);
  if (side == 141) return trsm_left_blocked<T>(stored_lower, trans, unit, m, n, A, lda, B, ldb, s);
  // Right side: X op(A) = B.  Column-block sweep: effective-upper op(A) runs left to right, effective-lower
  // right to left; each step solves a 32-wide column block against the diagonal block (one thread per row of
  // B) and updates the remaining columns with one GEMM.
  if (trsm_inv_pays(n, m)) return trsm_inv_device<T>(142, stored_lower, trans, unit, m, n, A, lda, B, ldb, s);
  const int eff_upper = (trans == 111) ? !stored_lower : stored_lower;
  const int nblk = (n + NB - 1) / NB;
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  if (unit || !eff_upper) {
    // less common variants: solve the transposed left-side system  op(A)^T X^T = B^T  through a scratch transpose
    T *W = nullptr;
    const int ldw = n + (n & 1);
    if (int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)ldw * m, s)) return rc;
    int rc = lacpy_device<T>(112, m, n, B, ldb, W, ldw, s);
    // (X op(A))^T = op(A)^T X^T : transposing flips the trans flag (conj-trans <-> plain conj is handled by
    // conjugating through the transpose copy for 113)
    int t2 = (trans == 111) ? 112 : 111;
    if (trans == 113) {   // op(A) = A^H: A^H^T = conj(A): solve conj(A) Y = B^T  <=>  A conj(Y) = conj(B^T)
      rc = rc ? rc : lacpy_device<T>(113, m, n, B, ldb, W, ldw, s);
      t2 = 111;
    }
    if (!rc) rc = trsm_left_blocked<T>(stored_lower, t2, unit, n, m, A, lda, W, ldw, s);
    if (!rc) rc = lacpy_device<T>(trans == 113 ? 113 : 112, n, m, W, ldw, B, ldb, s);
    ws_free(W, s);
    return rc;
  }
  for (int bi = 0; bi < nblk; bi++) {
    const int c0 = bi * NB, cb = (n - c0 < NB) ? (n - c0) : NB;
    trsm_small_right_upper_kernel<T><<<(m + 255) / 256, 256, 0, s>>>(trans, cb, m, A + c0 + (size_t)c0 * lda, lda,
                                                                    B + (size_t)c0 * ldb, ldb, nullptr);
    count_launch();
    const int right = n - c0 - cb;
    if (right > 0) {   // B[:, c0+cb:] -= X * op(A)[c0:c0+cb, c0+cb:]
      const T *Ab = (trans == 111) ? A + c0 + (size_t)(c0 + cb) * lda : A + (c0 + cb) + (size_t)c0 * lda;
      int rc = gemm_device<T>(111, trans, m, right, cb, minus1, B + (size_t)c0 * ldb, ldb, Ab, lda, one,
                              B + (size_t)(c0 + cb) * ldb, ldb, s, 0);
      if (rc) return rc;
    }
  }
  JB_CUDA(cudaGetLastError());
  return 0;
}

// forward row interchanges ipiv[k0..k1) (1-based, global rows = rows of A) on `ncols` columns of A
template <typename T>
int laswp_device(T *A, int lda, int ncols, const int *d_ipiv, int k0, int k1, cudaStream_t s) {
  if (int rc = laswp_launch<T>(A, lda, 0, ncols, d_ipiv, k0, k1, 0, s)) return rc;
  JB_CUDA(cudaGetLastError());
  return 0;
}
template int laswp_device<double>(double *, int, int, const int *, int, int, cudaStream_t);

#define INST(T)                                                                                        \
  template int getrf_device<T>(int, int, T *, int, int *, int *, cudaStream_t);                        \
  template int getrs_device<T>(char, int, int, const T *, int, const int *, T *, int, cudaStream_t);   \
  template int getri_device<T>(int, T *, int, const int *, int *, cudaStream_t);                       \
  template int potrf_device<T>(char, int, T *, int, int *, cudaStream_t);                              \
  template int potrs_device<T>(char, int, int, const T *, int, T *, int, cudaStream_t);              \
  template int trsm_device<T>(int, int, int, int, int, int, T, const T *, int, T *, int, cudaStream_t);
INST(float)
INST(double)
INST(cfloat)
INST(cdouble)

}  // namespace jb
