// Triangular solves with many right-hand sides as GEMMs: B <- op(A)^{-1} B (Left) or B op(A)^{-1} (Right), column-major —
// the level-3 step inside the blocked factorisations (the unit-lower block-row solve of ?getrf, the panel solve of ?potrf,
// ?getrs / ?potrs / ?getri with wide right-hand sides) and behind cblas_?trsm
// (/root/reference/Numeric/Jalla/Foreign/BlasOps.hs:57, Foreign/LAPACKE.chs:697-706,741-744).
//
// The 32-wide substitution kernels of lapack_blocked.cu pay two launch-bound kernels per 32 rows of the triangle
// (ncu, dgetrf 16384: 976 trsm_small launches + their rank-32 updates = 1 ms of every 3.3 ms outer step).  Here the 64 x 64
// diagonal blocks of the triangle are inverted once (one CTA per block, forward substitution on the identity in shared
// memory) and the sweep becomes two GEMMs per 64 rows on the tuned kernels:
//     X_c = op(inv(A_cc)) B_c            B_rest -= op(A)_{rest,c} X_c            (Left; Right mirrors it)
// X is collected in a scratch copy (a GEMM cannot run in place) and copied back once at the end.  Inverting only the small
// diagonal blocks keeps the error of plain blocked substitution up to the conditioning of a 64 x 64 block.
#include "jb_common.cuh"
#include <atomic>

namespace jb {

constexpr int NBI = 64, NBI_LD = NBI + 1;

// Dinv[b] (NBI x NBI, ld NBI, storage orientation, other triangle zero) = inverse of the b-th diagonal block of the
// triangular matrix A (order nt); a ragged last block is padded with the identity.
template <typename T>
__global__ void __launch_bounds__(256) trtri_diag_kernel(int stored_lower, int unit, int nt, const T *__restrict__ A, int lda, T *__restrict__ Dinv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *D = reinterpret_cast<T *>(smem_raw);          // effective lower block L (= A_bb or its transpose), row-major, padded
  T *X = D + NBI * NBI_LD;                          // inv(L)
  const int b = blockIdx.x, r0 = b * NBI, nb = min(NBI, nt - r0), tid = threadIdx.x;
  for (int idx = tid; idx < NBI * NBI; idx += blockDim.x) {
    const int i = idx % NBI, j = idx / NBI;         // storage element (i, j) of the block
    const int li = stored_lower ? i : j, lk = stored_lower ? j : i;
    if (li < lk) continue;
    T v = (li == lk) ? Num<T>::one() : Num<T>::zero();
    if (i < nb && j < nb && !(unit && li == lk)) v = A[(size_t)(r0 + i) + (size_t)(r0 + j) * lda];
    D[li * NBI_LD + lk] = v;
  }
  __syncthreads();
  if (tid < NBI) {
    // column j of inv(L): forward substitution of L x = e_j, every thread in lockstep over the rows (x_k = 0 above j)
    const int j = tid;
    for (int i = 0; i < NBI; i++) {
      T acc0 = (i == j) ? Num<T>::one() : Num<T>::zero(), acc1 = Num<T>::zero();
      int k = 0;
      for (; k + 1 < i; k += 2) {
        acc0 = Num<T>::sub(acc0, Num<T>::mul(D[i * NBI_LD + k], X[k * NBI_LD + j]));
        acc1 = Num<T>::sub(acc1, Num<T>::mul(D[i * NBI_LD + k + 1], X[(k + 1) * NBI_LD + j]));
      }
      if (k < i) acc0 = Num<T>::sub(acc0, Num<T>::mul(D[i * NBI_LD + k], X[k * NBI_LD + j]));
      T x = Num<T>::add(acc0, acc1);
      if (i < j) x = Num<T>::zero();
      else if (!unit) x = Num<T>::mul(x, Num<T>::recip(D[i * NBI_LD + i]));
      X[i * NBI_LD + j] = x;
    }
  }
  __syncthreads();
  T *out = Dinv + (size_t)b * NBI * NBI;
  for (int idx = tid; idx < NBI * NBI; idx += blockDim.x) {
    const int i = idx % NBI, j = idx / NBI;
    const int li = stored_lower ? i : j, lk = stored_lower ? j : i;
    out[idx] = (li >= lk) ? X[li * NBI_LD + lk] : Num<T>::zero();      // inv(U) = inv(U^T)^T
  }
}

// Triangles of order <= REC_BASE: the 64-row sweep.
template <typename T>
static int trsm_inv_base(int side, int stored_lower, int trans, int unit, int m, int n, const T *A, int lda, T *B, int ldb, cudaStream_t s) {
  const int nt = side == 141 ? m : n, nblk = (nt + NBI - 1) / NBI;
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one(), zero = Num<T>::zero();
  T *Dinv = nullptr, *W = nullptr;
  if (int rc = ws_alloc((void **)&Dinv, sizeof(T) * (size_t)nblk * NBI * NBI, s)) return rc;
  const int ldw = m + (m & 1);
  if (int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)ldw * n, s)) { ws_free(Dinv, s); return rc; }
  const int shm = 2 * NBI * NBI_LD * (int)sizeof(T);
  int rc = 0;
  {
    static_assert(2 * NBI * NBI_LD * 16 <= 200 * 1024, "shared memory of the block inversion");
    cudaError_t e = cudaSuccess;
    if (shm > 48 * 1024) e = cudaFuncSetAttribute(trtri_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm);
    if (e == cudaSuccess) {
      trtri_diag_kernel<T><<<nblk, 256, shm, s>>>(stored_lower, unit, nt, A, lda, Dinv);
      count_launch();
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "trtri_diag launch failed: %s", cudaGetErrorString(e)); rc = JB_ERR_CUDA; }
  }
  const int eff_lower = (trans == 111) ? stored_lower : !stored_lower;      // shape of op(A)
  // Left: op(A) lower runs top-down, upper bottom-up.  Right: X op(A) = B with op(A) upper runs left to right.
  const bool ascending = (side == 141) ? eff_lower : !eff_lower;
  for (int bi = 0; bi < nblk && !rc; bi++) {
    const int c = ascending ? bi : nblk - 1 - bi;
    const int p0 = c * NBI, pb = min(NBI, nt - p0);
    const int rest0 = ascending ? p0 + pb : 0, cnt = ascending ? nt - p0 - pb : p0;
    const T *Dc = Dinv + (size_t)c * NBI * NBI;
    if (side == 141) {
      rc = gemm_device<T>(trans, 111, pb, n, pb, one, Dc, NBI, B + p0, ldb, zero, W + p0, ldw, s, 0);
      if (!rc && cnt > 0) {
        // op(A)[rest, c]: stored block A[rest, c] (NoTrans) or op of the stored block A[c, rest]
        const T *Ab = (trans == 111) ? A + rest0 + (size_t)p0 * lda : A + p0 + (size_t)rest0 * lda;
        rc = gemm_device<T>(trans, 111, cnt, n, pb, minus1, Ab, lda, W + p0, ldw, one, B + rest0, ldb, s, 0);
      }
    } else {
      rc = gemm_device<T>(111, trans, m, pb, pb, one, B + (size_t)p0 * ldb, ldb, Dc, NBI, zero, W + (size_t)p0 * ldw, ldw, s, 0);
      if (!rc && cnt > 0) {
        // op(A)[c, rest]: stored block A[c, rest] (NoTrans) or op of the stored block A[rest, c]
        const T *Ab = (trans == 111) ? A + p0 + (size_t)rest0 * lda : A + rest0 + (size_t)p0 * lda;
        rc = gemm_device<T>(111, trans, m, cnt, pb, minus1, W + (size_t)p0 * ldw, ldw, Ab, lda, one, B + (size_t)rest0 * ldb, ldb, s, 0);
      }
    }
  }
  if (!rc) rc = lacpy_device<T>(111, m, n, W, ldw, B, ldb, s);
  ws_free(W, s);
  ws_free(Dinv, s);
  return rc;
}

// side 141 Left / 142 Right; A is the stored triangle (order m for Left, n for Right); alpha already applied by the caller.
// Larger triangles are halved recursively — solve one half, update the other with ONE GEMM whose inner dimension is the
// half (the efficient shape for the tuned kernels), solve the other half — down to REC_BASE, where the 64-row sweep runs
// (dtrsm n = nrhs = 8192: 63.6 ms with the 32-row kernels, 36.3 ms with the sweep alone).
constexpr int REC_BASE = 512;
template <typename T>
int trsm_inv_device(int side, int stored_lower, int trans, int unit, int m, int n, const T *A, int lda, T *B, int ldb, cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  if (!Num<T>::cplx && trans == 113) trans = 112;
  const int nt = side == 141 ? m : n;
  // narrow right-hand sides leave the coupling GEMM too few tiles (n = 8192, nrhs = 512: 6.2 ms halved, 5.0 ms swept)
  int rec_base = option("trsm.rec_base");        // experiments (tools/time_trsm.py): order from which triangles are halved
  if (rec_base < 2 * NBI) rec_base = REC_BASE;
  if (nt <= rec_base || (side == 141 ? n : m) < 1024) return trsm_inv_base<T>(side, stored_lower, trans, unit, m, n, A, lda, B, ldb, s);
  const int h = ((nt / 2 + NBI - 1) / NBI) * NBI, r = nt - h;
  const T minus1 = Num<T>::neg(Num<T>::one()), one = Num<T>::one();
  const int eff_lower = (trans == 111) ? stored_lower : !stored_lower;
  const bool ascending = (side == 141) ? eff_lower : !eff_lower;
  const T *A11 = A, *A22 = A + h + (size_t)h * lda;
  // the off-diagonal block of op(A) that couples the halves: stored below the diagonal (A + h) or right of it (A + h*lda)
  const T *Aoff = (stored_lower ? A + h : A + (size_t)h * lda);
  int rc;
  if (side == 141) {
    T *B1 = B, *B2 = B + h;
    if (ascending) {            // op(A) lower: X1, then B2 -= op(A)[h:, :h] X1, then X2
      if ((rc = trsm_inv_device<T>(side, stored_lower, trans, unit, h, n, A11, lda, B1, ldb, s))) return rc;
      if ((rc = gemm_device<T>(trans, 111, r, n, h, minus1, Aoff, lda, B1, ldb, one, B2, ldb, s, 0))) return rc;
      return trsm_inv_device<T>(side, stored_lower, trans, unit, r, n, A22, lda, B2, ldb, s);
    }
    if ((rc = trsm_inv_device<T>(side, stored_lower, trans, unit, r, n, A22, lda, B2, ldb, s))) return rc;
    if ((rc = gemm_device<T>(trans, 111, h, n, r, minus1, Aoff, lda, B2, ldb, one, B1, ldb, s, 0))) return rc;
    return trsm_inv_device<T>(side, stored_lower, trans, unit, h, n, A11, lda, B1, ldb, s);
  }
  T *B1 = B, *B2 = B + (size_t)h * ldb;
  if (ascending) {              // op(A) upper: X1, then B2 -= X1 op(A)[:h, h:], then X2
    if ((rc = trsm_inv_device<T>(side, stored_lower, trans, unit, m, h, A11, lda, B1, ldb, s))) return rc;
    if ((rc = gemm_device<T>(111, trans, m, r, h, minus1, B1, ldb, Aoff, lda, one, B2, ldb, s, 0))) return rc;
    return trsm_inv_device<T>(side, stored_lower, trans, unit, m, r, A22, lda, B2, ldb, s);
  }
  if ((rc = trsm_inv_device<T>(side, stored_lower, trans, unit, m, r, A22, lda, B2, ldb, s))) return rc;
  if ((rc = gemm_device<T>(111, trans, m, h, r, minus1, B2, ldb, Aoff, lda, one, B1, ldb, s, 0))) return rc;
  return trsm_inv_device<T>(side, stored_lower, trans, unit, m, h, A11, lda, B1, ldb, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// A handful of right-hand sides (the Jalla call: solveLinearSystem with a vector, /root/reference/Numeric/Jalla/Matrix.hs:495-508
// -> ?gesv nrhs = 1; ?getrs / ?potrs likewise).  With one column a triangular solve only STREAMS the triangle once — n^2/2
// entries, 1 GiB at n = 16384, 0.17 ms at the HBM rate — but the 32-row substitution kernels pay two launches per 32 rows
// (dgetrs n = 16384, nrhs = 1: 17 ms), and the GEMM-form sweep above has no tile to fill.  Here the same inverted 64 x 64
// diagonal blocks are used, and one launch per 64 rows does both steps of the sweep:
//     every CTA:  x_c = op(inv(A_cc)) b_c   (64 x 64, from L2; CTA 0 stores it)      then its rows:  b_r -= op(A)_{r,c} x_c
// a thread owns a row (NoTrans: coalesced column reads; Trans: 512 contiguous bytes per thread), so the launch is a
// matrix-vector product at memory speed.  x is collected in a scratch column and copied back at the end: within a launch
// the other CTAs still read b_c.
// A CTA of four warps owns 32 rows: warp q multiplies columns 16q .. 16q+15 of the block column (lane = row), and its
// sixteen loads are issued BEFORE x_c is formed — they do not depend on it — so a launch costs one memory round trip plus
// the 64 x 64 product (all 128 threads, Dinv from L2, fully unrolled loads) instead of several dependent ones
// (first version, a thread per row and 64 columns: 18.6 us per launch, dgetrs 16384 x 1 9.5 ms).
constexpr int TRSV_THREADS = 128, TRSV_ROWS = 32, TRSV_COLS = NBI / 4;
// Launched with programmatic stream serialisation (after the first step): a launch lets its successor start at once, and the
// successor loads its entries of A and Dinv — which no step writes — while this one is still running, then waits
// (griddepcontrol.wait = predecessor complete and visible) before it touches B.
template <typename T, int NR>
__global__ void __launch_bounds__(TRSV_THREADS)
trsv_step_kernel(int trans, int p0, int pb, int rest0, int cnt, const T *__restrict__ A, int lda, const T *__restrict__ Dc,
                 T *__restrict__ B, int ldb, T *__restrict__ W, int ldw, int nrhs) {
  __shared__ T bs[NR][NBI], xs[NR][NBI], part[2][NR][NBI], red[4][NR][TRSV_ROWS];
  const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
  asm volatile("griddepcontrol.launch_dependents;");
  // ---- this thread's sixteen entries of op(A)[row, p0 + 16q ..] (zero outside the block column / the rows)
  const int r = blockIdx.x * TRSV_ROWS + lane;
  const bool live = r < cnt;
  const size_t row = (size_t)rest0 + (live ? r : 0);
  T a[TRSV_COLS];
  const int cq = q * TRSV_COLS;
  if (trans == 111) {
    const T *Ar = A + row + (size_t)(p0 + cq) * lda;
#pragma unroll
    for (int u = 0; u < TRSV_COLS; u++) a[u] = (live && cq + u < pb) ? Ar[(size_t)u * lda] : Num<T>::zero();
  } else {
    const T *Ar = A + (p0 + cq) + row * lda;               // op(A)[row, p0 + i] = op(stored A[p0 + i, row])
#pragma unroll
    for (int u = 0; u < TRSV_COLS; u++) {
      T v = (live && cq + u < pb) ? Ar[u] : Num<T>::zero();
      a[u] = trans == 113 ? Num<T>::conj(v) : v;
    }
  }
  // ---- x_c = op(Dinv_c) b_c: thread (t, h) sums columns 32h .. 32h+31 of row t
  {
    const int t = tid & (NBI - 1), h = tid / NBI;
    T d[NBI / 2];
#pragma unroll
    for (int u = 0; u < NBI / 2; u++) {
      const int jj = h * (NBI / 2) + u;
      T v = (trans == 111) ? Dc[t + jj * NBI] : Dc[jj + t * NBI];
      d[u] = trans == 113 ? Num<T>::conj(v) : v;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");          // B as the previous step left it
    for (int idx = tid; idx < NR * NBI; idx += TRSV_THREADS) {
      const int c = idx / NBI, jj = idx % NBI;
      bs[c][jj] = (c < nrhs && jj < pb) ? B[(size_t)(p0 + jj) + (size_t)c * ldb] : Num<T>::zero();
    }
    __syncthreads();
    T acc[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) acc[c] = Num<T>::zero();
#pragma unroll
    for (int u = 0; u < NBI / 2; u++)
#pragma unroll
      for (int c = 0; c < NR; c++) acc[c] = Num<T>::add(acc[c], Num<T>::mul(d[u], bs[c][h * (NBI / 2) + u]));
#pragma unroll
    for (int c = 0; c < NR; c++) part[h][c][t] = acc[c];
  }
  __syncthreads();
  for (int idx = tid; idx < NR * NBI; idx += TRSV_THREADS) {
    const int c = idx / NBI, t = idx % NBI;
    const T x = Num<T>::add(part[0][c][t], part[1][c][t]);
    xs[c][t] = x;
    if (blockIdx.x == 0 && t < pb && c < nrhs) W[(size_t)(p0 + t) + (size_t)c * ldw] = x;
  }
  __syncthreads();
  if (cnt <= 0) return;
  // ---- rows: partial products per warp, summed by warp 0
  {
    T acc[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) acc[c] = Num<T>::zero();
#pragma unroll
    for (int u = 0; u < TRSV_COLS; u++)
#pragma unroll
      for (int c = 0; c < NR; c++) acc[c] = Num<T>::add(acc[c], Num<T>::mul(a[u], xs[c][cq + u]));
#pragma unroll
    for (int c = 0; c < NR; c++) red[q][c][lane] = acc[c];
  }
  __syncthreads();
  if (q == 0 && live) {
#pragma unroll
    for (int c = 0; c < NR; c++)
      if (c < nrhs) {
        const T sum = Num<T>::add(Num<T>::add(red[0][c][lane], red[1][c][lane]), Num<T>::add(red[2][c][lane], red[3][c][lane]));
        T *bp = B + row + (size_t)c * ldb;
        *bp = Num<T>::sub(*bp, sum);
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// One right-hand side, ONE launch per triangle (round 2d).  The step kernel above is a launch per 64 rows — 512 dependent
// launches at n = 16384 (2.9 ms for ?getrs against 0.35 ms of memory traffic), ~0.1 ms of the 0.9 ms of dgesv 512.  Here a
// cooperative grid is persistent over the whole sweep and every CTA owns row blocks (of 64, in sweep order: block k goes
// to CTA k mod P): it streams its rows' entries of op(A) block by block — left-looking, so nothing but x is ever
// exchanged — and subtracts op(A)(k, j) x_j as soon as x_j exists, then forms x_k = op(inv(A_kk)) (b_k - sum) and
// publishes it.  x travels as FLAGGED WORDS (4 bytes of payload + a 4-byte launch tag per 8-byte relaxed store / load,
// single-copy atomic: a reader needs no fence, it polls the words of x_j until their tags match — the scheme of the
// cooperative LU panel), so the dependency chain per block is one L2 round trip, one 64 x 64 product with entries that
// were loaded while waiting, and the product with the inverted diagonal block (also loaded ahead).  The grid is
// co-resident (cooperative launch, P <= occupancy x SMs) and a CTA walks its blocks in sweep order, so every wait ends.
constexpr int PIPE_THREADS = 256, PIPE_Q = PIPE_THREADS / NBI, PIPE_COLS = NBI / PIPE_Q;
__device__ __forceinline__ void pipe_store(unsigned long long *p, unsigned data, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | data;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// Thread (t, h): row t of the block, columns 16h .. 16h+15 of every block column.  PIPE_DEPTH block columns of op(A) are
// in flight per thread (a ring of register buffers with static indices: the sweep loop is unrolled by the depth), so a
// CTA streams its strip at memory speed instead of one DRAM round trip per 64 columns (first version, depth 1, 128
// threads: dgetrs 16384 x 1 2.47 ms).
// NR right-hand sides ride together (1, 2, 4 or 8 columns per launch; columns >= nrhs are zeros that nothing stores): the
// entries of op(A) are loaded once and feed NR accumulators, x_j travels as NR x 64 flagged elements.
template <typename T, int NR>
__global__ void __launch_bounds__(PIPE_THREADS, 1)
trsv_pipe_kernel(int eff_lower, int trans, int n, int nblk, const T *__restrict__ A, int lda, const T *__restrict__ Dinv,
                 T *__restrict__ B, int ldb, int nrhs, unsigned long long *__restrict__ flagged, unsigned tag) {
  constexpr int W = sizeof(T) / 4;
  constexpr int PIPE_DEPTH = (sizeof(T) == 16 || NR >= 4) ? 2 : 4;      // <= 128 registers of operands in flight per thread
  __shared__ __align__(16) T xs[2][NR][NBI], part[PIPE_Q][NR][NBI], bs[NR][NBI], xfin[NR][NBI];
  const int tid = threadIdx.x, t = tid & (NBI - 1), h = tid / NBI;
  for (int k = blockIdx.x; k < nblk; k += gridDim.x) {
    const int c = eff_lower ? k : nblk - 1 - k;
    const int p0 = c * NBI, pb = min(NBI, n - p0);
    const bool live = t < pb;
    const size_t row = (size_t)p0 + (live ? t : 0);
    // entries of op(A)(row, q0 + 16h ..) of block column cj (zero outside the matrix)
    auto load_block = [&](int jo, T (&a)[PIPE_COLS]) {
      const int cj = eff_lower ? jo : nblk - 1 - jo;
      const int q0 = cj * NBI + h * PIPE_COLS, qb = n - q0;
      if (trans == 111) {
        const T *Ar = A + row + (size_t)q0 * lda;
#pragma unroll
        for (int u = 0; u < PIPE_COLS; u++) a[u] = (live && u < qb) ? Ar[(size_t)u * lda] : Num<T>::zero();
      } else {
        const T *Ar = A + q0 + row * lda;
#pragma unroll
        for (int u = 0; u < PIPE_COLS; u++) {
          const T v = (live && u < qb) ? Ar[u] : Num<T>::zero();
          a[u] = trans == 113 ? Num<T>::conj(v) : v;
        }
      }
    };
    T acc[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) acc[r] = Num<T>::zero();
    T a[PIPE_DEPTH][PIPE_COLS];
#pragma unroll
    for (int s = 0; s < PIPE_DEPTH; s++)
      if (s < k) load_block(s, a[s]);
    // this block's row of the inverted diagonal block and its right-hand sides, long before they are needed
    T d[PIPE_COLS];
    {
      const T *Dc = Dinv + (size_t)c * NBI * NBI;
#pragma unroll
      for (int u = 0; u < PIPE_COLS; u++) {
        const int jj = h * PIPE_COLS + u;
        const T v = (trans == 111) ? Dc[t + jj * NBI] : Dc[jj + t * NBI];
        d[u] = trans == 113 ? Num<T>::conj(v) : v;
      }
    }
    T bk[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) bk[r] = (live && h == 0 && r < nrhs) ? B[row + (size_t)r * ldb] : Num<T>::zero();
    for (int j0 = 0; j0 < k; j0 += PIPE_DEPTH) {
#pragma unroll
      for (int s = 0; s < PIPE_DEPTH; s++) {
        const int jo = j0 + s;
        if (jo < k) {                                   // CTA-uniform
          // ---- x_jo: poll its flagged words
          unsigned *const xw = reinterpret_cast<unsigned *>(xs[s & 1]);
          const unsigned long long *const src = flagged + (size_t)jo * NR * NBI * W;
          for (int w = tid; w < NR * NBI * W; w += PIPE_THREADS) {
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(src + w) : "memory"); } while ((unsigned)(v >> 32) != tag);
            xw[w] = (unsigned)v;
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < PIPE_COLS; u++)
#pragma unroll
            for (int r = 0; r < NR; r++) acc[r] = Num<T>::add(acc[r], Num<T>::mul(a[s][u], xs[s & 1][r][h * PIPE_COLS + u]));
          if (jo + PIPE_DEPTH < k) load_block(jo + PIPE_DEPTH, a[s]);
        }
      }
    }
    // ---- x_k = op(inv(A_kk)) (b_k - sum)
#pragma unroll
    for (int r = 0; r < NR; r++) part[h][r][t] = acc[r];
    __syncthreads();
    if (h == 0) {
#pragma unroll
      for (int r = 0; r < NR; r++) {
        T sum = part[0][r][t];
#pragma unroll
        for (int g = 1; g < PIPE_Q; g++) sum = Num<T>::add(sum, part[g][r][t]);
        bs[r][t] = live ? Num<T>::sub(bk[r], sum) : Num<T>::zero();
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; r++) {
      T ax = Num<T>::zero();
#pragma unroll
      for (int u = 0; u < PIPE_COLS; u++) ax = Num<T>::add(ax, Num<T>::mul(d[u], bs[r][h * PIPE_COLS + u]));
      acc[r] = ax;
    }
#pragma unroll
    for (int r = 0; r < NR; r++) part[h][r][t] = acc[r];      // (part was last read before the barrier above)
    __syncthreads();
    if (h == 0) {
#pragma unroll
      for (int r = 0; r < NR; r++) {
        T x = part[0][r][t];
#pragma unroll
        for (int g = 1; g < PIPE_Q; g++) x = Num<T>::add(x, part[g][r][t]);
        if (!live) x = Num<T>::zero();
        xfin[r][t] = x;
        if (live && r < nrhs) B[row + (size_t)r * ldb] = x;
      }
    }
    __syncthreads();
    {
      const unsigned *const xw = reinterpret_cast<const unsigned *>(xfin);
      unsigned long long *const dst = flagged + (size_t)k * NR * NBI * W;
      for (int w = tid; w < NR * NBI * W; w += PIPE_THREADS) pipe_store(dst + w, xw[w], tag);
    }
    __syncthreads();          // xs / part / bs / xfin are reused by this CTA's next block
  }
}
static unsigned next_pipe_tag() {
  static std::atomic<unsigned> seq{1};
  unsigned v = seq.fetch_add(1);
  return v == 0 ? seq.fetch_add(1) : v;        // never the value the flag buffer is cleared to
}
// Up to eight right-hand sides through the persistent sweep; Dinv holds the inverted diagonal blocks.  Returns 0, an error
// code, or -1 when this shape is not served (Complex Double beyond four columns: shared memory).
template <typename T, int NR>
static int trsv_pipe_launch(int eff_lower, int trans, int n, int nblk, const T *A, int lda, const T *Dinv, T *B, int ldb, int nrhs, cudaStream_t s) {
  constexpr int W = sizeof(T) / 4;
  unsigned long long *flagged = nullptr;
  const size_t bytes = sizeof(unsigned long long) * (size_t)nblk * NR * NBI * W;
  if (int rc = ws_alloc((void **)&flagged, bytes, s)) return rc;
  cudaError_t e = cudaMemsetAsync(flagged, 0, bytes, s);
  int occ = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trsv_pipe_kernel<T, NR>, PIPE_THREADS, 0);
  if (e == cudaSuccess) {
    if (occ < 1) occ = 1;
    const long cap = (long)num_sms() * occ;
    int grid = (int)(nblk < cap ? nblk : cap);
    const int lim = option("trsv.pipe_ctas");              // tests: fewer CTAs than row blocks (a CTA then walks several blocks)
    if (lim > 0 && lim < grid) grid = lim;
    unsigned tag = next_pipe_tag();
    void *args[] = {&eff_lower, &trans, &n, &nblk, &A, &lda, &Dinv, &B, &ldb, &nrhs, &flagged, &tag};
    e = cudaLaunchCooperativeKernel((void *)trsv_pipe_kernel<T, NR>, dim3(grid), dim3(PIPE_THREADS), args, 0, s);
    count_launch();
  }
  ws_free(flagged, s);
  if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "trsv_pipe launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return JB_ERR_CUDA; }
  return 0;
}
template <typename T>
static int trsv_pipe_run(int eff_lower, int trans, int n, int nblk, const T *A, int lda, const T *Dinv, T *B, int ldb, int nrhs, cudaStream_t s) {
  if (nrhs == 1) return trsv_pipe_launch<T, 1>(eff_lower, trans, n, nblk, A, lda, Dinv, B, ldb, nrhs, s);
  if (nrhs == 2) return trsv_pipe_launch<T, 2>(eff_lower, trans, n, nblk, A, lda, Dinv, B, ldb, nrhs, s);
  if (nrhs <= 4) return trsv_pipe_launch<T, 4>(eff_lower, trans, n, nblk, A, lda, Dinv, B, ldb, nrhs, s);
  if constexpr (sizeof(T) < 16) return trsv_pipe_launch<T, 8>(eff_lower, trans, n, nblk, A, lda, Dinv, B, ldb, nrhs, s);
  return -1;
}

// op(Tri) X = B in place, Tri = the stored triangle of A (order n), nrhs <= TRSV_MAX_NRHS.
template <typename T>
int trsv_few_device(int stored_lower, int trans, int unit, int n, int nrhs, const T *A, int lda, T *B, int ldb, cudaStream_t s) {
  if (n <= 0 || nrhs <= 0) return 0;
  if (nrhs > TRSV_MAX_NRHS) { set_error(JB_ERR_ARG, "trsv_few: too many right-hand sides"); return JB_ERR_ARG; }
  if (!Num<T>::cplx && trans == 113) trans = 112;
  const int nblk = (n + NBI - 1) / NBI;
  T *Dinv = nullptr, *W = nullptr;
  if (int rc = ws_alloc((void **)&Dinv, sizeof(T) * (size_t)nblk * NBI * NBI, s)) return rc;
  if (int rc = ws_alloc((void **)&W, sizeof(T) * (size_t)n * nrhs, s)) { ws_free(Dinv, s); return rc; }
  const int shm = 2 * NBI * NBI_LD * (int)sizeof(T);
  cudaError_t e = cudaSuccess;
  if (shm > 48 * 1024) e = cudaFuncSetAttribute(trtri_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm);
  if (e == cudaSuccess) {
    trtri_diag_kernel<T><<<nblk, 256, shm, s>>>(stored_lower, unit, n, A, lda, Dinv);
    count_launch();
  }
  const int eff_lower = (trans == 111) ? stored_lower : !stored_lower;
  // the persistent sweep, in place (option trsv.pipe_max_nrhs: only up to that many columns; tools/time_trsv_pipe.py, dgetrs
  // 16384 x 4 / x 8: 3.22 / 4.93 -> 2.09 / 2.82 ms, 4096 x 8: 1.03 -> 0.61 ms against the stepped sweep below)
  const int pipe_max = option("trsv.pipe_max_nrhs") > 0 ? option("trsv.pipe_max_nrhs") : TRSV_MAX_NRHS;
  if (e == cudaSuccess && nrhs <= pipe_max && option("trsv.pipe") != 0) {
    const int rc = trsv_pipe_run<T>(eff_lower, trans, n, nblk, A, lda, Dinv, B, ldb, nrhs, s);
    if (rc >= 0) {
      ws_free(W, s);
      ws_free(Dinv, s);
      return rc;
    }
  }
  const bool pdl = option("trsv.pdl") != 0;
  for (int bi = 0; bi < nblk && e == cudaSuccess; bi++) {
    const int c = eff_lower ? bi : nblk - 1 - bi;
    const int p0 = c * NBI, pb = min(NBI, n - p0);
    const int rest0 = eff_lower ? p0 + pb : 0, cnt = eff_lower ? n - p0 - pb : p0;
    const int grid = cnt > 0 ? (cnt + TRSV_ROWS - 1) / TRSV_ROWS : 1;
    const T *Dc = Dinv + (size_t)c * NBI * NBI;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TRSV_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (bi > 0 && pdl) ? 1 : 0;        // the first step follows trtri_diag in plain stream order
    T *Bp = B, *Wp = W;
    if (nrhs == 1) e = cudaLaunchKernelEx(&cfg, trsv_step_kernel<T, 1>, trans, p0, pb, rest0, cnt, A, lda, Dc, Bp, ldb, Wp, n, nrhs);
    else if (nrhs == 2) e = cudaLaunchKernelEx(&cfg, trsv_step_kernel<T, 2>, trans, p0, pb, rest0, cnt, A, lda, Dc, Bp, ldb, Wp, n, nrhs);
    else if (nrhs <= 4) e = cudaLaunchKernelEx(&cfg, trsv_step_kernel<T, 4>, trans, p0, pb, rest0, cnt, A, lda, Dc, Bp, ldb, Wp, n, nrhs);
    else e = cudaLaunchKernelEx(&cfg, trsv_step_kernel<T, 8>, trans, p0, pb, rest0, cnt, A, lda, Dc, Bp, ldb, Wp, n, nrhs);
    count_launch();
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  int rc = 0;
  if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "trsv_few launch failed: %s", cudaGetErrorString(e)); rc = JB_ERR_CUDA; }
  if (!rc) rc = lacpy_device<T>(111, n, nrhs, W, n, B, ldb, s);
  ws_free(W, s);
  ws_free(Dinv, s);
  return rc;
}

#define INST(T) template int trsm_inv_device<T>(int, int, int, int, int, int, const T *, int, T *, int, cudaStream_t); \
  template int trsv_few_device<T>(int, int, int, int, int, const T *, int, T *, int, cudaStream_t);
INST(float)
INST(double)
INST(cfloat)
INST(cdouble)

}  // namespace jb
