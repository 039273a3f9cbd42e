// Batched 32x32 Double LU factorise-and-solve on the FP64 tensor pipe (lu32.variant 50; SURVEY.md §8d config 3, the
// batched form of /root/reference/Numeric/Jalla/Matrix.hs:495-508 gesv and :763-767 getrf).
//
// The lane = row kernels of batched_lu32p.cu are bound by the LSU pipe: every entry of every pivot row has to reach 32
// lanes (a broadcast load per entry and lane-load) after ONE lane has published it (a shared-memory store costs the pipe
// ~4.4 cycles whatever its predicate).  Here the trailing matrix lives in the accumulator layout of DMMA.8x8x4 instead —
// lane (q, m) of the warp holds rows q + 8t and columns 8s + 2m, 8s + 2m + 1 of every 8x8 tile (t, s): 32 doubles — and a
// block of FOUR elimination steps is applied to it with one mma.sync per tile: C(t,s) += (-L)(t) x U(s).  A pivot-row
// entry is then loaded by ONE lane (the B fragment) instead of 32, a multiplier by one lane (the A fragment, shared by
// the four column tiles), and the only full-warp traffic left is the image of the trailing matrix that every lane stores
// after each block (row-major by PHYSICAL row in shared memory: rows are never interchanged, only `pos` moves), which is
// what lets the next block find its pivot rows at run-time addresses — something registers cannot do.
//
// Per block of four columns K0 .. K0+3:
//   1. lane = row: every lane reads its row's four panel entries from the image, and the four steps run as in the
//      block-of-four kernel — redux.max on the high word of |a| among unused rows, ballot, speculative correctly-rounded
//      reciprocal, the pivot lane's reciprocal / position / <= 3 panel entries / right-hand side by shfl.idx — producing
//      the multipliers l(row, 0..3) (an arithmetic zero from the step a row was used at on).
//   2. the panel entries go back to the image (final L and U11 values), the negated multipliers to a [32][4] table.
//   3. U12: every lane loads the four pivot rows' entries of ITS column (col = 8s + q) as they stood before the block and
//      rebuilds u0 = r0, u1 = fma(-l10,u0,r1), u2 = fma(-l21,u1,fma(-l20,u0,r2)), u3 = ... — the very operations the pivot
//      lanes' rows undergo — and keeps u_m as its B-fragment entry.
//   4. one DMMA per tile.  mma.sync.m8n8k4.f64 is the chain fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c)))) with k = 0
//      first (tools/smem_probe.cu: 262144 / 262144 samples), i.e. exactly the four rank-1 updates in step order; the pivot
//      rows themselves become their rows of U through their own (truncated) multipliers.
//   5. the lanes store the updated tiles to the image (rows used before this block are skipped: their image is final).
// The tail stays lane = physical row: every lane streams its row of the image out to its final position in global memory
// and the back substitution runs on the same loads, x_k broadcast by shfl from the lane that holds position k (inverse
// permutation through shared memory).  Generic steps only: a tie in the high word or a pivot outside the guarded
// exponent range flags the system, and flagged systems are redone from their untouched inputs by the exact one-warp
// routine (as in batched_lu32p.cu).  The operation sequence is the oracle's, so LU, ipiv and x are bit-identical to it
// (tools/time_lu32.py on 10^6 systems, tests/test_batched.py variant 50).
//
// MEASURED (B200, 10^6 systems): 7.19 ms against 5.83 ms of the shipped pair-lookahead kernel — NOT the default.  ncu
// (profiles/ncu_lu32d_r2d.txt): 64 DMMA per system keep the tensor pipe 12.6 % busy and the broadcast traffic is gone,
// but the LSU data pipe is still what the kernel sits on (70.6 % busy): per system 318 shared-memory load wavefronts,
// 414 store wavefronts (the image: 80 full-warp 128-bit stores), 320 shuffles (the lane = row panel: reciprocal,
// position, right-hand side and <= 3 panel entries per step; the substitution) and ~410 global wavefronts (the
// accumulator-layout loads touch four 128-byte lines per instruction) = 1466 against 1214 of the pair-lookahead kernel,
// and both kernels run at the same ~71 % of that pipe, i.e. time follows the LSU wavefront count.  2920 instructions per
// system at 36 % issue utilisation; 12 instead of 16 warps per SM changes 4 % (not latency-bound).
#include "batched_common.cuh"
#include "dmma_common.cuh"

namespace jb {

// The image: row r of the trailing matrix / factors is 16 chunks of 16 bytes (two columns each) at chunk address
//   r * 16 + (ch & 8) + ((ch + f(r)) & 7),   f(r) = ((r >> 1) & 3) + 4 (r & 1)
// — a rotation inside each half row.  A 128-bit shared-memory access is resolved per QUARTER warp (8 lanes), and both
// full-warp patterns of the kernel are conflict-free under it: lane = row (8 consecutive rows, one chunk: f is a
// bijection of Z8 on them) and the accumulator layout (rows q0, q0 + 1 with four consecutive chunks each: f differs by 4).
// With a plain 272-byte row stride the second pattern ran at 8 wavefronts per store instead of 4 (first ncu capture:
// 379 bank-conflict wavefronts per system).
constexpr int D_LM = 32 * 16 * 16;               // negated multipliers of the current block: chunk h * 36 + row = {-l(row, 2h), -l(row, 2h+1)}
constexpr int D_ROWOF = D_LM + (36 + 32) * 16;   // [32] physical row at position i
constexpr int D_BYTES = D_ROWOF + 32 * 4;        // 9408

__device__ __forceinline__ int d_f(int r) { return ((r >> 1) & 3) + 4 * (r & 1); }
__device__ __forceinline__ void d_sts2(uint32_t addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void d_sts2_pred(uint32_t addr, double x, double y, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(addr), "d"(x), "d"(y), "r"(pred) : "memory");
}
__device__ __forceinline__ double2 d_lds2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double d_lds1(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double d_sel(bool p, double x, double y) {
  double r;
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(x), "d"(y), "r"((int)p));
  return r;
}

// Per-lane address constants (bytes, shared window).
struct Lu32dAddr {
  uint32_t img;        // the image
  uint32_t lm;         // the multiplier table
  uint32_t row;        // this lane's own row (lane = row): img + lane * 256
  int fr;              // f(lane)
  uint32_t frag;       // accumulator layout: img + q * 256 + 16 * ((m + f(q)) & 7); tile (t, s) at + t * 2048 + 128 * (s >> 1), ^ 64 for odd s
};

template <int MODE, int J>
__device__ __forceinline__ void lu32d_block(double (&c)[32], double &bb, int &pos, int &dm, int &ip, int &bad, const Lu32dAddr &ad,
                                            const int lane) {
  constexpr unsigned RLO = 64u << 20, RSPAN = ((1981u - 64u) << 20) - 1u;   // high words whose reciprocal rcp_fast rounds correctly
  constexpr int K0 = 4 * J, S0 = J >> 1, S_LO = (K0 + 4) >> 3, CH0 = K0 / 2;
  const int q = lane >> 2, m = lane & 3;
  // ---- 1. lane = row: the panel
  const uint32_t pa0 = ad.row + 16 * ((CH0 & 8) + ((CH0 + ad.fr) & 7)), pa1 = ad.row + 16 * ((CH0 & 8) + ((CH0 + 1 + ad.fr) & 7));
  double p[4];
  {
    const double2 v0 = d_lds2(pa0), v1 = d_lds2(pa1);
    p[0] = v0.x; p[1] = v0.y; p[2] = v1.x; p[3] = v1.y;
  }
  const int fresh = dm == 0;                                   // this row has not been used before this block
  const unsigned dead0 = __ballot_sync(0xffffffffu, dm != 0);
  double l[4];
  int src[4];
#pragma unroll
  for (int jj = 0; jj < 4; jj++) {
    const int k = K0 + jj;
    const int hk = (__double2hiint(p[jj]) & 0x7fffffff) | dm;
    const int Mx = __reduce_max_sync(0xffffffffu, hk);
    const bool cl = hk == Mx;
    const unsigned ball = __ballot_sync(0xffffffffu, cl);
    bad |= ((ball & (ball - 1u)) != 0u) | ((unsigned)Mx - RLO > RSPAN);
    src[jj] = __ffs(ball) - 1;
    const double rk = rcp_fast(p[jj]);
    const double rinv = __shfl_sync(0xffffffffu, rk, src[jj]);
    const int ppos = __shfl_sync(0xffffffffu, pos, src[jj]);
    if (lane == k) ip = ppos + 1;
    pos = cl ? k : (pos == k ? ppos : pos);
    dm = cl ? -1 : dm;
    const bool act = dm == 0;
    mul_pred(p[jj], rinv, act);                                // l = a * (1/pivot) on the rows still in play
    l[jj] = d_sel(act, p[jj], 0.0);
#pragma unroll
    for (int mm = jj + 1; mm < 4; mm++) {
      const double u = __shfl_sync(0xffffffffu, p[mm], src[jj]);
      p[mm] = __fma_rn(-l[jj], u, p[mm]);
    }
    if (MODE != 0) {
      const double y = __shfl_sync(0xffffffffu, bb, src[jj]);
      bb = __fma_rn(-l[jj], y, bb);
    }
  }
  // ---- 2. final panel entries to the image (rows used before this block keep theirs), multipliers to the table
  d_sts2_pred(pa0, p[0], p[1], fresh);
  d_sts2_pred(pa1, p[2], p[3], fresh);
  if (J < 7) {
    d_sts2(ad.lm + 16 * lane, -l[0], -l[1]);
    d_sts2(ad.lm + 16 * (36 + lane), -l[2], -l[3]);
    __syncwarp();
    // ---- 3. fragments: A = (-L)(rows q + 8t, k = m); B = U12(k = m, col = 8s + q) rebuilt from the pivot rows' images
    double af[4];
    const uint32_t afa = ad.lm + 16 * (36 * (m >> 1) + q) + 8 * (m & 1);
#pragma unroll
    for (int t = 0; t < 4; t++) af[t] = d_lds1(afa + 128 * t);
    const double nl10 = d_lds1(ad.lm + 16 * src[1]);
    const double2 nl2 = d_lds2(ad.lm + 16 * src[2]);
    const double2 nl3 = d_lds2(ad.lm + 16 * src[3]);
    const double nl32 = d_lds1(ad.lm + 16 * (36 + src[3]));
    // pivot row k, column 8s + q: chunk 4s + (q >> 1), word q & 1
    uint32_t ra[4];
#pragma unroll
    for (int k = 0; k < 4; k++) ra[k] = ad.img + 256 * src[k] + 16 * (((q >> 1) + d_f(src[k])) & 7) + 8 * (q & 1);
    int ok[4];
#pragma unroll
    for (int t = 0; t < 4; t++) ok[t] = ((dead0 >> (q + 8 * t)) & 1u) == 0u;
#pragma unroll
    for (int s = S_LO; s < 4; s++) {
      const uint32_t so = 128 * (s >> 1), sx = (s & 1) ? 64u : 0u;      // half row, rotation by four chunks
      const double u0 = d_lds1((ra[0] ^ sx) + so), v1 = d_lds1((ra[1] ^ sx) + so), v2 = d_lds1((ra[2] ^ sx) + so), v3 = d_lds1((ra[3] ^ sx) + so);
      const double u1 = __fma_rn(nl10, u0, v1);
      const double u2 = __fma_rn(nl2.y, u1, __fma_rn(nl2.x, u0, v2));
      const double u3 = __fma_rn(nl32, u2, __fma_rn(nl3.y, u1, __fma_rn(nl3.x, u0, v3)));
      const double bf = d_sel(m == 0, u0, d_sel(m == 1, u1, d_sel(m == 2, u2, u3)));
      // ---- 4. the four rank-1 updates of the block, one DMMA per tile
#pragma unroll
      for (int t = 0; t < 4; t++) dmma::dmma884(c[(t * 4 + s) * 2], c[(t * 4 + s) * 2 + 1], af[t], bf);
      // ---- 5. updated tiles to the image (in the panel's own tile only the lanes right of the panel)
      const int keep = s > S0 || m >= 2;
#pragma unroll
      for (int t = 0; t < 4; t++)
        d_sts2_pred((ad.frag ^ sx) + so + 2048 * t, c[(t * 4 + s) * 2], c[(t * 4 + s) * 2 + 1], ok[t] && keep);
    }
  }
  __syncwarp();
}
template <int MODE, int J> struct Lu32dBlocks {
  static __device__ __forceinline__ void run(double (&c)[32], double &bb, int &pos, int &dm, int &ip, int &bad, const Lu32dAddr &ad, const int lane) {
    lu32d_block<MODE, J>(c, bb, pos, dm, ip, bad, ad, lane);
    Lu32dBlocks<MODE, J + 1>::run(c, bb, pos, dm, ip, bad, ad, lane);
  }
};
template <int MODE> struct Lu32dBlocks<MODE, 8> {
  static __device__ __forceinline__ void run(double (&)[32], double &, int &, int &, int &, int &, const Lu32dAddr &, int) {}
};

// Tail (lane = physical row): column pair (C, C+1) of this lane's row goes out to its final position, then the
// substitution steps k = C+1, C with x_k broadcast from the lane that holds position k.
template <int MODE, int C>
struct Lu32dTail {
  static __device__ __forceinline__ void run(const Lu32dAddr &ad, const int *rowof, double *Aout, double &y, const double rinv, const int pos, int4 &ro4) {
    constexpr int CH = C / 2;
    const double2 v = d_lds2(ad.row + 16 * ((CH & 8) + ((CH + ad.fr) & 7)));
    if (MODE != 2) { __stcs(Aout + N32 * C, v.x); __stcs(Aout + N32 * (C + 1), v.y); }
    if (MODE != 0) {
      if ((C & 3) == 2) ro4 = *reinterpret_cast<const int4 *>(&rowof[C - 2]);      // rows at positions C-2 .. C+1
      const int s1 = (C & 3) == 2 ? ro4.w : ro4.y, s0 = (C & 3) == 2 ? ro4.z : ro4.x;
      const double x1 = __shfl_sync(0xffffffffu, __dmul_rn(y, rinv), s1);
      fma_pred(y, -v.y, x1, pos < C + 1);
      const double x0 = __shfl_sync(0xffffffffu, __dmul_rn(y, rinv), s0);
      fma_pred(y, -v.x, x0, pos < C);
    }
    Lu32dTail<MODE, C - 2>::run(ad, rowof, Aout, y, rinv, pos, ro4);
  }
};
template <int MODE> struct Lu32dTail<MODE, -2> {
  static __device__ __forceinline__ void run(const Lu32dAddr &, const int *, double *, double &, const double, const int, int4 &) {}
};

template <int MODE, int MINB>
__global__ void __launch_bounds__(32, MINB)
lu32d_kernel(double *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, double *__restrict__ Bg, long long stride_b,
             double *__restrict__ Xg, long long stride_x, int *__restrict__ info_g, int *__restrict__ redo, int nsys) {
  __shared__ __align__(128) unsigned char sm[D_BYTES];
  const int lane = threadIdx.x, q = lane >> 2, m = lane & 3;
  int *const rowof = reinterpret_cast<int *>(sm + D_ROWOF);
  Lu32dAddr ad;
  ad.img = (uint32_t)__cvta_generic_to_shared(sm);
  ad.lm = ad.img + D_LM;
  ad.row = ad.img + 256 * lane;
  ad.fr = d_f(lane);
  ad.frag = ad.img + 256 * q + 16 * ((m + d_f(q)) & 7);
  for (long long sys = blockIdx.x; sys < nsys; sys += gridDim.x) {
    double *const A = Ag + sys * stride_a;
    // ---- load straight into the accumulator layout (8 rows x 4 column pairs per instruction: 64-byte runs) and lay the
    //      image down for the first block
    double c[32];
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
      for (int s = 0; s < 4; s++) {
        c[(t * 4 + s) * 2] = __ldcs(A + (q + 8 * t) + N32 * (8 * s + 2 * m));
        c[(t * 4 + s) * 2 + 1] = __ldcs(A + (q + 8 * t) + N32 * (8 * s + 2 * m + 1));
      }
    double bb = 0.0;
    if (MODE != 0) bb = __ldcs(Bg + sys * stride_b + lane);
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
      for (int s = 0; s < 4; s++)
        d_sts2((ad.frag ^ ((s & 1) ? 64u : 0u)) + 128 * (s >> 1) + 2048 * t, c[(t * 4 + s) * 2], c[(t * 4 + s) * 2 + 1]);
    __syncwarp();
    int pos = lane, dm = 0, bad = 0, ip = 0;
    Lu32dBlocks<MODE, 0>::run(c, bb, pos, dm, ip, bad, ad, lane);
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    if (badm == 0) {
      rowof[pos] = lane;
      __syncwarp();
      double y = bb, rinv = 0.0;
      if (MODE != 0) rinv = rcp_fast(d_lds1(ad.row + 16 * ((((pos >> 1) & 8)) + (((pos >> 1) + ad.fr) & 7)) + 8 * (pos & 1)));   // 1 / u(pos, pos)
      int4 ro4 = make_int4(0, 0, 0, 0);
      Lu32dTail<MODE, N32 - 2>::run(ad, rowof, A + pos, y, rinv, pos, ro4);
      if (MODE != 2) ipiv_g[sys * N32 + lane] = ip;
      if (MODE == 1) Bg[sys * stride_b + pos] = __dmul_rn(y, rinv);
      if (MODE == 2) Xg[sys * stride_x + pos] = __dmul_rn(y, rinv);
      if (lane == 0 && info_g) info_g[sys] = 0;
    } else if (lane == 0) {
      redo[1 + atomicAdd(redo, 1)] = (int)sys;
    }
    __syncwarp();
  }
}

template <int MODE>
void lu32d_run(int shape, double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info, int *redo,
               int nsys, cudaStream_t s) {
  const int lim = option("lu32.ctas_per_sm");
  if (shape == 12) {
    auto fn = lu32d_kernel<MODE, 12>;
    JB_ONCE_PER_DEVICE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    const long cap = (long)num_sms() * (lim > 0 && lim < 12 ? lim : 12);
    fn<<<(int)(nsys < cap ? nsys : cap), 32, 0, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo, nsys);
  } else {
    auto fn = lu32d_kernel<MODE, 16>;
    JB_ONCE_PER_DEVICE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    const long cap = (long)num_sms() * (lim > 0 && lim < 16 ? lim : 16);
    fn<<<(int)(nsys < cap ? nsys : cap), 32, 0, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo, nsys);
  }
  count_launch();
}
template void lu32d_run<0>(int, double *, long long, int *, double *, long long, double *, long long, int *, int *, int, cudaStream_t);
template void lu32d_run<1>(int, double *, long long, int *, double *, long long, double *, long long, int *, int *, int, cudaStream_t);
template void lu32d_run<2>(int, double *, long long, int *, double *, long long, double *, long long, int *, int *, int, cudaStream_t);

}  // namespace jb
