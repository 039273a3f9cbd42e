// Batched 32x32 Double LU factorise-and-solve, the headline batched case (the batched form of
// /root/reference/Numeric/Jalla/Matrix.hs:495-508 gesv and :763-767 getrf).  Third-generation kernel.
//
// What bounds the one-warp-per-system kernel (batched_small.cu, lu32v2) is not HBM but the shared-memory pipe and the
// issue slots: every entry of every pivot row is broadcast to the 32 rows of its system once (528 doubles per system,
// one shared-memory wavefront per double and lane-load) and every elimination step carries ~80 instructions of
// selection / bookkeeping.  This kernel changes the mapping instead of the schedule:
//   * S = 2: a HALF warp owns a system, a lane owns two adjacent rows (2r, 2r+1) — so a broadcast operand feeds two FMAs
//     per lane, one instruction stream serves two systems (per-system selection overhead halves) and the column loads
//     are 128-bit (16 lanes x 16 B = one 256-byte column).  S = 1 (lane = row) is kept as a template point for comparison.
//   * the fast path handles only the generic step: one redux.max on the high word of |a|, a ballot, and a count.  A tie
//     in the high word, a pivot outside the correctly-rounded range of the fast reciprocal (zero, denormal, Inf, NaN)
//     flags the system; flagged systems are redone from their untouched inputs by the exact one-warp routine
//     (lu32_one, the ?getf2 sequence with lowest-index tie-break and info).  For continuous data that is ~5e-4 of systems.
//   * everything a step shares — the pivot row, its right-hand-side entry, its reciprocal, its position — goes through
//     shared memory with predicated 128-bit stores by the pivot lane and broadcast loads by all: no select trees.
//   * finished rows are predicated off (not multiplied by zero), so NaN/Inf never leak across rows.
//   * the per-step records stay in shared memory: position order y (then x), reciprocals, interchanges, so the back
//     substitution needs no per-lane saved scalars and x / ipiv leave as vector stores.
// Arithmetic is the oracle's sequence (r = 1/pivot correctly rounded, l = a*r, a_ij = fma(-l,u,a_ij), x_k = y_k*r_k), so
// factors, pivots and solutions are bit-identical to oracle/oracle_impl.inc.
#include "batched_common.cuh"

namespace jb {

constexpr int P_ROW = 0;        // 2 x 32 doubles: the published pivot row, double buffered by step parity
constexpr int P_HDR = 512;      // 32 x {y_k (later x_k), 1/u_kk}
constexpr int P_POS = 1024;     // 32 x int: position of the pivot row before step k's interchange (ipiv[k] - 1)
constexpr int P_SYS_BYTES = 1216;   // + 64 B so the two systems of a warp sit in different banks

__device__ __forceinline__ void sts1_pred(void *p, double x, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.f64 [%0], %1; }"
               ::"r"((uint32_t)__cvta_generic_to_shared(p)), "d"(x), "r"(pred) : "memory");
}
__device__ __forceinline__ void sts_int_pred(void *p, int x, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.b32 [%0], %1; }"
               ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(x), "r"(pred) : "memory");
}
__device__ __forceinline__ void stg_cs_pred(double *p, double x, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.global.cs.f64 [%0], %1; }" ::"l"(p), "d"(x), "r"(pred) : "memory");
}
// select on a double as two 32-bit selects (what ptxas emits anyway), kept explicit so that the masks below stay
// two instructions per row and step
__device__ __forceinline__ double dsel(bool p, double x, double y) {
  double r;       // (asm: with `p ? a[1][j] : a[0][j]` nvcc turns the register array into a dynamically indexed local one)
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(x), "d"(y), "r"((int)p));
  return r;
}

// S systems per warp (32/S lanes each, S rows per lane); `ngroups` groups of S consecutive systems.
// REALIGN: the warps of a CTA re-converge once per group so they walk the unrolled body together (instruction cache);
// ngroups must then be a multiple of WARPS.  Systems the fast path cannot finish exactly are appended to `redo`
// (count in redo[0], indices from redo[1]) with their outputs untouched.
// ABL (profiling only, results are then wrong): 1 skip the row publication, 2 skip the broadcast loads, 4 skip the
// trailing FMAs, 8 skip the global stores, 16 skip the speculative reciprocal.
template <int S, int MODE, int WARPS, int MINB, bool REALIGN, int ABL = 0>
__global__ void __launch_bounds__(32 * WARPS, MINB)
lu32p_kernel(double *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, double *__restrict__ Bg, long long stride_b,
             double *__restrict__ Xg, long long stride_x, int *__restrict__ info_g, int *__restrict__ redo, int ngroups) {
  static_assert(S == 1 || S == 2, "one or two systems per warp");
  constexpr int LPS = 32 / S;
  __shared__ __align__(16) unsigned char smem_all[WARPS * S * P_SYS_BYTES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane / LPS, r = lane % LPS;
  unsigned char *sm = smem_all + (warp * S + h) * P_SYS_BYTES;
  double *const srow = reinterpret_cast<double *>(sm + P_ROW);
  double2 *const shdr = reinterpret_cast<double2 *>(sm + P_HDR);
  int *const spos = reinterpret_cast<int *>(sm + P_POS);
  const unsigned mym = S == 2 ? (h ? 0xffff0000u : 0x0000ffffu) : 0xffffffffu;
  const int other = S == 2 ? (h ? (int)0x80000000 : 0) : 0;      // OR-mask that makes system 1's lanes negative in system 0's maximum
  constexpr unsigned RLO = 64u << 20, RSPAN = ((1981u - 64u) << 20) - 1u;   // high words whose reciprocal rcp_fast rounds correctly

  for (long long g0 = (long long)blockIdx.x * WARPS; g0 < ngroups; g0 += (long long)gridDim.x * WARPS) {
    long long g = g0 + warp;
    if (REALIGN) __syncthreads();
    else if (g >= ngroups) g = ngroups - 1;          // a surplus warp repeats the last group (same values rewritten)
    const long long sys = g * S + h;
    double *const A = Ag + sys * stride_a;
    double a[S][N32], bb[S];
    int pos[S], dm[S];
    // ---- load: lane r of a system holds rows S*r .. S*r+S-1; one 256-byte column per system and instruction
#pragma unroll
    for (int j = 0; j < N32; j++) {
      if (S == 2) { const double2 v = __ldcs(reinterpret_cast<const double2 *>(A + 2 * r + N32 * j)); a[0][j] = v.x; a[S - 1][j] = v.y; }
      else a[0][j] = __ldcs(A + r + N32 * j);
    }
#pragma unroll
    for (int s = 0; s < S; s++) { bb[s] = 0.0; pos[s] = S * r + s; dm[s] = 0; }
    if (MODE != 0) {
      if (S == 2) { const double2 v = __ldcs(reinterpret_cast<const double2 *>(Bg + sys * stride_b + 2 * r)); bb[0] = v.x; bb[S - 1] = v.y; }
      else bb[0] = __ldcs(Bg + sys * stride_b + r);
    }
    int bad = 0;
#pragma unroll
    for (int k = 0; k < N32; k++) {
      double *const sr = srow + (k & 1) * N32;
      // ---- selection: maximum of the high words of |a| over the rows not yet used (dm = sign bit for used rows)
      int hk[S];
#pragma unroll
      for (int s = 0; s < S; s++) hk[s] = (__double2hiint(a[s][k]) & 0x7fffffff) | dm[s];
      int hm = hk[0];
      if (S == 2) hm = max(hk[0], hk[S - 1]);
      int M;
      if (S == 2) {
        const int m0 = __reduce_max_sync(0xffffffffu, hm | other);                    // system 0: lanes 16..31 are negative
        const int m1 = __reduce_max_sync(0xffffffffu, hm | (other ^ (int)0x80000000));   // system 1
        M = h ? m1 : m0;
      } else {
        M = __reduce_max_sync(0xffffffffu, hm);
      }
      // exactly one winner per system, else (a tie in the high word) the system is redone exactly
      const bool sb = S == 2 && hk[S - 1] > hk[0];          // this lane's better candidate sits in slot 1
      const bool cl = hm == M;                               // this lane holds the pivot row (in its better slot)
      const unsigned ball = __ballot_sync(0xffffffffu, cl) & mym;
      bad |= ((ball & (ball - 1u)) != 0u) | ((unsigned)M - RLO > RSPAN);
      if (S == 2) bad |= (hk[0] == hk[S - 1]) & cl;
      int c[S];
#pragma unroll
      for (int s = 0; s < S; s++) c[s] = cl && (S == 1 || sb == (s == 1));
      // ---- speculative reciprocal of this lane's better candidate; only the pivot lane's is published
      const double rk = (ABL & 16) ? 1.25 : rcp_fast(S == 2 ? dsel(sb, a[S - 1][k], a[0][k]) : a[0][k]);
      // ---- the pivot lane publishes its row (columns > k), right-hand side, reciprocal and position.  A shared-memory
      //      store occupies the LSU front end for ~4.4 cycles whatever its predicate says (tools/smem_probe.cu), so with
      //      two rows per lane the row is selected first (it depends only on this lane's own comparison, not on the
      //      reduction) and ONE store per column pair serves both systems of the warp.
      if (!(ABL & 1))
#pragma unroll
      for (int j = (k + 1) & ~1; j < N32; j += 2)
        sts2_pred(&sr[j], S == 2 ? dsel(sb, a[S - 1][j], a[0][j]) : a[0][j], S == 2 ? dsel(sb, a[S - 1][j + 1], a[0][j + 1]) : a[0][j + 1], cl);
      sts2_pred(reinterpret_cast<double *>(&shdr[k]), S == 2 ? dsel(sb, bb[S - 1], bb[0]) : bb[0], rk, cl);
      sts_int_pred(&spos[k], sb ? pos[S - 1] : pos[0], cl);
      __syncwarp();
      const double2 hd = shdr[k];
      const int ppos = spos[k];
      // ---- implicit interchange: the row sitting at position k moves to the pivot's old position, the pivot row to k.
      //      Rows already used keep their registers: their multiplier is an arithmetic zero (a non-finite entry that
      //      0 * x could leak into them also reaches a later pivot search and flags the system).
      double lm[S];
#pragma unroll
      for (int s = 0; s < S; s++) {
        pos[s] = c[s] ? k : (pos[s] == k ? ppos : pos[s]);
        dm[s] = c[s] ? -1 : dm[s];
        const bool act = dm[s] == 0;
        const double t = __dmul_rn(a[s][k], hd.y);             // l = a * (1/pivot)
        a[s][k] = dsel(act, t, a[s][k]);
        lm[s] = dsel(act, t, 0.0);
      }
      if (k + 1 < N32) {
#pragma unroll
        for (int j = (k + 1) & ~1; j < N32; j += 2) {
          const double2 u = (ABL & 2) ? make_double2(hd.x, hd.y) : *reinterpret_cast<const double2 *>(&sr[j]);
          if ((ABL & 4) && j > k + 1) continue;
#pragma unroll
          for (int s = 0; s < S; s++) {
            if (j > k) a[s][j] = __fma_rn(-lm[s], u.x, a[s][j]);
            a[s][j + 1] = __fma_rn(-lm[s], u.y, a[s][j + 1]);
          }
        }
      }
      if (MODE != 0) {
#pragma unroll
        for (int s = 0; s < S; s++) bb[s] = __fma_rn(-lm[s], hd.x, bb[s]);
      }
    }
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    const int ok = (badm & mym) == 0;
    if (MODE != 0) {
      // ---- back substitution, column sweep on the implicitly permuted U; shdr[k].x turns from y_k into x_k.  A row's
      //      right-hand side is dead once its x is published, so nothing here needs a mask.
      double rinv[S];
#pragma unroll
      for (int s = 0; s < S; s++) { const double2 t = shdr[pos[s]]; bb[s] = t.x; rinv[s] = t.y; }
      __syncwarp();
#pragma unroll
      for (int k = N32 - 1; k >= 0; k--) {
        if (S == 2) {
          const bool o1 = pos[S - 1] == k;
          sts1_pred(&shdr[k].x, dsel(o1, __dmul_rn(bb[S - 1], rinv[S - 1]), __dmul_rn(bb[0], rinv[0])), o1 || pos[0] == k);
        } else {
          sts1_pred(&shdr[k].x, __dmul_rn(bb[0], rinv[0]), pos[0] == k);
        }
        __syncwarp();
        const double xk = shdr[k].x;
#pragma unroll
        for (int s = 0; s < S; s++) bb[s] = __fma_rn(-a[s][k], xk, bb[s]);
      }
    }
    // ---- results: rows to their final positions, interchanges and x as vector stores from the per-step records
    if (MODE != 2) {
      if (ABL & 8) {
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int j = 0; j < N32; j++) acc += a[s][j];
        if (acc == 1.2345) A[0] = acc;
      } else if (badm == 0) {
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int j = 0; j < N32; j++) __stcs(A + pos[s] + N32 * j, a[s][j]);
      } else {
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int j = 0; j < N32; j++) stg_cs_pred(A + pos[s] + N32 * j, a[s][j], ok);
      }
      if (ok) {
        if (S == 2) {
          const int2 p = *reinterpret_cast<const int2 *>(&spos[2 * r]);
          *reinterpret_cast<int2 *>(ipiv_g + sys * N32 + 2 * r) = make_int2(p.x + 1, p.y + 1);
        } else {
          ipiv_g[sys * N32 + r] = spos[r] + 1;
        }
      }
    }
    if (ok) {
      if (MODE != 0) {
        double *const xo = MODE == 1 ? Bg + sys * stride_b : Xg + sys * stride_x;
        if (S == 2) *reinterpret_cast<double2 *>(xo + 2 * r) = make_double2(shdr[2 * r].x, shdr[2 * r + 1].x);
        else xo[r] = shdr[r].x;
      }
      if (r == 0 && info_g) info_g[sys] = 0;
    } else if (r == 0 && (REALIGN || g0 + warp < ngroups)) {
      redo[1 + atomicAdd(redo, 1)] = (int)sys;
    }
    __syncwarp();
  }
}

// One elimination step of the shuffle-broadcast kernel, K a compile-time constant (the step loop is unrolled by template
// recursion: with `#pragma unroll` nvcc left the column loops of the two-rows-per-lane instantiation rolled and the
// register arrays went to local memory).
template <int S, int MODE, int K>
__device__ __forceinline__ void lu32s_step(double (&a)[S][N32], double (&bb)[S], double (&my_rinv)[S], int (&pos)[S], int (&dm)[S], int (&ip)[S],
                                           int &bad, const int h, const int r, const int lane, const unsigned mym, const int other) {
  constexpr int k = K;
  constexpr unsigned RLO = 64u << 20, RSPAN = ((1981u - 64u) << 20) - 1u;
  int hk[S];
#pragma unroll
  for (int s = 0; s < S; s++) hk[s] = (__double2hiint(a[s][k]) & 0x7fffffff) | dm[s];
  const bool sb = S == 2 && hk[S - 1] > hk[0];          // this lane's better candidate sits in slot 1
  const int hm = sb ? hk[S - 1] : hk[0];
  int M;
  if (S == 2) {
    const int m0 = __reduce_max_sync(0xffffffffu, hm | other);
    const int m1 = __reduce_max_sync(0xffffffffu, hm | (other ^ (int)0x80000000));
    M = h ? m1 : m0;
  } else {
    M = __reduce_max_sync(0xffffffffu, hm);
  }
  const bool cl = hm == M;
  const unsigned ball = __ballot_sync(0xffffffffu, cl) & mym;
  bad |= (ball & (ball - 1u)) != 0u | ((unsigned)M - RLO > RSPAN);
  if (S == 2) bad |= (hk[0] == hk[S - 1]) & cl;
  const int src = __ffs(ball) - 1;                       // the pivot lane of this lane's system
  // ---- the pivot row as the source lane sees it (slot select), its reciprocal, right-hand side and position
  const double rk = rcp_fast(S == 2 ? dsel(sb, a[S - 1][k], a[0][k]) : a[0][k]);
  const double rinv = __shfl_sync(0xffffffffu, rk, src);
  const int ppos = __shfl_sync(0xffffffffu, sb ? pos[S - 1] : pos[0], src);
  double lm[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    const bool c = cl && (S == 1 || sb == (s == 1));
    pos[s] = c ? k : (pos[s] == k ? ppos : pos[s]);
    dm[s] = c ? -1 : dm[s];
    if (MODE != 0) my_rinv[s] = dsel(c, rk, my_rinv[s]);
    const bool act = dm[s] == 0;
    const double t = __dmul_rn(a[s][k], rinv);
    a[s][k] = dsel(act, t, a[s][k]);
    lm[s] = dsel(act, t, 0.0);
  }
  if (S == 1) { if (lane == k) ip[0] = ppos + 1; }
  else if ((k >> 1) == r) ip[k & 1] = ppos + 1;
#pragma unroll
  for (int j = k + 1; j < N32; j++) {
    const double u = __shfl_sync(0xffffffffu, S == 2 ? dsel(sb, a[S - 1][j], a[0][j]) : a[0][j], src);
#pragma unroll
    for (int s = 0; s < S; s++) a[s][j] = __fma_rn(-lm[s], u, a[s][j]);
  }
  if (MODE != 0) {
    const double y = __shfl_sync(0xffffffffu, S == 2 ? dsel(sb, bb[S - 1], bb[0]) : bb[0], src);
#pragma unroll
    for (int s = 0; s < S; s++) bb[s] = __fma_rn(-lm[s], y, bb[s]);
  }

}
template <int S, int MODE, int K> struct Lu32sSteps {
  static __device__ __forceinline__ void run(double (&a)[S][N32], double (&bb)[S], double (&my_rinv)[S], int (&pos)[S], int (&dm)[S], int (&ip)[S],
                                             int &bad, const int h, const int r, const int lane, const unsigned mym, const int other) {
    lu32s_step<S, MODE, K>(a, bb, my_rinv, pos, dm, ip, bad, h, r, lane, mym, other);
    Lu32sSteps<S, MODE, K + 1>::run(a, bb, my_rinv, pos, dm, ip, bad, h, r, lane, mym, other);
  }
};
template <int S, int MODE> struct Lu32sSteps<S, MODE, N32> {
  static __device__ __forceinline__ void run(double (&)[S][N32], double (&)[S], double (&)[S], int (&)[S], int (&)[S], int (&)[S], int &, int, int, int, unsigned, int) {}
};

// ---------------------------------------------------------------------------------------------------------------------
// Shuffle-broadcast variant.  Measured on B200 (tools/smem_probe.cu, profiles/smem_probe_r2.jsonl): a shared-memory
// store costs the LSU front end ~4.4 cycles per STS.128 however many lanes are predicated on (2.2 cycles per double
// published by ONE lane), a broadcast LDS.128 2.0 cycles, a SHFL 1.0 — so publishing a pivot row through shared memory
// is 3.2 LSU cycles per double and the lu32p / lu32v2 kernels are bound by exactly that (1820 cycles per system and SM
// measured, 1840 predicted).  Here the pivot row never touches shared memory: every lane reads it straight out of the
// pivot lane's registers with shfl.idx (2 cycles per double; with S = 2 the two half-warps read two different source
// lanes in the same instruction, 1 cycle per double and system).
template <int S, int MODE, int WARPS, int MINB, bool REALIGN>
__global__ void __launch_bounds__(32 * WARPS, MINB)
lu32s_kernel(double *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, double *__restrict__ Bg, long long stride_b,
             double *__restrict__ Xg, long long stride_x, int *__restrict__ info_g, int *__restrict__ redo, int ngroups) {
  static_assert(S == 1 || S == 2, "one or two systems per warp");
  constexpr int LPS = 32 / S;
  __shared__ __align__(16) double xs_all[WARPS * S][N32 + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane / LPS, r = lane % LPS;
  double *const xs = xs_all[warp * S + h];
  const unsigned mym = S == 2 ? (h ? 0xffff0000u : 0x0000ffffu) : 0xffffffffu;
  const int other = S == 2 ? (h ? (int)0x80000000 : 0) : 0;
  constexpr unsigned RLO = 64u << 20, RSPAN = ((1981u - 64u) << 20) - 1u;

  for (long long g0 = (long long)blockIdx.x * WARPS; g0 < ngroups; g0 += (long long)gridDim.x * WARPS) {
    long long g = g0 + warp;
    if (REALIGN) __syncthreads();
    else if (g >= ngroups) g = ngroups - 1;
    const long long sys = g * S + h;
    double *const A = Ag + sys * stride_a;
    double a[S][N32], bb[S], my_rinv[S];
    int pos[S], dm[S], ip[S];
#pragma unroll
    for (int j = 0; j < N32; j++) {
      if (S == 2) { const double2 v = __ldcs(reinterpret_cast<const double2 *>(A + 2 * r + N32 * j)); a[0][j] = v.x; a[S - 1][j] = v.y; }
      else a[0][j] = __ldcs(A + r + N32 * j);
    }
#pragma unroll
    for (int s = 0; s < S; s++) { bb[s] = 0.0; my_rinv[s] = 0.0; pos[s] = S * r + s; dm[s] = 0; ip[s] = 0; }
    if (MODE != 0) {
      if (S == 2) { const double2 v = __ldcs(reinterpret_cast<const double2 *>(Bg + sys * stride_b + 2 * r)); bb[0] = v.x; bb[S - 1] = v.y; }
      else bb[0] = __ldcs(Bg + sys * stride_b + r);
    }
    int bad = 0;
    Lu32sSteps<S, MODE, 0>::run(a, bb, my_rinv, pos, dm, ip, bad, h, r, lane, mym, other);
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    const int ok = (badm & mym) == 0;
    if (MODE != 0) {
      // back substitution (column sweep): a row's right-hand side is dead once its x is published
#pragma unroll
      for (int k = N32 - 1; k >= 0; k--) {
        if (S == 2) {
          const bool o1 = pos[S - 1] == k;
          sts1_pred(&xs[k], dsel(o1, __dmul_rn(bb[S - 1], my_rinv[S - 1]), __dmul_rn(bb[0], my_rinv[0])), o1 || pos[0] == k);
        } else {
          sts1_pred(&xs[k], __dmul_rn(bb[0], my_rinv[0]), pos[0] == k);
        }
        __syncwarp();
        const double xk = xs[k];
#pragma unroll
        for (int s = 0; s < S; s++) bb[s] = __fma_rn(-a[s][k], xk, bb[s]);
      }
    }
    if (MODE != 2) {
      if (badm == 0) {
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int j = 0; j < N32; j++) __stcs(A + pos[s] + N32 * j, a[s][j]);
      } else {
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int j = 0; j < N32; j++) stg_cs_pred(A + pos[s] + N32 * j, a[s][j], ok);
      }
      if (ok) {
        if (S == 2) *reinterpret_cast<int2 *>(ipiv_g + sys * N32 + 2 * r) = make_int2(ip[0], ip[S - 1]);
        else ipiv_g[sys * N32 + r] = ip[0];
      }
    }
    if (ok) {
      if (MODE != 0) {
        double *const xo = MODE == 1 ? Bg + sys * stride_b : Xg + sys * stride_x;
        if (S == 2) *reinterpret_cast<double2 *>(xo + 2 * r) = *reinterpret_cast<const double2 *>(&xs[2 * r]);
        else xo[r] = xs[r];
      }
      if (r == 0 && info_g) info_g[sys] = 0;
    } else if (r == 0 && (REALIGN || g0 + warp < ngroups)) {
      redo[1 + atomicAdd(redo, 1)] = (int)sys;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Pair-lookahead variant (lane = row, one system per warp).  The one-system kernels above sit on the LSU front end
// (tools/smem_probe.cu: a shared-memory store costs ~4.4 cycles per STS.128 whatever its predicate, a broadcast LDS.128
// 2.0; occupancy sweep in profiles/lu32_r2_notes.md: 8 -> 16 warps per SM buys 15 %), and what fills it is the pivot row
// published by ONE lane, two doubles per store.  Here two elimination steps share one publication: after step k has
// updated only column k+1, step k+1's pivot is already known, and the two pivot lanes store their rows with the same
// instruction — lane p0 its final row k, lane p1 its row as it stands BEFORE step k.  Every lane then rebuilds row k+1
// itself, u1_j = fma(-lambda, u0_j, r1_j) with lambda = l(p1,k) — the very operation lane p1 would have done, so the
// result is bit-identical — and applies both updates a_ij = fma(-l1, u1_j, fma(-l0, u0_j, a_ij)).  Half the row stores
// (the expensive half of the traffic) for one extra FMA per column and step pair.
constexpr int Q_ROWS = 0;          // [parity][2 rows][34]: columns + {rhs, pad}, double buffered per step pair
constexpr int Q_ROW_LD = 34;
constexpr int Q_HDR = 2 * 2 * Q_ROW_LD * 8;          // 32 x {1/pivot, u0(k+1) | lambda}
constexpr int Q_POS = Q_HDR + 32 * 16;               // 32 x int
constexpr int Q_XS = Q_POS + 32 * 4;                 // 32 x double (back substitution)
constexpr int Q_BYTES = Q_XS + 32 * 8 + 64;

template <int MODE, int K>
__device__ __forceinline__ void lu32q_pair(double (&a)[N32], double &bb, double &my_rinv, int &pos, int &dm, int &bad, unsigned char *sm) {
  constexpr unsigned RLO = 64u << 20, RSPAN = ((1981u - 64u) << 20) - 1u;
  double *const rows = reinterpret_cast<double *>(sm + Q_ROWS) + ((K >> 1) & 1) * 2 * Q_ROW_LD;
  double2 *const shdr = reinterpret_cast<double2 *>(sm + Q_HDR);
  int *const spos = reinterpret_cast<int *>(sm + Q_POS);
  // ---- step K: select, publish {1/pivot, pivot row's entry in column K+1}, update column K+1 only
  const int hk0 = (__double2hiint(a[K]) & 0x7fffffff) | dm;
  const int M0 = __reduce_max_sync(0xffffffffu, hk0);
  const bool c0 = hk0 == M0;
  const unsigned ball0 = __ballot_sync(0xffffffffu, c0);
  bad |= ((ball0 & (ball0 - 1u)) != 0u) | ((unsigned)M0 - RLO > RSPAN);
  const double rk0 = rcp_fast(a[K]);
  sts2_pred(reinterpret_cast<double *>(&shdr[K]), rk0, a[K + 1], c0);
  sts_int_pred(&spos[K], pos, c0);
  __syncwarp();
  const double2 h0 = shdr[K];
  const int ppos0 = spos[K];
  pos = c0 ? K : (pos == K ? ppos0 : pos);
  dm = c0 ? -1 : dm;
  const bool act0 = dm == 0;
  mul_pred(a[K], h0.x, act0);                       // l = a * (1/pivot) on the rows still in play
  const double l0 = dsel(act0, a[K], 0.0);
  a[K + 1] = __fma_rn(-l0, h0.y, a[K + 1]);
  // ---- step K+1: select, publish {1/pivot, lambda}; both pivot lanes publish their rows with one store per column pair
  const int hk1 = (__double2hiint(a[K + 1]) & 0x7fffffff) | dm;
  const int M1 = __reduce_max_sync(0xffffffffu, hk1);
  const bool c1 = hk1 == M1;
  const unsigned ball1 = __ballot_sync(0xffffffffu, c1);
  bad |= ((ball1 & (ball1 - 1u)) != 0u) | ((unsigned)M1 - RLO > RSPAN);
  const double rk1 = rcp_fast(a[K + 1]);
  sts2_pred(reinterpret_cast<double *>(&shdr[K + 1]), rk1, l0, c1);
  sts_int_pred(&spos[K + 1], pos, c1);
  double *const mine = rows + (c1 ? Q_ROW_LD : 0);
  const bool cp = c0 || c1;
#pragma unroll
  for (int j = K + 2; j < N32; j += 2) sts2_pred(&mine[j], a[j], a[j + 1], cp);
  if (MODE != 0) sts2_pred(&mine[N32], bb, 0.0, cp);
  __syncwarp();
  const double2 h1 = shdr[K + 1];
  const int ppos1 = spos[K + 1];
  pos = c1 ? K + 1 : (pos == K + 1 ? ppos1 : pos);
  dm = c1 ? -1 : dm;
  const bool act1 = dm == 0;
  mul_pred(a[K + 1], h1.x, act1);
  const double l1 = dsel(act1, a[K + 1], 0.0);
  const double lambda = h1.y;
#pragma unroll
  for (int j = K + 2; j < N32; j += 2) {
    const double2 u0 = *reinterpret_cast<const double2 *>(&rows[j]);
    const double2 r1 = *reinterpret_cast<const double2 *>(&rows[Q_ROW_LD + j]);
    const double u1x = __fma_rn(-lambda, u0.x, r1.x), u1y = __fma_rn(-lambda, u0.y, r1.y);
    a[j] = __fma_rn(-l1, u1x, __fma_rn(-l0, u0.x, a[j]));
    a[j + 1] = __fma_rn(-l1, u1y, __fma_rn(-l0, u0.y, a[j + 1]));
  }
  if (MODE != 0) {
    const double y0 = rows[N32], y1 = __fma_rn(-lambda, y0, rows[Q_ROW_LD + N32]);
    bb = __fma_rn(-l1, y1, __fma_rn(-l0, y0, bb));
  }
}
template <int MODE, int K> struct Lu32qPairs {
  static __device__ __forceinline__ void run(double (&a)[N32], double &bb, double &my_rinv, int &pos, int &dm, int &bad, unsigned char *sm) {
    lu32q_pair<MODE, K>(a, bb, my_rinv, pos, dm, bad, sm);
    Lu32qPairs<MODE, K + 2>::run(a, bb, my_rinv, pos, dm, bad, sm);
  }
};
template <int MODE> struct Lu32qPairs<MODE, N32> {
  static __device__ __forceinline__ void run(double (&)[N32], double &, double &, int &, int &, int &, unsigned char *) {}
};

// IOV == 3 (experiment, not the default): the results leave through a shared-memory image of the system's outputs (LU at
// its final row positions, x, ipiv) and three bulk asynchronous copies issued by one lane.  ncu on the plain kernel
// (profiles/ncu_lu32q_r2.txt) counts ~4.2 LSU data-pipe wavefronts per STG.64 against 2 per conflict-free STS.64, which
// suggested the detour; measured it is SLOWER at equal occupancy (6.44 against 5.82 ms per 10^6 systems, 16 warps per SM)
// and only level at 19 warps per SM / 96 registers (5.83) — a full-warp STS.64 costs this kernel more than the STG.64 it
// replaces (profiles/lu32_stage_r2d.jsonl).  Needs 16-byte aligned outputs and even strides (checked by the launcher).
constexpr int Q_STAGE_X = 1024, Q_STAGE_IPIV = 1056, Q_STAGE_DOUBLES = 1072;
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
template <int MODE, int WARPS, int MINB, bool REALIGN, int IOV = 0>
__global__ void __launch_bounds__(32 * WARPS, MINB)
lu32q_kernel(double *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, double *__restrict__ Bg, long long stride_b,
             double *__restrict__ Xg, long long stride_x, int *__restrict__ info_g, int *__restrict__ redo, int nsys) {
  __shared__ __align__(16) unsigned char smem_all[WARPS * Q_BYTES];
  __shared__ __align__(128) double stage_all[IOV == 3 ? WARPS * Q_STAGE_DOUBLES : 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *const sm = smem_all + warp * Q_BYTES;
  double *const st = stage_all + (IOV == 3 ? warp * Q_STAGE_DOUBLES : 0);
  double *const xs = IOV == 3 ? st + Q_STAGE_X : reinterpret_cast<double *>(sm + Q_XS);
  const int *const spos = reinterpret_cast<const int *>(sm + Q_POS);
  for (long long g0 = (long long)blockIdx.x * WARPS; g0 < nsys; g0 += (long long)gridDim.x * WARPS) {
    long long sys = g0 + warp;
    if (REALIGN) __syncthreads();
    else if (sys >= nsys) sys = nsys - 1;
    double *const A = Ag + sys * stride_a;
    double a[N32], bb = 0.0, my_rinv = 0.0;
#pragma unroll
    for (int j = 0; j < N32; j++) a[j] = IOV == 1 ? A[lane + N32 * j] : __ldcs(A + lane + N32 * j);
    if (MODE != 0) bb = __ldcs(Bg + sys * stride_b + lane);
    if (IOV == 2) {       // pull the next system of this warp towards L2 while this one is factorised
      const long long nxt = sys + (long long)gridDim.x * WARPS;
      if (nxt < nsys) {
        const char *pa = reinterpret_cast<const char *>(Ag + nxt * stride_a) + lane * 256;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 128));
      }
    }
    int pos = lane, dm = 0, bad = 0;
    Lu32qPairs<MODE, 0>::run(a, bb, my_rinv, pos, dm, bad, sm);
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    // 1 / u_kk of this lane's row: the header of the step that chose it (one load instead of two selects per step)
    if (MODE != 0) my_rinv = reinterpret_cast<const double2 *>(sm + Q_HDR)[pos].x;
    if (IOV == 3) {       // the previous system's copies have long finished reading the image; make it formal
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    if (MODE != 0) {
#pragma unroll
      for (int k = N32 - 1; k >= 0; k--) {
        sts1_pred(&xs[k], __dmul_rn(bb, my_rinv), pos == k);
        __syncwarp();
        bb = __fma_rn(-a[k], xs[k], bb);
      }
    }
    if (badm == 0) {
      if (IOV == 3) {
        if (MODE != 2) {
#pragma unroll
          for (int j = 0; j < N32; j++) st[pos + N32 * j] = a[j];
          reinterpret_cast<int *>(st + Q_STAGE_IPIV)[lane] = spos[lane] + 1;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (MODE != 2) {
            bulk_s2g(A, st, N32 * N32 * 8);
            bulk_s2g(ipiv_g + sys * N32, st + Q_STAGE_IPIV, N32 * 4);
          }
          if (MODE == 1) bulk_s2g(Bg + sys * stride_b, xs, N32 * 8);
          if (MODE == 2) bulk_s2g(Xg + sys * stride_x, xs, N32 * 8);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (info_g) info_g[sys] = 0;
        }
      } else {
        if (MODE != 2) {
#pragma unroll
          for (int j = 0; j < N32; j++) { if (IOV == 1) A[pos + N32 * j] = a[j]; else __stcs(A + pos + N32 * j, a[j]); }
          ipiv_g[sys * N32 + lane] = spos[lane] + 1;
        }
        if (MODE == 1) Bg[sys * stride_b + lane] = xs[lane];
        if (MODE == 2) Xg[sys * stride_x + lane] = xs[lane];
        if (lane == 0 && info_g) info_g[sys] = 0;
      }
    } else if (lane == 0 && (REALIGN || g0 + warp < nsys)) {
      redo[1 + atomicAdd(redo, 1)] = (int)sys;
    }
    __syncwarp();
  }
  if (IOV == 3) {         // the image must outlive the copies that read it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Block-of-four variant (lane = row, one system per warp).  The kernels above pay ~4.4 LSU cycles for every 128-bit
// shared-memory store although only the pivot lane's predicate is on (tools/smem_probe.cu) — the row publication is the
// largest item on the saturated pipe (profiles/lu32_r2_notes.md).  Here NOBODY publishes a pivot row.  Once per block
// of four elimination steps every lane stores what is left of its own row (columns right of the block and the
// right-hand side): the same store instruction now carries 32 rows instead of one.  The four steps of the block touch
// only the block's own columns — pivot value, interchange position and the pivot row's (at most three) block entries
// travel by shfl.idx from the pivot lane, which the ballot of the search already names.  After the block every lane
// reads the four pivot rows AS THEY STOOD BEFORE THE BLOCK (broadcast 128-bit loads), rebuilds the rows of U with the
// multipliers lambda(a,b) = l(pivot lane a, step b) —
//     u0 = r0, u1 = fma(-l10,u0,r1), u2 = fma(-l21,u1,fma(-l20,u0,r2)), u3 = fma(-l32,u2,fma(-l31,u1,fma(-l30,u0,r3)))
// which are the very operations the pivot lanes perform on their own registers — and applies the four updates in step
// order, a_ij = fma(-l3,u3,fma(-l2,u2,fma(-l1,u1,fma(-l0,u0,a_ij)))): bit-identical to the one-step-at-a-time sequence.
// Row stores per system: 64 full-warp instructions instead of 136 one-lane ones; the header stores and loads of the
// pair kernel become 36 shuffles per block; the back substitution broadcasts x_k by shuffle as well.
constexpr int B_LD = 34;                    // doubles per stored row: 32 columns, rhs, pad -> 272-byte row stride, so the
constexpr int B_BYTES = 32 * B_LD * 8;      // 128-bit stores of 8 consecutive lanes cover all 32 banks once

__device__ __forceinline__ void sts2(double *p, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ double2 lds2(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

template <int MODE, int KB>
__device__ __forceinline__ void lu32b_block(double (&a)[N32], double &bb, double &my_rinv, int &pos, int &dm, int &ip, int &bad,
                                            double *rows, const int lane) {
  constexpr unsigned RLO = 64u << 20, RSPAN = ((1981u - 64u) << 20) - 1u;
  constexpr int KE = KB + 4;                 // first column right of the block
  double *const mine = rows + lane * B_LD;
#pragma unroll
  for (int j = KE; j < N32; j += 2) sts2(&mine[j], a[j], a[j + 1]);
  if (MODE != 0) mine[N32] = bb;
  double l[4];
  int src[4];
#pragma unroll
  for (int jj = 0; jj < 4; jj++) {
    constexpr int k0 = KB;
    const int k = k0 + jj;
    const int hk = (__double2hiint(a[k]) & 0x7fffffff) | dm;
    const int M = __reduce_max_sync(0xffffffffu, hk);
    const bool c = hk == M;
    const unsigned ball = __ballot_sync(0xffffffffu, c);
    bad |= ((ball & (ball - 1u)) != 0u) | ((unsigned)M - RLO > RSPAN);
    src[jj] = __ffs(ball) - 1;
    const double rk = rcp_fast(a[k]);
    const double rinv = __shfl_sync(0xffffffffu, rk, src[jj]);
    const int ppos = __shfl_sync(0xffffffffu, pos, src[jj]);
    if (lane == k) ip = ppos + 1;
    pos = c ? k : (pos == k ? ppos : pos);
    dm = c ? -1 : dm;
    if (MODE != 0) my_rinv = dsel(c, rk, my_rinv);
    const bool act = dm == 0;
    const double t = __dmul_rn(a[k], rinv);
    a[k] = dsel(act, t, a[k]);
    l[jj] = dsel(act, t, 0.0);
#pragma unroll
    for (int m = jj + 1; m < 4; m++) {
      const double u = __shfl_sync(0xffffffffu, a[k0 + m], src[jj]);
      a[k0 + m] = __fma_rn(-l[jj], u, a[k0 + m]);
    }
  }
  if (KE < N32 || MODE != 0) {
    const double l10 = __shfl_sync(0xffffffffu, l[0], src[1]);
    const double l20 = __shfl_sync(0xffffffffu, l[0], src[2]), l21 = __shfl_sync(0xffffffffu, l[1], src[2]);
    const double l30 = __shfl_sync(0xffffffffu, l[0], src[3]), l31 = __shfl_sync(0xffffffffu, l[1], src[3]), l32 = __shfl_sync(0xffffffffu, l[2], src[3]);
    const double *const r0 = rows + src[0] * B_LD, *const r1 = rows + src[1] * B_LD, *const r2 = rows + src[2] * B_LD, *const r3 = rows + src[3] * B_LD;
    __syncwarp();
#pragma unroll
    for (int j = KE; j < N32; j += 2) {
      const double2 u0 = lds2(&r0[j]), v1 = lds2(&r1[j]), v2 = lds2(&r2[j]), v3 = lds2(&r3[j]);
      const double u1x = __fma_rn(-l10, u0.x, v1.x), u1y = __fma_rn(-l10, u0.y, v1.y);
      const double u2x = __fma_rn(-l21, u1x, __fma_rn(-l20, u0.x, v2.x)), u2y = __fma_rn(-l21, u1y, __fma_rn(-l20, u0.y, v2.y));
      const double u3x = __fma_rn(-l32, u2x, __fma_rn(-l31, u1x, __fma_rn(-l30, u0.x, v3.x)));
      const double u3y = __fma_rn(-l32, u2y, __fma_rn(-l31, u1y, __fma_rn(-l30, u0.y, v3.y)));
      a[j] = __fma_rn(-l[3], u3x, __fma_rn(-l[2], u2x, __fma_rn(-l[1], u1x, __fma_rn(-l[0], u0.x, a[j]))));
      a[j + 1] = __fma_rn(-l[3], u3y, __fma_rn(-l[2], u2y, __fma_rn(-l[1], u1y, __fma_rn(-l[0], u0.y, a[j + 1]))));
    }
    if (MODE != 0) {
      const double y0 = r0[N32];
      const double y1 = __fma_rn(-l10, y0, r1[N32]);
      const double y2 = __fma_rn(-l21, y1, __fma_rn(-l20, y0, r2[N32]));
      const double y3 = __fma_rn(-l32, y2, __fma_rn(-l31, y1, __fma_rn(-l30, y0, r3[N32])));
      bb = __fma_rn(-l[3], y3, __fma_rn(-l[2], y2, __fma_rn(-l[1], y1, __fma_rn(-l[0], y0, bb))));
    }
    __syncwarp();
  }
}
template <int MODE, int KB> struct Lu32bBlocks {
  static __device__ __forceinline__ void run(double (&a)[N32], double &bb, double &my_rinv, int &pos, int &dm, int &ip, int &bad, double *rows, const int lane) {
    lu32b_block<MODE, KB>(a, bb, my_rinv, pos, dm, ip, bad, rows, lane);
    Lu32bBlocks<MODE, KB + 4>::run(a, bb, my_rinv, pos, dm, ip, bad, rows, lane);
  }
};
template <int MODE> struct Lu32bBlocks<MODE, N32> {
  static __device__ __forceinline__ void run(double (&)[N32], double &, double &, int &, int &, int &, int &, double *, int) {}
};

template <int MODE, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
lu32b_kernel(double *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, double *__restrict__ Bg, long long stride_b,
             double *__restrict__ Xg, long long stride_x, int *__restrict__ info_g, int *__restrict__ redo, int nsys) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *const rows = reinterpret_cast<double *>(smem_dyn + warp * B_BYTES);
  for (long long g0 = (long long)blockIdx.x * WARPS; g0 < nsys; g0 += (long long)gridDim.x * WARPS) {
    long long sys = g0 + warp;
    if (sys >= nsys) sys = nsys - 1;          // a surplus warp repeats the last system (same values rewritten)
    double *const A = Ag + sys * stride_a;
    double a[N32], bb = 0.0, my_rinv = 0.0;
#pragma unroll
    for (int j = 0; j < N32; j++) a[j] = __ldcs(A + lane + N32 * j);
    if (MODE != 0) bb = __ldcs(Bg + sys * stride_b + lane);
    int pos = lane, dm = 0, bad = 0, ip = 0;
    Lu32bBlocks<MODE, 0>::run(a, bb, my_rinv, pos, dm, ip, bad, rows, lane);
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    double myx = 0.0;
    if (MODE != 0) {
      // back substitution, column sweep on the implicitly permuted U; x_k = y_k * (1/u_kk) leaves the lane at position k by shuffle
#pragma unroll
      for (int k = N32 - 1; k >= 0; k--) {
        const double xc = __dmul_rn(bb, my_rinv);
        const bool own = pos == k;
        const int o = __ffs(__ballot_sync(0xffffffffu, own)) - 1;
        const double xk = __shfl_sync(0xffffffffu, xc, o);
        myx = dsel(own, xc, myx);
        bb = __fma_rn(-a[k], xk, bb);
      }
    }
    if (badm == 0) {
      if (MODE != 2) {
#pragma unroll
        for (int j = 0; j < N32; j++) __stcs(A + pos + N32 * j, a[j]);
        ipiv_g[sys * N32 + lane] = ip;
      }
      if (MODE == 1) Bg[sys * stride_b + pos] = myx;
      if (MODE == 2) Xg[sys * stride_x + pos] = myx;
      if (lane == 0 && info_g) info_g[sys] = 0;
    } else if (lane == 0 && g0 + warp < nsys) {
      redo[1 + atomicAdd(redo, 1)] = (int)sys;
    }
    __syncwarp();
  }
}

// The exact one-warp routine on the systems the fast kernel flagged.
template <int MODE>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA, 4)
lu32_redo_kernel(double *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, double *__restrict__ Bg, long long stride_b,
                 double *__restrict__ Xg, long long stride_x, int *__restrict__ info_g, const int *__restrict__ redo) {
  __shared__ __align__(16) double srow_all[WARPS_PER_CTA][N32 + 2];
  const int warp = threadIdx.x >> 5, n = redo[0];
  for (int i = blockIdx.x * WARPS_PER_CTA + warp; i < n; i += gridDim.x * WARPS_PER_CTA) {
    const long long sy = redo[1 + i];
    lu32_one<double, MODE, true>(N32, Ag + sy * stride_a, N32, ipiv_g ? ipiv_g + sy * N32 : nullptr, Bg ? Bg + sy * stride_b : nullptr,
                                 Xg ? Xg + sy * stride_x : nullptr, info_g ? info_g + sy : nullptr, srow_all[warp]);
  }
}

// Launch: `batch` systems, lda == 32.  variant selects the mapping / CTA shape (tools/time_lu32.py).
template <int MODE, int WARPS, int MINB, bool REALIGN, int IOV = 0>
static void launch_q(double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info, int *redo, int nsys, cudaStream_t s) {
  const int lim = option("lu32.ctas_per_sm");
  const long ctas = (nsys + WARPS - 1) / WARPS, cap = (long)num_sms() * (lim > 0 && lim < MINB ? lim : MINB);
  if (IOV == 3) {        // 16 one-warp CTAs per SM need 16 x 10.6 KB of shared memory
    auto fn = lu32q_kernel<MODE, WARPS, MINB, REALIGN, IOV>;
    JB_ONCE_PER_DEVICE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  }
  lu32q_kernel<MODE, WARPS, MINB, REALIGN, IOV><<<(int)(ctas < cap ? ctas : cap), 32 * WARPS, 0, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo, nsys);
  count_launch();
}
template <int MODE, int WARPS, int MINB>
static int launch_b(double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info, int *redo, int nsys, cudaStream_t s) {
  JB_ONCE_PER_DEVICE(JB_CUDA(cudaFuncSetAttribute(lu32b_kernel<MODE, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * B_BYTES)));
  const int lim = option("lu32.ctas_per_sm");
  const long ctas = (nsys + WARPS - 1) / WARPS, cap = (long)num_sms() * (lim > 0 && lim < MINB ? lim : MINB);
  lu32b_kernel<MODE, WARPS, MINB><<<(int)(ctas < cap ? ctas : cap), 32 * WARPS, WARPS * B_BYTES, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo, nsys);
  count_launch();
  return 0;
}
template <int S, int MODE, int WARPS, int MINB, bool REALIGN>
static void launch_s(double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info, int *redo, int ngroups, cudaStream_t s) {
  const int lim = option("lu32.ctas_per_sm");
  const long ctas = (ngroups + WARPS - 1) / WARPS, cap = (long)num_sms() * (lim > 0 && lim < MINB ? lim : MINB);
  lu32s_kernel<S, MODE, WARPS, MINB, REALIGN><<<(int)(ctas < cap ? ctas : cap), 32 * WARPS, 0, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo, ngroups);
  count_launch();
}
template <int S, int MODE, int WARPS, int MINB, bool REALIGN, int ABL = 0>
static void launch_p(double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info, int *redo, int ngroups, cudaStream_t s) {
  const int lim = option("lu32.ctas_per_sm");        // occupancy experiments (tools/time_lu32.py)
  const long ctas = (ngroups + WARPS - 1) / WARPS, cap = (long)num_sms() * (lim > 0 && lim < MINB ? lim : MINB);
  lu32p_kernel<S, MODE, WARPS, MINB, REALIGN, ABL><<<(int)(ctas < cap ? ctas : cap), 32 * WARPS, 0, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo, ngroups);
  count_launch();
}

template <int MODE>
void lu32d_run(int shape, double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info, int *redo,
               int nsys, cudaStream_t s);      // batched_lu32d.cu: trailing updates on the FP64 tensor pipe

// Handles the leading *done systems of the batch (a multiple of the launch granule); the caller runs the rest.
// variant (jb_set_option "lu32.variant", tools/time_lu32.py; 10^6 systems on B200, all bit-identical to lu32_one):
//   30 pair-lookahead, 16 one-warp CTAs per SM   5.83 ms   <- default
//   31 pair-lookahead, 16 warps re-aligned       6.53 ms
//   32 / 33 pair-lookahead, outputs staged in shared memory + bulk copies, 16 / 19 warps per SM   6.44 / 5.83 ms
//   35 pair-lookahead, 20 warps per SM at 96 registers (172 B of spills)   5.87 ms
//   50 / 51 trailing updates on the FP64 tensor pipe (batched_lu32d.cu), 16 / 12 warps per SM   7.19 / 7.47 ms
//    7 lane = row, shared-memory broadcast       6.14 ms      17 lane = row, shuffle broadcast        6.78 ms
//    6 two rows per lane, shared-memory          7.05 ms      16 two rows per lane, shuffle           8.00 ms
//  100+m / 200+m: ablations of 7 / 6 (-DJB_LU32_ABLATIONS builds only)
template <int MODE>
int lu32p_launch(int variant, double *a, long long sa, int *ipiv, double *b, long long sb, double *x, long long sx, int *info,
                 int batch, cudaStream_t s, int *done) {
  const bool two = variant == 6 || variant == 16 || (variant >= 200 && variant < 300);
  const int per = variant == 31 ? 16 : (two ? 2 : 1);
  const int n = (batch / per) * per, ng = n / (two ? 2 : 1);
  *done = n;
  if (n == 0) return 0;
  int *redo = nullptr;
  if (int rc = ws_alloc((void **)&redo, sizeof(int) * ((size_t)n + 1), s)) return rc;
  cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(int), s);
  if (e == cudaSuccess) {
    switch (variant) {
      case 6: launch_p<2, MODE, 1, 12, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 7: launch_p<1, MODE, 1, 16, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 16: launch_s<2, MODE, 1, 12, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 17: launch_s<1, MODE, 1, 16, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 31: launch_q<MODE, 16, 1, true>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 50: lu32d_run<MODE>(16, a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 51: lu32d_run<MODE>(12, a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 35: launch_q<MODE, 1, 20, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;     // 20 warps per SM at 96 registers
      case 32: case 33: {   // staged outputs + bulk copies; 16-byte aligned outputs and even strides, else the plain kernel
        const bool al = (((uintptr_t)a | (uintptr_t)ipiv | (uintptr_t)b | (uintptr_t)x) & 15) == 0 && ((sa | sb | sx) & 1) == 0;
        if (!al) launch_q<MODE, 1, 16, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s);
        else if (variant == 32) launch_q<MODE, 1, 16, false, 3>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s);
        else launch_q<MODE, 1, 20, false, 3>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s);
        break;
      }
      case 40: launch_b<MODE, 1, 16>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 41: launch_b<MODE, 4, 4>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      case 42: launch_b<MODE, 1, 12>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
#ifdef JB_LU32_ABLATIONS
#define ABL1(v) case 100 + v: if (MODE == 1) launch_p<1, 1, 1, 16, false, v>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break; \
                case 200 + v: if (MODE == 1) launch_p<2, 1, 1, 12, false, v>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
      ABL1(1) ABL1(2) ABL1(3) ABL1(4) ABL1(7) ABL1(8) ABL1(15)
#endif
      default: launch_q<MODE, 1, 16, false>(a, sa, ipiv, b, sb, x, sx, info, redo, ng, s); break;
    }
    const long want = ((long)n + WARPS_PER_CTA - 1) / WARPS_PER_CTA, cap = (long)num_sms() * 4;
    lu32_redo_kernel<MODE><<<(int)(want < cap ? want : cap), 32 * WARPS_PER_CTA, 0, s>>>(a, sa, ipiv, b, sb, x, sx, info, redo);
    count_launch();
    e = cudaGetLastError();
  }
  ws_free(redo, s);
  if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "lu32p launch failed: %s", cudaGetErrorString(e)); return JB_ERR_CUDA; }
  return 0;
}
template int lu32p_launch<0>(int, double *, long long, int *, double *, long long, double *, long long, int *, int, cudaStream_t, int *);
template int lu32p_launch<1>(int, double *, long long, int *, double *, long long, double *, long long, int *, int, cudaStream_t, int *);
template int lu32p_launch<2>(int, double *, long long, int *, double *, long long, double *, long long, int *, int, cudaStream_t, int *);

}  // namespace jb
