// Batched 32x32 LU factorise-and-solve for Float, Complex Float and Complex Double (the batched form of
// /root/reference/Numeric/Jalla/Matrix.hs:495-508 gesv and :763-767 getrf; FFI shapes of
// /root/reference/Numeric/Jalla/Foreign/LAPACKE.chs:679-694).  The Double kernels live in batched_lu32p.cu.
//
// Same plan as the Double fast path, written over Num<T>: one warp per system, lane = row, the row in registers
// (32 / 64 / 128 registers for s / c / z), coalesced column loads and stores, row interchanges never executed (`pos`
// tracking).  The fast path handles only the generic step — one redux.max on the order-preserving integer image of
// |re| + |im| (the whole word for Float magnitudes, the high word for Double ones), a ballot, and the pivot lane's
// predicated 128-bit stores of its row / right-hand side / reciprocal / position into shared memory, which all lanes
// read back as broadcast 128-bit loads (4, 2 or 1 entries each).  A tie, a zero, infinite or NaN maximum flags the
// system; flagged systems are redone from their untouched inputs by the exact one-warp routine (lu32_one: lowest-index
// tie-break, zero pivots -> info, the NaN rule of i?amax).  The arithmetic is lu32_one's sequence (r = 1/pivot by IEEE
// division, l = a*r, a_ij = a_ij - l*u with Num<T>::msub's rounding contract, x_k = y_k*r_k), so factors, pivots and
// solutions are bit-identical to oracle/oracle_impl.inc.
#include "batched_common.cuh"

namespace jb {

template <typename T> struct alignas(16) Vec128 { T v[16 / sizeof(T)]; };

__device__ __forceinline__ int mag_key(float v) { return __float_as_int(v); }
__device__ __forceinline__ int mag_key(double v) { return __double2hiint(v); }
template <typename R> struct KeyRange;      // keys 1 .. MAX are finite and non-zero
template <> struct KeyRange<float> { static constexpr unsigned MAX = 0x7f7fffffu; };
template <> struct KeyRange<double> { static constexpr unsigned MAX = 0x7fefffffu; };

// The row of a lane: real and imaginary parts in separate register arrays (one array of 32 16-byte entries is left in
// local memory by nvcc — that is what holds lu32_one<cdouble> back).
template <typename T> struct Row32 {
  using R = typename Num<T>::R;
  R re[N32], im[Num<T>::cplx ? N32 : 1];
  __device__ __forceinline__ T get(int j) const { if constexpr (Num<T>::cplx) return Num<T>::make(re[j], im[j]); else return re[j]; }
  __device__ __forceinline__ void set(int j, T x) { if constexpr (Num<T>::cplx) { re[j] = x.x; im[j] = x.y; } else re[j] = x; }
};
template <typename T> struct Lu32gSmem {
  alignas(16) T srow[2][N32];     // the published pivot row, double buffered by step parity
  alignas(16) T shy[N32];         // y_k, later x_k
  alignas(16) T shr[N32];         // 1 / u_kk
  int spos[N32];                  // position of the pivot row before step k's interchange (ipiv[k] - 1)
};

// One elimination step, K a compile-time constant: the step loop is unrolled by template recursion (with `#pragma unroll`
// nvcc unrolls the Complex Double body too late and the row stays in local memory).
template <typename T, int MODE, int K>
__device__ __forceinline__ void lu32g_step(Row32<T> &a, T &bb, int &pos, int &dm, int &bad, Lu32gSmem<T> &sm) {
  using R = typename Num<T>::R;
  using V = Vec128<T>;
  constexpr int E = 16 / sizeof(T), k = K, J0 = (K + 1) & ~(E - 1);
  T *const sr = sm.srow[k & 1];
  // ---- selection among the rows not yet used (dm = sign bit for used rows)
  const T ak = a.get(k);
  const int key = mag_key(Num<T>::abs1(ak)) | dm;
  const int M = __reduce_max_sync(0xffffffffu, key);
  const bool cl = key == M;
  const unsigned ball = __ballot_sync(0xffffffffu, cl);
  bad |= ((ball & (ball - 1u)) != 0u) | ((unsigned)(M - 1) >= KeyRange<R>::MAX);
  const T rk = Num<T>::recip(ak);            // speculative; only the pivot lane's is published
  if (cl) {
#pragma unroll
    for (int j = J0; j < N32; j += E) {
      V v;
#pragma unroll
      for (int e = 0; e < E; e++) v.v[e] = a.get(j + e);
      *reinterpret_cast<V *>(&sr[j]) = v;
    }
    sm.shy[k] = bb; sm.shr[k] = rk; sm.spos[k] = pos;
  }
  __syncwarp();
  const T yk = sm.shy[k], rinv = sm.shr[k];
  const int ppos = sm.spos[k];
  // ---- implicit interchange: the row sitting at position k moves to the pivot's old position, the pivot row to k.
  //      Used rows keep their registers; their multiplier is an arithmetic zero (a non-finite entry that 0 * x could
  //      leak into them also reaches a later pivot search and flags the system).
  pos = cl ? k : (pos == k ? ppos : pos);
  dm = cl ? (int)0x80000000 : dm;
  const bool act = dm == 0;
  const T t = Num<T>::cmul(ak, rinv);        // l = a * (1/pivot)
  a.set(k, act ? t : ak);
  const T lm = act ? t : Num<T>::zero();
#pragma unroll
  for (int j = J0; j < N32; j += E) {
    const V u = *reinterpret_cast<const V *>(&sr[j]);
#pragma unroll
    for (int e = 0; e < E; e++)
      if (j + e > k) a.set(j + e, Num<T>::msub(a.get(j + e), lm, u.v[e]));
  }
  if (MODE != 0) bb = Num<T>::msub(bb, lm, yk);
}
template <typename T, int MODE, int K> struct Lu32gSteps {
  static __device__ __forceinline__ void run(Row32<T> &a, T &bb, int &pos, int &dm, int &bad, Lu32gSmem<T> &sm) {
    lu32g_step<T, MODE, K>(a, bb, pos, dm, bad, sm);
    Lu32gSteps<T, MODE, K + 1>::run(a, bb, pos, dm, bad, sm);
  }
};
template <typename T, int MODE> struct Lu32gSteps<T, MODE, N32> {
  static __device__ __forceinline__ void run(Row32<T> &, T &, int &, int &, int &, Lu32gSmem<T> &) {}
};
template <typename T, int MODE, int K> struct Lu32gBack {       // back substitution, column sweep K = 31 .. 0
  static __device__ __forceinline__ void run(const Row32<T> &a, T &bb, const T my_rinv, const int pos, Lu32gSmem<T> &sm) {
    if (pos == K) sm.shy[K] = Num<T>::cmul(bb, my_rinv);
    __syncwarp();
    bb = Num<T>::msub(bb, a.get(K), sm.shy[K]);
    Lu32gBack<T, MODE, K - 1>::run(a, bb, my_rinv, pos, sm);
  }
};
template <typename T, int MODE> struct Lu32gBack<T, MODE, -1> {
  static __device__ __forceinline__ void run(const Row32<T> &, T &, const T, const int, Lu32gSmem<T> &) {}
};

// MODE 0: getrf (LU + ipiv), 1: gesv with one right-hand side (LU + ipiv + x in place of b)
template <typename T, int MODE, int MINB>
__global__ void __launch_bounds__(32, MINB)
lu32g_kernel(T *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, T *__restrict__ Bg, long long stride_b,
             int *__restrict__ info_g, int *__restrict__ redo, int nsys) {
  __shared__ Lu32gSmem<T> sm;
  const int lane = threadIdx.x;
  for (long long sys = blockIdx.x; sys < nsys; sys += gridDim.x) {
    T *const A = Ag + sys * stride_a;
    Row32<T> a;
    T bb = Num<T>::zero();
#pragma unroll
    for (int j = 0; j < N32; j++) a.set(j, A[lane + N32 * j]);
    if (MODE != 0) bb = Bg[sys * stride_b + lane];
    int pos = lane, dm = 0, bad = 0;
    Lu32gSteps<T, MODE, 0>::run(a, bb, pos, dm, bad, sm);
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    if (MODE != 0) {
      // shy[k] turns from y_k into x_k.  A row's right-hand side is dead once its x is published, so nothing needs a mask.
      const T my_rinv = sm.shr[pos];
      __syncwarp();
      Lu32gBack<T, MODE, N32 - 1>::run(a, bb, my_rinv, pos, sm);
    }
    if (badm == 0) {
#pragma unroll
      for (int j = 0; j < N32; j++) A[pos + N32 * j] = a.get(j);
      ipiv_g[sys * N32 + lane] = sm.spos[lane] + 1;
      if (MODE != 0) Bg[sys * stride_b + lane] = sm.shy[lane];
      if (lane == 0 && info_g) info_g[sys] = 0;
    } else if (lane == 0) {
      redo[1 + atomicAdd(redo, 1)] = (int)sys;
    }
    __syncwarp();
  }
}

// The exact one-warp routine on the systems the fast kernel flagged.
template <typename T, int MODE>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
lu32g_redo_kernel(T *__restrict__ Ag, long long stride_a, int *__restrict__ ipiv_g, T *__restrict__ Bg, long long stride_b,
                  int *__restrict__ info_g, const int *__restrict__ redo) {
  __shared__ __align__(16) T srow_all[WARPS_PER_CTA][N32 + 2];
  const int warp = threadIdx.x >> 5, n = redo[0];
  for (int i = blockIdx.x * WARPS_PER_CTA + warp; i < n; i += gridDim.x * WARPS_PER_CTA) {
    const long long sy = redo[1 + i];
    lu32_one<T, MODE, true>(N32, Ag + sy * stride_a, N32, ipiv_g + sy * N32, Bg ? Bg + sy * stride_b : nullptr, nullptr,
                            info_g ? info_g + sy : nullptr, srow_all[warp]);
  }
}

// Resident one-warp CTAs per SM, set by the registers the row needs.  Measured on 10^6 systems: Float at 24 CTAs / 80
// registers 4.16 ms against 4.13 at 32 / 64; Complex Double at 12 / 168 (spilling) 22.4 against 22.7 at 8 / 226 — the
// kernels are bound by their instruction streams (6 rounded operations per complex multiply-subtract, IEEE divisions
// for the reciprocal), not by occupancy (profiles/batched_others_r2d.jsonl).
template <typename T> struct Lu32gShape;
template <> struct Lu32gShape<float> { static constexpr int MINB = 32; };
template <> struct Lu32gShape<cfloat> { static constexpr int MINB = 16; };
template <> struct Lu32gShape<cdouble> { static constexpr int MINB = 8; };

// `batch` systems of order 32 with lda == 32 (any stride between systems).
template <typename T, int MODE>
int lu32g_launch(T *a, long long sa, int *ipiv, T *b, long long sb, int *info, int batch, cudaStream_t s) {
  constexpr int MINB = Lu32gShape<T>::MINB;
  if (batch <= 0) return 0;
  int *redo = nullptr;
  if (int rc = ws_alloc((void **)&redo, sizeof(int) * ((size_t)batch + 1), s)) return rc;
  cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(int), s);
  if (e == cudaSuccess) {
    const int lim = option("lu32.ctas_per_sm");
    const long cap = (long)num_sms() * (lim > 0 && lim < MINB ? lim : MINB);
    lu32g_kernel<T, MODE, MINB><<<(int)(batch < cap ? batch : cap), 32, 0, s>>>(a, sa, ipiv, b, sb, info, redo, batch);
    count_launch();
    const long want = ((long)batch + WARPS_PER_CTA - 1) / WARPS_PER_CTA, rcap = (long)num_sms() * 4;
    lu32g_redo_kernel<T, MODE><<<(int)(want < rcap ? want : rcap), 32 * WARPS_PER_CTA, 0, s>>>(a, sa, ipiv, b, sb, info, redo);
    count_launch();
    e = cudaGetLastError();
  }
  ws_free(redo, s);
  if (e != cudaSuccess) { set_error(JB_ERR_CUDA, "lu32g launch failed: %s", cudaGetErrorString(e)); return JB_ERR_CUDA; }
  return 0;
}
#define JB_LU32G_INST(T)                                                                                   \
  template int lu32g_launch<T, 0>(T *, long long, int *, T *, long long, int *, int, cudaStream_t);       \
  template int lu32g_launch<T, 1>(T *, long long, int *, T *, long long, int *, int, cudaStream_t);
JB_LU32G_INST(float)
JB_LU32G_INST(cfloat)
JB_LU32G_INST(cdouble)

}  // namespace jb
