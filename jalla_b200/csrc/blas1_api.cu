// Device-aware level-1 BLAS: cblas_?{copy,swap,scal,axpy,dot*,nrm2,asum} and cblas_i?amax.
//
// Why they are here: Jalla's top-level matrixCopy — the step immediately before every
// solveLinearSystem / invert (/root/reference/Numeric/Jalla/Matrix.hs:519-520,539) — is not a class method; it loops
// unsafeCopyVector over rows (Matrix.hs:855-878), and that is cblas_?copy with a stride (CVector.hs:284-288 through
// BlasOps.copy, Foreign/BlasOps.hs:40,96,120,144,168; raw imports Foreign/BLAS.chs:38-82).  A GPU-resident CMatrix
// therefore hands DEVICE pointers to these symbols; with libjalla_b200 in place of `cblas` they resolve here and
// dispatch on pointer residency.  The same holds for scaleColumn -> cblas_?scal (matrixMultDiag, Matrix.hs:660-665).
// Host-only operands are served too (staged through the GPU for arithmetic; plain moves for copy/swap), so the stock
// host Matrix type keeps working against this library.
#include "jb_common.cuh"

namespace jb {

// BLAS increment convention: element i of (x, incx) is x[i*incx] for incx > 0 and x[(i-(n-1))*incx] for incx < 0.
__host__ __device__ inline long long at(long long i, int n, int inc) { return inc >= 0 ? i * inc : (i - (n - 1)) * (long long)inc; }

template <typename T> __global__ void copy_kernel(int n, const T *__restrict__ x, int incx, T *__restrict__ y, int incy) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[at(i, n, incy)] = x[at(i, n, incx)];
}
template <typename T> __global__ void swap_kernel(int n, T *x, int incx, T *y, int incy) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T a = x[at(i, n, incx)], b = y[at(i, n, incy)];
    x[at(i, n, incx)] = b; y[at(i, n, incy)] = a;
  }
}
template <typename T, typename S> __global__ void scal_kernel(int n, S alpha, T *x, int incx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T v = x[i * (long long)incx];
    if constexpr (std::is_same<S, T>::value) x[i * (long long)incx] = Num<T>::mul(alpha, v);
    else x[i * (long long)incx] = Num<T>::scale(v, alpha);
  }
}
template <typename T> __global__ void axpy_kernel(int n, T alpha, const T *__restrict__ x, int incx, T *y, int incy) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[at(i, n, incy)] = Num<T>::fma(alpha, x[at(i, n, incx)], y[at(i, n, incy)]);
}

// ---- reductions: fixed grid, fixed tree — the result does not depend on scheduling -------------------------------
constexpr int RED_T = 256, RED_B = 296;
template <typename A> __device__ inline A block_reduce(A v, A (*comb)(A, A), A *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = RED_T / 2; s > 0; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] = comb(sh[threadIdx.x], sh[threadIdx.x + s]); __syncthreads(); }
  return sh[0];
}
struct Amax { double v; long long i; };
struct Ssq { double scale, ssq; };
template <typename A> __device__ inline A comb_add(A a, A b) { return a + b; }
__device__ inline cdouble comb_addz(cdouble a, cdouble b) { return {a.x + b.x, a.y + b.y}; }
__device__ inline Amax comb_amax(Amax a, Amax b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }   // first maximum wins
__device__ inline Ssq comb_ssq(Ssq a, Ssq b) {        // (scale, ssq) pairs as in ?lassq: value = scale * sqrt(ssq)
  if (a.scale < b.scale) { Ssq t = a; a = b; b = t; }
  if (a.scale == 0.0) return a;
  double r = b.scale / a.scale;
  return {a.scale, a.ssq + b.ssq * r * r};
}
// MODE 0: dot (conj(x) if CONJ)   1: asum (|re|+|im|)   2: nrm2   3: iamax
template <typename T, int MODE, bool CONJ, typename A, A (*COMB)(A, A)>
__global__ void reduce_kernel(int n, const T *__restrict__ x, int incx, const T *__restrict__ y, int incy, A *part, int stage2) {
  __shared__ A sh[RED_T];
  A acc;
  if constexpr (MODE == 3) acc = {-1.0, 0x7fffffffffffffffLL};
  else if constexpr (MODE == 2) acc = {0.0, 1.0};
  else if constexpr (std::is_same<A, cdouble>::value) acc = {0.0, 0.0};
  else acc = 0.0;
  if (stage2) {
    for (int i = threadIdx.x; i < n; i += RED_T) acc = COMB(acc, part[i]);
  } else {
    for (long long i = blockIdx.x * (long long)RED_T + threadIdx.x; i < n; i += (long long)gridDim.x * RED_T) {
      const T xv = x[at(i, n, incx)];
      if constexpr (MODE == 0) {
        const T yv = y[at(i, n, incy)];
        if constexpr (Num<T>::cplx) {
          const double xr = xv.x, xi = CONJ ? -(double)xv.y : (double)xv.y, yr = yv.x, yi = yv.y;
          acc = {acc.x + (xr * yr - xi * yi), acc.y + (xr * yi + xi * yr)};
        } else acc += (double)xv * (double)yv;
      } else if constexpr (MODE == 1) {
        acc += (double)Num<T>::abs1(xv);
      } else if constexpr (MODE == 2) {
        const double parts[2] = {fabs((double)Num<T>::re(xv)), fabs((double)Num<T>::im(xv))};
        for (int q = 0; q < (Num<T>::cplx ? 2 : 1); q++)
          if (parts[q] != 0.0) acc = comb_ssq(acc, Ssq{parts[q], 1.0});
      } else {
        const double v = (double)Num<T>::abs1(xv);
        if (v == v) acc = comb_amax(acc, Amax{v, i});     // NaNs never win, as the reference i?amax
      }
    }
  }
  A r = block_reduce<A>(acc, COMB, sh);
  if (threadIdx.x == 0) part[stage2 ? 0 : blockIdx.x] = r;
}

// ---- operand views: host vectors are staged into compact device buffers ---------------------------------------------
template <typename T> struct Vec {
  T *dev = nullptr; int inc = 0; bool staged = false; T *host = nullptr; int host_inc = 0; int n = 0;
  int open(const T *p, int n_, int inc_, bool copy_in, cudaStream_t s) {
    n = n_;
    if (classify(p) == Residency::Device) { dev = const_cast<T *>(p); inc = inc_; return 0; }
    staged = true; host = const_cast<T *>(p); host_inc = inc_; inc = inc_ < 0 ? -1 : 1;
    if (int rc = ws_alloc((void **)&dev, sizeof(T) * (size_t)n, s)) return rc;
    const size_t pitch = sizeof(T) * (size_t)(inc_ < 0 ? -inc_ : inc_);
    if (copy_in) JB_CUDA(cudaMemcpy2DAsync(dev, sizeof(T), host, pitch, sizeof(T), n, cudaMemcpyHostToDevice, s));
    return 0;
  }
  int close(bool copy_out, cudaStream_t s) {
    if (!staged) return 0;
    int rc = 0;
    const size_t pitch = sizeof(T) * (size_t)(host_inc < 0 ? -host_inc : host_inc);
    if (copy_out && cudaMemcpy2DAsync(host, pitch, dev, sizeof(T), sizeof(T), n, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = JB_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = JB_ERR_CUDA;
    ws_free(dev, s);
    if (rc) set_error(JB_ERR_CUDA, "level-1 staging copy failed");
    return rc;
  }
};
static inline int grid_for_n(int n) { long b = ((long)n + 255) / 256; long cap = (long)num_sms() * 8; return (int)(b < cap ? (b > 0 ? b : 1) : cap); }
static inline bool fail(const char *name, cudaError_t e) {
  if (e == cudaSuccess) return false;
  set_error(JB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  return true;
}

template <typename T> static void copy_entry(const char *name, int n, const T *x, int incx, T *y, int incy, bool swap) {
  if (n <= 0) return;
  if (incx == 0 || incy == 0) { set_error(JB_ERR_ARG, "%s: zero increment", name); return; }
  const bool xd = classify(x) == Residency::Device, yd = classify(y) == Residency::Device;
  if (!xd && !yd) {          // host to host: a move, no arithmetic
    T *xm = const_cast<T *>(x);
    for (long long i = 0; i < n; i++) {
      T &a = xm[at(i, n, incx)], &b = y[at(i, n, incy)];
      if (swap) { T t = a; a = b; b = t; } else b = a;
    }
    return;
  }
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  Vec<T> vx, vy;
  if (vx.open(x, n, incx, true, s)) return;
  if (vy.open(y, n, incy, swap || false, s)) { vx.close(false, s); return; }
  if (swap) swap_kernel<T><<<grid_for_n(n), 256, 0, s>>>(n, vx.dev, vx.inc, vy.dev, vy.inc);
  else copy_kernel<T><<<grid_for_n(n), 256, 0, s>>>(n, vx.dev, vx.inc, vy.dev, vy.inc);
  count_launch();
  fail(name, cudaGetLastError());
  vx.close(swap, s);
  vy.close(true, s);
}
template <typename T, typename S> static void scal_entry(const char *name, int n, S alpha, T *x, int incx) {
  if (n <= 0 || incx <= 0) return;            // reference ?scal: no-op for non-positive increments
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  Vec<T> vx;
  if (vx.open(x, n, incx, true, s)) return;
  scal_kernel<T, S><<<grid_for_n(n), 256, 0, s>>>(n, alpha, vx.dev, vx.inc);
  count_launch();
  fail(name, cudaGetLastError());
  vx.close(true, s);
}
template <typename T> static void axpy_entry(const char *name, int n, T alpha, const T *x, int incx, T *y, int incy) {
  if (n <= 0) return;
  if (ensure_device()) return;
  cudaStream_t s = tls().stream;
  Vec<T> vx, vy;
  if (vx.open(x, n, incx, true, s)) return;
  if (vy.open(y, n, incy, true, s)) { vx.close(false, s); return; }
  axpy_kernel<T><<<grid_for_n(n), 256, 0, s>>>(n, alpha, vx.dev, vx.inc, vy.dev, vy.inc);
  count_launch();
  fail(name, cudaGetLastError());
  vx.close(false, s);
  vy.close(true, s);
}
// Runs the two-stage reduction and brings the result to the host (these entry points return by value).
template <typename T, int MODE, bool CONJ, typename A, A (*COMB)(A, A)>
static int reduce_entry(const char *name, int n, const T *x, int incx, const T *y, int incy, A *out) {
  if (int rc = ensure_device()) return rc;
  cudaStream_t s = tls().stream;
  Vec<T> vx, vy;
  if (int rc = vx.open(x, n, incx, true, s)) return rc;
  if (y) { if (int rc = vy.open(y, n, incy, true, s)) { vx.close(false, s); return rc; } }
  A *part = nullptr;
  int rc = ws_alloc((void **)&part, sizeof(A) * RED_B, s);
  if (!rc) {
    const int blocks = (int)std::min<long>(RED_B, ((long)n + RED_T - 1) / RED_T);
    reduce_kernel<T, MODE, CONJ, A, COMB><<<blocks, RED_T, 0, s>>>(n, vx.dev, vx.inc, y ? vy.dev : nullptr, vy.inc, part, 0);
    reduce_kernel<T, MODE, CONJ, A, COMB><<<1, RED_T, 0, s>>>(blocks, nullptr, 0, nullptr, 0, part, 1);
    count_launch(2);
    if (fail(name, cudaGetLastError()) || fail(name, cudaMemcpyAsync(out, part, sizeof(A), cudaMemcpyDeviceToHost, s)) ||
        fail(name, cudaStreamSynchronize(s))) rc = JB_ERR_CUDA;
    ws_free(part, s);
  }
  vx.close(false, s);
  if (y) vy.close(false, s);
  return rc;
}
template <typename T> static double dot_real(const char *name, int n, const T *x, int incx, const T *y, int incy) {
  double r = 0.0;
  if (n > 0) reduce_entry<T, 0, false, double, comb_add<double>>(name, n, x, incx, y, incy, &r);
  return r;
}
template <typename T, bool CONJ> static void dot_cplx(const char *name, int n, const T *x, int incx, const T *y, int incy, T *out) {
  cdouble r = {0.0, 0.0};
  if (n > 0) reduce_entry<T, 0, CONJ, cdouble, comb_addz>(name, n, x, incx, y, incy, &r);
  T res = Num<T>::make((typename Num<T>::R)r.x, (typename Num<T>::R)r.y);
  if (classify(out) == Residency::Device) cudaMemcpy(out, &res, sizeof(T), cudaMemcpyHostToDevice);
  else *out = res;
}
template <typename T> static double asum_of(const char *name, int n, const T *x, int incx) {
  double r = 0.0;
  if (n > 0 && incx > 0) reduce_entry<T, 1, false, double, comb_add<double>>(name, n, x, incx, nullptr, 0, &r);
  return r;
}
template <typename T> static double nrm2_of(const char *name, int n, const T *x, int incx) {
  Ssq r = {0.0, 1.0};
  if (n > 0 && incx > 0) reduce_entry<T, 2, false, Ssq, comb_ssq>(name, n, x, incx, nullptr, 0, &r);
  return r.scale * sqrt(r.ssq);
}
template <typename T> static size_t iamax_of(const char *name, int n, const T *x, int incx) {
  Amax r = {-1.0, 0};
  if (n > 0 && incx > 0) reduce_entry<T, 3, false, Amax, comb_amax>(name, n, x, incx, nullptr, 0, &r);
  return (r.v < 0.0) ? 0 : (size_t)r.i;
}
template <typename T> static int scalar_in(const void *p, T *out) {
  if (!p) { set_error(JB_ERR_ARG, "null scalar pointer"); return 1; }
  if (classify(p) == Residency::Device) {
    if (ensure_device()) return 1;
    if (cudaMemcpy(out, p, sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error(JB_ERR_CUDA, "scalar read failed"); return 1; }
  } else memcpy(out, p, sizeof(T));
  return 0;
}
}  // namespace jb

using namespace jb;
extern "C" {
#define L1_REAL(t, T)                                                                                                          \
  void jb_cblas_##t##copy(int n, const T *x, int incx, T *y, int incy) { copy_entry<T>("cblas_" #t "copy", n, x, incx, y, incy, false); } \
  void jb_cblas_##t##swap(int n, T *x, int incx, T *y, int incy) { copy_entry<T>("cblas_" #t "swap", n, x, incx, y, incy, true); }        \
  void jb_cblas_##t##scal(int n, T alpha, T *x, int incx) { scal_entry<T, T>("cblas_" #t "scal", n, alpha, x, incx); }                     \
  void jb_cblas_##t##axpy(int n, T alpha, const T *x, int incx, T *y, int incy) { axpy_entry<T>("cblas_" #t "axpy", n, alpha, x, incx, y, incy); } \
  T jb_cblas_##t##dot(int n, const T *x, int incx, const T *y, int incy) { return (T)dot_real<T>("cblas_" #t "dot", n, x, incx, y, incy); } \
  T jb_cblas_##t##nrm2(int n, const T *x, int incx) { return (T)nrm2_of<T>("cblas_" #t "nrm2", n, x, incx); }                              \
  T jb_cblas_##t##asum(int n, const T *x, int incx) { return (T)asum_of<T>("cblas_" #t "asum", n, x, incx); }                              \
  size_t jb_cblas_i##t##amax(int n, const T *x, int incx) { return iamax_of<T>("cblas_i" #t "amax", n, x, incx); }                         \
  void cblas_##t##copy(int n, const T *x, int incx, T *y, int incy) { jb_cblas_##t##copy(n, x, incx, y, incy); }                           \
  void cblas_##t##swap(int n, T *x, int incx, T *y, int incy) { jb_cblas_##t##swap(n, x, incx, y, incy); }                                 \
  void cblas_##t##scal(int n, T alpha, T *x, int incx) { jb_cblas_##t##scal(n, alpha, x, incx); }                                          \
  void cblas_##t##axpy(int n, T alpha, const T *x, int incx, T *y, int incy) { jb_cblas_##t##axpy(n, alpha, x, incx, y, incy); }           \
  T cblas_##t##dot(int n, const T *x, int incx, const T *y, int incy) { return jb_cblas_##t##dot(n, x, incx, y, incy); }                   \
  T cblas_##t##nrm2(int n, const T *x, int incx) { return jb_cblas_##t##nrm2(n, x, incx); }                                                \
  T cblas_##t##asum(int n, const T *x, int incx) { return jb_cblas_##t##asum(n, x, incx); }                                                \
  size_t cblas_i##t##amax(int n, const T *x, int incx) { return jb_cblas_i##t##amax(n, x, incx); }
L1_REAL(s, float)
L1_REAL(d, double)
double jb_cblas_dsdot(int n, const float *x, int incx, const float *y, int incy) { return dot_real<float>("cblas_dsdot", n, x, incx, y, incy); }
double cblas_dsdot(int n, const float *x, int incx, const float *y, int incy) { return jb_cblas_dsdot(n, x, incx, y, incy); }

#define L1_CPLX(t, T, R, rt)                                                                                                   \
  void jb_cblas_##t##copy(int n, const void *x, int incx, void *y, int incy) { copy_entry<T>("cblas_" #t "copy", n, (const T *)x, incx, (T *)y, incy, false); } \
  void jb_cblas_##t##swap(int n, void *x, int incx, void *y, int incy) { copy_entry<T>("cblas_" #t "swap", n, (const T *)x, incx, (T *)y, incy, true); }        \
  void jb_cblas_##t##scal(int n, const void *alpha, void *x, int incx) { T a; if (scalar_in(alpha, &a)) return; scal_entry<T, T>("cblas_" #t "scal", n, a, (T *)x, incx); } \
  void jb_cblas_##t##rt##scal(int n, R alpha, void *x, int incx) { scal_entry<T, R>("cblas_" #t #rt "scal", n, alpha, (T *)x, incx); }     \
  void jb_cblas_##t##axpy(int n, const void *alpha, const void *x, int incx, void *y, int incy) {                                           \
    T a; if (scalar_in(alpha, &a)) return; axpy_entry<T>("cblas_" #t "axpy", n, a, (const T *)x, incx, (T *)y, incy); }                    \
  void jb_cblas_##t##dotu_sub(int n, const void *x, int incx, const void *y, int incy, void *r) { dot_cplx<T, false>("cblas_" #t "dotu_sub", n, (const T *)x, incx, (const T *)y, incy, (T *)r); } \
  void jb_cblas_##t##dotc_sub(int n, const void *x, int incx, const void *y, int incy, void *r) { dot_cplx<T, true>("cblas_" #t "dotc_sub", n, (const T *)x, incx, (const T *)y, incy, (T *)r); }  \
  R jb_cblas_##rt##t##nrm2(int n, const void *x, int incx) { return (R)nrm2_of<T>("cblas_" #rt #t "nrm2", n, (const T *)x, incx); }        \
  R jb_cblas_##rt##t##asum(int n, const void *x, int incx) { return (R)asum_of<T>("cblas_" #rt #t "asum", n, (const T *)x, incx); }        \
  size_t jb_cblas_i##t##amax(int n, const void *x, int incx) { return iamax_of<T>("cblas_i" #t "amax", n, (const T *)x, incx); }           \
  void cblas_##t##copy(int n, const void *x, int incx, void *y, int incy) { jb_cblas_##t##copy(n, x, incx, y, incy); }                     \
  void cblas_##t##swap(int n, void *x, int incx, void *y, int incy) { jb_cblas_##t##swap(n, x, incx, y, incy); }                           \
  void cblas_##t##scal(int n, const void *alpha, void *x, int incx) { jb_cblas_##t##scal(n, alpha, x, incx); }                             \
  void cblas_##t##rt##scal(int n, R alpha, void *x, int incx) { jb_cblas_##t##rt##scal(n, alpha, x, incx); }                               \
  void cblas_##t##axpy(int n, const void *alpha, const void *x, int incx, void *y, int incy) { jb_cblas_##t##axpy(n, alpha, x, incx, y, incy); } \
  void cblas_##t##dotu_sub(int n, const void *x, int incx, const void *y, int incy, void *r) { jb_cblas_##t##dotu_sub(n, x, incx, y, incy, r); } \
  void cblas_##t##dotc_sub(int n, const void *x, int incx, const void *y, int incy, void *r) { jb_cblas_##t##dotc_sub(n, x, incx, y, incy, r); } \
  R cblas_##rt##t##nrm2(int n, const void *x, int incx) { return jb_cblas_##rt##t##nrm2(n, x, incx); }                                     \
  R cblas_##rt##t##asum(int n, const void *x, int incx) { return jb_cblas_##rt##t##asum(n, x, incx); }                                     \
  size_t cblas_i##t##amax(int n, const void *x, int incx) { return jb_cblas_i##t##amax(n, x, incx); }
L1_CPLX(c, cfloat, float, s)
L1_CPLX(z, cdouble, double, d)
}
