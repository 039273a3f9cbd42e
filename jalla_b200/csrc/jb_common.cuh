// Shared declarations for libjalla_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <atomic>
#include "../../include/jalla_b200.h"

namespace jb {

// ---- element types -------------------------------------------------------------
// Complex numbers are interleaved (re,im) pairs, the layout of Haskell's
// `Storable (Complex a)` and of LAPACKE's lapack_complex_* structs
// (/root/reference/lapacke/include/lapacke_config.h:60-65).
struct cfloat { float x, y; };
struct cdouble { double x, y; };

template <typename T> struct Num;
template <> struct Num<float> {
  using R = float; static constexpr bool cplx = false;
  __host__ __device__ static float zero() { return 0.f; }
  __host__ __device__ static float one() { return 1.f; }
  __host__ __device__ static float make(float r, float) { return r; }
  __host__ __device__ static float re(float a) { return a; }
  __host__ __device__ static float im(float) { return 0.f; }
  __host__ __device__ static float conj(float a) { return a; }
  __host__ __device__ static float neg(float a) { return -a; }
  __host__ __device__ static float add(float a, float b) { return a + b; }
  __host__ __device__ static float sub(float a, float b) { return a - b; }
  __host__ __device__ static float mul(float a, float b) { return a * b; }
  __host__ __device__ static float scale(float a, float s) { return a * s; }
  __device__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
  // c - l*u with the oracle's rounding contract (oracle/oracle_impl.inc msub)
  __device__ static float msub(float c, float l, float u) { return __fmaf_rn(-l, u, c); }
  __device__ static float cmul(float a, float b) { return __fmul_rn(a, b); }
  __device__ static float recip(float a) { return __fdiv_rn(1.0f, a); }
  __host__ __device__ static float abs1(float a) { return fabsf(a); }
  __host__ __device__ static bool is_zero(float a) { return a == 0.f; }
  __host__ __device__ static bool is_nan(float a) { return a != a; }
};
template <> struct Num<double> {
  using R = double; static constexpr bool cplx = false;
  __host__ __device__ static double zero() { return 0.0; }
  __host__ __device__ static double one() { return 1.0; }
  __host__ __device__ static double make(double r, double) { return r; }
  __host__ __device__ static double re(double a) { return a; }
  __host__ __device__ static double im(double) { return 0.0; }
  __host__ __device__ static double conj(double a) { return a; }
  __host__ __device__ static double neg(double a) { return -a; }
  __host__ __device__ static double add(double a, double b) { return a + b; }
  __host__ __device__ static double sub(double a, double b) { return a - b; }
  __host__ __device__ static double mul(double a, double b) { return a * b; }
  __host__ __device__ static double scale(double a, double s) { return a * s; }
  __device__ static double fma(double a, double b, double c) { return ::fma(a, b, c); }
  __device__ static double msub(double c, double l, double u) { return __fma_rn(-l, u, c); }
  __device__ static double cmul(double a, double b) { return __dmul_rn(a, b); }
  __device__ static double recip(double a) { return __ddiv_rn(1.0, a); }
  __host__ __device__ static double abs1(double a) { return fabs(a); }
  __host__ __device__ static bool is_zero(double a) { return a == 0.0; }
  __host__ __device__ static bool is_nan(double a) { return a != a; }
};
template <> struct Num<cfloat> {
  using R = float; static constexpr bool cplx = true;
  __host__ __device__ static cfloat zero() { return {0.f, 0.f}; }
  __host__ __device__ static cfloat one() { return {1.f, 0.f}; }
  __host__ __device__ static cfloat make(float r, float i) { return {r, i}; }
  __host__ __device__ static float re(cfloat a) { return a.x; }
  __host__ __device__ static float im(cfloat a) { return a.y; }
  __host__ __device__ static cfloat conj(cfloat a) { return {a.x, -a.y}; }
  __host__ __device__ static cfloat neg(cfloat a) { return {-a.x, -a.y}; }
  __host__ __device__ static cfloat add(cfloat a, cfloat b) { return {a.x + b.x, a.y + b.y}; }
  __host__ __device__ static cfloat sub(cfloat a, cfloat b) { return {a.x - b.x, a.y - b.y}; }
  __host__ __device__ static cfloat mul(cfloat a, cfloat b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
  __host__ __device__ static cfloat scale(cfloat a, float s) { return {a.x * s, a.y * s}; }
  // four real multiplies per complex MAC ("4M"), never 3M: SURVEY.md §7.3 H6
  __device__ static cfloat fma(cfloat a, cfloat b, cfloat c) {
    return {fmaf(-a.y, b.y, fmaf(a.x, b.x, c.x)), fmaf(a.y, b.x, fmaf(a.x, b.y, c.y))};
  }
  __device__ static cfloat cmul(cfloat a, cfloat b) {
    return {__fmaf_rn(-a.y, b.y, __fmul_rn(a.x, b.x)), __fmaf_rn(a.y, b.x, __fmul_rn(a.x, b.y))};
  }
  __device__ static cfloat msub(cfloat c, cfloat l, cfloat u) {
    cfloat p = cmul(l, u); return {__fsub_rn(c.x, p.x), __fsub_rn(c.y, p.y)};
  }
  __device__ static cfloat recip(cfloat a) {
    float d = __fmaf_rn(a.y, a.y, __fmul_rn(a.x, a.x)); return {__fdiv_rn(a.x, d), __fdiv_rn(-a.y, d)};
  }
  __host__ __device__ static float abs1(cfloat a) { return fabsf(a.x) + fabsf(a.y); }
  __host__ __device__ static bool is_zero(cfloat a) { return a.x == 0.f && a.y == 0.f; }
  __host__ __device__ static bool is_nan(cfloat a) { return a.x != a.x || a.y != a.y; }
};
template <> struct Num<cdouble> {
  using R = double; static constexpr bool cplx = true;
  __host__ __device__ static cdouble zero() { return {0.0, 0.0}; }
  __host__ __device__ static cdouble one() { return {1.0, 0.0}; }
  __host__ __device__ static cdouble make(double r, double i) { return {r, i}; }
  __host__ __device__ static double re(cdouble a) { return a.x; }
  __host__ __device__ static double im(cdouble a) { return a.y; }
  __host__ __device__ static cdouble conj(cdouble a) { return {a.x, -a.y}; }
  __host__ __device__ static cdouble neg(cdouble a) { return {-a.x, -a.y}; }
  __host__ __device__ static cdouble add(cdouble a, cdouble b) { return {a.x + b.x, a.y + b.y}; }
  __host__ __device__ static cdouble sub(cdouble a, cdouble b) { return {a.x - b.x, a.y - b.y}; }
  __host__ __device__ static cdouble mul(cdouble a, cdouble b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
  __host__ __device__ static cdouble scale(cdouble a, double s) { return {a.x * s, a.y * s}; }
  __device__ static cdouble fma(cdouble a, cdouble b, cdouble c) {
    return {::fma(-a.y, b.y, ::fma(a.x, b.x, c.x)), ::fma(a.y, b.x, ::fma(a.x, b.y, c.y))};
  }
  __device__ static cdouble cmul(cdouble a, cdouble b) {
    return {__fma_rn(-a.y, b.y, __dmul_rn(a.x, b.x)), __fma_rn(a.y, b.x, __dmul_rn(a.x, b.y))};
  }
  __device__ static cdouble msub(cdouble c, cdouble l, cdouble u) {
    cdouble p = cmul(l, u); return {__dsub_rn(c.x, p.x), __dsub_rn(c.y, p.y)};
  }
  __device__ static cdouble recip(cdouble a) {
    double d = __fma_rn(a.y, a.y, __dmul_rn(a.x, a.x)); return {__ddiv_rn(a.x, d), __ddiv_rn(-a.y, d)};
  }
  __host__ __device__ static double abs1(cdouble a) { return fabs(a.x) + fabs(a.y); }
  __host__ __device__ static bool is_zero(cdouble a) { return a.x == 0.0 && a.y == 0.0; }
  __host__ __device__ static bool is_nan(cdouble a) { return a.x != a.x || a.y != a.y; }
};

// ---- per-thread state (Haskell `ccall unsafe` may enter from several OS threads: H11) ----
struct ThreadState {
  cudaStream_t stream = nullptr;
  int err = 0;
  char msg[256] = {0};
  int device = -1;          // >= 0: entry points on this thread bind this device instead of the library default (jb_dist_*)
  int gemm_cta_cap = 0;     // > 0: the persistent DMMA GEMMs launched by this thread use at most this many CTAs (SMs)
};
ThreadState &tls();
void set_error(int code, const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int option(const char *key);   // -1 when unset; `key` must be a string literal (values are cached per thread by pointer)

// Runs `stmt` the first time a call site is reached on each device (cudaFuncSetAttribute is per device / context).
#define JB_ONCE_PER_DEVICE(stmt)                                                                   \
  do {                                                                                             \
    static std::atomic<uint64_t> done__{0};                                                        \
    int dev__ = 0;                                                                                 \
    cudaGetDevice(&dev__);                                                                         \
    const uint64_t bit__ = 1ull << (dev__ & 63);                                                   \
    if (!(done__.load(std::memory_order_acquire) & bit__)) { stmt; done__.fetch_or(bit__, std::memory_order_release); } \
  } while (0)

#define JB_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      jb::set_error(JB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return JB_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

int ensure_device();   // 0 when a CUDA device is usable, else JB_ERR_CUDA (and error recorded)
int num_sms();

// ---- pointer residency -----------------------------------
enum class Residency { Host, Device };
Residency classify(const void *p);

// Stream-ordered scratch (cudaMallocAsync).  Freed on the same stream.
int ws_alloc(void **p, size_t bytes, cudaStream_t s);
void ws_free(void *p, cudaStream_t s);

// Host <-> device 2-D copies of one library call.  Pinned / registered host memory: plain cudaMemcpy2DAsync.  Pageable
// memory (what a stock Jalla Matrix is): the library's own staging through pinned slots filled by several host threads
// (host_stage.cu) — h2d blocks only for the host-side copy of a slot, d2h returns once the device side is queued;
// finish() (or the destructor) returns when every d2h has reached the caller's memory.
bool is_pageable(const void *p);
class HostStage {
 public:
  HostStage();
  ~HostStage();
  HostStage(const HostStage &) = delete;
  HostStage &operator=(const HostStage &) = delete;
  int h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t st);
  int d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t st);
  int finish();
 private:
  int acquire();
  bool locked_ = false;
};

// Brings a possibly-host matrix onto the device for the duration of one call.
// Compact copy (ld = rows) for host data; device data is used in place.
struct Staged {
  void *dev = nullptr;     // device-usable pointer
  int ld = 0;              // its leading dimension
  void *host = nullptr;    // original pointer when staged
  int host_ld = 0;
  size_t esz = 0;
  long rows = 0, cols = 0; // storage shape (column-major: rows contiguous)
  bool staged = false;
  HostStage stage;         // pageable host memory goes through the library's pinned slots
  int open(const void *p, long rows, long cols, int ld, size_t esz, bool copy_in, cudaStream_t s);
  int close(bool copy_out, cudaStream_t s);   // D2H + free; host memory is complete when the caller has synchronised s
};

// ---- internal device-pointer entry points -----------------------------------------
// C <- alpha*op(A)*op(B) + beta*C, column-major, all pointers device memory.
// tri: 0 full, 'L'/'U' compute only tiles touching that triangle of C (syrk-like updates).
template <typename T>
int gemm_device(int ta, int tb, int m, int n, int k, T alpha, const T *A, int lda, const T *B, int ldb,
                T beta, T *C, int ldc, cudaStream_t s, int tri = 0);
// fast paths return 1 when they took the call, 0 when the shape/alignment is not theirs
int dgemm_dmma_try(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda,
                   const double *B, int ldb, double beta, double *C, int ldc, cudaStream_t s, int tri);
int sgemm_tma_try(int ta, int tb, int m, int n, int k, float alpha, const float *A, int lda,
                  const float *B, int ldb, float beta, float *C, int ldc, cudaStream_t s, int tri);
int sgemm_simt_try(int ta, int tb, int m, int n, int k, float alpha, const float *A, int lda,
                   const float *B, int ldb, float beta, float *C, int ldc, cudaStream_t s, int tri);
int cgemm_tma_try(int ta, int tb, int m, int n, int k, cfloat alpha, const cfloat *A, int lda,
                  const cfloat *B, int ldb, cfloat beta, cfloat *C, int ldc, cudaStream_t s, int tri);
int zgemm_dmma_try(int ta, int tb, int m, int n, int k, cdouble alpha, const cdouble *A, int lda,
                   const cdouble *B, int ldb, cdouble beta, cdouble *C, int ldc, cudaStream_t s, int tri);

template <typename T> int lacpy_device(int trans, int m, int n, const T *a, int lda, T *b, int ldb, cudaStream_t s);
template <typename T> int getrs_upper_device(int n, int nrhs, const T *A, int lda, T *B, int ldb, cudaStream_t s);   // lapack_blocked.cu
// counts NaNs in an m x n column-major block (tri: 0 full, 'L'/'U'); result in *d_flag (device int, pre-zeroed)
template <typename T> int nancheck_device(int tri, int m, int n, const T *a, int lda, int *d_flag, cudaStream_t s);

// LAPACK-level device routines (column-major, device pointers; ipiv/info device ints)
template <typename T> int getrf_device(int m, int n, T *A, int lda, int *d_ipiv, int *d_info, cudaStream_t s);
template <typename T> int laswp_device(T *A, int lda, int ncols, const int *d_ipiv, int k0, int k1, cudaStream_t s);
template <typename T> int getrs_device(char trans, int n, int nrhs, const T *A, int lda, const int *d_ipiv, T *B, int ldb, cudaStream_t s);
template <typename T> int getri_device(int n, T *A, int lda, const int *d_ipiv, int *d_info, cudaStream_t s);
template <typename T> int potrf_device(char uplo, int n, T *A, int lda, int *d_info, cudaStream_t s);
template <typename T> int potrs_device(char uplo, int n, int nrhs, const T *A, int lda, T *B, int ldb, cudaStream_t s);
// side 141 left / 142 right, uplo 121 upper / 122 lower, trans 111/112/113, diag 131 non-unit / 132 unit
template <typename T> int trsm_device(int side, int uplo, int trans, int diag, int m, int n, T alpha, const T *A, int lda, T *B, int ldb, cudaStream_t s);
// the same solve as two GEMMs per 64 rows of the triangle (inverted diagonal blocks; trsm_inv.cu), alpha == 1
template <typename T> int trsm_inv_device(int side, int stored_lower, int trans, int unit, int m, int n, const T *A, int lda, T *B, int ldb, cudaStream_t s);
// wide enough for the GEMM form to pay (order of the triangle, number of right-hand sides / rows)
// the same solve for at most TRSV_MAX_NRHS right-hand sides: one bandwidth-bound launch per 64 rows (trsm_inv.cu)
constexpr int TRSV_MAX_NRHS = 8;
template <typename T> int trsv_few_device(int stored_lower, int trans, int unit, int n, int nrhs, const T *A, int lda, T *B, int ldb, cudaStream_t s);
inline bool trsv_few_pays(int nt, int nrhs) { return nrhs <= TRSV_MAX_NRHS && nt >= 256 && option("trsv.few") != 0; }
inline bool trsm_inv_pays(int nt, int other) { return nt >= 128 && other >= 256 && option("trsm.inv") != 0; }

}  // namespace jb
