// Auxiliary device kernels: copy / transpose-copy (the device replacement for the
// per-row cblas_?copy loop of unsafeMatrixCopy, /root/reference/Numeric/Jalla/Matrix.hs:855-878),
// LAPACKE-style NaN screening, counter-based synthetic inputs.
#include "jb_common.cuh"

namespace jb {

// ---------------------------------------------------------------- lacpy / transpose
template <typename T>
__global__ void lacpy_kernel(int m, int n, const T *__restrict__ a, int lda, T *__restrict__ b, int ldb) {
  // 2-D grid-stride copy; x walks rows (contiguous) so both sides are coalesced
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
      b[(size_t)i + (size_t)j * ldb] = a[(size_t)i + (size_t)j * lda];
}

// b (n x m) = a^T or a^H, a is m x n.  32x32 tiles through padded shared memory.
template <typename T>
__global__ void transpose_kernel(int conj, int m, int n, const T *__restrict__ a, int lda, T *__restrict__ b, int ldb) {
  __shared__ T tile[32][33];
  int tiles_m = (m + 31) / 32, tiles_n = (n + 31) / 32;
  for (long t = blockIdx.x; t < (long)tiles_m * tiles_n; t += gridDim.x) {
    int i0 = (int)(t % tiles_m) * 32, j0 = (int)(t / tiles_m) * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
      int i = i0 + threadIdx.x, j = j0 + jj;
      if (i < m && j < n) tile[jj][threadIdx.x] = a[(size_t)i + (size_t)j * lda];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
      int j = j0 + threadIdx.x, i = i0 + ii;
      if (i < m && j < n) {
        T v = tile[threadIdx.x][ii];
        b[(size_t)j + (size_t)i * ldb] = conj ? Num<T>::conj(v) : v;
      }
    }
    __syncthreads();
  }
}

template <typename T>
int lacpy_device(int trans, int m, int n, const T *a, int lda, T *b, int ldb, cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  if (trans == 111) {
    dim3 grid((unsigned)min((m + 255) / 256, 1024), (unsigned)min(n, 4096));
    lacpy_kernel<T><<<grid, 256, 0, s>>>(m, n, a, lda, b, ldb);
  } else {
    long tiles = (long)((m + 31) / 32) * ((n + 31) / 32);
    transpose_kernel<T><<<(unsigned)min(tiles, (long)num_sms() * 16), dim3(32, 8), 0, s>>>(trans == 113, m, n, a, lda, b, ldb);
  }
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 0;
}
template int lacpy_device<float>(int, int, int, const float *, int, float *, int, cudaStream_t);
template int lacpy_device<double>(int, int, int, const double *, int, double *, int, cudaStream_t);
template int lacpy_device<cfloat>(int, int, int, const cfloat *, int, cfloat *, int, cudaStream_t);
template int lacpy_device<cdouble>(int, int, int, const cdouble *, int, cdouble *, int, cudaStream_t);

// ---------------------------------------------------------------- NaN screening
// LAPACKE's high-level wrappers reject NaN inputs with info = -(argument index)
// (lapacke_utils.h:110,126 helpers; verified on OpenBLAS: gesv A -> -4, B -> -7).
template <typename T>
__global__ void nancheck_kernel(int tri, int m, int n, const T *__restrict__ a, int lda, int *flag) {
  bool bad = false;
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
      if (tri == 'L' && i < j) continue;
      if (tri == 'U' && i > j) continue;
      if (Num<T>::is_nan(a[(size_t)i + (size_t)j * lda])) bad = true;
    }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}
template <typename T>
int nancheck_device(int tri, int m, int n, const T *a, int lda, int *d_flag, cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  dim3 grid((unsigned)min((m + 255) / 256, 1024), (unsigned)min(n, 4096));
  nancheck_kernel<T><<<grid, 256, 0, s>>>(tri, m, n, a, lda, d_flag);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 0;
}
template int nancheck_device<float>(int, int, int, const float *, int, int *, cudaStream_t);
template int nancheck_device<double>(int, int, int, const double *, int, int *, cudaStream_t);
template int nancheck_device<cfloat>(int, int, int, const cfloat *, int, int *, cudaStream_t);
template int nancheck_device<cdouble>(int, int, int, const cdouble *, int, int *, cudaStream_t);

// ---------------------------------------------------------------- synthetic inputs
// value(seed,i,j): two rounds of the splitmix64 finaliser over (seed,i) then j; the top
// 53 (24) bits become U[-1,1).  jalla_b200/synth.py holds the bit-identical numpy twin.
__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31; return x;
}
__host__ __device__ inline uint64_t key2(uint64_t seed, uint64_t i, uint64_t j) {
  uint64_t x = mix64(seed + i * 0x9E3779B97F4A7C15ull);
  return mix64(x + j * 0xC2B2AE3D27D4EB4Full);
}
__device__ inline double u_double(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0; }
__device__ inline float u_float(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f) * 2.0f - 1.0f; }

template <typename T>
__global__ void fill_uniform_kernel(T *a, int lda, int m, int n, long long row0, long long col0, uint64_t seed, int sym,
                                    typename Num<T>::R diag) {
  for (int j = blockIdx.y; j < n; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
      uint64_t gi = (uint64_t)(row0 + i), gj = (uint64_t)(col0 + j);
      T v;
      if constexpr (Num<T>::cplx) {
        uint64_t hr = key2(seed, gi, 2 * gj), hi = key2(seed, gi, 2 * gj + 1);
        if constexpr (sizeof(typename Num<T>::R) == 8) v = Num<T>::make(u_double(hr), u_double(hi));
        else v = Num<T>::make(u_float(hr), u_float(hi));
      } else {
        uint64_t ki = gi, kj = gj;
        if (sym && ki > kj) { uint64_t t = ki; ki = kj; kj = t; }
        uint64_t h = key2(seed, ki, kj);
        if constexpr (sizeof(T) == 8) v = u_double(h); else v = u_float(h);
        if (sym && gi == gj) v += diag;
      }
      a[(size_t)i + (size_t)j * lda] = v;
    }
}
template <typename T>
static int fill_uniform(T *a, int lda, int m, int n, long long row0, long long col0, uint64_t seed, int sym,
                        typename Num<T>::R diag) {
  if (int rc = ensure_device()) return rc;
  if (m <= 0 || n <= 0) return 0;
  if (lda < m) { set_error(JB_ERR_ARG, "fill_uniform: lda < m"); return JB_ERR_ARG; }
  dim3 grid((unsigned)min((m + 255) / 256, 1024), (unsigned)min(n, 8192));
  fill_uniform_kernel<T><<<grid, 256, 0, tls().stream>>>(a, lda, m, n, row0, col0, seed, sym, diag);
  count_launch();
  JB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace jb

using namespace jb;
extern "C" {
#define LACPY_API(name, T)                                                                   \
  int name(int trans, int m, int n, const T *a, int lda, T *b, int ldb) {                    \
    if (int rc = ensure_device()) return rc;                                                 \
    if (trans < 111 || trans > 113 || m < 0 || n < 0) { set_error(JB_ERR_ARG, #name ": bad argument"); return JB_ERR_ARG; } \
    return lacpy_device<T>(trans, m, n, a, lda, b, ldb, tls().stream);                       \
  }
LACPY_API(jb_slacpy, float)
LACPY_API(jb_dlacpy, double)
int jb_clacpy(int trans, int m, int n, const void *a, int lda, void *b, int ldb) {
  if (int rc = ensure_device()) return rc;
  if (trans < 111 || trans > 113 || m < 0 || n < 0) { set_error(JB_ERR_ARG, "jb_clacpy: bad argument"); return JB_ERR_ARG; }
  return lacpy_device<cfloat>(trans, m, n, (const cfloat *)a, lda, (cfloat *)b, ldb, tls().stream);
}
int jb_zlacpy(int trans, int m, int n, const void *a, int lda, void *b, int ldb) {
  if (int rc = ensure_device()) return rc;
  if (trans < 111 || trans > 113 || m < 0 || n < 0) { set_error(JB_ERR_ARG, "jb_zlacpy: bad argument"); return JB_ERR_ARG; }
  return lacpy_device<cdouble>(trans, m, n, (const cdouble *)a, lda, (cdouble *)b, ldb, tls().stream);
}
int jb_dfill_uniform(double *a, int lda, int m, int n, long long row0, long long col0, uint64_t seed, int sym, double diag) {
  return fill_uniform<double>(a, lda, m, n, row0, col0, seed, sym, diag);
}
int jb_sfill_uniform(float *a, int lda, int m, int n, long long row0, long long col0, uint64_t seed, int sym, float diag) {
  return fill_uniform<float>(a, lda, m, n, row0, col0, seed, sym, diag);
}
int jb_zfill_uniform(void *a, int lda, int m, int n, long long row0, long long col0, uint64_t seed) {
  return fill_uniform<cdouble>((cdouble *)a, lda, m, n, row0, col0, seed, 0, 0.0);
}
int jb_cfill_uniform(void *a, int lda, int m, int n, long long row0, long long col0, uint64_t seed) {
  return fill_uniform<cfloat>((cfloat *)a, lda, m, n, row0, col0, seed, 0, 0.0f);
}
}
