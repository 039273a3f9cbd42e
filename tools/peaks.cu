// Micro-benchmarks that establish the roofline denominators this repo quotes:
// FP64 DMMA (mma.sync f64) issue rate per shape, DFMA, FFMA and an HBM copy.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/peaks tools/peaks.cu
// Output: one JSON object per line on stdout.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NACC independent accumulator chains per warp, m8n8k4. flops per mma per warp = 2*8*8*4 = 512
template <int NACC>
__global__ void k_dmma884(double *out, int iters, double seed) {
  double c0[NACC], c1[NACC];
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1684(double *out, int iters, double seed) {
  double c[NACC][4]; double a[2] = {seed, seed + 1e-9 * threadIdx.x}; double b = seed * 0.5;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma1684(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1688(double *out, int iters, double seed) {
  double c[NACC][4]; double a[4] = {seed, seed + 1e-9 * threadIdx.x, seed, seed}; double b[2] = {seed * 0.5, seed};
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma16816(double *out, int iters, double seed) {
  double c[NACC][4]; double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed + i * 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double *out, int iters, double seed) {
  double c[NACC]; double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_ffma(float *out, int iters, float seed) {
  float c[NACC]; float a = seed + threadIdx.x * 1e-6f, b = seed * 0.5f;
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fmaf(a, c[i], b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_copy(const double4 *__restrict__ in, double4 *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = in[i];
}

template <typename F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
#define RUN_DMMA(kern, name, nacc, flops_per_mma, ctas_per_sm, threads) do { \
    float ms = time_ms([&] { kern<nacc><<<sms * ctas_per_sm, threads>>>(out, iters, 1.0); }); \
    CK(cudaGetLastError()); \
    double fl = (double)sms * ctas_per_sm * (threads / 32) * (double)iters * nacc * flops_per_mma; \
    printf("{\"bench\": \"%s\", \"nacc\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", \
           name, nacc, ctas_per_sm * threads / 32, ms, fl / ms * 1e-9); fflush(stdout); } while (0)
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 8, 512.0, 1, 128);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 8, 512.0, 1, 256);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 16, 512.0, 1, 256);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 32, 512.0, 1, 256);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 16, 512.0, 2, 256);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 4, 512.0, 1, 128);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 2, 512.0, 1, 128);
  RUN_DMMA(k_dmma884, "dmma_m8n8k4", 1, 512.0, 1, 128);
  RUN_DMMA(k_dmma1684, "dmma_m16n8k4", 8, 1024.0, 1, 256);
  RUN_DMMA(k_dmma1684, "dmma_m16n8k4", 16, 1024.0, 1, 256);
  RUN_DMMA(k_dmma1688, "dmma_m16n8k8", 8, 2048.0, 1, 256);
  RUN_DMMA(k_dmma1688, "dmma_m16n8k8", 16, 2048.0, 1, 256);
  RUN_DMMA(k_dmma16816, "dmma_m16n8k16", 8, 4096.0, 1, 256);
  RUN_DMMA(k_dmma16816, "dmma_m16n8k16", 16, 4096.0, 1, 256);
  RUN_DMMA(k_dfma, "dfma", 16, 64.0, 1, 256);
  RUN_DMMA(k_dfma, "dfma", 16, 64.0, 2, 512);
  RUN_DMMA(k_dfma, "dfma", 8, 64.0, 4, 512);
  {
    float *fo = (float *)out;
    for (int cfg = 0; cfg < 2; cfg++) {
      int cps = cfg ? 4 : 2, th = 512;
      float ms = time_ms([&] { k_ffma<16><<<sms * cps, th>>>(fo, iters, 1.0f); });
      double fl = (double)sms * cps * th * (double)iters * 16 * 2;
      printf("{\"bench\": \"ffma\", \"nacc\": 16, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n",
             cps * th / 32, ms, fl / ms * 1e-9);
    }
  }
  {
    size_t bytes = (size_t)4 << 30; double4 *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 1, bytes));
    size_t n = bytes / sizeof(double4);
    float ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, n); });
    printf("{\"bench\": \"hbm_copy_kernel\", \"bytes_rw\": %zu, \"ms\": %.4f, \"gbs\": %.1f}\n", 2 * bytes, ms, 2.0 * bytes / ms * 1e-6);
    float ms2 = time_ms([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); });
    printf("{\"bench\": \"hbm_copy_memcpy\", \"bytes_rw\": %zu, \"ms\": %.4f, \"gbs\": %.1f}\n", 2 * bytes, ms2, 2.0 * bytes / ms2 * 1e-6);
  }
  return 0;
}
