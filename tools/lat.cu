// Latency / throughput probes for the instructions on the batched-LU critical path (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(long long *out, double seed, int *ibuf) {
  __shared__ double sh[64];
  double x = seed + threadIdx.x * 1e-3, y = seed * 0.5;
  int iv = threadIdx.x + (int)seed; unsigned uv = iv;
  sh[threadIdx.x & 63] = x;
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) {
    if (OP == 0) x = __fma_rn(x, y, y);                       // dependent DFMA
    if (OP == 1) x = __dadd_rn(x, y);                         // dependent DADD
    if (OP == 2) x = __dmul_rn(x, y);                         // dependent DMUL
    if (OP == 3) { x = (x > y) ? x + 1.0 : x - 1.0; }         // DSETP + select-ish
    if (OP == 4) iv = __reduce_max_sync(0xffffffffu, iv) + threadIdx.x;   // CREDUX chain
    if (OP == 5) uv = __ballot_sync(0xffffffffu, uv & 1) + threadIdx.x;   // VOTE chain
    if (OP == 6) iv = __shfl_sync(0xffffffffu, iv, (iv + 1) & 31);        // SHFL chain
    if (OP == 7) { sh[threadIdx.x & 31] = x; __syncwarp(); x = sh[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }  // STS->sync->LDS->DADD
    if (OP == 8) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }   // MUFU.RCP64H chain
    if (OP == 9) x = fabs(x) * y;                              // DMUL w/ abs modifier
    if (OP == 10) { float f = __int_as_float(iv); f = f * 1.0001f + 1.0f; iv = __float_as_int(f); }  // FFMA chain
    if (OP == 11) iv = iv * 3 + 1;                              // IMAD chain
  }
  long long t1 = clock64();
  if (x == 123.0 || iv == 77 || uv == 99) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int OP> void run(const char *name, int warps) {
  long long *d; cudaMalloc(&d, 16); int *ib; cudaMalloc(&ib, 4096);
  k<OP><<<1, 32 * warps>>>(d, 1.000001, ib); cudaDeviceSynchronize();
  k<OP><<<1, 32 * warps>>>(d, 1.000001, ib);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("{\"op\": \"%s\", \"warps\": %d, \"cycles_per_iter\": %.2f}\n", name, warps, (double)h[0] / N);
  cudaFree(d); cudaFree(ib);
}
int main() {
  for (int w : {1, 4, 16}) {
    run<0>("dfma_dep", w); run<1>("dadd_dep", w); run<2>("dmul_dep", w); run<3>("dsetp_sel_dadd", w);
    run<4>("credux_dep(+iadd)", w); run<5>("vote_dep(+iadd)", w); run<6>("shfl_dep", w);
    run<7>("sts_sync_lds_dadd_sync", w); run<8>("mufu_rcp64h_dep", w); run<10>("ffma_dep", w); run<11>("imad_dep", w);
  }
  return 0;
}
