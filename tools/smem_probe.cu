// Throughput probes for the shared-memory broadcast patterns the batched LU kernel chooses between, and an
// exactness probe of DMMA.8x8x4 against sequential fma chains (sm_100a).  One CTA on one SM; cycles per
// warp-instruction at SM level = elapsed / (N * warps).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/smem_probe tools/smem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 2048

__device__ __forceinline__ void lds128(double &x, double &y, uint32_t addr) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds64(double &x, uint32_t addr) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr) : "memory"); }
__device__ __forceinline__ void sts64p(uint32_t addr, double x, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(addr), "d"(x), "r"(pred) : "memory");
}
__device__ __forceinline__ void sts32p(uint32_t addr, int x, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.b32 [%0], %1; }" ::"r"(addr), "r"(x), "r"(pred) : "memory");
}
__device__ __forceinline__ void sts128p(uint32_t addr, double x, double y, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(addr), "d"(x), "d"(y), "r"(pred) : "memory");
}

// OP 0: LDS.64 broadcast (one address)      1: LDS.128 broadcast (one address)
//    2: LDS.128, two addresses (half-warps)  3: LDS.128, four addresses (quarter-warps)
//    4: LDS.128 all lanes distinct           5: STS.128 predicated, one active lane
//    6: STS.128 predicated, two active lanes (lanes 3 and 19)   7: STS.128 with no active lane
//    8: CREDUX.MAX (independent)             9: SHFL.IDX (independent)
//   10: LDS.64, two addresses (half-warps)  11: LDS.64 all lanes distinct (conflict-free)
//   12: DFMA independent (8 accumulators)   13: LDS.128 2-address + 4 DFMA per load (the A2 inner loop shape)
//   14: LDS.128 1-address + 2 DFMA per load (the A inner loop shape)
template <int OP> __global__ void k(long long *out, double seed) {
  extern __shared__ __align__(16) double sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = seed + i;
  __syncthreads();
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sh) + warp * 1024;
  uint32_t a1 = base, a2 = base + (lane >> 4) * 528, a4 = base + (lane >> 3) * 272, aall = base + lane * 16, a2_64 = base + (lane >> 4) * 264, aall64 = base + lane * 8;
  double acc[8] = {seed, seed, seed, seed, seed, seed, seed, seed};
  int iv = lane + (int)seed;
  int ivs[8] = {iv, iv + 1, iv + 2, iv + 3, iv + 4, iv + 5, iv + 6, iv + 7};
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      double x = 0, y = 0;
      const uint32_t off = u * 16 + (it & 3) * 128;
      if (OP == 0) { lds64(x, a1 + off); acc[u] += x; }
      if (OP == 1) { lds128(x, y, a1 + off); acc[u] += x + y; }
      if (OP == 2) { lds128(x, y, a2 + off); acc[u] += x + y; }
      if (OP == 3) { lds128(x, y, a4 + off); acc[u] += x + y; }
      if (OP == 4) { lds128(x, y, aall + (u & 1) * 512); acc[u] += x + y; }
      if (OP == 5) sts128p(a1 + off, acc[u], acc[u], lane == 3);
      if (OP == 6) sts128p(a2 + off, acc[u], acc[u], (lane & 15) == 3);
      if (OP == 7) sts128p(a1 + off, acc[u], acc[u], lane == 99);
      if (OP == 8) ivs[u] = __reduce_max_sync(0xffffffffu, ivs[u]) + lane;
      if (OP == 9) ivs[u] = __shfl_sync(0xffffffffu, ivs[u], 5) + lane;
      if (OP == 10) { lds64(x, a2_64 + off); acc[u] += x; }
      if (OP == 11) { lds64(x, aall64 + (u & 1) * 256); acc[u] += x; }
      if (OP == 12) acc[u] = __fma_rn(acc[u], 1.0000001, seed);
      if (OP == 13) { lds128(x, y, a2 + off); acc[u] = __fma_rn(acc[u], x, y); acc[(u + 1) & 7] = __fma_rn(acc[(u + 1) & 7], y, x);
                      acc[(u + 2) & 7] = __fma_rn(acc[(u + 2) & 7], x, y); acc[(u + 3) & 7] = __fma_rn(acc[(u + 3) & 7], y, x); }
      if (OP == 15) sts128p(aall + (u & 1) * 512, acc[u], acc[u], 1);
      if (OP == 16) sts64p(a1 + off, acc[u], lane == 3);
      if (OP == 17) sts32p(a1 + off, iv, lane == 3);
      if (OP == 18) sts128p(a1 + off, acc[u], acc[u], (lane & 7) == 3);
      if (OP == 14) { lds128(x, y, a1 + off); acc[u] = __fma_rn(acc[u], x, y); acc[(u + 1) & 7] = __fma_rn(acc[(u + 1) & 7], y, x); }
    }
  }
  long long t1 = clock64();
  double s = 0; int si = 0;
  for (int u = 0; u < 8; u++) { s += acc[u]; si += ivs[u]; }
  if (s == 123.0 || si == 77) out[1] = 1;
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
}
template <int OP> void run(const char *name, int warps) {
  long long *d; cudaMalloc(&d, 16);
  k<OP><<<1, 32 * warps, 32768>>>(d, 1.000001); cudaDeviceSynchronize();
  k<OP><<<1, 32 * warps, 32768>>>(d, 1.000001);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("{\"op\": \"%s\", \"warps\": %d, \"sm_cycles_per_warp_instr\": %.3f}\n", name, warps, (double)h[0] / N / warps);
  cudaFree(d);
}

// ---- DMMA exactness: is mma.sync.m8n8k4.f64 the chain fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c)))) bit for bit? ----
__global__ void dmma_probe(const double *A, const double *B, const double *Cin, double *D) {   // A 8x4 row-major, B 4x8 row-major (k x n), C 8x8 row-major
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a = A[g * 4 + t], b = B[t * 8 + g], c0 = Cin[g * 8 + 2 * t], c1 = Cin[g * 8 + 2 * t + 1], d0, d1;
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
  D[g * 8 + 2 * t] = d0; D[g * 8 + 2 * t + 1] = d1;
}
static void dmma_check() {
  const int T = 4096;
  double hA[32], hB[32], hC[64], hD[64];
  double *dA, *dB, *dC, *dD; cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512); cudaMalloc(&dD, 512);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0; };
  long fwd = 0, bwd = 0, pair = 0, total = 0;
  for (int it = 0; it < T; it++) {
    for (auto &v : hA) v = rnd(); for (auto &v : hB) v = rnd(); for (auto &v : hC) v = rnd();
    cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice); cudaMemcpy(dC, hC, 512, cudaMemcpyHostToDevice);
    dmma_probe<<<1, 32>>>(dA, dB, dC, dD);
    cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {
      double f = hC[i * 8 + j], r = hC[i * 8 + j];
      for (int k = 0; k < 4; k++) f = __builtin_fma(hA[i * 4 + k], hB[k * 8 + j], f);
      for (int k = 3; k >= 0; k--) r = __builtin_fma(hA[i * 4 + k], hB[k * 8 + j], r);
      double p = __builtin_fma(hA[i * 4 + 1], hB[8 + j], hA[i * 4] * hB[j]) + __builtin_fma(hA[i * 4 + 3], hB[24 + j], hA[i * 4 + 2] * hB[16 + j]) + hC[i * 8 + j];
      total++; fwd += (hD[i * 8 + j] == f); bwd += (hD[i * 8 + j] == r); pair += (hD[i * 8 + j] == p);
    }
  }
  printf("{\"dmma_m8n8k4\": {\"samples\": %ld, \"equals_fma_chain_k0_first\": %ld, \"equals_fma_chain_k3_first\": %ld, \"equals_pairwise\": %ld}}\n", total, fwd, bwd, pair);
}

int main() {
  for (int w : {4, 16}) {
    run<0>("lds64_bcast_1addr", w); run<1>("lds128_bcast_1addr", w); run<2>("lds128_2addr_halfwarp", w); run<3>("lds128_4addr_quarterwarp", w);
    run<4>("lds128_distinct", w); run<10>("lds64_2addr_halfwarp", w); run<11>("lds64_distinct", w);
    run<5>("sts128_pred_1lane", w); run<6>("sts128_pred_2lanes", w); run<7>("sts128_pred_0lanes", w);
    run<8>("credux_max", w); run<9>("shfl_idx", w); run<12>("dfma_indep", w);
    run<13>("lds128_2addr+4dfma", w); run<14>("lds128_1addr+2dfma", w);
    run<15>("sts128_all_lanes", w); run<16>("sts64_pred_1lane", w); run<17>("sts32_pred_1lane", w); run<18>("sts128_pred_4lanes", w);
  }
  dmma_check();
  return 0;
}
