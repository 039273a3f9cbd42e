// mbarrier / TMA / DMMA primitives shared by the FP64 tensor-pipe GEMM kernels (sm_100a).
#pragma once
#include "jb_common.cuh"

namespace jb {
namespace dmma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// 2-D tensor map over FP64 words: dim0 = `rows` contiguous words, dim1 = `cols`, pitch in BYTES, 128B swizzle.
inline int make_map_f64(CUtensorMap *map, const void *ptr, long rows, long cols, size_t pitch_bytes, int box0, int box1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error(JB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable"); return JB_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(JB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r); return JB_ERR_CUDA; }
  return 0;
}


// Tile sequence of one CTA of a persistent GEMM: band rasterisation (GROUP_M tile rows per band, column-major inside a band,
// for L2 reuse), CTA c takes entries c, c + grid, c + 2 grid, ... of the sequence.  In triangle mode (tri = 'L' / 'U': only
// tiles that touch that triangle are computed — Cholesky trailing updates, syrk / herk) the sequence holds ONLY the kept
// tiles, so every CTA gets the same number of them (+-1); striding over all tiles and skipping left the CTAs of a
// 52-wave update with 49..55 tiles each.  The walker advances a (band, tile column) cursor monotonically, O(columns) per
// CTA and launch in total.  Square tiles (BM == BN) assumed for the triangle test.
template <int GROUP_M>
struct TileWalker {
  int tiles_m, tiles_n, tri, balanced;
  int band = 0, tn = 0, base = 0;       // cursor: kept tiles before column tn of this band
  __device__ TileWalker(int tm_, int tn_, int tri_, int balanced_ = 1) : tiles_m(tm_), tiles_n(tn_), tri(tri_), balanced(balanced_) {}
  // coordinates of the idx-th tile of the sequence (idx must not decrease between calls); 0 when past the end, 1 = compute,
  // 2 = skip (only with balanced == 0: the round-1 order, every tile in the sequence and the unwanted ones skipped)
  __device__ __forceinline__ int at(int idx, int &tm_out, int &tn_out) {
    if (tri == 0 || !balanced) {
      if (idx >= tiles_m * tiles_n) return 0;
      const int per_band = GROUP_M * tiles_n, b = idx / per_band, in_band = idx % per_band;
      const int first_m = b * GROUP_M, gsz = min(tiles_m - first_m, GROUP_M);
      tm_out = first_m + in_band % gsz; tn_out = in_band / gsz;
      if (tri == 'L' && tm_out < tn_out) return 2;
      if (tri == 'U' && tn_out < tm_out) return 2;
      return 1;
    }
    while (band * GROUP_M < tiles_m) {
      const int first_m = band * GROUP_M, end_m = min(tiles_m, first_m + GROUP_M);
      // rows of column tn in this band that touch the triangle: 'L' tm >= tn, 'U' tm <= tn
      const int lo = tri == 'L' ? max(first_m, tn) : first_m;
      const int hi = tri == 'L' ? end_m : min(end_m, tn + 1);
      const int cnt = max(hi - lo, 0);
      if (idx < base + cnt) { tm_out = lo + (idx - base); tn_out = tn; return 1; }
      base += cnt;
      if (++tn == tiles_n) { tn = 0; ++band; }
    }
    return 0;
  }
};

}  // namespace dmma
// Asynchronous L2 prefetch of `bytes` (multiple of 16) at a 16-byte aligned global address.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

}  // namespace jb
