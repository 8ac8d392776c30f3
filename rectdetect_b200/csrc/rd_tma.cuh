// rd_tma.cuh - tensor-map (TMA) tile loads: HBM -> shared memory by the copy engine (cp.async.bulk.tensor -> UTMALDG), completion on an
// mbarrier.  Used by the tile-staged kernels for the CTAs whose aproned tile lies inside the frame; CTAs on the frame border keep
// their plain staging (the reference mirrors / clamps coordinates there, the copy engine can only fill zeros).
// Rules learnt on the device (tools/probes/tma_probe.cu): the first element of a box must be 16-byte aligned - a box starting at
// x = 27 (32-bit elements) raises "illegal instruction", x = 28 works - so every box starts at a multiple of four columns and the
// kernels index into it with an offset; the driver wants each stride to be a multiple of the one before it.
// A/B of every adopting kernel: profiles/r04g_*.  RD_TMA=0 switches the tensor-map paths off.
#ifndef RD_TMA_CUH
#define RD_TMA_CUH
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

__device__ __forceinline__ void rd_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rd_mbar_expect(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rd_mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
      ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void rd_tma_load3(void *smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

static inline bool rd_tma_enabled() {
  static const bool on = !(getenv("RD_TMA") && atoi(getenv("RD_TMA")) == 0);
  return on;
}
// tensor maps of one plane of the arenas: rank 3 = (x, y, frame), frame stride fs
static inline bool rd_tma_make_map(CUtensorMap *m, const void *base, CUtensorMapDataType dt, int iw, int ih, int nb, size_t fs, int boxw, int boxh) {
  typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_t enc = NULL;
  if (!enc) {
    void *fn = NULL;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
    enc = (encode_t)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)iw, (cuuint64_t)ih, (cuuint64_t)nb};
  const cuuint64_t strides[2] = {(cuuint64_t)iw * 4, fs ? (cuuint64_t)fs : (cuuint64_t)iw * ih * 4};
  const cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1}, es[3] = {1, 1, 1};
  return enc(m, dt, 3, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// can planes of iw x ih 32-bit elements, frames fs bytes apart, be described by a tensor map?
static inline bool rd_tma_ok(const void *base, int iw, size_t fs) {
  return rd_tma_enabled() && (iw & 3) == 0 && ((uintptr_t)base & 15) == 0 && (fs & 15) == 0 && (fs == 0 || fs % ((size_t)iw * 4) == 0);
}
#endif
