// TMA / mbarrier helpers shared by the gather kernels (sm_100a): tensor maps are encoded on
// the host inside the C-ABI call (driver entry point looked up through the runtime, no link
// against libcuda) and handed to the kernels by value as __grid_constant__ parameters.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dfine {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TMA_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TMA_DONE;\n\t"
      "bra TMA_WAIT;\n\t"
      "TMA_DONE:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
// one box of a 3-D tensor map -> shared memory; completion is counted (in bytes) on `bar`
__device__ __forceinline__ void load_3d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0,
                                        int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Un-swizzled 3-D map over a dense-inner tensor: dims {d0, d1, d2} elements, byte strides
// {s1, s2} of dims 1 and 2, box {b0, b1, 1}.  Out-of-range box rows are filled with zeros.
// Returns 0 or a negative DFINE_E_* (message set).
inline int encode_3d_plain(CUtensorMap* map, bool bf16, const void* base, uint64_t d0, uint64_t d1,
                           uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1,
                           const char* what) {
  // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context current on this thread.
  // PyTorch's autograd threads may never have made a runtime call that binds it (201 =
  // CUDA_ERROR_INVALID_CONTEXT otherwise); cudaSetDevice on the current device does, is legal
  // under stream capture and costs nothing once bound.
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled is not available from the CUDA driver", what);
    return DFINE_E_UNSUPPORTED;
  }
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {s1, s2};
  const cuuint32_t box[3] = {b0, b1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                         3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
    return DFINE_E_SHAPE;
  }
  return 0;
}

}  // namespace tma
}  // namespace dfine
