// TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a, and the host-side
// tensor-map encoder.  The driver entry points are fetched through the runtime
// (cudaGetDriverEntryPoint) so the library does not link libcuda and still
// loads on a machine without a driver (the CPU-only build check).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace sb200 {
namespace tma {

// ---- host ---------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using GetAddressRangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);

inline void* driver_entry_point(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult status;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &status) != cudaSuccess ||
      status != cudaDriverEntryPointSuccess) {
    while (cudaGetLastError() != cudaSuccess) {
    }
    return nullptr;
  }
  return fn;
}

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_entry_point("cuTensorMapEncodeTiled"));
  return fn;
}

inline GetAddressRangeFn get_address_range_fn() {
  static GetAddressRangeFn fn =
      reinterpret_cast<GetAddressRangeFn>(driver_entry_point("cuMemGetAddressRange"));
  return fn;
}

// True if [begin, end) lies inside one device allocation.
inline bool range_is_allocated(const void* begin, const void* end) {
  GetAddressRangeFn fn = get_address_range_fn();
  if (fn == nullptr) return false;
  CUdeviceptr base = 0;
  size_t size = 0;
  if (fn(&base, &size, reinterpret_cast<CUdeviceptr>(begin)) != CUDA_SUCCESS) return false;
  return reinterpret_cast<CUdeviceptr>(end) <= base + size;
}

// 3-D tiled tensor map, no swizzle, zero OOB fill; dims in elements, strides in bytes.
inline bool encode_3d(CUtensorMap* map, CUtensorMapDataType type, const void* base, uint64_t d0,
                      uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                      uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t elem[3] = {1, 1, 1};
  return fn(map, type, 3, const_cast<void*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class T>
constexpr CUtensorMapDataType tensor_type() {
  return sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

// 3-D tiled tensor map over 8-byte elements (used for float64 data and for
// pairs of float32): dims/strides in elements/bytes, no swizzle, zero OOB fill.
inline bool encode_3d_u64(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                          uint32_t b2) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t elem[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SB200_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SB200_DONE;\n"
      "bra SB200_WAIT;\n"
      "SB200_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// global -> shared, 3-D box at coordinates (c0, c1, c2), completion on `bar`
__device__ __forceinline__ void load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                        uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

}  // namespace tma
}  // namespace sb200
