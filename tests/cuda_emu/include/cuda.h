// Stand-in for the driver API header (tensor maps only): the emulated matcher never dereferences a tensor map.
#pragma once
#include "../cuda_emu.h"
typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
struct CUtensorMap { alignas(64) uint64_t opaque[16]; };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
constexpr unsigned long long cudaEnableDefault = 0;
inline CUresult emu_tensor_map_encode(CUtensorMap* tm, CUtensorMapDataType, cuuint32_t, void* base, const cuuint64_t* gdim,
                                      const cuuint64_t*, const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  tm->opaque[0] = (uint64_t)base; tm->opaque[1] = gdim[1]; tm->opaque[2] = box[1];
  return CUDA_SUCCESS;
}
inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q = nullptr) {
  *fn = (void*)&emu_tensor_map_encode;
  if (q) *q = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}
