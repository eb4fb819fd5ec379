// TEST INFRASTRUCTURE -- the few driver-API types csrc/tcgen05_util.cuh and csrc/aggregate.cu name (tensor maps); the
// host build never encodes one.
#pragma once
#include <cstdint>

typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
struct alignas(64) CUtensorMap {
    uint64_t opaque[16];
};
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
