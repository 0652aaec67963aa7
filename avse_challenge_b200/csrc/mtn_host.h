// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/mtn_b200.h"

namespace mtn {

void set_error(const char* fmt, ...);
int num_sms();

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no -lcuda link).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1.
bool encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                 CUtensorMapL2promotion l2_promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_128B);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to one (function, DEVICE) pair, so the "already set"
// cache is a per-call-site bit mask over device ordinals (atomic: the C ABI is re-entrant).  Devices >= 64 are
// simply set on every launch.
int ensure_dyn_smem(const void* kern, int bytes, std::atomic<unsigned long long>& done_mask, const char* what);

#define MTN_REQUIRE(cond, ...)           \
    do {                                 \
        if (!(cond)) {                   \
            mtn::set_error(__VA_ARGS__); \
            return MTN_EINVAL;           \
        }                                \
    } while (0)

#define MTN_CUDA_LAUNCH_CHECK(what)                                                \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            mtn::set_error("%s: launch failed: %s", what, cudaGetErrorString(e__)); \
            return MTN_ECUDA;                                                      \
        }                                                                          \
    } while (0)

}  // namespace mtn
