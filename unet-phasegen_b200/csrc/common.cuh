// Shared helpers for the phasegen CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/phasegen.h"

namespace pg {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PG_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            pg::set_error(__VA_ARGS__);       \
            return PG_ERR_INVALID;            \
        }                                     \
    } while (0)

// x = hi + lo + O(2^-16 |x|): the two bf16 planes every tensor-core operand is stored as.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

}  // namespace pg
