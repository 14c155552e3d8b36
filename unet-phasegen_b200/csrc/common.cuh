// Shared helpers for the phasegen CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/phasegen.h"

namespace pg {

void set_error(const char* fmt, ...);
// cudaFuncSetAttribute is per device: opt-in bookkeeping is keyed by the current device ordinal
constexpr int kMaxDevices = 64;
inline int current_device_slot() { int dev = 0; cudaGetDevice(&dev); return dev >= 0 && dev < kMaxDevices ? dev : 0; }
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

// Same split on raw 16-bit words for either operand format (PG_FMT_BF16: 8+8 significant bits,
// x = hi + lo to 2^-16; PG_FMT_F16: 11+11 bits, to 2^-22 while |x| stays inside the fp16 range).
__device__ __forceinline__ void split16(float x, int fmt, uint16_t& hi, uint16_t& lo) {
    if (fmt == PG_FMT_F16) {
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        hi = __half_as_ushort(h); lo = __half_as_ushort(l);
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
        hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(l);
    }
}
__device__ __forceinline__ int fmt_of_dtype(int dtype) { return (dtype == PG_DT_F16_SPLIT || dtype == PG_DT_F16) ? PG_FMT_F16 : PG_FMT_BF16; }

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

}  // namespace pg
