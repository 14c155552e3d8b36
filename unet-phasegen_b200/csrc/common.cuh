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

// fp16 operand planes: the hi plane saturates to inf beyond 65504 (and NaN/inf inputs poison the MMAs)
__device__ __forceinline__ bool f16_fits(float x) { return fabsf(x) <= 65504.f; }   // false for NaN too

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// Destination of an activated tensor: the operand buffer of a consumer convolution (the skip concat of model.py:113 is a
// write at a channel offset) or an fp32 tensor.  Device-side mirror of pg_act_dst.
struct ActDst {
    void* hi; void* lo; long long batch_stride; int ld; int ch_off; int dtype; float slope; int* range_flag;
};

// returns true when a value written to an fp16 plane does not fit the fp16 range (or is not finite)
__device__ __forceinline__ bool store_act4(const ActDst& d, int b, int l, int c, float4 v) {
    v.x = leaky(v.x, d.slope); v.y = leaky(v.y, d.slope); v.z = leaky(v.z, d.slope); v.w = leaky(v.w, d.slope);
    const size_t o = (size_t)b * d.batch_stride + (size_t)l * d.ld + d.ch_off + c;
    bool bad = false;
    if (d.dtype == PG_DT_F32) {
        *reinterpret_cast<float4*>(static_cast<float*>(d.hi) + o) = v;
    } else {
        __align__(8) uint16_t h[4], lo[4];
        const int fmt = fmt_of_dtype(d.dtype);
        if (fmt == PG_FMT_F16) bad = !f16_fits(v.x) || !f16_fits(v.y) || !f16_fits(v.z) || !f16_fits(v.w);
        split16(v.x, fmt, h[0], lo[0]); split16(v.y, fmt, h[1], lo[1]);
        split16(v.z, fmt, h[2], lo[2]); split16(v.w, fmt, h[3], lo[3]);
        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(d.hi) + o) = *reinterpret_cast<uint2*>(h);
        if (d.dtype == PG_DT_BF16_SPLIT || d.dtype == PG_DT_F16_SPLIT) *reinterpret_cast<uint2*>(static_cast<uint16_t*>(d.lo) + o) = *reinterpret_cast<uint2*>(lo);
    }
    return bad;
}
// one element (the tensor-core epilogue owns one channel per thread)
__device__ __forceinline__ bool store_act1(const ActDst& d, int b, int l, int c, float v) {
    v = leaky(v, d.slope);
    const size_t o = (size_t)b * d.batch_stride + (size_t)l * d.ld + d.ch_off + c;
    if (d.dtype == PG_DT_F32) { static_cast<float*>(d.hi)[o] = v; return false; }
    const int fmt = fmt_of_dtype(d.dtype);
    uint16_t h, lo;
    split16(v, fmt, h, lo);
    static_cast<uint16_t*>(d.hi)[o] = h;
    if (d.dtype == PG_DT_BF16_SPLIT || d.dtype == PG_DT_F16_SPLIT) static_cast<uint16_t*>(d.lo)[o] = lo;
    return fmt == PG_FMT_F16 && !f16_fits(v);
}

// validate + copy a pg_act_dst (C = channels written); `who` names the entry point in the error message
inline int to_act_dst(const pg_act_dst* s, int C, ActDst* o, const char* who, const char* which) {
    o->dtype = 0; o->hi = o->lo = nullptr; o->batch_stride = 0; o->ld = 0; o->ch_off = 0; o->slope = 1.f; o->range_flag = nullptr;
    if (!s || s->dtype == PG_DT_NONE) return PG_OK;
    PG_REQUIRE(s->dtype >= PG_DT_F32 && s->dtype <= PG_DT_F16, "%s: %s: bad dtype %d", who, which, s->dtype);
    PG_REQUIRE(s->hi && ((s->dtype != PG_DT_BF16_SPLIT && s->dtype != PG_DT_F16_SPLIT) || s->lo), "%s: %s: null plane", who, which);
    PG_REQUIRE(s->ld % 4 == 0 && s->ch_off % 4 == 0 && s->batch_stride % 4 == 0 && s->ch_off + C <= s->ld, "%s: %s: misaligned or too narrow destination", who, which);
    o->hi = s->hi; o->lo = s->lo; o->batch_stride = s->batch_stride; o->ld = s->ld; o->ch_off = s->ch_off; o->dtype = s->dtype; o->slope = s->slope;
    o->range_flag = s->range_flag;
    return PG_OK;
}

}  // namespace pg
