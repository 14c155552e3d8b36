// Microbenchmark (B200): issue / element rate of packed fp32x2 arithmetic (FFMA2, FADD2, FMUL2) against the scalar
// forms, 8 independent dependency chains per thread, full occupancy.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    const float m = 1.0000001f, c = 1e-7f;
    uint64_t p[8], pm, pc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i] = fmaf(a[i], m, c); }                                        // 8 FFMA
                if (MODE == 1) { asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pm), "l"(pc)); }   // 8 FFMA2
                if (MODE == 2) { a[i] = a[i] + c; }                                                // 8 FADD
                if (MODE == 3) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc)); }               // 8 FADD2
                if (MODE == 4) { a[i] = fmaf(a[i], m, c); a[i + 8] = fmaf(a[i + 8], m, c); }       // 16 FFMA (same elements as mode 1)
                if (MODE == 5) { asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm)); }               // 8 FMUL2
                if (MODE == 6) { a[i] = a[i] * m; }                                                // 8 FMUL
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
    if (s == 123.456f) out[0] = s;
}

template <int MODE> void run(const char* name, int instr_per_inner, int elems_per_instr) {
    float* out; cudaMalloc(&out, 4);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 4000, grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 10, 1.f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)grid * 8 * iters * 8.0 * instr_per_inner;      // 8 warps per CTA, 8 unrolled groups
    const double per_sm_per_ns = warp_instr / sms / (ms * 1e6);
    printf("%-28s %8.3f ms  %6.3f warp-instr/ns/SM  %7.1f elements/ns/SM   (max SM clock %.2f GHz)\n", name, ms, per_sm_per_ns,
           per_sm_per_ns * 32 * elems_per_instr, clk / 1e6);
    cudaFree(out);
}

int main() {
    run<0>("FFMA  x8 chains", 8, 1);
    run<4>("FFMA  x16 chains", 16, 1);
    run<1>("FFMA2 x8 chains", 8, 2);
    run<2>("FADD  x8 chains", 8, 1);
    run<3>("FADD2 x8 chains", 8, 2);
    run<6>("FMUL  x8 chains", 8, 1);
    run<5>("FMUL2 x8 chains", 8, 2);
    return 0;
}
