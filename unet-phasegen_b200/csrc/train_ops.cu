// Training-step kernels around the convolutions (train.py:37-62), all HBM-bound:
//   * pg_phase_loss     MSE(cos p, cos phi) + MSE(sin p, sin phi) + 0.2 MSE(m, logmag) (train.py:45-60)
//                       and its gradient w.r.t. the network output, one pass
//   * pg_bn_bwd_reduce / pg_bn_bwd_apply
//                       backward of [train-mode batch norm -> ReLU / LeakyReLU fan-out] (model.py:80-83):
//                       dH = sum_j dA_j * act_j'(h);  dgamma = sum dH*zhat, dbeta = sum dH,
//                       dZ = gamma*invstd*(dH - dbeta/n - zhat*dgamma/n), written as the bf16 operand
//                       planes the dgrad / wgrad convolutions consume
//   * pg_wgrad_simt     exact-fp32 weight gradient (small channel counts, GPU-side check of wgrad_tc)
//   * pg_unpack_grad    packed [k][C_out][C_in] gradient -> torch Conv / ConvT weight layout
//   * pg_adam_step      torch.optim.Adam defaults (train.py:26-27), fused, fp32 state
#include "common.cuh"
#include "conv_plan.h"

namespace pg {

// ------------------------------------------------------------------------------------ loss
// out [rows][2C] channels-last (rows = B*T), logmag / phase [rows][C].  partial[block][3] =
// sums of squared errors (cos, sin, mag); d_out [rows][2C].
__global__ void __launch_bounds__(256)
phase_loss_kernel(const float* __restrict__ out, const float* __restrict__ logmag, const float* __restrict__ phase,
                  long long rows, int C, float inv_n, float mag_weight, float* __restrict__ d_out,
                  double* __restrict__ partial) {
    const long long total = rows * C;
    float s_cos = 0.f, s_sin = 0.f, s_mag = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / C; const int c = (int)(i % C);
        const float p = out[r * 2 * C + c], m = out[r * 2 * C + C + c];
        const float phi = phase[i], lm = logmag[i];
        float sp, cp, st, ct;
        sincosf(p, &sp, &cp);
        sincosf(phi, &st, &ct);
        const float dc = cp - ct, ds = sp - st, dm = m - lm;
        s_cos += dc * dc; s_sin += ds * ds; s_mag += dm * dm;
        if (d_out) {
            // d/dp [ (cos p - cos phi)^2 + (sin p - sin phi)^2 ] / n = 2 sin(p - phi) / n
            d_out[r * 2 * C + c] = 2.f * inv_n * (sp * ct - cp * st);
            d_out[r * 2 * C + C + c] = 2.f * inv_n * mag_weight * dm;
        }
    }
    __shared__ float red[3][8];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s_cos += __shfl_xor_sync(0xffffffffu, s_cos, o);
        s_sin += __shfl_xor_sync(0xffffffffu, s_sin, o);
        s_mag += __shfl_xor_sync(0xffffffffu, s_mag, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_cos; red[1][threadIdx.x >> 5] = s_sin; red[2][threadIdx.x >> 5] = s_mag; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256)
phase_loss_final_kernel(const double* __restrict__ partial, int n_blocks, double inv_n, float mag_weight,
                        float* __restrict__ loss3) {
    // deterministic: thread t sums the block partials t, t + 256, ... in order, then a fixed tree over the 256 sums
    // (one thread walking all partials took 57 us per step)
    __shared__ double sh[3][256];
    const int t = threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int b = t; b < n_blocks; b += 256) {
        s0 += partial[(size_t)b * 3 + 0]; s1 += partial[(size_t)b * 3 + 1]; s2 += partial[(size_t)b * 3 + 2];
    }
    sh[0][t] = s0; sh[1][t] = s1; sh[2][t] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) { sh[0][t] += sh[0][t + o]; sh[1][t] += sh[1][t + o]; sh[2][t] += sh[2][t + o]; }
        __syncthreads();
    }
    if (t == 0) {
        const float a = (float)(sh[0][0] * inv_n), b = (float)(sh[1][0] * inv_n), c = (float)(sh[2][0] * inv_n);
        loss3[1] = a; loss3[2] = b; loss3[3] = c;
        loss3[0] = a + b + mag_weight * c;
    }
}

// ------------------------------------------------------------------------------ BN backward
struct GradSrc { const float* g; int ld; int ch_off; float slope; };   // upstream dA_j, act_j slope

__device__ __forceinline__ float act_grad(float h, float slope) { return h > 0.f ? 1.f : slope; }

// partial[chunk][C] = {sum dH, sum dH*zhat} over the rows of the chunk.  block = (32 channel quads, 8 row lanes).
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ z, long long rows, int C, const float2* __restrict__ scale_shift,
                     const float2* __restrict__ mean_var, float eps, GradSrc g0, GradSrc g1, int rows_per_chunk,
                     float2* __restrict__ partial) {
    const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk; if (r1 > rows) r1 = rows;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (c < C) {
        float sc[4], sh[4], mu[4], is[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 ss = scale_shift[c + i], mv = mean_var[c + i];
            sc[i] = ss.x; sh[i] = ss.y; mu[i] = mv.x; is[i] = rsqrtf(mv.y + eps);
        }
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
            const float4 zv = *reinterpret_cast<const float4*>(z + r * C + c);
            const float4 a = *reinterpret_cast<const float4*>(g0.g + r * g0.ld + g0.ch_off + c);
            float4 b = make_float4(0, 0, 0, 0);
            if (g1.g) b = *reinterpret_cast<const float4*>(g1.g + r * g1.ld + g1.ch_off + c);
            const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float h = fmaf(zz[i], sc[i], sh[i]);
                const float dh = aa[i] * act_grad(h, g0.slope) + bb[i] * act_grad(h, g1.slope);
                s1[i] += dh;
                s2[i] += dh * (zz[i] - mu[i]) * is[i];
            }
        }
    }
    __shared__ float red[8][32][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[threadIdx.y][threadIdx.x][i] = s1[i]; red[threadIdx.y][threadIdx.x][4 + i] = s2[i]; }
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a = 0.f, b = 0.f;
            for (int y = 0; y < 8; ++y) { a += red[y][threadIdx.x][i]; b += red[y][threadIdx.x][4 + i]; }
            partial[(size_t)blockIdx.y * C + c + i] = make_float2(a, b);
        }
    }
}

// coef[c] = {sum dH / n, sum dH*zhat / n}; dgamma = sum dH*zhat, dbeta = sum dH
__global__ void bn_bwd_final_kernel(const float2* __restrict__ partial, int n_chunks, int C, double inv_n,
                                    float2* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
    for (int k = 0; k < n_chunks; ++k) { const float2 p = partial[(size_t)k * C + c]; a += p.x; b += p.y; }
    coef[c] = make_float2((float)(a * inv_n), (float)(b * inv_n));
    if (dgamma) dgamma[c] = (float)b;
    if (dbeta) dbeta[c] = (float)a;
}

// dZ -> operand planes [b][rows_alloc][C] (bf16 hi / hi+lo, or fp32 for the SIMT path)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ z, int L, int C, const float2* __restrict__ scale_shift,
                    const float2* __restrict__ mean_var, float eps, const float2* __restrict__ coef, GradSrc g0, GradSrc g1,
                    void* __restrict__ out_hi, void* __restrict__ out_lo, int out_rows, int out_dtype) {
    const int c4 = C >> 2;
    const int b = blockIdx.y;
    const size_t total = (size_t)L * c4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(i / c4), c = (int)(i % c4) * 4;
        const long long r = (long long)b * L + l;
        const float4 zv = *reinterpret_cast<const float4*>(z + r * C + c);
        const float4 a = *reinterpret_cast<const float4*>(g0.g + r * g0.ld + g0.ch_off + c);
        float4 bq = make_float4(0, 0, 0, 0);
        if (g1.g) bq = *reinterpret_cast<const float4*>(g1.g + r * g1.ld + g1.ch_off + c);
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {bq.x, bq.y, bq.z, bq.w};
        float dz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (scale_shift) {
                const float2 ss = scale_shift[c + k], mv = mean_var[c + k], cf = coef[c + k];
                const float h = fmaf(zz[k], ss.x, ss.y);
                const float dh = aa[k] * act_grad(h, g0.slope) + bb[k] * act_grad(h, g1.slope);
                const float zh = (zz[k] - mv.x) * rsqrtf(mv.y + eps);
                dz[k] = ss.x * (dh - cf.x - zh * cf.y);
            } else {
                dz[k] = aa[k] * act_grad(zz[k], g0.slope) + bb[k] * act_grad(zz[k], g1.slope);
            }
        }
        const size_t o = ((size_t)b * out_rows + l) * C + c;
        if (out_dtype == PG_DT_F32) {
            *reinterpret_cast<float4*>(static_cast<float*>(out_hi) + o) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        } else {
            __nv_bfloat16 h[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(dz[k], h[k], lo[k]);
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out_hi) + o) = *reinterpret_cast<uint2*>(h);
            if (out_dtype == PG_DT_BF16_SPLIT) *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out_lo) + o) = *reinterpret_cast<uint2*>(lo);
        }
    }
}

// ------------------------------------------------------------------------------ SIMT wgrad
// dW[tap][co][ci] = sum_{b,m} G[b][m*OS + phase][co] * X[b][(m + d)*IS + parity][ci]
__global__ void wgrad_simt_kernel(const __grid_constant__ ConvPlan pl, const float* __restrict__ x, int in_rows, int in_ld,
                                  const float* __restrict__ g, int g_rows, float* __restrict__ dw) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    const int co = blockIdx.y * blockDim.y + threadIdx.y;
    if (ci >= pl.C_in || co >= pl.C_out) return;
    int t = blockIdx.z, phase = 0;
    while (t >= pl.n_taps[phase]) { t -= pl.n_taps[phase]; ++phase; }
    const ConvTap tp = pl.taps[phase][t];
    const int l_phase = (pl.L_out - phase + pl.OS - 1) / pl.OS;
    float acc = 0.f;
    for (int b = 0; b < pl.B; ++b)
        for (int m = 0; m < l_phase; ++m) {
            const int row = (m + tp.d) * pl.IS + tp.parity;
            if (row < 0 || row >= pl.L_in) continue;
            acc = fmaf(g[((size_t)b * g_rows + m * pl.OS + phase) * pl.C_out + co], x[((size_t)b * in_rows + row) * in_ld + ci], acc);
        }
    dw[((size_t)tp.w_idx * pl.C_out + co) * pl.C_in + ci] = acc;
}

// ---------------------------------------------------------------------------- grad unpack
__global__ void unpack_grad_kernel(const float* __restrict__ packed, int transposed, int C_in, int C_out, int k,
                                   float* __restrict__ out) {
    const size_t total = (size_t)k * C_out * C_in;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % C_in);
        const int co = (int)((i / C_in) % C_out);
        const int t = (int)(i / ((size_t)C_in * C_out));
        const size_t o = transposed ? ((size_t)ci * C_out + co) * k + t : ((size_t)co * C_in + ci) * k + t;
        out[o] = packed[i];
    }
}

// ------------------------------------------------------------------------------------ Adam
__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float lr_bc1, float b1, float b2, float eps,
                                          float bc2_sqrt_inv, float grad_scale) {
    const float gi = g * grad_scale;
    m = b1 * m + (1.f - b1) * gi;
    v = b2 * v + (1.f - b2) * gi * gi;
    // torch.optim.Adam: p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
    p = p - lr_bc1 * m / (sqrtf(v) * bc2_sqrt_inv + eps);
    return p;
}

// HBM-bound: 16 B read + 12 B written per parameter (+2..4 B of refreshed operand planes); 128-bit accesses.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale,
            uint16_t* __restrict__ w_hi, uint16_t* __restrict__ w_lo, int g_bf16) {
    const __nv_bfloat16* g16 = reinterpret_cast<const __nv_bfloat16*>(g);
    const float lr_bc1 = lr / bc1, bc2_sqrt_inv = 1.f / bc2_sqrt;
    const size_t n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                      ((reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo)) & 7) == 0 ? n / 4 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        // streaming (evict-first) accesses: 28 B/parameter pass through exactly once per step
        float4 pi = __ldcs(reinterpret_cast<const float4*>(p) + i);
        float4 gi;
        if (g_bf16) {
            const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(g16) + i);
            const float2 lo2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
            const float2 hi2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
            gi = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
        } else {
            gi = __ldcs(reinterpret_cast<const float4*>(g) + i);
        }
        float4 mi = __ldcs(reinterpret_cast<const float4*>(m) + i), vi = __ldcs(reinterpret_cast<const float4*>(v) + i);
        adam_one(pi.x, gi.x, mi.x, vi.x, lr_bc1, b1, b2, eps, bc2_sqrt_inv, grad_scale);
        adam_one(pi.y, gi.y, mi.y, vi.y, lr_bc1, b1, b2, eps, bc2_sqrt_inv, grad_scale);
        adam_one(pi.z, gi.z, mi.z, vi.z, lr_bc1, b1, b2, eps, bc2_sqrt_inv, grad_scale);
        adam_one(pi.w, gi.w, mi.w, vi.w, lr_bc1, b1, b2, eps, bc2_sqrt_inv, grad_scale);
        __stcs(reinterpret_cast<float4*>(p) + i, pi);
        __stcs(reinterpret_cast<float4*>(m) + i, mi);
        __stcs(reinterpret_cast<float4*>(v) + i, vi);
        if (w_hi) {
            __align__(8) uint16_t h[4], l[4];
            split16(pi.x, PG_FMT_BF16, h[0], l[0]); split16(pi.y, PG_FMT_BF16, h[1], l[1]);
            split16(pi.z, PG_FMT_BF16, h[2], l[2]); split16(pi.w, PG_FMT_BF16, h[3], l[3]);
            __stcs(reinterpret_cast<uint2*>(w_hi) + i, *reinterpret_cast<uint2*>(h));
            if (w_lo) __stcs(reinterpret_cast<uint2*>(w_lo) + i, *reinterpret_cast<uint2*>(l));
        }
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_one(pi, g_bf16 ? __bfloat162float(g16[i]) : g[i], mi, vi, lr_bc1, b1, b2, eps, bc2_sqrt_inv, grad_scale);
        p[i] = pi; m[i] = mi; v[i] = vi;
        if (w_hi) {
            uint16_t h, l;
            split16(pi, PG_FMT_BF16, h, l);
            w_hi[i] = h;
            if (w_lo) w_lo[i] = l;
        }
    }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_phase_loss(const float* out, const float* logmag, const float* phase, int64_t rows, int C, float mag_weight,
                             float* d_out, double* partial, int n_blocks, float* loss3, pg_stream stream) {
    PG_REQUIRE(out && logmag && phase && partial && loss3 && rows > 0 && C > 0 && n_blocks > 0 && n_blocks <= 4096, "pg_phase_loss: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double inv_n = 1.0 / ((double)rows * C);
    phase_loss_kernel<<<n_blocks, 256, 0, st>>>(out, logmag, phase, rows, C, (float)inv_n, mag_weight, d_out, partial);
    phase_loss_final_kernel<<<1, 256, 0, st>>>(partial, n_blocks, inv_n, mag_weight, loss3);
    return check_launch("phase_loss_kernel");
}

static GradSrc to_src(const pg_grad_src* s) {
    GradSrc o; o.g = nullptr; o.ld = 0; o.ch_off = 0; o.slope = 0.f;
    if (s && s->g) { o.g = s->g; o.ld = s->ld; o.ch_off = s->ch_off; o.slope = s->slope; }
    return o;
}

extern "C" int pg_bn_bwd(const float* z, int B, int L, int C, const float* scale_shift, const float* mean_var, float eps,
                         const pg_grad_src* g0, const pg_grad_src* g1, float* partial, int n_chunks, float* coef,
                         float* dgamma, float* dbeta, void* dz_hi, void* dz_lo, int dz_rows, int dz_dtype, pg_stream stream) {
    PG_REQUIRE(z && g0 && g0->g && dz_hi && B > 0 && L > 0 && C > 0 && C % 4 == 0 && B <= 65535, "pg_bn_bwd: bad arguments");
    PG_REQUIRE(dz_dtype == PG_DT_F32 || dz_dtype == PG_DT_BF16 || (dz_dtype == PG_DT_BF16_SPLIT && dz_lo), "pg_bn_bwd: bad output dtype");
    PG_REQUIRE(g0->ld % 4 == 0 && g0->ch_off % 4 == 0 && (!g1 || !g1->g || (g1->ld % 4 == 0 && g1->ch_off % 4 == 0)), "pg_bn_bwd: misaligned gradient source");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const GradSrc s0 = to_src(g0), s1 = to_src(g1);
    const long long rows = (long long)B * L;
    const float2* ss = reinterpret_cast<const float2*>(scale_shift);
    const float2* mv = reinterpret_cast<const float2*>(mean_var);
    if (ss) {
        PG_REQUIRE(mv && partial && coef && n_chunks > 0, "pg_bn_bwd: statistics / workspace missing");
        const int rpc = (int)((rows + n_chunks - 1) / n_chunks);
        dim3 grid((C / 4 + 31) / 32, n_chunks);
        bn_bwd_reduce_kernel<<<grid, dim3(32, 8), 0, st>>>(z, rows, C, ss, mv, eps, s0, s1, rpc, reinterpret_cast<float2*>(partial));
        bn_bwd_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float2*>(partial), n_chunks, C, 1.0 / (double)rows,
                                                           reinterpret_cast<float2*>(coef), dgamma, dbeta);
    }
    const size_t total = (size_t)L * (C / 4);
    int gx = (int)((total + 255) / 256); if (gx > 1024) gx = 1024;
    bn_bwd_apply_kernel<<<dim3(gx, B), 256, 0, st>>>(z, L, C, ss, mv, eps, reinterpret_cast<const float2*>(coef), s0, s1,
                                                    dz_hi, dz_lo, dz_rows, dz_dtype);
    return check_launch("bn_bwd kernels");
}

extern "C" int pg_wgrad_simt(const pg_conv_desc* d, const float* x, const float* g, int g_rows, float* dw_packed, pg_stream stream) {
    PG_REQUIRE(d && x && g && dw_packed, "pg_wgrad_simt: null pointer");
    ConvPlan pl;
    pg_conv_desc dd = *d; dd.taps_per_group = 1;
    int rc = conv_plan_build(&dd, &pl);
    if (rc != PG_OK) return rc;
    dim3 block(32, 8);
    dim3 grid((pl.C_in + 31) / 32, (pl.C_out + 7) / 8, pl.n_taps[0] + pl.n_taps[1]);
    wgrad_simt_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pl, x, d->in_rows, d->in_ld, g, g_rows, dw_packed);
    return check_launch("wgrad_simt_kernel");
}

extern "C" int pg_unpack_grad(const float* packed, int kind, int C_in, int C_out, int k, float* out, pg_stream stream) {
    PG_REQUIRE(packed && out && C_in > 0 && C_out > 0 && k > 0, "pg_unpack_grad: bad arguments");
    const size_t total = (size_t)k * C_out * C_in;
    int gx = (int)((total + 255) / 256); if (gx > 148 * 16) gx = 148 * 16;
    unpack_grad_kernel<<<gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(packed, kind == PG_CONV_TRANSPOSE, C_in, C_out, k, out);
    return check_launch("unpack_grad_kernel");
}

extern "C" int pg_adam_step(float* p, const void* g, int g_dtype, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                            float eps, int step, float grad_scale, uint16_t* w_hi, uint16_t* w_lo, pg_stream stream) {
    PG_REQUIRE(g_dtype == PG_DT_F32 || g_dtype == PG_DT_BF16, "pg_adam_step: gradient dtype must be PG_DT_F32 or PG_DT_BF16");
    PG_REQUIRE(p && g && m && v && n > 0 && step >= 1, "pg_adam_step: bad arguments");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    int gx = (int)((n / 4 + 255) / 256); if (gx > 148 * 16) gx = 148 * 16; if (gx < 1) gx = 1;
    adam_kernel<<<gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, static_cast<const float*>(g), m, v, (size_t)n, lr, beta1, beta2, eps, bc1, bc2s, grad_scale, w_hi, w_lo, g_dtype == PG_DT_BF16 ? 1 : 0);
    return check_launch("adam_kernel");
}
