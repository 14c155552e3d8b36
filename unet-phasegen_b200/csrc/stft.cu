// STFT front end and ISTFT back end (hand-written, sm_100a).
//
// Reference path replaced:
//   * librosa.stft call + DC-row drop + re/im split   preproc_mdb.py:84-97 (call at :93)
//   * abs / log1p / angle                             data.py:39-47
//   * expm1(mag) * exp(j*phase)                       demo.py:39, train.py:83
//   * zero DC row + librosa.istft + peak normalise    utils.py:34-42
//
// Both kernels keep everything between the HBM read and the HBM write on chip:
//   STFT : wave span -> smem (reflect padding resolved while loading) -> Hann window ->
//          n_fft-point real FFT as an n_fft/2-point complex Stockham FFT (fft_core.cuh) ->
//          Hermitian post-processing -> |X|, log1p, atan2 -> fp32 frame-major [B][T][C]
//          (+ the bf16 hi/lo operand planes the first convolution consumes).
//   ISTFT: (log-mag, phase) or (re, im) frame-major -> X -> half-length inverse FFT ->
//          synthesis window -> GATHER overlap-add of the 4 frames that cover each output
//          sample + window-sum-square normalisation -> fp32 wave.  No atomics on the output;
//          the per-clip peak (for utils.py:42) is one atomicMax per CTA on a scalar.
// hop must be n_fft/4 (true for every configuration of the reference and of BASELINE.json).
#include "common.cuh"
#include "fft_core.cuh"

namespace pg {
using namespace pgfft;

template <int NC, int FR_> struct StftCfg {
    static constexpr int TG = NC / 16;                      // threads per transform
    static constexpr int FR = FR_;                          // frames per CTA
    static constexpr int THREADS = TG * FR;
    static constexpr int NFFT = 2 * NC;
    static constexpr int HOP = NFFT / 4;
    static constexpr int PL = padded_len(NC);
    static constexpr int SPAN = (FR - 1) * HOP + NFFT;      // wave samples a CTA touches
    static constexpr size_t smem_stft() { return sizeof(float) * (3 * NFFT + SPAN + 2 * PL * FR); }
    static constexpr size_t smem_istft() { return sizeof(float) * (3 * NFFT + 2 * PL * FR); }
};

// Forward transform whose first pass takes its inputs straight from the windowed wave span
// (z[m] = x[2m] w[2m] + i x[2m+1] w[2m+1]) instead of a staged copy.
template <int NC>
__device__ __forceinline__ void fft_forward_from_span(float* sre, float* sim, const cpx* tw, const float* x, const float* win, int t) {
    using P = Plan<NC>;
    {
        Pass<NC, P::R0, 1, false> p;
        static_assert(P::R0 == 16, "first pass is one radix-16 butterfly per thread");
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int m = t + r * (NC / 16);
            const float2 xv = *reinterpret_cast<const float2*>(x + 2 * m);
            const float2 wv = *reinterpret_cast<const float2*>(win + 2 * m);
            p.v[r] = {xv.x * wv.x, xv.y * wv.y};
        }
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
    {
        Pass<NC, P::R1, P::R0, false> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
    if (P::R2 > 1) {
        Pass<NC, (P::R2 > 1 ? P::R2 : 2), P::R0 * P::R1, false> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
}

// Inverse transform whose last pass applies the synthesis window while storing.
template <int NC>
__device__ __forceinline__ void fft_inverse_windowed(float* sre, float* sim, const cpx* tw, const float* win, int t) {
    using P = Plan<NC>;
    {
        Pass<NC, P::R0, 1, true> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
    {
        Pass<NC, P::R1, P::R0, true> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t);
        if (P::R2 > 1) p.store(sre, sim, t); else p.store_windowed(sre, sim, t, win);
        __syncthreads();
    }
    if (P::R2 > 1) {
        Pass<NC, (P::R2 > 1 ? P::R2 : 2), P::R0 * P::R1, true> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store_windowed(sre, sim, t, win); __syncthreads();
    }
}

template <int NC, bool INV>
__device__ __forceinline__ void fft_inplace(float* sre, float* sim, const cpx* tw, int t) {
    using P = Plan<NC>;
    {
        Pass<NC, P::R0, 1, INV> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
    {
        Pass<NC, P::R1, P::R0, INV> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
    if (P::R2 > 1) {
        Pass<NC, (P::R2 > 1 ? P::R2 : 2), P::R0 * P::R1, INV> p;
        p.load(sre, sim, t); __syncthreads();
        p.twiddle_butterfly(tw, t); p.store(sre, sim, t); __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ STFT
template <int NC, int FR>
__global__ void __launch_bounds__(StftCfg<NC, FR>::THREADS)
stft_kernel(const float* __restrict__ wave, int N, int T, const float2* __restrict__ tw_g, int mode,
            float* __restrict__ out_a, float* __restrict__ out_b,
            uint16_t* __restrict__ op_hi, uint16_t* __restrict__ op_lo,
            long long op_batch_stride, int op_fmt) {
    using Cfg = StftCfg<NC, FR>;
    extern __shared__ float smem[];
    cpx* tw = reinterpret_cast<cpx*>(smem);                 // [NFFT] exp(-2 pi i m / n_fft)
    float* win = smem + 2 * Cfg::NFFT;                      // [NFFT] periodic Hann
    float* span = win + Cfg::NFFT;                          // [SPAN]
    float* bufs = span + Cfg::SPAN;                         // FR x (re[PL], im[PL])

    const int b = blockIdx.y;
    const int t0 = blockIdx.x * Cfg::FR;                    // first frame of this CTA
    const int tid = threadIdx.x;
    const float* w = wave + (size_t)b * N;

    for (int i = tid; i < Cfg::NFFT; i += Cfg::THREADS) {
        float2 v = __ldg(tw_g + i);
        tw[i] = {v.x, v.y};
        win[i] = 0.5f - 0.5f * v.x;                         // Hann(n) = 0.5 - 0.5 cos(2 pi n / n_fft)
    }
    // padded sample p <-> wave index p - n_fft/2, reflected at both ends (librosa center=True)
    const int p0 = t0 * Cfg::HOP - NC;
    for (int i = tid; i < Cfg::SPAN; i += Cfg::THREADS) {
        int n = p0 + i;
        if (n < 0) n = -n;
        if (n >= N) n = 2 * (N - 1) - n;
        span[i] = (n >= 0 && n < N) ? __ldg(w + n) : 0.f;
    }
    __syncthreads();

    const int f = tid / Cfg::TG, t = tid % Cfg::TG;
    float* sre = bufs + f * 2 * Cfg::PL;
    float* sim = sre + Cfg::PL;
    fft_forward_from_span<NC>(sre, sim, tw, span + f * Cfg::HOP, win, t);

    const int frame = t0 + f;
    if (frame >= T) return;
    const size_t row = (size_t)b * T + frame;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int k = 1 + t + i * Cfg::TG;                        // bins 1..NC (DC dropped)
        int ka = k & (NC - 1), kb = (NC - k) & (NC - 1);
        cpx zk = {sre[pad(ka)], sim[pad(ka)]};
        cpx zm = {sre[pad(kb)], -sim[pad(kb)]};             // conj Z[NC-k]
        cpx e = {0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y)};
        cpx d = {0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y)};
        cpx o = {d.y, -d.x};                                // d / i
        cpx x = cadd(e, cmul(tw[k], o));
        if (k == NC) x.y = 0.f;
        float a, ph;
        if (mode == PG_STFT_LOGMAG) {
            a = log1pf(sqrtf(x.x * x.x + x.y * x.y));
            ph = atan2f(x.y, x.x);
        } else {
            a = x.x; ph = x.y;
        }
        size_t o_idx = row * NC + (k - 1);
        if (out_a) out_a[o_idx] = a;
        if (out_b) out_b[o_idx] = ph;
        if (op_hi) {
            uint16_t hi, lo;
            split16(a, op_fmt, hi, lo);
            size_t q = (size_t)b * op_batch_stride + (size_t)frame * NC + (k - 1);
            op_hi[q] = hi;
            if (op_lo) op_lo[q] = lo;
        }
    }
}

// ----------------------------------------------------------------------------------- ISTFT
template <int NC, int FR>
__global__ void __launch_bounds__(StftCfg<NC, FR>::THREADS)
istft_kernel(const float* __restrict__ in_a, const float* __restrict__ in_b, int mode, int T,
             const float2* __restrict__ tw_g, float* __restrict__ wave,
             unsigned* __restrict__ peak_bits, int* __restrict__ nonfinite) {
    using Cfg = StftCfg<NC, FR>;
    constexpr int H = Cfg::FR - 3;                          // output hop-blocks per CTA
    extern __shared__ float smem[];
    cpx* tw = reinterpret_cast<cpx*>(smem);
    float* win = smem + 2 * Cfg::NFFT;
    float* bufs = win + Cfg::NFFT;

    const int b = blockIdx.y;
    const int j0 = blockIdx.x * H;                          // first output hop-block
    const int tid = threadIdx.x;
    const int n_out = (T - 1) * Cfg::HOP;

    for (int i = tid; i < Cfg::NFFT; i += Cfg::THREADS) {
        float2 v = __ldg(tw_g + i);
        tw[i] = {v.x, v.y};
        win[i] = 0.5f - 0.5f * v.x;
    }

    const int f = tid / Cfg::TG, t = tid % Cfg::TG;
    float* sre = bufs + f * 2 * Cfg::PL;
    float* sim = sre + Cfg::PL;
    const int frame = j0 - 1 + f;                           // frames j0-1 .. j0+H+1
    const bool live = frame >= 0 && frame < T;
    const size_t row = ((size_t)b * T + (live ? frame : 0)) * NC;

    // X[k], k = 1..NC from the inputs; X[0] = 0 (the zero DC row of utils.py:38-39)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int k = 1 + t + i * Cfg::TG;
        float xr = 0.f, xi = 0.f;
        if (live) {
            float a = __ldg(in_a + row + k - 1);
            float p = in_b ? __ldg(in_b + row + k - 1) : 0.f;
            if (mode == PG_SPEC_CARTESIAN) { xr = a; xi = p; }
            else {
                float mag = mode == PG_SPEC_POLAR_LOG ? expm1f(a) : a;
                float s, c;
                sincosf(p, &s, &c);
                xr = mag * c; xi = mag * s;
            }
        }
        if (k == NC) xi = 0.f;                              // irfft ignores Im of the Nyquist bin
        sre[pad(k)] = xr;
        sim[pad(k)] = xi;
    }
    if (t == 0) { sre[pad(0)] = 0.f; sim[pad(0)] = 0.f; }
    __syncthreads();

    // Z[k] = E[k] + i O[k],  E = (X[k] + conj X[NC-k]) / 2,  O = (X[k] - conj X[NC-k]) / 2 * W^-k.
    // Thread handles the pair (k, NC-k) so the update is in place.  1/(2*NC) folded in here.
    const float scale = 0.5f / NC;
    for (int k = t; k <= NC / 2; k += Cfg::TG) {
        int km = NC - k;
        cpx xk = {sre[pad(k)], sim[pad(k)]};
        cpx xm = {sre[pad(km)], sim[pad(km)]};
        cpx wk = tw[k];  wk.y = -wk.y;                      // W^-k
        cpx wm = tw[km]; wm.y = -wm.y;
        // for k
        cpx e1 = {xk.x + xm.x, xk.y - xm.y};
        cpx o1 = cmul({xk.x - xm.x, xk.y + xm.y}, wk);
        cpx z1 = {(e1.x - o1.y) * scale, (e1.y + o1.x) * scale};
        // for NC-k
        cpx e2 = {xm.x + xk.x, xm.y - xk.y};
        cpx o2 = cmul({xm.x - xk.x, xm.y + xk.y}, wm);
        cpx z2 = {(e2.x - o2.y) * scale, (e2.y + o2.x) * scale};
        sre[pad(k)] = z1.x; sim[pad(k)] = z1.y;
        if (km < NC && km != k) { sre[pad(km)] = z2.x; sim[pad(km)] = z2.y; }
    }
    __syncthreads();
    // x[2m] = Re z[m], x[2m+1] = Im z[m]; the last pass stores them already multiplied by the
    // synthesis window.
    fft_inverse_windowed<NC>(sre, sim, tw, win, t);

    // gather overlap-add: output hop-block jb (trimmed coordinates) is covered by frames
    // jb-1 .. jb+2; within frame jb-1+q the sample sits at offset (3-q)*hop + i.
    float* wv = wave + (size_t)b * n_out;
    float pk = 0.f;
    bool bad = false;
    for (int s = tid; s < H * Cfg::HOP; s += Cfg::THREADS) {
        int jl = s / Cfg::HOP, i = s % Cfg::HOP;
        int jb = j0 + jl;
        if (jb >= T - 1) break;
        float acc = 0.f, wss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int fr = jb - 1 + q;
            if (fr < 0 || fr >= T) continue;
            int o = (3 - q) * Cfg::HOP + i;
            const float* fb = bufs + (jl + q) * 2 * Cfg::PL + ((o & 1) ? Cfg::PL : 0);
            acc += fb[pad(o >> 1)];
            const float wn = win[o];
            wss += wn * wn;
        }
        float y = wss > 1.17549435e-38f ? acc / wss : acc;
        wv[(size_t)jb * Cfg::HOP + i] = y;
        if (!(fabsf(y) <= 3.402823466e+38f)) bad = true;
        pk = fmaxf(pk, fabsf(y));
    }
    if (peak_bits) {
#pragma unroll
        for (int o = 16; o; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        if ((tid & 31) == 0 && pk > 0.f) atomicMax(peak_bits + b, __float_as_uint(pk));
    }
    if (nonfinite && bad) atomicOr(nonfinite + b, 1);
}

__global__ void peak_normalize_kernel(float* __restrict__ wave, const unsigned* __restrict__ peak_bits, int n) {
    const int b = blockIdx.y;
    float pk = __uint_as_float(peak_bits[b]);
    if (!(pk >= 1.17549435e-38f)) return;                   // librosa.util.normalize: tiny -> unchanged
    float inv = 1.0f / pk;
    float* w = wave + (size_t)b * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = w[i] * inv;
}

// frames per CTA: 256 threads for the forward transform; the inverse recomputes a 3-frame halo
// per CTA, so it takes more frames per CTA to amortise it.
template <int NC> struct FramesPerCta {
    static constexpr int STFT = NC >= 1024 ? 8 : 256 / (NC / 16);
    static constexpr int ISTFT = NC >= 1024 ? 8 : 16;
};

template <int NC>
static int launch_stft(const float* wave, int B, int N, int T, const float* tw, int mode, float* a, float* bq,
                       uint16_t* hi, uint16_t* lo, long long bs, int fmt, cudaStream_t st) {
    constexpr int FR = FramesPerCta<NC>::STFT;
    using Cfg = StftCfg<NC, FR>;
    auto k = stft_kernel<NC, FR>;
    size_t sm = Cfg::smem_stft();
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    dim3 grid((T + Cfg::FR - 1) / Cfg::FR, B);
    k<<<grid, Cfg::THREADS, sm, st>>>(wave, N, T, reinterpret_cast<const float2*>(tw), mode, a, bq, hi, lo, bs, fmt);
    return check_launch("stft_kernel");
}

template <int NC>
static int launch_istft(const float* a, const float* bq, int mode, int B, int T, const float* tw, float* wave,
                        float* peak, int* nonfinite, cudaStream_t st) {
    constexpr int FR = FramesPerCta<NC>::ISTFT;
    using Cfg = StftCfg<NC, FR>;
    constexpr int H = Cfg::FR - 3;
    auto k = istft_kernel<NC, FR>;
    size_t sm = Cfg::smem_istft();
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    dim3 grid((T - 1 + H - 1) / H, B);
    k<<<grid, Cfg::THREADS, sm, st>>>(a, bq, mode, T, reinterpret_cast<const float2*>(tw), wave,
                                     reinterpret_cast<unsigned*>(peak), nonfinite);
    return check_launch("istft_kernel");
}

}  // namespace pg

extern "C" int pg_stft_num_frames(int n_samples, int hop) { return hop > 0 ? 1 + n_samples / hop : 0; }

extern "C" int pg_stft(const float* wave, int B, int N, int n_fft, int hop, const float* twiddle, int mode,
                       float* out_a, float* out_b, uint16_t* op_hi, uint16_t* op_lo,
                       int64_t op_batch_stride, int op_fmt, pg_stream stream) {
    PG_REQUIRE(wave && twiddle && B > 0 && N > 0, "pg_stft: null pointer or empty batch");
    PG_REQUIRE(hop * 4 == n_fft, "pg_stft: hop must be n_fft/4 (got n_fft=%d hop=%d)", n_fft, hop);
    PG_REQUIRE(N > n_fft / 2, "pg_stft: reflect padding needs more than n_fft/2 samples (N=%d)", N);
    PG_REQUIRE(mode == PG_STFT_LOGMAG || mode == PG_STFT_REIM, "pg_stft: bad mode %d", mode);
    PG_REQUIRE(B <= 65535, "pg_stft: batch too large for one launch");
    const int T = 1 + N / hop;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (n_fft) {
        case 256:  return pg::launch_stft<128>(wave, B, N, T, twiddle, mode, out_a, out_b, op_hi, op_lo, op_batch_stride, op_fmt, st);
        case 512:  return pg::launch_stft<256>(wave, B, N, T, twiddle, mode, out_a, out_b, op_hi, op_lo, op_batch_stride, op_fmt, st);
        case 1024: return pg::launch_stft<512>(wave, B, N, T, twiddle, mode, out_a, out_b, op_hi, op_lo, op_batch_stride, op_fmt, st);
        case 2048: return pg::launch_stft<1024>(wave, B, N, T, twiddle, mode, out_a, out_b, op_hi, op_lo, op_batch_stride, op_fmt, st);
    }
    pg::set_error("pg_stft: n_fft must be 256, 512, 1024 or 2048 (got %d)", n_fft);
    return PG_ERR_UNSUPPORTED;
}

extern "C" int pg_istft(const float* in_a, const float* in_b, int mode, int B, int T, int n_fft, int hop,
                        const float* twiddle, float* wave, float* peak, int* nonfinite, pg_stream stream) {
    PG_REQUIRE(in_a && twiddle && wave && B > 0 && T > 1, "pg_istft: null pointer or empty input");
    PG_REQUIRE(hop * 4 == n_fft, "pg_istft: hop must be n_fft/4 (got n_fft=%d hop=%d)", n_fft, hop);
    PG_REQUIRE(mode >= PG_SPEC_POLAR_LOG && mode <= PG_SPEC_POLAR_MAG, "pg_istft: bad mode %d", mode);
    PG_REQUIRE(in_b || mode == PG_SPEC_POLAR_MAG, "pg_istft: second plane missing");
    PG_REQUIRE(B <= 65535, "pg_istft: batch too large for one launch");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (peak) cudaMemsetAsync(peak, 0, sizeof(float) * B, st);
    if (nonfinite) cudaMemsetAsync(nonfinite, 0, sizeof(int) * B, st);
    switch (n_fft) {
        case 256:  return pg::launch_istft<128>(in_a, in_b, mode, B, T, twiddle, wave, peak, nonfinite, st);
        case 512:  return pg::launch_istft<256>(in_a, in_b, mode, B, T, twiddle, wave, peak, nonfinite, st);
        case 1024: return pg::launch_istft<512>(in_a, in_b, mode, B, T, twiddle, wave, peak, nonfinite, st);
        case 2048: return pg::launch_istft<1024>(in_a, in_b, mode, B, T, twiddle, wave, peak, nonfinite, st);
    }
    pg::set_error("pg_istft: n_fft must be 256, 512, 1024 or 2048 (got %d)", n_fft);
    return PG_ERR_UNSUPPORTED;
}

extern "C" int pg_peak_normalize(float* wave, const float* peak, int B, int N, pg_stream stream) {
    PG_REQUIRE(wave && peak && B > 0 && N > 0 && B <= 65535, "pg_peak_normalize: bad arguments");
    dim3 grid((N + 1023) / 1024 < 64 ? (N + 1023) / 1024 : 64, B);
    pg::peak_normalize_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        wave, reinterpret_cast<const unsigned*>(peak), N);
    return pg::check_launch("peak_normalize_kernel");
}
