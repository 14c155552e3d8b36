// Shared-memory Stockham FFT building blocks used by the STFT and ISTFT kernels.
//
// A complex FFT of length NC (= n_fft/2, the "real FFT through a half-length complex FFT"
// trick) is owned by a group of TG = NC/16 threads, 16 complex points per thread.  A pass of
// radix R has every thread do 16/R radix-R butterflies on registers; passes exchange data
// through a padded split re/im shared-memory buffer, in place (load -> barrier -> store ->
// barrier).  Everything is PG_HD so the index algebra and the codelets can be exercised on
// the host (csrc/fft_selftest.cpp) where there is no GPU.
#pragma once
#ifndef PG_HD
#ifdef __CUDACC__
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD inline
#endif
#endif

namespace pgfft {

struct cpx { float x, y; };

PG_HD cpx cmul(cpx a, cpx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
PG_HD cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
PG_HD cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
// multiply by -i (forward) or +i (inverse)
template <bool INV> PG_HD cpx mul_mi(cpx a) { return INV ? cpx{-a.y, a.x} : cpx{a.y, -a.x}; }

// 32-bit bank skew: one pad word every 32 keeps stride-R writes of a Stockham pass and the
// mirrored reads of the real-FFT post-processing conflict-free (or 2-way at worst).
PG_HD int pad(int i) { return i + (i >> 5); }
PG_HD constexpr int padded_len(int n) { return n + (n >> 5) + 1; }

#define PG_SQH 0.70710678118654752440f
#define PG_C8 0.92387953251128675613f   // cos(pi/8)
#define PG_S8 0.38268343236508977173f   // sin(pi/8)

template <bool INV> PG_HD void dft2(cpx& a, cpx& b) { cpx t = a; a = cadd(t, b); b = csub(t, b); }

// natural-order in, natural-order out
template <bool INV> PG_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
    cpx s02 = cadd(a0, a2), d02 = csub(a0, a2);
    cpx s13 = cadd(a1, a3), d13 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(s02, s13); a2 = csub(s02, s13);
    a1 = cadd(d02, d13); a3 = csub(d02, d13);
}

// W16^m = exp(-+ 2 pi i m / 16)
template <bool INV> PG_HD cpx w16(int m) {
    const float c[8] = {1.f, PG_C8, PG_SQH, PG_S8, 0.f, -PG_S8, -PG_SQH, -PG_C8};
    const float s[8] = {0.f, PG_S8, PG_SQH, PG_C8, 1.f, PG_C8, PG_SQH, PG_S8};
    m &= 15;
    float cc = m < 8 ? c[m] : -c[m - 8];
    float ss = m < 8 ? s[m] : -s[m - 8];
    return {cc, INV ? ss : -ss};
}

template <bool INV> PG_HD void dft8(cpx* v) {
    // 8 = 2 (n1) x 4 (n2):  n = 4*n1 + n2, k = k1 + 2*k2
    cpx a[2][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        a[0][n2] = cadd(v[n2], v[n2 + 4]);
        a[1][n2] = cmul(csub(v[n2], v[n2 + 4]), w16<INV>(2 * n2));
    }
#pragma unroll
    for (int k1 = 0; k1 < 2; ++k1) {
        dft4<INV>(a[k1][0], a[k1][1], a[k1][2], a[k1][3]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 2 * k2] = a[k1][k2];
    }
}

template <bool INV> PG_HD void dft16(cpx* v) {
    // 16 = 4 (n1) x 4 (n2):  n = 4*n1 + n2, k = k1 + 4*k2
    cpx a[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        cpx t0 = v[n2], t1 = v[n2 + 4], t2 = v[n2 + 8], t3 = v[n2 + 12];
        dft4<INV>(t0, t1, t2, t3);
        a[0][n2] = t0;
        a[1][n2] = n2 ? cmul(t1, w16<INV>(n2)) : t1;
        a[2][n2] = n2 ? cmul(t2, w16<INV>(2 * n2)) : t2;
        a[3][n2] = n2 ? cmul(t3, w16<INV>(3 * n2)) : t3;
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        dft4<INV>(a[k1][0], a[k1][1], a[k1][2], a[k1][3]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = a[k1][k2];
    }
}

template <int R, bool INV> PG_HD void dftR(cpx* v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else dft16<INV>(v);
}

// One Stockham pass (radix R, NS = product of the radices of earlier passes) for thread t of
// the TG = NC/16 threads that own this transform.  tw[m] = exp(-2 pi i m / (2*NC)), m < 2*NC.
// load(): shared -> registers; store(): twiddle, butterfly, registers -> shared.
template <int NC, int R, int NS, bool INV>
struct Pass {
    static constexpr int TG = NC / 16;
    static constexpr int NB = 16 / R;
    cpx v[16];

    PG_HD void load(const float* sre, const float* sim, int t) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            int j = t + i * TG;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int p = pad(j + r * (NC / R));
                v[i * R + r] = {sre[p], sim[p]};
            }
        }
    }
    PG_HD void twiddle_butterfly(const cpx* tw, int t) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            int j = t + i * TG;
            if (NS > 1) {
                int k = j % NS;
                int step = 2 * k * (NC / (NS * R));   // index into the 2*NC-point table
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    cpx w = tw[r * step];
                    if (INV) w.y = -w.y;
                    v[i * R + r] = cmul(v[i * R + r], w);
                }
            }
            dftR<R, INV>(&v[i * R]);
        }
    }
    // store with a per-sample gain: real sample 2m gets win[2m], 2m+1 gets win[2m+1] (synthesis window
    // of the inverse real transform folded into the last pass)
    PG_HD void store_windowed(float* sre, float* sim, int t, const float* win) const {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            int j = t + i * TG;
            int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int m = base + r * NS;
                int p = pad(m);
                sre[p] = v[i * R + r].x * win[2 * m];
                sim[p] = v[i * R + r].y * win[2 * m + 1];
            }
        }
    }
    PG_HD void store(float* sre, float* sim, int t) const {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            int j = t + i * TG;
            int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int p = pad(base + r * NS);
                sre[p] = v[i * R + r].x;
                sim[p] = v[i * R + r].y;
            }
        }
    }
};

// Radix plan per transform length: NC = R0 * R1 * R2 (R2 may be 1).
template <int NC> struct Plan;
template <> struct Plan<128>  { static constexpr int R0 = 16, R1 = 8,  R2 = 1; };
template <> struct Plan<256>  { static constexpr int R0 = 16, R1 = 16, R2 = 1; };
template <> struct Plan<512>  { static constexpr int R0 = 16, R1 = 16, R2 = 2; };
template <> struct Plan<1024> { static constexpr int R0 = 16, R1 = 16, R2 = 4; };

}  // namespace pgfft
