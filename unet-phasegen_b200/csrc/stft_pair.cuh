// Frame-PAIR transforms for n_fft 1024 (NC = 512): one warp owns TWO frames and every lane carries the same
// point of both, as a structure of arrays -- f2 x = (Re frame A, Re frame B), f2 y = (Im A, Im B).  Every
// butterfly, twiddle product and Hermitian step of stft_core.cuh then becomes ONE packed fp32x2 instruction
// for both frames (sm_100: add/sub/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2), the shared-memory exchanges move
// 16 bytes per lane (LDS.128 / STS.128: both frames in one access), and the index arithmetic and twiddle loads
// are paid once per pair.  The scalar kernels spend 781 (STFT) / 1,015 (ISTFT) of their 1,511 / 1,956 warp
// instructions per frame on FADD/FMUL/FFMA (profiles/r02_stft_istft_instruction_mix.txt); this form halves
// those, the LDS/STS count and most of the address arithmetic.
// No swaps between the halves are ever needed: a pair's partner is always the same lane position of the other
// frame, so multiplication by -+i is a renaming of (x, y) plus a subtraction instead of an addition.
// Everything is PG_HD: csrc/fft_selftest2.cpp runs the same thread programs on host threads (the packed
// operations fall back to two scalar ones there) against a naive DFT.
#pragma once
#include <cstring>
#include "stft_core.cuh"

namespace pgfft {

struct alignas(8) f2 { unsigned long long v; };

PG_HD f2 f2_make(float a, float b) {
    f2 r;
#ifdef __CUDA_ARCH__
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
#else
    float t[2] = {a, b}; std::memcpy(&r.v, t, 8);
#endif
    return r;
}
PG_HD void f2_split(f2 p, float& a, float& b) {
#ifdef __CUDA_ARCH__
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
#else
    float t[2]; std::memcpy(t, &p.v, 8); a = t[0]; b = t[1];
#endif
}
PG_HD f2 f2_bcast(float a) { return f2_make(a, a); }
PG_HD f2 operator+(f2 p, f2 q) {
#ifdef __CUDA_ARCH__
    f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(p.v), "l"(q.v)); return r;
#else
    float a, b, c, d; f2_split(p, a, b); f2_split(q, c, d); return f2_make(a + c, b + d);
#endif
}
PG_HD f2 operator-(f2 p, f2 q) {
#ifdef __CUDA_ARCH__
    f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(p.v), "l"(q.v)); return r;
#else
    float a, b, c, d; f2_split(p, a, b); f2_split(q, c, d); return f2_make(a - c, b - d);
#endif
}
PG_HD f2 operator*(f2 p, f2 q) {
#ifdef __CUDA_ARCH__
    f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(p.v), "l"(q.v)); return r;
#else
    float a, b, c, d; f2_split(p, a, b); f2_split(q, c, d); return f2_make(a * c, b * d);
#endif
}
// p * q + r
PG_HD f2 fma2(f2 p, f2 q, f2 r) {
#ifdef __CUDA_ARCH__
    f2 o; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o.v) : "l"(p.v), "l"(q.v), "l"(r.v)); return o;
#else
    float a, b, c, d, e, f; f2_split(p, a, b); f2_split(q, c, d); f2_split(r, e, f);
    return f2_make(a * c + e, b * d + f);
#endif
}
PG_HD f2 f2_neg(f2 p) { return f2_make(0.f, 0.f) - p; }

// two complex numbers, structure of arrays.  As a TWIDDLE the two halves of x (and of y) hold the same value.
struct alignas(16) cpx2 { f2 x, y; };

PG_HD cpx2 cadd(cpx2 a, cpx2 b) { return {a.x + b.x, a.y + b.y}; }
PG_HD cpx2 csub(cpx2 a, cpx2 b) { return {a.x - b.x, a.y - b.y}; }
PG_HD cpx2 cmul(cpx2 v, cpx2 w) { return {v.x * w.x - v.y * w.y, fma2(v.x, w.y, v.y * w.x)}; }    // v * w
PG_HD cpx2 cmulc(cpx2 v, cpx2 w) { return {fma2(v.x, w.x, v.y * w.y), v.y * w.x - v.x * w.y}; }   // v * conj(w)
PG_HD cpx2 tw_bcast(cpx w) { return {f2_bcast(w.x), f2_bcast(w.y)}; }
PG_HD cpx2 cpx2_of(cpx a, cpx b) { return {f2_make(a.x, b.x), f2_make(a.y, b.y)}; }
PG_HD void cpx2_split(cpx2 v, cpx& a, cpx& b) { f2_split(v.x, a.x, b.x); f2_split(v.y, a.y, b.y); }

// natural-order in, natural-order out; forward multiplies the odd difference by -i, inverse by +i
template <bool INV> PG_HD void dft4p(cpx2& a0, cpx2& a1, cpx2& a2, cpx2& a3) {
    const cpx2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(s02, s13); a2 = csub(s02, s13);
    if (!INV) { a1 = {d02.x + d.y, d02.y - d.x}; a3 = {d02.x - d.y, d02.y + d.x}; }
    else      { a1 = {d02.x - d.y, d02.y + d.x}; a3 = {d02.x + d.y, d02.y - d.x}; }
}

// v * W16^M (forward: exp(-2 pi i M / 16); inverse: the conjugate), M a compile-time constant in {0,1,2,3,4,6,9}
template <bool INV, int M> PG_HD cpx2 mul_w16(cpx2 v) {
    if (M == 0) return v;
    if (M == 4) return INV ? cpx2{f2_neg(v.y), v.x} : cpx2{v.y, f2_neg(v.x)};
    if (M == 2) {                                    // sqh (1 -+ i)
        const f2 h = f2_bcast(PG_SQH);
        return INV ? cpx2{(v.x - v.y) * h, (v.x + v.y) * h} : cpx2{(v.x + v.y) * h, (v.y - v.x) * h};
    }
    if (M == 6) {                                    // -sqh (1 +- i)
        const f2 h = f2_bcast(-PG_SQH);
        return INV ? cpx2{(v.x + v.y) * h, (v.y - v.x) * h} : cpx2{(v.x - v.y) * h, (v.x + v.y) * h};
    }
    // general: w = c + i s
    const float c = M == 1 ? PG_C8 : M == 3 ? PG_S8 : -PG_C8;                       // M = 9: cos(9 pi / 8) = -C8
    const float sf = M == 1 ? -PG_S8 : M == 3 ? -PG_C8 : PG_S8;                     // forward sine: -sin(2 pi M / 16)
    const float s = INV ? -sf : sf;
    const f2 cc = f2_bcast(c), ss = f2_bcast(s), ns = f2_bcast(-s);
    return {fma2(v.y, ns, v.x * cc), fma2(v.x, ss, v.y * cc)};
}

template <bool INV> PG_HD void dft16p(cpx2* v) {
    // 16 = 4 (n1) x 4 (n2):  n = 4*n1 + n2, k = k1 + 4*k2   (same decomposition as dft16 of fft_core.cuh)
    cpx2 a[4][4];
#define PG_DFT16P_COL(n2)                                                                    \
    {                                                                                        \
        cpx2 t0 = v[n2], t1 = v[n2 + 4], t2 = v[n2 + 8], t3 = v[n2 + 12];                    \
        dft4p<INV>(t0, t1, t2, t3);                                                          \
        a[0][n2] = t0;                                                                       \
        a[1][n2] = mul_w16<INV, n2>(t1);                                                     \
        a[2][n2] = mul_w16<INV, 2 * n2>(t2);                                                 \
        a[3][n2] = mul_w16<INV, 3 * n2>(t3);                                                 \
    }
    PG_DFT16P_COL(0) PG_DFT16P_COL(1) PG_DFT16P_COL(2) PG_DFT16P_COL(3)
#undef PG_DFT16P_COL
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        dft4p<INV>(a[k1][0], a[k1][1], a[k1][2], a[k1][3]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = a[k1][k2];
    }
}

// One radix-16 Stockham pass of the 512-point transform pair (one butterfly per lane).  The pair buffer has the
// element order and padding of stft_core.cuh's frame buffer with 16-byte elements: every access below is a
// conflict-free 128-bit access per quarter warp (the stride-2 stores of the inverse first pass are 2-way).
template <int NS, bool INV>
struct PassP {
    static constexpr int NC = 512;
    cpx2 v[16];
    PG_HD void load(const cpx2* s, int t) {
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = s[pad2(t + r * (NC / 16))];
    }
    // tab: this pass's compact table (15 * NS broadcast twiddles, entry (r-1)*NS + k); ignored when NS == 1
    PG_HD void twiddle_butterfly(const cpx2* tab, int t) {
        if (NS > 1) {
            const int k = t % NS;
#pragma unroll
            for (int r = 1; r < 16; ++r) {
                const cpx2 w = tab[(r - 1) * NS + k];
                v[r] = INV ? cmulc(v[r], w) : cmul(v[r], w);
            }
        }
        dft16p<INV>(v);
    }
    PG_HD void store(cpx2* s, int t) const {
        const int base = (t / NS) * (NS * 16) + (t % NS);
#pragma unroll
        for (int r = 0; r < 16; ++r) s[pad2(base + r * NS)] = v[r];
    }
    // last inverse pass: window the samples and write the two frames to their own (scalar-layout) buffers:
    // sample 2m = Re z[m] * win[m].x, sample 2m+1 = Im z[m] * win[m].y
    PG_HD void store_windowed_split(cpx* sa, cpx* sb, int t, const cpx* win) const {
        const int base = (t / NS) * (NS * 16) + (t % NS);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int m = base + r * NS;
            const cpx w = win[m];
            cpx a, b;
            cpx2_split(v[r], a, b);
            sa[pad2(m)] = {a.x * w.x, a.y * w.y};
            sb[pad2(m)] = {b.x * w.x, b.y * w.y};
        }
    }
};

// Compact broadcast-twiddle tables of the pair programs: same entries and offsets as Radix<512, INV>.
template <bool INV> PG_HD cpx2 pair_table_entry(const cpx* tw, int e) { return tw_bcast(table_entry<512, INV>(tw, e)); }

// za = Z[a], zb = Z[NC - a], u = exp(-2 pi i a / (2 NC)) -> 2 X[a], 2 X[NC - a]   (herm_post of stft_core.cuh)
PG_HD void herm_post_p(cpx2 za, cpx2 zb, cpx2 u, cpx2& xa, cpx2& xb) {
    const cpx2 s = {za.x + zb.x, za.y - zb.y};
    const cpx2 d = {za.x - zb.x, za.y + zb.y};
    // t = u * (d.y - i d.x)
    const cpx2 t = {fma2(u.x, d.y, u.y * d.x), u.y * d.y - u.x * d.x};
    xa = {s.x + t.x, s.y + t.y};
    xb = {s.x - t.x, t.y - s.y};
}
// X[a], X[NC - a] -> Z[a], Z[NC - a]   (herm_pre of stft_core.cuh without its scale, which the window table carries)
PG_HD void herm_pre_p(cpx2 xa, cpx2 xb, cpx2 u, cpx2& za, cpx2& zb) {
    const cpx2 e = {xa.x + xb.x, xa.y - xb.y};
    const cpx2 d = {xa.x - xb.x, xa.y + xb.y};
    const cpx2 o = cmulc(d, u);
    za = {e.x - o.y, e.y + o.x};
    zb = {e.x + o.y, o.x - e.y};
}

// ---------------------------------------------------------------------------------------- forward
// v_in[r] = windowed z[t + 32 r] of both frames
PG_HD void fwd_pair_phase0(cpx2* s, int t, const cpx2* v_in) {
    PassP<1, false> p;
#pragma unroll
    for (int r = 0; r < 16; ++r) p.v[r] = v_in[r];
    p.twiddle_butterfly(nullptr, t);
    p.store(s, t);
}
template <class Sync>
PG_HD void fwd_pair_phase1(cpx2* s, int t, const cpx2* tabs, Sync&& sync) {
    using R = Radix<512, false>;
    PassP<16, false> p;
    p.load(s, t);
    sync();
    p.twiddle_butterfly(tabs + R::OFF1, t);
    p.store(s, t);
}
// last radix-2 pass + Hermitian post-processing (fwd_fused_last_512).  emit(bin, X) receives bins 1 .. 512 of both frames.
template <class Emit>
PG_HD void fwd_pair_fused_last(const cpx2* s, int t, const cpx2* tabs, Emit&& emit) {
    using R = Radix<512, false>;
    const cpx2* w2 = tabs + R::OFF2;
    const cpx2* up = tabs + R::OFFP;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = t + 32 * i;
        const int kq = k == 0 ? 128 : 256 - k;
        const cpx2 p0 = s[pad2(k)], p1 = s[pad2(k + 256)], q0 = s[pad2(kq)], q1 = s[pad2(kq + 256)];
        const cpx2 tp = cmul(p1, w2[k]), tq = cmul(q1, w2[kq]);
        const cpx2 zk = cadd(p0, tp), zk2 = csub(p0, tp);
        const cpx2 zq = cadd(q0, tq), zq2 = csub(q0, tq);
        const bool z0 = (i == 0) && (k == 0);
        cpx2 a1 = zq2, b1 = zk2;
        if (i == 0) { a1 = z0 ? zk : zq2; b1 = z0 ? zq2 : zk2; }
        cpx2 xa, xb, ya, yb;
        herm_post_p(zk, a1, up[k], xa, xb);
        herm_post_p(zq, b1, up[kq], ya, yb);
        if (i == 0) {
            cpx2 ma, mb;
            herm_post_p(zk2, zk2, up[256], ma, mb);
            emit(z0 ? 256 : k, z0 ? ma : xa);
            emit(512 - k, xb);
            emit(kq, ya);
            emit(z0 ? 384 : k + 256, yb);
        } else {
            emit(k, xa); emit(512 - k, xb);
            emit(kq, ya); emit(k + 256, yb);
        }
    }
}

// ---------------------------------------------------------------------------------------- inverse
// x[j] = X[inv_bin<512>(t, j)] of both frames.  Pre-processing + first radix-2 pass (inv_fused_first_512).
PG_HD void inv_pair_fused_first(cpx2* s, int t, const cpx2* tabs, const cpx2* x) {
    using R = Radix<512, true>;
    const cpx2* up = tabs + R::OFFP;
    const cpx2 zero = {f2_make(0.f, 0.f), f2_make(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = t + 32 * i;
        const int kq = k == 0 ? 128 : 256 - k;
        const cpx2 x0 = x[4 * i], x1 = x[4 * i + 1], x2 = x[4 * i + 2], x3 = x[4 * i + 3];
        const bool z0 = (i == 0) && (k == 0);
        cpx2 a0 = x0, a1 = x1;
        if (i == 0) {
            a0 = z0 ? zero : x0;                                   // X[0] = 0 (utils.py:38-39)
            a1 = z0 ? cpx2{x0.x, zero.y} : x1;                     // X[512], imaginary part ignored like numpy's irfft
        }
        cpx2 pa, pb, qa, qb;
        herm_pre_p(a0, a1, up[k], pa, pb);
        herm_pre_p(x2, x3, up[kq], qa, qb);
        cpx2 zk = pa, zq = qa, zq2 = pb, zk2 = qb;
        if (i == 0) {
            cpx2 ma, mb;
            herm_pre_p(x1, x1, up[256], ma, mb);                   // X[256] (self-mirrored) -> Z[256]
            zq2 = z0 ? qb : pb;
            zk2 = z0 ? ma : qb;
        }
        s[pad2(2 * k)] = cadd(zk, zk2);
        s[pad2(2 * k + 1)] = csub(zk, zk2);
        s[pad2(2 * kq)] = cadd(zq, zq2);
        s[pad2(2 * kq + 1)] = csub(zq, zq2);
    }
}
// The two radix-16 passes; the last one windows and writes frame A to sa, frame B to sb (scalar frame buffers, which
// may alias the pair buffer s: every lane has loaded its points before any lane stores).
template <class Sync>
PG_HD void inv_pair_passes(cpx2* s, cpx* sa, cpx* sb, int t, const cpx2* tabs, const cpx* win, Sync&& sync) {
    using R = Radix<512, true>;
    {
        PassP<2, true> p;
        p.load(s, t);
        sync();
        p.twiddle_butterfly(tabs + R::OFF1, t);
        p.store(s, t);
        sync();
    }
    {
        PassP<32, true> p;
        p.load(s, t);
        sync();
        p.twiddle_butterfly(tabs + R::OFF2, t);
        p.store_windowed_split(sa, sb, t, win);
        sync();
    }
}

}  // namespace pgfft
