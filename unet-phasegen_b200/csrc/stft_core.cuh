// Per-thread building blocks of the STFT / ISTFT frame transforms (second generation).
//
// One frame = an n_fft-point real transform done as an NC = n_fft/2 point complex Stockham FFT owned by
// TG = NC/16 threads (16 complex points per thread; TG <= 32 means the owners are lanes of ONE warp, so
// the passes are separated by __syncwarp only).  Differences from fft_core.cuh's first version:
//   * interleaved float2 shared buffers (LDS.64/STS.64: half the shared-memory instructions), one pad
//     element per 16 -- every pass below is bank-conflict-free for 64-bit accesses;
//   * compact per-pass twiddle tables laid out [r][k] so lanes read consecutive entries (the 2*NC table
//     indexed by r*k*stride gave 4-way conflicts);
//   * NC = 512 (n_fft 1024, the BASELINE shape): the last radix-2 pass of the forward transform is fused
//     with the Hermitian post-processing, and the first radix-2 pass of the inverse transform with the
//     Hermitian pre-processing: a thread that owns {k, k+256, 256-k, 512-k} owns both mirror pairs, so one
//     shared-memory round trip disappears in each direction.
// Everything is PG_HD so csrc/fft_selftest.cpp runs the exact thread programs on the host.
#pragma once
#include "fft_core.cuh"

namespace pgfft {

PG_HD int pad2(int i) { return i + (i >> 4); }
PG_HD constexpr int padded_len2(int n) { return n + (n >> 4) + 2; }

// conj(a) * b style helpers
PG_HD cpx conj(cpx a) { return {a.x, -a.y}; }

// Radix order.  Forward: R0, R1, R2 (R2 = 1: two passes).  The inverse of NC = 512 runs 2, 16, 16 so that
// its radix-2 pass comes first (fused with the pre-processing); other sizes run the forward order.
template <int NC, bool INV> struct Radix {
    using P = Plan<NC>;
    static constexpr bool SWAP = INV && NC == 512;
    static constexpr int R0 = SWAP ? 2 : P::R0;
    static constexpr int R1 = SWAP ? 16 : P::R1;
    static constexpr int R2 = SWAP ? 16 : P::R2;
    static constexpr int NS1 = R0, NS2 = R0 * R1;
    // compact twiddle tables: pass p (radix R, NS): entry (r-1)*NS + k = exp(-2 pi i r k / (NS R))
    static constexpr int N1 = (R1 - 1) * NS1;
    static constexpr int N2 = R2 > 1 ? (R2 - 1) * NS2 : 0;
    static constexpr int NPOST = NC / 2 + 1;                  // exp(-2 pi i a / (2 NC)), a = 0 .. NC/2
    static constexpr int OFF1 = 0, OFF2 = N1, OFFP = N1 + N2, TOTAL = N1 + N2 + NPOST;
};

// Fill the compact tables from the 2*NC-entry table tw[m] = exp(-2 pi i m / (2 NC)); e in [0, TOTAL).
template <int NC, bool INV>
PG_HD cpx table_entry(const cpx* tw, int e) {
    using R = Radix<NC, INV>;
    if (e < R::OFF2) {
        const int r = e / R::NS1 + 1, k = e % R::NS1;
        return tw[r * k * (2 * NC / (R::NS1 * R::R1))];
    }
    if (e < R::OFFP) {
        const int q = e - R::OFF2;
        const int r = q / R::NS2 + 1, k = q % R::NS2;
        return tw[r * k * (2 * NC / (R::NS2 * (R::R2 > 1 ? R::R2 : 1)))];
    }
    return tw[e - R::OFFP];
}

template <int NC, int R, int NS, bool INV>
struct Pass2 {
    static constexpr int TG = NC / 16;
    static constexpr int NB = 16 / R;
    cpx v[16];

    PG_HD void load(const cpx* s, int t) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = t + i * TG;
#pragma unroll
            for (int r = 0; r < R; ++r) v[i * R + r] = s[pad2(j + r * (NC / R))];
        }
    }
    // tab: this pass's compact table ((R-1)*NS entries); ignored when NS == 1
    PG_HD void twiddle_butterfly(const cpx* tab, int t) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = t + i * TG;
            if (NS > 1) {
                const int k = j % NS;
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    cpx w = tab[(r - 1) * NS + k];
                    if (INV) w.y = -w.y;
                    v[i * R + r] = cmul(v[i * R + r], w);
                }
            }
            dftR<R, INV>(&v[i * R]);
        }
    }
    PG_HD void store(cpx* s, int t) const {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = t + i * TG;
            const int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
            for (int r = 0; r < R; ++r) s[pad2(base + r * NS)] = v[i * R + r];
        }
    }
    // inverse real transform: sample 2m = Re z[m] * win[m].x, sample 2m+1 = Im z[m] * win[m].y
    PG_HD void store_windowed(cpx* s, int t, const cpx* win) const {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = t + i * TG;
            const int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int m = base + r * NS;
                const cpx w = win[m];
                s[pad2(m)] = {v[i * R + r].x * w.x, v[i * R + r].y * w.y};
            }
        }
    }
};

// One mirror pair of the real-FFT post-processing.  za = Z[a], zb = Z[NC - a], u = exp(-2 pi i a / (2 NC)).
// Returns 2 X[a] in xa and 2 X[NC - a] in xb (the factor 2 is folded into the analysis window).
PG_HD void herm_post(cpx za, cpx zb, cpx u, cpx& xa, cpx& xb) {
    const cpx s = {za.x + zb.x, za.y - zb.y};                 // Za + conj Zb
    const cpx d = {za.x - zb.x, za.y + zb.y};                 // Za - conj Zb
    const cpx t = cmul(u, cpx{d.y, -d.x});                    // u * (D / i)
    xa = {s.x + t.x, s.y + t.y};
    xb = {s.x - t.x, -(s.y - t.y)};                           // conj(S - t)
}
// Inverse: X[a], X[NC - a] -> Z[a], Z[NC - a] (times `scale`).  u as above.
PG_HD void herm_pre(cpx xa, cpx xb, cpx u, float scale, cpx& za, cpx& zb) {
    const cpx e = {xa.x + xb.x, xa.y - xb.y};                 // Xa + conj Xb          (= 2 E)
    const cpx d = {xa.x - xb.x, xa.y + xb.y};                 // Xa - conj Xb          (= 2 u O)
    const cpx o = cmul(conj(u), d);                           // 2 O
    za = {(e.x - o.y) * scale, (e.y + o.x) * scale};          // E + i O
    zb = {(e.x + o.y) * scale, (-e.y + o.x) * scale};         // conj(E) + i conj(O)
}

// ---------------------------------------------------------------------------------------- forward
// Thread program of one frame.  `sync()` separates the threads' passes (__syncwarp, a named barrier, or a
// std::barrier in the host self-test).  A phase expects its input to be visible at entry and does not
// synchronise after its last store: the caller does, between phases.
//   phase 0: v[r] = windowed z[t + r*TG] supplied by the caller; radix-R0 butterflies; store.
template <int NC>
PG_HD void fwd_phase0(cpx* s, int t, const cpx* v_in) {
    using R = Radix<NC, false>;
    static_assert(R::R0 == 16, "first forward pass is one radix-16 butterfly per thread");
    Pass2<NC, 16, 1, false> p;
#pragma unroll
    for (int r = 0; r < 16; ++r) p.v[r] = v_in[r];
    p.twiddle_butterfly(nullptr, t);
    p.store(s, t);
}
template <int NC, class Sync>
PG_HD void fwd_phase1(cpx* s, int t, const cpx* tabs, Sync&& sync) {
    using R = Radix<NC, false>;
    Pass2<NC, R::R1, R::NS1, false> p;
    p.load(s, t);
    sync();
    p.twiddle_butterfly(tabs + R::OFF1, t);
    p.store(s, t);
}
template <int NC, class Sync>
PG_HD void fwd_phase2_unfused(cpx* s, int t, const cpx* tabs, Sync&& sync) {
    using R = Radix<NC, false>;
    if (R::R2 > 1) {
        Pass2<NC, (R::R2 > 1 ? R::R2 : 2), R::NS2, false> p;
        p.load(s, t);
        sync();
        p.twiddle_butterfly(tabs + R::OFF2, t);
        p.store(s, t);
    }
}
// Generic post-processing: thread t emits bins of the pairs a = t + i*TG (a < NC/2) and, thread 0, a = NC/2.
// emit(bin, X) receives bins 1 .. NC (the DC bin is dropped: preproc_mdb.py:93).
template <int NC, class Emit>
PG_HD void fwd_post_generic(const cpx* s, int t, const cpx* tabs, Emit&& emit) {
    using R = Radix<NC, false>;
    constexpr int TG = NC / 16;
    const cpx* up = tabs + R::OFFP;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int a = t + i * TG;
        const cpx za = s[pad2(a)], zb = s[pad2((NC - a) & (NC - 1))];
        cpx xa, xb;
        herm_post(za, zb, up[a], xa, xb);
        if (a > 0) emit(a, xa);
        emit(NC - a, xb);                                     // a = 0: the Nyquist bin NC
    }
    if (t == 0) {
        const cpx z = s[pad2(NC / 2)];
        cpx xa, xb;
        herm_post(z, z, up[NC / 2], xa, xb);
        emit(NC / 2, xa);
    }
}
// NC = 512: last radix-2 pass + post-processing.  s holds the output P of pass 1 (Stockham order):
//   Z[k] = P[k] + w P[k+256],  Z[k+256] = P[k] - w P[k+256],  w = exp(-2 pi i k / 512)  (pass-2 table)
// Item k (k = t + 32 i, i < 4) owns bins {k, k+256, 256-k, 512-k}; item 0 owns {128, 256, 384, 512}.
template <class Emit>
PG_HD void fwd_fused_last_512(const cpx* s, int t, const cpx* tabs, Emit&& emit) {
    using R = Radix<512, false>;
    const cpx* w2 = tabs + R::OFF2;                           // W_512^k, k < 256
    const cpx* up = tabs + R::OFFP;                           // W_1024^a, a <= 256
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = t + 32 * i;
        const int kq = k == 0 ? 128 : 256 - k;                // second index of the item
        const cpx p0 = s[pad2(k)], p1 = s[pad2(k + 256)], q0 = s[pad2(kq)], q1 = s[pad2(kq + 256)];
        const cpx w = w2[k], wq = w2[kq];
        const cpx tp = cmul(w, p1), tq = cmul(wq, q1);
        const cpx zk = cadd(p0, tp), zk2 = csub(p0, tp);      // Z[k], Z[k+256]
        const cpx zq = cadd(q0, tq), zq2 = csub(q0, tq);      // Z[kq], Z[kq+256]
        // General item: pairs (k, 512-k) = (Z[k], Z[kq+256]) and (256-k, 256+k) = (Z[kq], Z[k+256]).  Item k = 0 owns
        // {128, 256, 384, 512}: DC/Nyquist from (Z[0], Z[0]), bin 256 from (Z[256], Z[256]), pair (128, 384) from
        // (Z[128], Z[384]).  Selects instead of a divergent branch (see inv_fused_first_512); the third pair and the
        // bin selects exist in the first item only.
        const bool z0 = (i == 0) && (k == 0);
        cpx a1 = zq2, b1 = zk2;
        if (i == 0) { a1 = z0 ? zk : zq2; b1 = z0 ? zq2 : zk2; }
        cpx xa, xb, ya, yb;
        herm_post(zk, a1, up[k], xa, xb);                     // k = 0: (Z[0], Z[0], up[0]) -> xb = X[512]
        herm_post(zq, b1, up[kq], ya, yb);                    // k = 0: (Z[128], Z[384], up[128]) -> X[128], X[384]
        if (i == 0) {
            cpx ma, mb;
            herm_post(zk2, zk2, up[256], ma, mb);             // bin 256 (self-mirrored), used by item 0 only
            // bins: general {k, 512-k, kq, k+256}; item 0 {256, 512, 128, 384}
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
// The 16 spectrum bins (1 .. NC) thread t of a frame consumes, in the order the pre-processing reads them
// (the kernel prefetches them one iteration ahead).  X[0] = 0 (utils.py:38-39) is never loaded: the generic
// program's slot for it carries bin NC/2, which only thread 0 needs.
template <int NC>
PG_HD int inv_bin(int t, int j) {
    constexpr int TG = NC / 16;
    if (NC == 512) {
        const int k = t + 32 * (j >> 2), q = j & 3;
        if (k == 0) return q == 0 ? 512 : q == 1 ? 256 : q == 2 ? 128 : 384;
        return q == 0 ? k : q == 1 ? 512 - k : q == 2 ? 256 - k : k + 256;
    }
    const int a = t + (j >> 1) * TG;
    if (j & 1) return NC - a;
    return a == 0 ? NC / 2 : a;
}
// Generic pre-processing from x[j] = X[inv_bin(t, j)].  The imaginary part of the Nyquist bin is ignored
// like numpy's irfft does.
template <int NC>
PG_HD void inv_pre_generic(cpx* s, int t, const cpx* tabs, float scale, const cpx* x) {
    using R = Radix<NC, true>;
    constexpr int TG = NC / 16;
    const cpx* up = tabs + R::OFFP;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int a = t + i * TG;
        cpx xa = a > 0 ? x[2 * i] : cpx{0.f, 0.f};
        cpx xb = x[2 * i + 1];
        if (a == 0) xb.y = 0.f;
        cpx za, zb;
        herm_pre(xa, xb, up[a], scale, za, zb);
        s[pad2(a)] = za;
        if (a > 0) s[pad2(NC - a)] = zb;
    }
    if (t == 0) {
        const cpx xm = x[0];                                  // bin NC/2
        cpx za, zb;
        herm_pre(xm, xm, up[NC / 2], scale, za, zb);
        s[pad2(NC / 2)] = za;
    }
}
// NC = 512: pre-processing + first radix-2 pass (NS = 1): out[2j] = Z[j] + Z[j+256], out[2j+1] = Z[j] - Z[j+256].
PG_HD void inv_fused_first_512(cpx* s, int t, const cpx* tabs, float scale, const cpx* x) {
    using R = Radix<512, true>;
    const cpx* up = tabs + R::OFFP;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = t + 32 * i;
        const int kq = k == 0 ? 128 : 256 - k;
        cpx zk, zk2, zq, zq2;                                 // Z[k], Z[k+256], Z[kq], Z[kq+256]
        const cpx x0 = x[4 * i], x1 = x[4 * i + 1], x2 = x[4 * i + 2], x3 = x[4 * i + 3];   // order: inv_bin<512>
        // General item: (X[k], X[512-k]) -> (Z[k], Z[kq+256]) and (X[256-k], X[256+k]) -> (Z[kq], Z[k+256]).
        // Item k = 0 (thread 0, i = 0) owns {X[512], X[256], X[128], X[384]} instead and needs a third pair.  It is
        // handled with selects, not a branch: a divergent lane 0 made every warp walk both code paths (the ISTFT spent
        // ~70 of its 1,956 instructions per frame on the reconvergence alone); only i = 0 pays for the extra pair.
        const bool z0 = (i == 0) && (k == 0);
        cpx a0 = x0, a1 = x1;
        if (i == 0) {                                         // compile-time: the selects exist in the first item only
            a0 = z0 ? cpx{0.f, 0.f} : x0;                     // X[0] = 0 (utils.py:38-39)
            a1 = z0 ? cpx{x0.x, 0.f} : x1;                    // X[512], imaginary part ignored like numpy's irfft
        }
        cpx pa, pb, qa, qb;
        herm_pre(a0, a1, up[k], scale, pa, pb);               // k = 0: up[0], pa = Z[0]
        herm_pre(x2, x3, up[kq], scale, qa, qb);              // k = 0: (X[128], X[384]) -> (Z[128], Z[384])
        zk = pa; zq = qa; zq2 = pb; zk2 = qb;
        if (i == 0) {
            cpx ma, mb;
            herm_pre(x1, x1, up[256], scale, ma, mb);         // X[256] (self-mirrored) -> Z[256]
            zq2 = z0 ? qb : pb;
            zk2 = z0 ? ma : qb;
        }
        s[pad2(2 * k)] = cadd(zk, zk2);
        s[pad2(2 * k + 1)] = csub(zk, zk2);
        s[pad2(2 * kq)] = cadd(zq, zq2);
        s[pad2(2 * kq + 1)] = csub(zq, zq2);
    }
}
// Remaining inverse passes.  FIRST_DONE: pass 0 was fused into the pre-processing (NC = 512).
template <int NC, class Sync>
PG_HD void inv_passes(cpx* s, int t, const cpx* tabs, const cpx* win, Sync&& sync) {
    using R = Radix<NC, true>;
    if (!R::SWAP) {
        Pass2<NC, R::R0, 1, true> p;
        p.load(s, t);
        sync();
        p.twiddle_butterfly(nullptr, t);
        p.store(s, t);
        sync();
    }
    {
        Pass2<NC, R::R1, R::NS1, true> p;
        p.load(s, t);
        sync();
        p.twiddle_butterfly(tabs + R::OFF1, t);
        if (R::R2 > 1) p.store(s, t); else p.store_windowed(s, t, win);
        sync();
    }
    if (R::R2 > 1) {
        Pass2<NC, (R::R2 > 1 ? R::R2 : 2), R::NS2, true> p;
        p.load(s, t);
        sync();
        p.twiddle_butterfly(tabs + R::OFF2, t);
        p.store_windowed(s, t, win);
        sync();
    }
}

}  // namespace pgfft
