// Tap tables that turn a Conv1d / ConvTranspose1d into "output position m of phase phi reads
// input row (m + d) of parity q with weight slab w" -- shared by the tensor-core kernel
// (conv_tc.cu) and the fp32 SIMT kernel (conv_simt.cu).
//
//   Conv1d(k, s, p)            (model.py:77):  y[l]      = sum_k W[k] x[l*s + k - p]
//       one output phase; input viewed as rows of parity q = (k-p) mod s, row m + (k-p-q)/s.
//   ConvTranspose1d(k, s, p)   (model.py:88,94,101):  y[l_in*s - p + k] += W[k] x[l_in]
//       s output phases; phase phi = l mod s uses the taps with (phi + p - k) mod s == 0 and
//       reads row m + (phi + p - k)/s of the (stride-1) input, l = m*s + phi.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include "../../include/phasegen.h"
#include "common.cuh"

namespace pg {

constexpr int kMaxTaps = 32;

struct ConvTap { int8_t w_idx; int8_t shift; int8_t parity; int8_t pad_; int32_t d; };
struct ConvGroup { int32_t row0; int8_t parity; int8_t n_taps; int8_t first_tap; int8_t pad_; };

struct ConvPlan {
    int B, C_in, C_out, L_in, L_out, k;
    int IS, OS;                 // input row-parity count (conv stride), output phases (convT stride)
    int n_tile, n_ntiles;       // positions per tile (multiple of 16, <= 256), tiles per phase
    int n_chunks, n_cotiles;    // C_in / 64, C_out / 128
    int strip_rows;             // rows of one activation strip (n_tile + largest shift, rounded to 8)
    int nb;                     // clips per tile: short time axes pack several clips into one 128 x (nb*n_tile) tile
    int acc_stages;             // TMEM accumulator stages: 2 (nb*n_tile <= 256, epilogue overlaps MMA) or 1 (up to 512 columns)
    int pair;                   // tensor-core path: tiles are 256 channels wide, owned by a CTA pair (cta_group::2); strip_rows
                                // is then the half strip (n_tile/2 + largest shift) each CTA of the pair loads
    int mgroups;                // merged tiles: MMAs per weight tile (1..4); nb = mgroups * clips per MMA
    int merged;                 // tensor-core path, short time axes: the nb/mgroups clips of a group are ONE MMA of N = (nb/mgroups)*strip_rows
                                // (accumulator column pitch strip_rows per clip, the strip_rows - n_tile columns between clips
                                // are junk); strip_rows is the full strip, a CTA pair splits the tile by clips
    int clip_group;             // tensor-core tile order: clips per L2-resident group (set by the launcher)
    int whole_clip;             // tensor-core path: a tile holds ALL output positions of its nb clips for its 128/256 channels --
                                // every phase x position tile ("part") side by side in the accumulator, part (phase, nt) of
                                // clip c at column ((c*OS + phase)*n_ntiles + nt)*n_tile -- so per-clip norm statistics are
                                // complete inside the CTA and the epilogue can normalise + activate + write operand planes
    int out_rows, out_ld;
    int n_groups[2], n_taps[2];
    ConvGroup groups[2][kMaxTaps];
    ConvTap taps[2][kMaxTaps];
};

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Returns PG_OK or PG_ERR_INVALID (message via set_error by the caller's PG_REQUIRE wrappers).
static inline int conv_plan_build(const pg_conv_desc* d, ConvPlan* p) {
    const int s = d->stride, k = d->k, pad = d->pad;
    if (s < 1 || s > 2 || k < 1 || k > kMaxTaps) { set_error("conv plan: stride must be 1 or 2 and k <= %d (got s=%d k=%d)", kMaxTaps, s, k); return PG_ERR_UNSUPPORTED; }
    const bool tr = d->kind == PG_CONV_TRANSPOSE;
    const int L_expect = tr ? (d->L_in - 1) * s - 2 * pad + k : (d->L_in + 2 * pad - k) / s + 1;
    // a transposed convolution may be asked for up to s-1 extra rows (output_padding): that is the
    // data gradient of a strided convolution whose input length was not "natural"
    const bool len_ok = tr ? (d->L_out >= L_expect && d->L_out < L_expect + s) : d->L_out == L_expect;
    if (!len_ok || d->L_out < 1) { set_error("conv plan: L_out %d does not match geometry (expected %d)", d->L_out, L_expect); return PG_ERR_INVALID; }
    p->B = d->B; p->C_in = d->C_in; p->C_out = d->C_out; p->L_in = d->L_in; p->L_out = d->L_out; p->k = k;
    p->IS = tr ? 1 : s;
    p->OS = tr ? s : 1;
    p->out_rows = d->out_rows; p->out_ld = d->out_ld;
    p->n_chunks = d->C_in / 64; p->n_cotiles = d->C_out / 128;
    const int l_max = (d->L_out + p->OS - 1) / p->OS;
    p->n_ntiles = (l_max + 255) / 256;
    p->n_tile = (((l_max + p->n_ntiles - 1) / p->n_ntiles) + 15) / 16 * 16;
    int tpg = d->taps_per_group <= 0 ? 16 : d->taps_per_group;
    int max_shift = 0;
    for (int phi = 0; phi < 2; ++phi) { p->n_groups[phi] = 0; p->n_taps[phi] = 0; }
    for (int phi = 0; phi < p->OS; ++phi) {
        // collect (parity, d, w) and sort by (parity, d)
        ConvTap t[kMaxTaps]; int n = 0;
        for (int kk = 0; kk < k; ++kk) {
            ConvTap e; e.w_idx = (int8_t)kk; e.shift = 0; e.pad_ = 0;
            if (tr) {
                int num = phi + pad - kk;
                if (((num % s) + s) % s != 0) continue;
                e.parity = 0; e.d = floordiv(num, s);
            } else {
                int off = kk - pad;
                int q = ((off % s) + s) % s;
                e.parity = (int8_t)q; e.d = (off - q) / s;
            }
            t[n++] = e;
        }
        for (int i = 1; i < n; ++i) {           // insertion sort
            ConvTap e = t[i]; int j = i - 1;
            while (j >= 0 && (t[j].parity > e.parity || (t[j].parity == e.parity && t[j].d > e.d))) { t[j + 1] = t[j]; --j; }
            t[j + 1] = e;
        }
        int ng = 0, i = 0;
        while (i < n) {
            ConvGroup g; g.parity = t[i].parity; g.row0 = t[i].d; g.first_tap = (int8_t)i; g.pad_ = 0;
            int j = i;
            while (j < n && t[j].parity == g.parity && (j - i) < tpg && (t[j].d - g.row0) < tpg &&
                   (p->n_tile + (t[j].d - g.row0) <= 256)) {
                t[j].shift = (int8_t)(t[j].d - g.row0);
                if (t[j].shift > max_shift) max_shift = t[j].shift;
                ++j;
            }
            g.n_taps = (int8_t)(j - i);
            p->groups[phi][ng++] = g;
            i = j;
        }
        for (int q = 0; q < n; ++q) p->taps[phi][q] = t[q];
        p->n_groups[phi] = ng; p->n_taps[phi] = n;
    }
    // CTA pairs: 0 = auto (whenever the output channels split into 256-wide slabs), 1 = off, 2 = required
    p->pair = (d->tc_cta_pair != 1 && d->C_out % 256 == 0 && d->precision != PG_PREC_FP32_SIMT) ? 1 : 0;
    if (d->tc_cta_pair == 2 && !p->pair) { set_error("conv plan: CTA pairs need C_out %% 256 == 0 (got %d)", d->C_out); return PG_ERR_INVALID; }
    const int n_cta = p->pair ? p->n_tile / 2 : p->n_tile;        // positions of a clip whose rows one CTA loads
    p->strip_rows = (n_cta + max_shift + 7) / 8 * 8;
    if (p->strip_rows > 256) p->strip_rows = 256;
    // Several clips per tile when one clip's positions leave the 256-column accumulator mostly
    // empty: the weight tile is then shared by nb MMAs (one per clip).
    p->nb = 1;
    // Fewer than three products per MAC (F16X2, BF16): the tensor pipe outruns the weight stream even for
    // full-width tiles (measured: 72 % pipe-active at N = 176, profiles/r01_conv_tc_ncu_full_v2.csv), so two
    // clips share every weight tile (2 x 176 accumulator columns, one TMEM stage); the three-product forms
    // are pipe/power-bound there and keep the double-buffered accumulator.
    const bool weight_bound = d->precision == PG_PREC_F16X2 || d->precision == PG_PREC_BF16 || d->precision == PG_PREC_F16;
    if ((p->n_ntiles == 1 || weight_bound) && d->max_clips_per_tile != 1) {
        // Up to 512 accumulator columns (one TMEM stage) when the strips of that many clips still leave
        // room for the weight ring: the weight tile, the dominant L2->SM stream for short time axes, is
        // then shared by twice as many MMAs.  Otherwise up to 256 columns, double-buffered.
        const int planes = (d->precision == PG_PREC_BF16X3 || d->precision == PG_PREC_F16X3 || d->precision == PG_PREC_F16X2) ? 2 : 1;
        const int max_rows = (72 * 1024) / (planes * 128);          // rows of all clips' strips in one slot
        int nb = 512 / p->n_tile;
        if (nb * p->strip_rows > max_rows) nb = max_rows / p->strip_rows;
        if (nb * p->n_tile <= 256 || (p->n_tile > 128 && !weight_bound)) {   // not worth giving up the second stage
            nb = 256 / p->n_tile;
            if (nb * p->strip_rows > 256) nb = 256 / p->strip_rows;
        }
        if (nb > 16) nb = 16;
        if (d->max_clips_per_tile > 1 && nb > d->max_clips_per_tile) nb = d->max_clips_per_tile;
        if (nb > d->B) nb = d->B;
        if (nb < 1) nb = 1;
        p->nb = nb;
    }
    if (const char* e = getenv("PG_TC_NB")) {                      // experiment hook: force the bundle size
        int nb = atoi(e);
        if (nb >= 1 && nb * p->n_tile <= 512 && nb <= d->B) p->nb = nb;
    }
    // Merged clips.  With one MMA per clip a short time axis means tiny MMAs (N = 16..96) whose fixed operand-fetch
    // cost dominates (measured ~32 + N/2 cycles per MMA) and few tiles.  Laying the clips' strips end to end
    // makes them one N = nb*strip_rows operand: a tap shifted by s rows reads rows [s, s + N), so clip c's
    // outputs land in columns [c*strip_rows, c*strip_rows + n_tile) and the columns in between (fed by rows that
    // straddle two clips) are never read back.  nb minimises waves x (32 + N/2) over the persistent grid.
    p->merged = 0; p->mgroups = 1;
    {
        const int strip_full = (p->n_tile + max_shift + 7) / 8 * 8;
        const int step = p->pair ? 2 : 1;
        int nb_max = 256 / strip_full;
        if (nb_max > d->B) nb_max = d->B;
        if (d->max_clips_per_tile > 1 && nb_max > d->max_clips_per_tile) nb_max = d->max_clips_per_tile;
        nb_max -= nb_max % step;
        const char* off = getenv("PG_TC_MERGED");
        if (p->n_ntiles == 1 && nb_max >= 2 && d->max_clips_per_tile != 1 && !(off && atoi(off) == 0)) {
            const int units = (d->tc_max_ctas > 0 ? d->tc_max_ctas : 148) / (p->pair ? 2 : 1);
            const int slabs = (p->pair ? p->n_cotiles / 2 : p->n_cotiles) * p->OS;
            long best = -1; int best_nb = step;
            for (int nb = step; nb <= nb_max; nb += step) {
                const long tiles = (long)slabs * ((d->B + nb - 1) / nb);
                const long cost = ((tiles + units - 1) / units) * (64 + (long)nb * strip_full);
                if (best < 0 || cost <= best) { best = cost; best_nb = nb; }
            }
            p->merged = 1; p->nb = best_nb; p->strip_rows = strip_full;
            // Several MMAs (up to 512 accumulator columns in all, one TMEM stage) per weight tile while that still leaves
            // a tile for >= 80 % of the CTA pairs: the weight stream from L2 -- the bound of these layers, whose
            // weights are read once per tile -- halves.
            const char* g2 = getenv("PG_TC_MGROUPS");           // experiment hook: cap the number of groups
            const int g_cap = g2 ? atoi(g2) : 2;                // 3 groups measured worse (u1 train shape: 1.92 vs 1.42 ms: the tile count drops to 1.3 waves)
            const int planes_b = (d->precision == PG_PREC_BF16X3 || d->precision == PG_PREC_F16X3 || d->precision == PG_PREC_F16X2) ? 2 : 1;
            const char* fr = getenv("PG_TC_MG_FRAC");           // experiment hook: fraction of the grid multi-group tiles must fill
            const double need = (fr ? atof(fr) : 0.8) * units;    // 0.8: measured optimum (profiles/r01_conv_mgroups_ab.log)
            // (only where the weight stream is the bound: with three products per MAC the pipe is, and the lost
            //  epilogue overlap costs 3 %: profiles/r01_conv_mgroups_ab.log)
            for (int g = 4; g >= 2 && weight_bound; --g) {
                if (g > g_cap || g * best_nb > d->B || g * best_nb * strip_full > 512) continue;
                const long tiles_g = (long)slabs * ((d->B + g * best_nb - 1) / (g * best_nb));
                const long slot_bytes = (long)planes_b * (g * best_nb / (p->pair ? 2 : 1)) * strip_full * 128;   // per CTA, one strip slot
                if ((double)tiles_g < need || 2 * slot_bytes > 120 * 1024) continue;
                p->mgroups = g; p->nb = g * best_nb;
                break;
            }
        }
    }
    // Whole-clip tiles (requested for the fused per-clip norm epilogue): OS x n_ntiles parts of one clip in <= 512 columns.
    // Merged tiles already hold whole clips when there is one phase; otherwise they are given up for this layer.
    p->whole_clip = 0;
    if (d->tc_whole_clip) {
        const int parts = p->OS * p->n_ntiles;
        if (p->merged && p->OS > 1) {
            // short two-phase layers: un-merging them costs more (N = 96 MMAs: measured 0.96 -> 1.58 ms on u4 at the
            // BASELINE shape) than the separate normalising pass saves, so they keep the two-pass form
            set_error("conv plan: merged two-phase tiles do not hold a whole clip");
            return PG_ERR_UNSUPPORTED;
        }
        if (!p->merged) {
            if (parts * p->n_tile > 512) { set_error("conv plan: a whole clip needs %d accumulator columns (> 512)", parts * p->n_tile); return PG_ERR_UNSUPPORTED; }
            if (parts > 1) {
                p->whole_clip = 1;
                int nb = 1;
                if (weight_bound && 2 * parts * p->n_tile <= 512) nb = 2;      // two clips per weight tile where the weight stream is the bound
                if (nb > d->B) nb = d->B;
                const int planes = (d->precision == PG_PREC_BF16X3 || d->precision == PG_PREC_F16X3 || d->precision == PG_PREC_F16X2) ? 2 : 1;
                if (nb * p->n_ntiles * p->strip_rows * planes * 128 > 72 * 1024) nb = 1;
                // two strip slots plus at least two weight slots must fit the 227 KB of shared memory
                if (p->n_ntiles * p->strip_rows * planes * 128 > 80 * 1024) {
                    set_error("conv plan: the %d strips of a whole clip (%d rows each) do not fit shared memory", p->n_ntiles, p->strip_rows);
                    return PG_ERR_UNSUPPORTED;
                }
                p->nb = nb;
            }
        }
    }
    {
        const int cols = p->merged ? p->nb * p->strip_rows : p->nb * p->n_tile * (p->whole_clip ? p->OS * p->n_ntiles : 1);
        if (cols > 512) { set_error("conv plan: tile needs %d accumulator columns (> 512)", cols); return PG_ERR_UNSUPPORTED; }
        p->acc_stages = cols <= 256 ? 2 : 1;
    }
    p->clip_group = d->B;
    return PG_OK;
}

}  // namespace pg
