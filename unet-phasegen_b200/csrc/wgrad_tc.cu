// Weight gradient of Conv1d / ConvTranspose1d on tcgen05 (backward of model.py:77,88,94,101;
// the reference gets it from autograd through cuDNN, train.py:61).
//
//   dW[tap][co][ci] = sum over clips b and output positions m of
//                     G[b][m*OS + phase(tap)][co] * X[b][(m + d(tap))*IS + parity(tap)][ci]
// with the same tap tables as the forward kernel (conv_plan.h).  GEMM view per work item:
// D[128 co][nci ci] (one accumulator per tap of a group, up to 4 x 128 TMEM columns) over
// K = (clip, position).  Both operands are channels-last, i.e. **MN-major** for this GEMM: a TMA
// box {64 channels, R rows} lands in shared memory exactly as a 128-byte-swizzled MN-major atom
// stack (64 elements x 8 K-rows per atom), so no transposition is needed.  The taps of a group
// read the same X rows shifted by one: the X strip is loaded once per K chunk and each tap's MMA
// descriptor starts `shift` rows (shift*128 B) further -- the K-direction analogue of the forward
// kernel's strip reuse.  Every work item owns its full K reduction: no split-K, no atomics,
// deterministic.  Output is the packed [tap][C_out][C_in] fp32 layout (128 B contiguous per thread).
#include "tc_ptx.cuh"
#include "conv_plan.h"

namespace pg {

struct WgradParams {
    ConvPlan plan;
    float* dw;
    int nci;            // ci tile width (MMA N): 64 or 128
    int R, RB;          // K rows per stage of G, of the X strip (R + 8)
    int n_mchunks;
    int n_stages;
    int n_terms;
    int a_plane, b_plane;   // bytes of one precision plane of A (2 co blocks) / B (nci/64 ci blocks)
};

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 4;

struct WgItem { int co_tile, ci_tile, phase, group; };

__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
    const ConvPlan& pl = p.plan;
    const int n_groups = pl.n_groups[0] + pl.n_groups[1];
    const int n_citiles = pl.C_in / p.nci;
    WgItem w;
    int g = item % n_groups; item /= n_groups;
    w.ci_tile = item % n_citiles; w.co_tile = item / n_citiles;
    w.phase = 0;
    if (g >= pl.n_groups[0]) { g -= pl.n_groups[0]; w.phase = 1; }
    w.group = g;
    return w;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ WgradParams prm) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const ConvPlan& pl = prm.plan;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const bool three = prm.n_terms == 3;
    const int planes = three ? 2 : 1;
    const int stage_bytes = planes * (prm.a_plane + prm.b_plane);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)prm.n_stages * stage_bytes);
    uint64_t* full = bars;                        // [kWgMaxStages]
    uint64_t* empty = bars + kWgMaxStages;        // [kWgMaxStages]
    uint64_t* accFull = bars + 2 * kWgMaxStages;
    uint64_t* accEmpty = accFull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accEmpty + 1);

    // shuffled warp index: provably warp-uniform role branches (see conv_tc.cu)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int n_groups = pl.n_groups[0] + pl.n_groups[1];
    const int n_items = pl.n_cotiles * (pl.C_in / prm.nci) * n_groups;
    const int n_kchunks = pl.B * prm.n_mchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_g_hi); tma_prefetch_desc(&map_x_hi);
        for (int i = 0; i < prm.n_stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(accFull, 1); mbar_init(accEmpty, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t bytes = (uint32_t)stage_bytes;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const WgItem w = wg_decode(prm, item);
                const ConvGroup grp = pl.groups[w.phase][w.group];
                for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
                    const int b = kc / prm.n_mchunks, m0 = (kc % prm.n_mchunks) * prm.R;
                    const int s = it % prm.n_stages; const uint32_t ph = (it / prm.n_stages) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, bytes);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    for (int pln = 0; pln < planes; ++pln) {
                        const CUtensorMap* mg = pln ? &map_g_lo : &map_g_hi;
                        const CUtensorMap* mx = pln ? &map_x_lo : &map_x_hi;
                        uint8_t* a = st + (size_t)pln * prm.a_plane;
                        uint8_t* bq = st + (size_t)planes * prm.a_plane + (size_t)pln * prm.b_plane;
                        for (int h = 0; h < 2; ++h)
                            tma_load_4d(a + (size_t)h * prm.R * 128, mg, full + s, w.co_tile * 128 + h * 64, w.phase, m0, b);
                        for (int h = 0; h < prm.nci / 64; ++h)
                            tma_load_4d(bq + (size_t)h * prm.RB * 128, mx, full + s, w.ci_tile * prm.nci + h * 64, grp.parity,
                                        m0 + grp.row0, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp, warp-uniform control flow; one elected lane issues the tcgen05 instructions
            uint32_t it = 0, n_it = 0;
            const uint32_t idesc = make_idesc_bf16_mn(prm.nci);
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc_flag) { if (elect_one()) umma_bf16(d, da, db, idesc, acc_flag); };
            auto commit = [&](uint64_t* bar) { if (elect_one()) umma_commit(bar); __syncwarp(); };
            const uint32_t lbo_a = (uint32_t)prm.R * 128u, lbo_b = (uint32_t)prm.RB * 128u;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_it) {
                const WgItem w = wg_decode(prm, item);
                const ConvGroup grp = pl.groups[w.phase][w.group];
                mbar_wait_sleep(accEmpty, (n_it & 1) ^ 1, 200);
                tc_fence_after();
                for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
                    const int s = it % prm.n_stages; const uint32_t ph = (it / prm.n_stages) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t a_hi = st, a_lo = st + prm.a_plane;
                    const uint32_t b_hi = st + planes * prm.a_plane, b_lo = b_hi + prm.b_plane;
                    for (int j = 0; j < grp.n_taps; ++j) {
                        const ConvTap tp = pl.taps[w.phase][grp.first_tap + j];
                        const uint32_t d_tmem = tmem_base + j * prm.nci;
                        for (int ks = 0; ks < prm.R / 16; ++ks) {
                            const uint32_t acc = (kc | ks) ? 1u : 0u;
                            const uint32_t ao = (uint32_t)ks * 2048u, bo = ((uint32_t)tp.shift + (uint32_t)ks * 16u) * 128u;
                            const uint64_t da_hi = make_desc_sw128_mn(a_hi + ao, lbo_a);
                            const uint64_t db_hi = make_desc_sw128_mn(b_hi + bo, lbo_b);
                            if (three) {
                                const uint64_t da_lo = make_desc_sw128_mn(a_lo + ao, lbo_a);
                                const uint64_t db_lo = make_desc_sw128_mn(b_lo + bo, lbo_b);
                                mma(d_tmem, da_lo, db_hi, acc);
                                mma(d_tmem, da_hi, db_lo, 1);
                                mma(d_tmem, da_hi, db_hi, 1);
                            } else {
                                mma(d_tmem, da_hi, db_hi, acc);
                            }
                        }
                    }
                    commit(empty + s);
                }
                commit(accFull);
            }
        }
    } else {
        const int q = warp & 3;
        uint32_t n_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_it) {
            const WgItem w = wg_decode(prm, item);
            const ConvGroup grp = pl.groups[w.phase][w.group];
            mbar_wait_sleep(accFull, n_it & 1, 1000);
            tc_fence_after();
            const int co = w.co_tile * 128 + q * 32 + lane;
            for (int j = 0; j < grp.n_taps; ++j) {
                const ConvTap tp = pl.taps[w.phase][grp.first_tap + j];
                float* dst = prm.dw + ((size_t)tp.w_idx * pl.C_out + co) * pl.C_in + (size_t)w.ci_tile * prm.nci;
                const uint32_t taddr = tmem_base + j * prm.nci + ((uint32_t)(q * 32) << 16);
                for (int c0 = 0; c0 < prm.nci; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace pg

extern "C" int pg_wgrad_tc(const pg_conv_desc* d, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* g_hi,
                           const uint16_t* g_lo, int g_rows, float* dw_packed, pg_stream stream) {
    using namespace pg;
    PG_REQUIRE(d && x_hi && g_hi && dw_packed, "pg_wgrad_tc: null pointer");
    PG_REQUIRE(d->precision == PG_PREC_BF16X3 || d->precision == PG_PREC_BF16, "pg_wgrad_tc: precision must be BF16X3 or BF16");
    const bool three = d->precision == PG_PREC_BF16X3;
    PG_REQUIRE(!three || (x_lo && g_lo), "pg_wgrad_tc: lo planes required for BF16X3");
    PG_REQUIRE(d->C_in % 64 == 0 && d->C_out % 128 == 0 && d->in_ld % 8 == 0, "pg_wgrad_tc: needs C_in %% 64 == 0 and C_out %% 128 == 0");
    WgradParams prm;
    pg_conv_desc dd = *d;
    dd.taps_per_group = 4;            // 4 accumulators of <= 128 columns fill TMEM
    dd.max_clips_per_tile = 1;
    int rc = conv_plan_build(&dd, &prm.plan);
    if (rc != PG_OK) return rc;
    const ConvPlan& pl = prm.plan;
    int sm_count, max_smem;
    device_limits(&sm_count, &max_smem);
    prm.dw = dw_packed;
    prm.n_terms = three ? 3 : 1;
    prm.nci = d->C_in % 128 == 0 ? 128 : 64;
    const int l_max = (pl.L_out + pl.OS - 1) / pl.OS;
    const int r_cap = three ? 64 : 128;
    prm.n_mchunks = (l_max + r_cap - 1) / r_cap;
    prm.R = ((l_max + prm.n_mchunks - 1) / prm.n_mchunks + 15) / 16 * 16;
    prm.RB = prm.R + 8;
    prm.a_plane = 2 * prm.R * 128;
    prm.b_plane = (prm.nci / 64) * prm.RB * 128;
    const int planes = three ? 2 : 1;
    const int stage_bytes = planes * (prm.a_plane + prm.b_plane);
    int n_stages = (max_smem - 2048) / stage_bytes;
    if (n_stages > kWgMaxStages) n_stages = kWgMaxStages;
    PG_REQUIRE(n_stages >= 2, "pg_wgrad_tc: stage of %d bytes does not fit twice in shared memory", stage_bytes);
    prm.n_stages = n_stages;
    const size_t smem_bytes = (size_t)n_stages * stage_bytes + 2048;
    PG_REQUIRE(g_rows >= ((pl.L_out + pl.OS - 1) / pl.OS) * pl.OS, "pg_wgrad_tc: gradient buffer rows %d too small for L_out %d", g_rows, pl.L_out);

    CUtensorMap mg_hi, mg_lo, mx_hi, mx_lo;
    {
        const int OS = pl.OS;
        uint64_t dims[4] = {(uint64_t)d->C_out, (uint64_t)OS, (uint64_t)((pl.L_out + OS - 1) / OS), (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->C_out * 2, (uint64_t)d->C_out * 2 * OS, (uint64_t)g_rows * d->C_out * 2};
        uint32_t box[4] = {64, 1, (uint32_t)prm.R, 1};
        if ((rc = encode_bf16_map(&mg_hi, g_hi, 4, dims, str, box, "g_hi")) != PG_OK) return rc;
        if ((rc = encode_bf16_map(&mg_lo, three ? g_lo : g_hi, 4, dims, str, box, "g_lo")) != PG_OK) return rc;
    }
    {
        const int IS = pl.IS;
        uint64_t dims[4] = {(uint64_t)d->C_in, (uint64_t)IS, (uint64_t)((d->L_in + IS - 1) / IS), (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->in_ld * 2, (uint64_t)d->in_ld * 2 * IS, (uint64_t)d->in_rows * d->in_ld * 2};
        uint32_t box[4] = {64, 1, (uint32_t)prm.RB, 1};
        PG_REQUIRE(d->in_rows >= ((d->L_in + IS - 1) / IS) * IS, "pg_wgrad_tc: in_rows too small");
        if ((rc = encode_bf16_map(&mx_hi, x_hi, 4, dims, str, box, "x_hi")) != PG_OK) return rc;
        if ((rc = encode_bf16_map(&mx_lo, three ? x_lo : x_hi, 4, dims, str, box, "x_lo")) != PG_OK) return rc;
    }
    static size_t configured = 0;
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) { set_error("wgrad_tc: cannot opt in to %zu bytes of shared memory: %s", smem_bytes, cudaGetErrorString(e)); return PG_ERR_CUDA; }
        configured = smem_bytes;
    }
    const int n_items = pl.n_cotiles * (pl.C_in / prm.nci) * (pl.n_groups[0] + pl.n_groups[1]);
    int grid = n_items < sm_count ? n_items : sm_count;
    if (d->tc_max_ctas > 0 && grid > d->tc_max_ctas) grid = d->tc_max_ctas;
    wgrad_tc_kernel<<<grid, kWgThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(mg_hi, mg_lo, mx_hi, mx_lo, prm);
    return check_launch("wgrad_tc_kernel");
}
