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
    int dw_bf16;        // dw is a bf16 buffer (same packed order): the gradient leaves for a bf16 all-reduce
    int nci;            // ci tile width (MMA N): 64 or 128
    int R, RB;          // K rows per stage of G, of the X strip (R + 8)
    int n_mchunks;
    int n_stages;
    int n_terms;
    int a_plane, b_plane;   // bytes of one precision plane of A (2 co blocks) / B (nci/64 ci blocks; half of them per CTA of a pair)
    int pair;               // work items are 256 output channels wide, owned by a CTA pair (cta_group::2, see conv_tc.cu)
};

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 6;

struct WgItem { int co_tile, ci_tile, phase, group; };

__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
    const ConvPlan& pl = p.plan;
    const int n_groups = pl.n_groups[0] + pl.n_groups[1];
    const int n_citiles = pl.C_in / p.nci;
    WgItem w;
    int g = item % n_groups; item /= n_groups;
    w.ci_tile = item % n_citiles; w.co_tile = item / n_citiles;     // pair mode: co_tile counts 256-channel slabs
    w.phase = 0;
    if (g >= pl.n_groups[0]) { g -= pl.n_groups[0]; w.phase = 1; }
    w.group = g;
    return w;
}

template <bool PAIR>
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
    const int n_items = (PAIR ? pl.n_cotiles / 2 : pl.n_cotiles) * (pl.C_in / prm.nci) * n_groups;
    const int n_kchunks = pl.B * prm.n_mchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_g_hi); tma_prefetch_desc(&map_x_hi);
        for (int i = 0; i < prm.n_stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(accFull, 1); mbar_init(accEmpty, PAIR ? 8 : 4);
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_bblk = PAIR ? prm.nci / 128 : prm.nci / 64;       // 64-channel X blocks this CTA loads

    if (warp == 0) {
        {   // warp-uniform producer: every lane runs the loop, one elected lane issues (see conv_tc.cu)
            int st_i = 0; uint32_t st_ph = 0;
            const uint32_t bytes = (uint32_t)stage_bytes;
            for (int item = item0; item < n_items; item += item_step) {
                const WgItem w = wg_decode(prm, item);
                const ConvGroup grp = pl.groups[w.phase][w.group];
                int b = 0, mc = 0;
                for (int kc = 0; kc < n_kchunks; ++kc) {
                    const int m0 = mc * prm.R;
                    const int s = st_i; const uint32_t ph = st_ph;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    if (elect_one()) {
                    if (!PAIR) mbar_expect_tx(full + s, bytes);
                    else if (rank == 0) mbar_expect_tx(full + s, 2 * bytes);      // both CTAs' operands
                    const int co0 = (PAIR ? w.co_tile * 2 + rank : w.co_tile) * 128;
                    const int ci0 = w.ci_tile * prm.nci + (PAIR ? rank * (prm.nci / 2) : 0);
                    for (int pln = 0; pln < planes; ++pln) {
                        const CUtensorMap* mg = pln ? &map_g_lo : &map_g_hi;
                        const CUtensorMap* mx = pln ? &map_x_lo : &map_x_hi;
                        uint8_t* a = st + (size_t)pln * prm.a_plane;
                        uint8_t* bq = st + (size_t)planes * prm.a_plane + (size_t)pln * prm.b_plane;
                        for (int h = 0; h < 2; ++h) {
                            if (PAIR) tma_load_4d_pair(a + (size_t)h * prm.R * 128, mg, full + s, co0 + h * 64, w.phase, m0, b);
                            else tma_load_4d(a + (size_t)h * prm.R * 128, mg, full + s, co0 + h * 64, w.phase, m0, b);
                        }
                        for (int h = 0; h < n_bblk; ++h) {
                            if (PAIR) tma_load_4d_pair(bq + (size_t)h * prm.RB * 128, mx, full + s, ci0 + h * 64, grp.parity, m0 + grp.row0, b);
                            else tma_load_4d(bq + (size_t)h * prm.RB * 128, mx, full + s, ci0 + h * 64, grp.parity, m0 + grp.row0, b);
                        }
                    }
                    }
                    __syncwarp();
                    if (++st_i == prm.n_stages) { st_i = 0; st_ph ^= 1; }
                    if (++mc == prm.n_mchunks) { mc = 0; ++b; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // whole warp, warp-uniform control flow; one elected lane issues the tcgen05 instructions
            uint32_t n_it = 0;
            int st_i = 0; uint32_t st_ph = 0;
            const uint32_t idesc = (make_idesc_bf16_mn(prm.nci) & ~(0x1Fu << 24)) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);   // M = 256 for pairs
            // descriptors as their low 32-bit words (tc_ptx.cuh): the issue loop only adds to start addresses
            auto mma = [&](uint32_t d, uint32_t da, uint32_t db, uint32_t acc_flag) {
                if (elect_one()) { if (PAIR) umma_words_pair(d, da, db, idesc, acc_flag); else umma_words(d, da, db, idesc, acc_flag); }
            };
            auto commit = [&](uint64_t* bar) {
                if (elect_one()) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); }
                __syncwarp();
            };
            const uint32_t lbo_a = (uint32_t)prm.R * 128u, lbo_b = (uint32_t)prm.RB * 128u;
            for (int item = item0; item < n_items; item += item_step, ++n_it) {
                const WgItem w = wg_decode(prm, item);
                const ConvGroup grp = pl.groups[w.phase][w.group];
                mbar_wait_sleep(accEmpty, (n_it & 1) ^ 1, 200);
                tc_fence_after();
                for (int kc = 0; kc < n_kchunks; ++kc) {
                    const int s = st_i; const uint32_t ph = st_ph;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t a_hi = st, a_lo = st + prm.a_plane;
                    const uint32_t b_hi = st + planes * prm.a_plane, b_lo = b_hi + prm.b_plane;
                    const uint32_t aw_hi0 = desc_lo_sw128_mn(a_hi, lbo_a), aw_lo0 = desc_lo_sw128_mn(a_lo, lbo_a);
                    const int n_ks = prm.R / 16;
                    for (int j = 0; j < grp.n_taps; ++j) {
                        const ConvTap tp = pl.taps[w.phase][grp.first_tap + j];
                        const uint32_t d_tmem = tmem_base + j * prm.nci;
                        // 16 K rows further = 2048 B in both operands = +128 descriptor units
                        uint32_t aw_hi = aw_hi0, aw_lo = aw_lo0;
                        uint32_t bw_hi = desc_lo_sw128_mn(b_hi + (uint32_t)tp.shift * 128u, lbo_b);
                        uint32_t bw_lo = desc_lo_sw128_mn(b_lo + (uint32_t)tp.shift * 128u, lbo_b);
                        for (int ks = 0; ks < n_ks; ++ks, aw_hi += 128, aw_lo += 128, bw_hi += 128, bw_lo += 128) {
                            const uint32_t acc = (kc | ks) ? 1u : 0u;
                            if (three) {
                                mma(d_tmem, aw_lo, bw_hi, acc);
                                mma(d_tmem, aw_hi, bw_lo, 1);
                                mma(d_tmem, aw_hi, bw_hi, 1);
                            } else {
                                mma(d_tmem, aw_hi, bw_hi, acc);
                            }
                        }
                    }
                    commit(empty + s);
                    if (++st_i == prm.n_stages) { st_i = 0; st_ph ^= 1; }
                }
                commit(accFull);
            }
        }
    } else {
        const int q = warp & 3;
        uint32_t n_it = 0;
        for (int item = item0; item < n_items; item += item_step, ++n_it) {
            const WgItem w = wg_decode(prm, item);
            const ConvGroup grp = pl.groups[w.phase][w.group];
            mbar_wait_sleep(accFull, n_it & 1, 1000);
            tc_fence_after();
            const int co = (PAIR ? w.co_tile * 2 + rank : w.co_tile) * 128 + q * 32 + lane;
            for (int j = 0; j < grp.n_taps; ++j) {
                const ConvTap tp = pl.taps[w.phase][grp.first_tap + j];
                const size_t off = ((size_t)tp.w_idx * pl.C_out + co) * pl.C_in + (size_t)w.ci_tile * prm.nci;
                float* dst = prm.dw + off;
                __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(prm.dw) + off;
                const uint32_t taddr = tmem_base + j * prm.nci + ((uint32_t)(q * 32) << 16);
                for (int c0 = 0; c0 < prm.nci; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
                    if (prm.dw_bf16) {
                        __align__(16) __nv_bfloat162 h[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                        *reinterpret_cast<uint4*>(dst16 + c0) = *reinterpret_cast<uint4*>(&h[0]);
                        *reinterpret_cast<uint4*>(dst16 + c0 + 8) = *reinterpret_cast<uint4*>(&h[4]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(accEmpty); else mbar_arrive(accEmpty); }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace pg

extern "C" int pg_wgrad_tc(const pg_conv_desc* d, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* g_hi,
                           const uint16_t* g_lo, int g_rows, void* dw_packed, int dw_dtype, pg_stream stream) {
    using namespace pg;
    PG_REQUIRE(d && x_hi && g_hi && dw_packed, "pg_wgrad_tc: null pointer");
    PG_REQUIRE(dw_dtype == PG_DT_F32 || dw_dtype == PG_DT_BF16, "pg_wgrad_tc: gradient dtype must be PG_DT_F32 or PG_DT_BF16");
    PG_REQUIRE(d->precision == PG_PREC_BF16X3 || d->precision == PG_PREC_BF16, "pg_wgrad_tc: precision must be BF16X3 or BF16");
    const bool three = d->precision == PG_PREC_BF16X3;
    PG_REQUIRE(!three || (x_lo && g_lo), "pg_wgrad_tc: lo planes required for BF16X3");
    PG_REQUIRE(d->C_in % 64 == 0 && d->C_out % 128 == 0 && d->in_ld % 8 == 0, "pg_wgrad_tc: needs C_in %% 64 == 0 and C_out %% 128 == 0");
    WgradParams prm;
    pg_conv_desc dd = *d;
    // TMEM holds 512 accumulator columns: 4 taps x 128 input channels, or (wide) 2 taps x 256 input channels.
    // The wide form reads 12 KB of shared-memory operands per 128 math cycles instead of 8 KB per 64.
    // Measured (train shapes, bf16): 12.03 -> 11.01 ms per step with the wide form.
    bool wide = d->C_in % 256 == 0;
    if (const char* e = getenv("PG_WG_NCI")) wide = wide && atoi(e) == 256;                 // A/B hook: PG_WG_NCI=128
    dd.taps_per_group = wide ? 2 : 4;
    dd.max_clips_per_tile = 1;
    int rc = conv_plan_build(&dd, &prm.plan);
    if (rc != PG_OK) return rc;
    const ConvPlan& pl = prm.plan;
    int sm_count, max_smem;
    device_limits(&sm_count, &max_smem);
    prm.dw = static_cast<float*>(dw_packed);
    prm.dw_bf16 = dw_dtype == PG_DT_BF16 ? 1 : 0;
    prm.n_terms = three ? 3 : 1;
    prm.nci = wide ? 256 : d->C_in % 128 == 0 ? 128 : 64;
    {
        int want = d->tc_cta_pair;
        if (const char* e = getenv("PG_WG_PAIR")) want = atoi(e);   // A/B hook: 1 = single CTAs
        // opt-in only (tc_cta_pair = 2): measured neutral-to-slightly-slower than single CTAs at the train shapes
        // (12.46 vs 12.29 ms per step) -- this kernel is not bound by its shared-memory operand reads
        prm.pair = (want == 2 && prm.nci >= 128 && d->C_out % 256 == 0) ? 1 : 0;
    }
    const int l_max = (pl.L_out + pl.OS - 1) / pl.OS;
    const int r_cap = three ? 64 : 128;
    prm.n_mchunks = (l_max + r_cap - 1) / r_cap;
    prm.R = ((l_max + prm.n_mchunks - 1) / prm.n_mchunks + 15) / 16 * 16;
    prm.RB = prm.R + 8;
    prm.a_plane = 2 * prm.R * 128;
    prm.b_plane = (prm.nci / 64 / (prm.pair ? 2 : 1)) * prm.RB * 128;
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
    const bool pair = prm.pair != 0;
    static size_t configured_dev[kMaxDevices][2] = {};
    size_t* configured = configured_dev[current_device_slot()];
    if (smem_bytes > configured[pair]) {
        cudaError_t e = pair ? cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)
                             : cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) { set_error("wgrad_tc: cannot opt in to %zu bytes of shared memory: %s", smem_bytes, cudaGetErrorString(e)); return PG_ERR_CUDA; }
        configured[pair] = smem_bytes;
    }
    const int n_items = (pair ? pl.n_cotiles / 2 : pl.n_cotiles) * (pl.C_in / prm.nci) * (pl.n_groups[0] + pl.n_groups[1]);
    int units = pair ? sm_count / 2 : sm_count;
    if (const char* e = getenv("PG_TC_MAX_CTAS")) {           // experiment hook: leave SMs free (e.g. for NCCL)
        const int cap = atoi(e);
        if (cap > 0 && units > (pair ? cap / 2 : cap)) units = pair ? cap / 2 : cap;
    }
    if (d->tc_max_ctas > 0 && units > (pair ? (d->tc_max_ctas + 1) / 2 : d->tc_max_ctas)) units = pair ? (d->tc_max_ctas + 1) / 2 : d->tc_max_ctas;
    if (units > n_items) units = n_items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pair ? 2 * units : units);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t le = pair ? cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<true>, mg_hi, mg_lo, mx_hi, mx_lo, prm)
                          : cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<false>, mg_hi, mg_lo, mx_hi, mx_lo, prm);
    if (le != cudaSuccess) { set_error("wgrad_tc_kernel launch failed: %s", cudaGetErrorString(le)); return PG_ERR_CUDA; }
    return check_launch("wgrad_tc_kernel");
}
