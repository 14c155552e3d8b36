// Error plumbing and device checks of the phasegen C ABI.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "conv_plan.h"

namespace pg {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return PG_ERR_CUDA;
    }
    return PG_OK;
}
}  // namespace pg

extern "C" const char* pg_last_error(void) { return pg::g_err; }
extern "C" int pg_abi_version(void) { return 4; }

extern "C" int pg_check_device(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        pg::set_error("no CUDA device: the phasegen path has no CPU fallback");
        return PG_ERR_CUDA;
    }
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { pg::set_error("cudaGetDeviceProperties failed"); return PG_ERR_CUDA; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (p.major != 10) {
        pg::set_error("device '%s' is sm_%d%d; this library is built for sm_100a (B200) only", p.name, p.major, p.minor);
        return PG_ERR_UNSUPPORTED;
    }
    return PG_OK;
}

extern "C" int pg_conv_stat_parts(const pg_conv_desc* d) {
    if (!d) return PG_ERR_INVALID;
    pg::ConvPlan pl;
    int rc = pg::conv_plan_build(d, &pl);
    if (rc != PG_OK) return rc;
    return pl.OS * pl.n_ntiles;
}

// Host-side view of the tensor-core tiling decisions for a layer (no GPU needed): lets the tests pin the plan
// (tile width, clips per tile, CTA pairs, merged clips, accumulator stages) for the BASELINE shapes.
extern "C" int pg_conv_tc_plan(const pg_conv_desc* d, int* out, int n_out) {
    PG_REQUIRE(d && out && n_out >= 16, "pg_conv_tc_plan: need a descriptor and room for 16 ints");
    pg::ConvPlan p;
    int rc = pg::conv_plan_build(d, &p);
    if (rc != PG_OK) return rc;
    const int v[16] = {p.n_tile, p.n_ntiles, p.nb, p.strip_rows, p.pair, p.merged, p.mgroups, p.acc_stages,
                       p.n_chunks, p.n_cotiles, p.OS, p.IS, p.n_taps[0], p.n_taps[1], p.n_groups[0], p.n_groups[1]};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    if (n_out >= 17) out[16] = p.whole_clip;
    return PG_OK;
}
