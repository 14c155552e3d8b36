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
extern "C" int pg_abi_version(void) { return 1; }

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
