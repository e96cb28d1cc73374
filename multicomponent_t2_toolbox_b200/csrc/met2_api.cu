// met2_api.cu — error state, launch accounting and diagnostics of the C ABI (include/met2.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "met2_host.h"

namespace met2 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(MET2_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return MET2_OK;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace met2

extern "C" const char* met2_last_error(void) { return met2::g_err; }
extern "C" int met2_version(void) { return MET2_VERSION; }
extern "C" int64_t met2_launch_count(void) { return (int64_t)met2::g_launches.load(); }
