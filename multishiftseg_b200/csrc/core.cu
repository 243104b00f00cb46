// Library-wide plumbing: thread-local error string, launch counter, device properties.
#include <stdarg.h>

#include "common.cuh"

namespace mss {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace mss

extern "C" int mss_abi_version(void) { return MSS_ABI_VERSION; }
extern "C" const char *mss_last_error(void) { return mss::t_err; }
extern "C" int64_t mss_launch_count(void) { return mss::g_launches.load(); }
