// Error plumbing, device check and launch accounting for librtsds_b200.
#include "common.cuh"
#include <atomic>
#include <cstring>

namespace rtsds {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return RTSDS_ECUDA;
    }
    return RTSDS_OK;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

}  // namespace rtsds

extern "C" {

int rtsds_abi_version(void) { return RTSDS_ABI_VERSION; }

const char* rtsds_last_error_string(void) { return rtsds::g_err; }

int64_t rtsds_launch_count(void) { return rtsds::g_launches.load(std::memory_order_relaxed); }

int rtsds_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        rtsds::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return RTSDS_ECUDA;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        rtsds::set_error("device %d is sm_%d%d; librtsds_b200 is built for sm_100a only", dev, major,
                         minor);
        return RTSDS_EARCH;
    }
    return RTSDS_OK;
}

}  // extern "C"
