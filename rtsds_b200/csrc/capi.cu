// Error plumbing, device check and launch accounting for librtsds_b200.
#include "common.cuh"
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>

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

// ---- deterministic mode (common.cuh: det_add) ------------------------------------------------------------------------
static std::atomic<int> g_det{-1};
bool det_mode() {
    int v = g_det.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("RTSDS_DETERMINISTIC");
        v = (e && e[0] == '1') ? 1 : 0;
        g_det.store(v, std::memory_order_relaxed);
    }
    return v == 1;
}

struct DetRegion { unsigned long long* p = nullptr; size_t n = 0; };
static std::mutex g_det_mu;
static std::map<std::pair<int, cudaStream_t>, DetRegion> g_det_regions;

unsigned long long* det_scratch(cudaStream_t st, size_t n) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_det_mu);
    DetRegion& r = g_det_regions[std::make_pair(dev, st)];
    if (r.n < n) {
        // grow: cudaFree waits for the work that still reads the old region
        if (r.p) cudaFree(r.p);
        size_t want = n < (size_t(1) << 20) ? (size_t(1) << 20) : n + n / 4;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&r.p), want * 16);
        if (e != cudaSuccess) {
            r.p = nullptr; r.n = 0;
            set_error("deterministic mode: cudaMalloc of %zu bytes failed: %s", want * 16, cudaGetErrorString(e));
            return nullptr;
        }
        r.n = want;
    }
    if (cudaMemsetAsync(r.p, 0, n * 16, st) != cudaSuccess) {
        set_error("deterministic mode: cudaMemsetAsync failed");
        return nullptr;
    }
    return r.p;
}

__global__ void det_finish_kernel(const unsigned long long* acc, float* dst, size_t n, int accumulate) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = static_cast<float>(det_value(acc + 2 * i));
    dst[i] = accumulate ? dst[i] + v : v;
}

int det_finish(const unsigned long long* acc, float* dst, size_t n, bool accumulate, cudaStream_t st) {
    if (n == 0) return RTSDS_OK;
    det_finish_kernel<<<static_cast<unsigned>(cdiv(static_cast<int64_t>(n), 256)), 256, 0, st>>>(acc, dst, n, accumulate ? 1 : 0);
    count_launch();
    return check_launch("det_finish_kernel");
}

}  // namespace rtsds

extern "C" {

void rtsds_set_deterministic(int on) { rtsds::g_det.store(on ? 1 : 0, std::memory_order_relaxed); }
int rtsds_get_deterministic(void) { return rtsds::det_mode() ? 1 : 0; }

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
