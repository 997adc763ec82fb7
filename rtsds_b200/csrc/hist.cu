// Confusion-matrix histogram for mIoU: utils.fast_hist (utils.py:52-58) and
// the torch.argmax feeding it (validation.py:51,120).
//
// HBM-bound integer kernel: 16 B/pixel (int64 label + int64 prediction), or
// 4*C + 8 B/pixel when fused with the argmax over fp32 NCHW logits.
// Per-warp privatised shared-memory histograms (uint32) with a per-thread
// run-length accumulator (segmentation maps are piecewise constant, so most
// consecutive pixels hit the same bin), flushed with one 64-bit global atomic
// per non-zero bin per block.  Integer arithmetic only: bit-exact.
#include "common.cuh"

namespace rtsds {

constexpr int HIST_THREADS = 256;
constexpr int HIST_WARPS = HIST_THREADS / 32;

struct RunAcc {
    int bin;
    unsigned cnt;
    __device__ __forceinline__ void push(int b, unsigned* sh) {
        if (b == bin) { ++cnt; return; }
        if (cnt) atomicAdd(&sh[bin], cnt);
        bin = b; cnt = 1;
    }
    __device__ __forceinline__ void flush(unsigned* sh) {
        if (cnt) atomicAdd(&sh[bin], cnt);
        cnt = 0;
    }
};

__device__ __forceinline__ void block_flush(unsigned* sh_all, int copies, int nbins,
                                            unsigned long long* hist) {
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        unsigned long long t = 0;
        for (int k = 0; k < copies; ++k) t += sh_all[k * nbins + b];
        if (t) atomicAdd(&hist[b], t);
    }
}

// label/pred int64; idx = n*a + b exactly as numpy computes it.
__global__ void __launch_bounds__(HIST_THREADS)
confusion_hist_kernel(const long long* __restrict__ label, const long long* __restrict__ pred,
                      long long n_pix, int n_cls, int copies, unsigned long long* hist,
                      unsigned long long* n_bad) {
    extern __shared__ unsigned sh_all[];
    const int nbins = n_cls * n_cls;
    for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh_all[i] = 0;
    __syncthreads();
    unsigned* sh = sh_all + ((threadIdx.x >> 5) % copies) * nbins;
    RunAcc acc{0, 0};
    unsigned bad = 0;

    const long long n2 = n_pix >> 1;   // pairs (both arrays are 8-byte typed; 16-byte loads need 16B alignment)
    const bool aligned = ((reinterpret_cast<uintptr_t>(label) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    auto one = [&](long long a, long long b) {
        if (a >= 0 && a < n_cls) {
            long long idx = static_cast<long long>(n_cls) * a + b;
            if (idx >= 0 && idx < nbins) acc.push(static_cast<int>(idx), sh);
            else ++bad;
        }
    };
    if (aligned) {
        const longlong2* l2 = reinterpret_cast<const longlong2*>(label);
        const longlong2* p2 = reinterpret_cast<const longlong2*>(pred);
        for (; i < n2; i += stride) {
            longlong2 a = __ldg(&l2[i]);
            longlong2 b = __ldg(&p2[i]);
            one(a.x, b.x);
            one(a.y, b.y);
        }
        if ((n_pix & 1) && blockIdx.x == 0 && threadIdx.x == 0) one(label[n_pix - 1], pred[n_pix - 1]);
    } else {
        for (; i < n_pix; i += stride) one(label[i], pred[i]);
    }
    acc.flush(sh);
    if (bad && n_bad) atomicAdd(n_bad, static_cast<unsigned long long>(bad));
    block_flush(sh_all, copies, nbins, hist);
}

// logits fp32 [n, C, hw]; each thread owns 4 consecutive pixels of one image.
template <int VEC>
__global__ void __launch_bounds__(HIST_THREADS)
argmax_hist_kernel(const float* __restrict__ logits, const long long* __restrict__ label, int n,
                   int n_cls, long long hw, int copies, long long* __restrict__ pred_out,
                   uint8_t* __restrict__ pred_u8, unsigned long long* hist) {
    extern __shared__ unsigned sh_all[];
    const int nbins = n_cls * n_cls;
    for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh_all[i] = 0;
    __syncthreads();
    unsigned* sh = sh_all + ((threadIdx.x >> 5) % copies) * nbins;
    RunAcc acc{0, 0};

    const long long groups_per_img = (hw + VEC - 1) / VEC;
    const long long total = groups_per_img * n;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
         g += stride) {
        const long long img = g / groups_per_img;
        const long long p0 = (g - img * groups_per_img) * VEC;
        const float* base = logits + img * n_cls * hw + p0;
        float best[VEC];
        int arg[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { best[v] = -INFINITY; arg[v] = 0; }
        const bool full = (p0 + VEC <= hw);
        // The class planes are hw floats apart: the loads of one pixel group are independent, so they are issued eight
        // planes at a time BEFORE the compare chain consumes them (one load per iteration of a runtime-bound loop left a
        // single 16-byte request in flight per thread: 2.3 TB/s; eight in flight: the DRAM pipe is the limit).
        for (int c0 = 0; c0 < n_cls; c0 += 8) {
            float x[8][VEC];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = c0 + k;
                if (VEC == 4 && full) {
                    float4 t = c < n_cls ? __ldcs(reinterpret_cast<const float4*>(base + c * hw)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    x[k][0] = t.x; x[k][1 % VEC] = t.y; x[k][2 % VEC] = t.z; x[k][3 % VEC] = t.w;
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[k][v] = (c < n_cls && p0 + v < hw) ? base[c * hw + v] : -INFINITY;
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (c0 + k < n_cls) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        // first maximum wins (torch.argmax); NaN is treated as maximal like torch
                        if (x[k][v] > best[v] || (x[k][v] != x[k][v] && best[v] == best[v])) { best[v] = x[k][v]; arg[v] = c0 + k; }
                    }
                }
            }
        }
        if (pred_u8) {                          // class map as one byte per pixel (n_cls <= 256): 8x fewer bytes back to the host
            if (VEC == 4 && full) {
                *reinterpret_cast<uint32_t*>(pred_u8 + img * hw + p0) =
                    static_cast<uint32_t>(arg[0]) | (static_cast<uint32_t>(arg[1 % VEC]) << 8) |
                    (static_cast<uint32_t>(arg[2 % VEC]) << 16) | (static_cast<uint32_t>(arg[3 % VEC]) << 24);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (p0 + v < hw) pred_u8[img * hw + p0 + v] = static_cast<uint8_t>(arg[v]);
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (p0 + v < hw) {
                const long long pix = img * hw + p0 + v;
                if (pred_out) pred_out[pix] = arg[v];
                if (label) {
                    long long a = __ldg(&label[pix]);
                    if (a >= 0 && a < n_cls) acc.push(static_cast<int>(a) * n_cls + arg[v], sh);
                }
            }
        }
    }
    acc.flush(sh);
    if (hist) block_flush(sh_all, copies, nbins, hist);
}

static int hist_copies(int n_cls) {
    int nbins = n_cls * n_cls;
    int c = (40 * 1024) / (nbins * 4);
    if (c > HIST_WARPS) c = HIST_WARPS;
    return c;
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_confusion_hist(const int64_t* label, const int64_t* pred, int64_t n_pix,
                                    int n_cls, int64_t* hist, int64_t* n_bad, rtsds_stream_t s) {
    RTSDS_REQUIRE(n_cls > 0 && n_cls <= 100, "confusion_hist: n_cls=%d out of range (1..100)", n_cls);
    RTSDS_REQUIRE(n_pix >= 0 && n_pix < (1ll << 40), "confusion_hist: n_pix out of range");
    RTSDS_REQUIRE(hist != nullptr, "confusion_hist: hist is NULL");
    if (n_pix == 0) return RTSDS_OK;
    RTSDS_REQUIRE(label && pred, "confusion_hist: NULL input");
    const int copies = hist_copies(n_cls);
    RTSDS_REQUIRE(copies >= 1, "confusion_hist: n_cls too large for shared memory");
    long long want = cdiv(n_pix, 2LL * HIST_THREADS * 8);
    int grid = static_cast<int>(want < 1 ? 1 : (want > 4LL * num_sms() ? 4LL * num_sms() : want));
    size_t smem = static_cast<size_t>(copies) * n_cls * n_cls * 4;
    confusion_hist_kernel<<<grid, HIST_THREADS, smem, as_stream(s)>>>(
        reinterpret_cast<const long long*>(label), reinterpret_cast<const long long*>(pred), n_pix,
        n_cls, copies, reinterpret_cast<unsigned long long*>(hist),
        reinterpret_cast<unsigned long long*>(n_bad));
    count_launch();
    return check_launch("confusion_hist");
}

static int argmax_hist_impl(const float* logits, const int64_t* label, int n, int n_cls, int64_t hw, int64_t* pred_out,
                            uint8_t* pred_u8, int64_t* hist, rtsds_stream_t s) {
    RTSDS_REQUIRE(n_cls > 0 && n_cls <= 100, "argmax_hist: n_cls=%d out of range (1..100)", n_cls);
    RTSDS_REQUIRE(n >= 0 && hw >= 0, "argmax_hist: negative size");
    RTSDS_REQUIRE((label == nullptr) == (hist == nullptr), "argmax_hist: label and hist go together");
    if (n == 0 || hw == 0) return RTSDS_OK;
    RTSDS_REQUIRE(logits, "argmax_hist: NULL logits");
    const int copies = hist_copies(n_cls);
    size_t smem = static_cast<size_t>(copies) * n_cls * n_cls * 4;
    const bool vec4 = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
    long long groups = (vec4 ? cdiv(hw, 4) : hw) * n;
    long long want = cdiv(groups, HIST_THREADS);
    int grid = static_cast<int>(want > 8LL * num_sms() ? 8LL * num_sms() : want);
    if (vec4)
        argmax_hist_kernel<4><<<grid, HIST_THREADS, smem, as_stream(s)>>>(
            logits, reinterpret_cast<const long long*>(label), n, n_cls, hw, copies,
            reinterpret_cast<long long*>(pred_out), pred_u8, reinterpret_cast<unsigned long long*>(hist));
    else
        argmax_hist_kernel<1><<<grid, HIST_THREADS, smem, as_stream(s)>>>(
            logits, reinterpret_cast<const long long*>(label), n, n_cls, hw, copies,
            reinterpret_cast<long long*>(pred_out), pred_u8, reinterpret_cast<unsigned long long*>(hist));
    count_launch();
    return check_launch("argmax_hist");
}

extern "C" int rtsds_argmax_hist(const float* logits, const int64_t* label, int n, int n_cls,
                                 int64_t hw, int64_t* pred_out, int64_t* hist, rtsds_stream_t s) {
    return argmax_hist_impl(logits, label, n, n_cls, hw, pred_out, nullptr, hist, s);
}

extern "C" int rtsds_argmax_hist_u8(const float* logits, const int64_t* label, int n, int n_cls,
                                    int64_t hw, uint8_t* pred_u8_out, int64_t* hist, rtsds_stream_t s) {
    RTSDS_REQUIRE(pred_u8_out && (hw % 4 != 0 || (reinterpret_cast<uintptr_t>(pred_u8_out) & 3) == 0), "argmax_hist_u8: pred_u8_out NULL or misaligned");
    return argmax_hist_impl(logits, label, n, n_cls, hw, nullptr, pred_u8_out, hist, s);
}
