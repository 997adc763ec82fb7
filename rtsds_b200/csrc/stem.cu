// Direct (CUDA-core) convolutions for the layers whose input channel count is
// too small for a tensor-core K tile: they read the API-boundary NCHW fp32
// tensor and write NHWC.
//   * Spatial path ConvBlock1: conv3x3 s2 p1 3->64  (models/bisenet/build_bisenet.py:24)
//   * Context path / DeepLab stem: conv7x7 s2 p3 3->64 (torchvision resnet conv1 via
//     models/bisenet/build_contextpath.py:19; models/deeplabv2/deeplabv2.py:73)
//   * Discriminator conv1: conv4x4 s2 p1 19->64 + bias + LeakyReLU, with the
//     F.softmax(dim=1) of train.py:225,245,256 fused into the loader
//     (models/domain_shift/adversarial/model.py:45,72).
// These are bandwidth/FFMA-bound (K = 27 / 147 / 304): one output pixel per
// thread, all 64 output channels in registers, the input patch and the
// [ci][tap][co] weight slice in shared memory (weights read as 16-byte
// broadcasts), coalesced row loads from NCHW, 128/256-byte contiguous NHWC
// stores per thread.  Also: MaxPool2d(3,2,1) on NHWC.
#include "common.cuh"

namespace rtsds {

constexpr int ST_TW = 32, ST_TH = 8, ST_THREADS = ST_TW * ST_TH, ST_COUT = 64, ST_CCH = 4;

struct StemParams {
    int n, cin, h, w, oh, ow, pad;
    int act;
    float slope;
    int softmax_in;
    int out_dtype;
};

__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            float send = upper ? v[i] : v[i + s];
            float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

template <int K, int S>
__global__ void __launch_bounds__(ST_THREADS)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ scale,
                 const float* __restrict__ shift, float* stats, unsigned long long* det, void* y, StemParams p) {
    constexpr int PH = (ST_TH - 1) * S + K, PW = (ST_TW - 1) * S + K;
    constexpr int PWP = PW | 1;     // odd pitch: stride-S column reads spread over banks
    extern __shared__ float smem[];
    float* s_patch = smem;                                  // [ST_CCH][PH][PWP]
    float* s_w = s_patch + ST_CCH * PH * PWP;               // [ST_CCH][K*K][64]
    float* s_max = s_w + ST_CCH * K * K * ST_COUT;          // [PH*PW] (softmax_in)
    float* s_inv = s_max + PH * PW;
    __shared__ float s_stat[2 * ST_COUT];

    const int tid = threadIdx.x;
    const int lx = tid % ST_TW, ly = tid / ST_TW;
    const int img = blockIdx.z;
    const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
    const int iy0 = oy0 * S - p.pad, ix0 = ox0 * S - p.pad;
    const long long plane = static_cast<long long>(p.h) * p.w;
    const float* ximg = x + static_cast<long long>(img) * p.cin * plane;

    if (tid < 2 * ST_COUT) s_stat[tid] = 0.f;

    if (p.softmax_in) {
        for (int i = tid; i < PH * PW; i += ST_THREADS) {
            const int py = i / PW, px = i - py * PW;
            const int iy = iy0 + py, ix = ix0 + px;
            float m = -INFINITY, sum = 0.f;
            if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
                const float* xp = ximg + static_cast<long long>(iy) * p.w + ix;
                for (int c = 0; c < p.cin; ++c) m = fmaxf(m, __ldg(xp + c * plane));
                for (int c = 0; c < p.cin; ++c) sum += expf(__ldg(xp + c * plane) - m);
            }
            s_max[i] = m;
            s_inv[i] = sum > 0.f ? 1.0f / sum : 0.f;
        }
    }

    float acc[ST_COUT];
#pragma unroll
    for (int i = 0; i < ST_COUT; ++i) acc[i] = 0.f;

    for (int c0 = 0; c0 < p.cin; c0 += ST_CCH) {
        const int nc = min(ST_CCH, p.cin - c0);
        __syncthreads();
        // patch: coalesced along W
        for (int i = tid; i < nc * PH * PW; i += ST_THREADS) {
            const int c = i / (PH * PW);
            const int r = i - c * PH * PW;
            const int py = r / PW, px = r - py * PW;
            const int iy = iy0 + py, ix = ix0 + px;
            float v = 0.f;
            if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
                v = __ldg(ximg + (c0 + c) * plane + static_cast<long long>(iy) * p.w + ix);
                if (p.softmax_in) v = expf(v - s_max[r]) * s_inv[r];
            }
            s_patch[(c * PH + py) * PWP + px] = v;
        }
        // weights OIHW -> [c][tap][co]
        for (int i = tid; i < nc * K * K * ST_COUT; i += ST_THREADS) {
            const int co = i % ST_COUT;
            const int r = i / ST_COUT;
            const int t = r % (K * K), c = r / (K * K);
            s_w[i] = __ldg(wgt + (static_cast<long long>(co) * p.cin + c0 + c) * (K * K) + t);
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
#pragma unroll 1
            for (int r = 0; r < K; ++r) {
                const float* prow = s_patch + (c * PH + ly * S + r) * PWP + lx * S;
                const float4* wrow = reinterpret_cast<const float4*>(s_w + (c * K * K + r * K) * ST_COUT);
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const float a = prow[q];
#pragma unroll
                    for (int g = 0; g < ST_COUT / 4; ++g) {
                        const float4 w4 = wrow[q * (ST_COUT / 4) + g];
                        acc[g * 4 + 0] = fmaf(a, w4.x, acc[g * 4 + 0]);
                        acc[g * 4 + 1] = fmaf(a, w4.y, acc[g * 4 + 1]);
                        acc[g * 4 + 2] = fmaf(a, w4.z, acc[g * 4 + 2]);
                        acc[g * 4 + 3] = fmaf(a, w4.w, acc[g * 4 + 3]);
                    }
                }
            }
        }
    }

    const int oy = oy0 + ly, ox = ox0 + lx;
    const bool valid = oy < p.oh && ox < p.ow;
    if (stats) {
        const int lane = tid & 31;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = valid ? acc[half * 32 + j] : 0.f;
            float s1 = warp_transpose_sum32(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = valid ? acc[half * 32 + j] * acc[half * 32 + j] : 0.f;
            float s2 = warp_transpose_sum32(t, lane);
            if (det) {            // deterministic mode: exact accumulators (common.cuh) in the layout of stats
                det_add(det + 2 * (half * 32 + lane), s1);
                det_add(det + 2 * (ST_COUT + half * 32 + lane), s2);
            } else {
                atomicAdd(&s_stat[half * 32 + lane], s1);
                atomicAdd(&s_stat[ST_COUT + half * 32 + lane], s2);
            }
        }
    }
    if (valid) {
        const long long pix = (static_cast<long long>(img) * p.oh + oy) * p.ow + ox;
#pragma unroll
        for (int j = 0; j < ST_COUT; ++j) {
            float v = acc[j] * (scale ? __ldg(scale + j) : 1.f) + (shift ? __ldg(shift + j) : 0.f);
            acc[j] = apply_act(v, p.act, p.slope);
        }
        if (p.out_dtype == RTSDS_BF16) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + pix * ST_COUT);
#pragma unroll
            for (int g = 0; g < ST_COUT / 8; ++g) {
                uint4 o;
                o.x = pack_bf16x2(acc[g * 8 + 0], acc[g * 8 + 1]);
                o.y = pack_bf16x2(acc[g * 8 + 2], acc[g * 8 + 3]);
                o.z = pack_bf16x2(acc[g * 8 + 4], acc[g * 8 + 5]);
                o.w = pack_bf16x2(acc[g * 8 + 6], acc[g * 8 + 7]);
                dst[g] = o;
            }
        } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + pix * ST_COUT);
#pragma unroll
            for (int g = 0; g < ST_COUT / 4; ++g)
                dst[g] = make_float4(acc[g * 4], acc[g * 4 + 1], acc[g * 4 + 2], acc[g * 4 + 3]);
        }
    }
    if (stats && !det) {
        __syncthreads();
        if (tid < 2 * ST_COUT) atomicAdd(&stats[tid], s_stat[tid]);
    }
}

template <int K, int S>
static int launch_stem(const float* x, const float* w, const float* scale, const float* shift, float* stats,
                       void* y, const StemParams& p, cudaStream_t st) {
    constexpr int PH = (ST_TH - 1) * S + K, PW = (ST_TW - 1) * S + K, PWP = PW | 1;
    size_t smem = sizeof(float) * (ST_CCH * PH * PWP + ST_CCH * K * K * ST_COUT + 2 * PH * PW);
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) { set_error("stem_conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
        done = true;
    }
    dim3 grid(static_cast<unsigned>(cdiv(p.ow, ST_TW)), static_cast<unsigned>(cdiv(p.oh, ST_TH)), static_cast<unsigned>(p.n));
    unsigned long long* det = nullptr;
    if (stats && det_mode()) {
        det = det_scratch(st, 2 * ST_COUT);
        if (!det) return RTSDS_ECUDA;
    }
    stem_conv_kernel<K, S><<<grid, ST_THREADS, smem, st>>>(x, w, scale, shift, stats, det, y, p);
    count_launch();
    int rc = check_launch("stem_conv_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, stats, 2 * ST_COUT, true, st);
    return rc;
}

// ---- MaxPool2d(kernel 3, stride 2, pad 1[, ceil_mode]) on NHWC --------------------
// one thread = one output pixel x 8 channels; 16-byte (bf16) / 2x16-byte (fp32) vector loads.
struct Vec8 { float v[8]; };
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    Vec8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ Vec8 load8(const __half* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    Vec8 r;
    float2 a = unpack_f16x2(u.x), b = unpack_f16x2(u.y), c = unpack_f16x2(u.z), d = unpack_f16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ void store8(__half* p, const Vec8& r) {
    uint4 u;
    u.x = pack_f16x2(r.v[0], r.v[1]); u.y = pack_f16x2(r.v[2], r.v[3]);
    u.z = pack_f16x2(r.v[4], r.v[5]); u.w = pack_f16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ Vec8 load8(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    Vec8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const Vec8& r) {
    reinterpret_cast<float4*>(p)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool_kernel(const T* __restrict__ x, int n, int h, int w, int c, int x_ld, int oh, int ow, T* __restrict__ y, Div3 dv) {
    pdl_wait();
    const int cg = c / 8;
    const long long total = static_cast<long long>(n) * oh * ow * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned ug, ux, uy;
        unsigned r = fdivmod(static_cast<unsigned>(i), dv.a, &ug);
        r = fdivmod(r, dv.b, &ux);
        const int img = static_cast<int>(fdivmod(r, dv.c, &uy));
        const int g = static_cast<int>(ug), ox = static_cast<int>(ux), oy = static_cast<int>(uy);
        Vec8 m;
#pragma unroll
        for (int j = 0; j < 8; ++j) m.v[j] = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= w) continue;
                const Vec8 t = load8(x + ((static_cast<long long>(img) * h + iy) * w + ix) * x_ld + g * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) m.v[j] = fmaxf(m.v[j], t.v[j]);
            }
        }
        store8(y + ((static_cast<long long>(img) * oh + oy) * ow + ox) * c + g * 8, m);
    }
}

// Training form: also records, per 8-channel group, the window position (r*3+q, one nibble per channel) of the FIRST
// maximum in ATen's row-major scan (a later NaN overrides) -- the backward pass then never touches x.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_idx_kernel(const T* __restrict__ x, int n, int h, int w, int c, int oh, int ow, T* __restrict__ y,
                   uint32_t* __restrict__ idx, Div3 dv) {
    const int cg = c / 8;
    const long long total = static_cast<long long>(n) * oh * ow * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned ug, ux, uy;
        unsigned r = fdivmod(static_cast<unsigned>(i), dv.a, &ug);
        r = fdivmod(r, dv.b, &ux);
        const int img = static_cast<int>(fdivmod(r, dv.c, &uy));
        const int g = static_cast<int>(ug), ox = static_cast<int>(ux), oy = static_cast<int>(uy);
        Vec8 m;
        int pos[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m.v[j] = -INFINITY; pos[j] = 15; }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= w) continue;
                const Vec8 t = load8(x + ((static_cast<long long>(img) * h + iy) * w + ix) * c + g * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (t.v[j] > m.v[j] || pos[j] == 15 || t.v[j] != t.v[j]) { m.v[j] = t.v[j]; pos[j] = dy * 3 + dx; }
            }
        }
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) packed |= static_cast<uint32_t>(pos[j]) << (4 * j);
        store8(y + ((static_cast<long long>(img) * oh + oy) * ow + ox) * c + g * 8, m);
        idx[i] = packed;
    }
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_stem_conv_fwd(const float* x, const float* w_oihw, int n, int cin, int h, int w, int cout,
                                   int k, int stride, int pad, const float* scale, const float* shift, int act,
                                   float slope, int softmax_in, float* stats, int out_dtype, void* y,
                                   rtsds_stream_t s) {
    RTSDS_REQUIRE(x && w_oihw && y, "stem_conv_fwd: NULL argument");
    RTSDS_REQUIRE(cout == ST_COUT, "stem_conv_fwd: cout=%d (only 64 supported)", cout);
    RTSDS_REQUIRE(cin >= 1 && cin <= 32, "stem_conv_fwd: cin=%d out of range 1..32", cin);
    RTSDS_REQUIRE(n > 0 && h > 0 && w > 0, "stem_conv_fwd: empty tensor");
    RTSDS_REQUIRE(out_dtype == RTSDS_BF16 || out_dtype == RTSDS_F32, "stem_conv_fwd: bad out_dtype");
    RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "stem_conv_fwd: y must be 16-byte aligned");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    StemParams p;
    p.n = n; p.cin = cin; p.h = h; p.w = w; p.pad = pad;
    p.oh = (h + 2 * pad - k) / stride + 1;
    p.ow = (w + 2 * pad - k) / stride + 1;
    RTSDS_REQUIRE(p.oh > 0 && p.ow > 0, "stem_conv_fwd: empty output");
    p.act = act; p.slope = slope; p.softmax_in = softmax_in; p.out_dtype = out_dtype;
    cudaStream_t st = as_stream(s);
    if (k == 3 && stride == 2) return launch_stem<3, 2>(x, w_oihw, scale, shift, stats, y, p, st);
    if (k == 4 && stride == 2) return launch_stem<4, 2>(x, w_oihw, scale, shift, stats, y, p, st);
    if (k == 7 && stride == 2) return launch_stem<7, 2>(x, w_oihw, scale, shift, stats, y, p, st);
    set_error("stem_conv_fwd: k=%d stride=%d unsupported (3/4/7 with stride 2)", k, stride);
    return RTSDS_EUNSUP;
}

static int maxpool_fwd_impl(const void* x, int n, int h, int w, int c, int dtype, int ceil_mode, void* y, uint32_t* idx,
                            rtsds_stream_t s, int x_ld = 0) {
    if (x_ld == 0) x_ld = c;
    RTSDS_REQUIRE(x_ld >= c && x_ld % 8 == 0 && (idx == nullptr || x_ld == c), "maxpool: input pitch must be >= c and a multiple of 8");
    RTSDS_REQUIRE(x && y, "maxpool: NULL argument");
    RTSDS_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool: bad shape (c must be a multiple of 8)");
    RTSDS_REQUIRE(dtype == RTSDS_BF16 || dtype == RTSDS_F32 || dtype == RTSDS_F16, "maxpool: bad dtype");
    RTSDS_REQUIRE(!(idx && dtype == RTSDS_F16), "maxpool: the training form (idx) is bf16 / fp32 only");
    auto osz = [&](int in) {
        int o = ceil_mode ? (in + 2 - 3 + 1) / 2 + 1 : (in + 2 - 3) / 2 + 1;
        if (ceil_mode && (o - 1) * 2 >= in + 1) --o;   // last window must start inside input+left pad (torch rule)
        return o;
    };
    const int oh = osz(h), ow = osz(w);
    const long long total = static_cast<long long>(n) * oh * ow * (c / 8);
    RTSDS_REQUIRE(total < (1LL << 31), "maxpool: tensor too large");
    int grid = static_cast<int>(cdiv(total, 256) > 16LL * num_sms() ? 16LL * num_sms() : cdiv(total, 256));
    const Div3 dv = {make_fastdiv(c / 8), make_fastdiv(ow), make_fastdiv(oh)};
    if (idx) {
        if (dtype == RTSDS_BF16)
            maxpool_idx_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c, oh,
                                                                              ow, reinterpret_cast<__nv_bfloat16*>(y), idx, dv);
        else
            maxpool_idx_kernel<float><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const float*>(x), n, h, w, c, oh, ow,
                                                                      reinterpret_cast<float*>(y), idx, dv);
        count_launch();
        return check_launch("maxpool_idx_kernel");
    }
    if (dtype == RTSDS_F16)
        launch_pdl(maxpool_kernel<__half>, dim3(grid), dim3(256), 0, as_stream(s), reinterpret_cast<const __half*>(x), n, h, w, c, x_ld, oh, ow,
                   reinterpret_cast<__half*>(y), dv);
    else if (dtype == RTSDS_BF16)
        launch_pdl(maxpool_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(s), reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c,
                   x_ld, oh, ow, reinterpret_cast<__nv_bfloat16*>(y), dv);
    else
        launch_pdl(maxpool_kernel<float>, dim3(grid), dim3(256), 0, as_stream(s), reinterpret_cast<const float*>(x), n, h, w, c, x_ld, oh, ow,
                   reinterpret_cast<float*>(y), dv);
    count_launch();
    return check_launch("maxpool_kernel");
}

extern "C" int rtsds_maxpool3x3s2_fwd(const void* x, int n, int h, int w, int c, int dtype, int ceil_mode,
                                      void* y, rtsds_stream_t s) {
    return maxpool_fwd_impl(x, n, h, w, c, dtype, ceil_mode, y, nullptr, s);
}

// same, reading an input whose pixels are x_ld elements apart (a channel slice of a wider buffer: the fused stems write
// the context-path and spatial-path maps side by side)
extern "C" int rtsds_maxpool3x3s2_fwd_ld(const void* x, int n, int h, int w, int c, int x_ld, int dtype, int ceil_mode,
                                         void* y, rtsds_stream_t s) {
    return maxpool_fwd_impl(x, n, h, w, c, dtype, ceil_mode, y, nullptr, s, x_ld);
}

extern "C" int rtsds_maxpool3x3s2_fwd_idx(const void* x, int n, int h, int w, int c, int dtype, int ceil_mode,
                                          void* y, uint32_t* idx, rtsds_stream_t s) {
    RTSDS_REQUIRE(idx, "maxpool_fwd_idx: NULL idx");
    return maxpool_fwd_impl(x, n, h, w, c, dtype, ceil_mode, y, idx, s);
}


// ---- space-to-depth stems (see conv_tc.cu: rtsds_stem_s2d_conv_*) -------------------------------------------------
namespace rtsds {

// P[n, i, j, (py*2+px)*3 + c] = a_c * x[n, c, 2(i-2)+py, 2(j-2)+px] + b_c inside the image, 0 outside (the conv's zero
// padding applies to the NORMALISED image); channels 12..15 = 0; one thread per (n, i, j).  TI: float image or RAW uint8
// frame (the uint8 -> float conversion and transforms.Normalize cost nothing extra here); TO: bf16 or fp16.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
stem_s2d_pack_kernel(const TI* __restrict__ x, int n, int h, int w, int hp, int wp, float a0, float a1, float a2, float b0, float b1,
                     float b2, TO* __restrict__ P) {
    const long long total = static_cast<long long>(n) * hp * wp;
    const long long plane = static_cast<long long>(h) * w;
    const float av[3] = {a0, a1, a2}, bv[3] = {b0, b1, b2};
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(idx % wp);
        const long long r = idx / wp;
        const int i = static_cast<int>(r % hp);
        const int img = static_cast<int>(r / hp);
        const int y0 = 2 * (i - 2), x0 = 2 * (j - 2);
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0.f;
        const TI* xi = x + static_cast<long long>(img) * 3 * plane;
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            const int yy = y0 + py;
            if (yy < 0 || yy >= h) continue;
            const bool a = x0 >= 0 && x0 < w, b = x0 + 1 >= 0 && x0 + 1 < w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const TI* row = xi + c * plane + static_cast<long long>(yy) * w;
                if (a) v[(py * 2 + 0) * 3 + c] = fmaf(static_cast<float>(__ldg(row + x0)), av[c], bv[c]);
                if (b) v[(py * 2 + 1) * 3 + c] = fmaf(static_cast<float>(__ldg(row + x0 + 1)), av[c], bv[c]);
            }
        }
        uint4 lo, hi;
        lo.x = pack_x2<TO>(v[0], v[1]); lo.y = pack_x2<TO>(v[2], v[3]); lo.z = pack_x2<TO>(v[4], v[5]); lo.w = pack_x2<TO>(v[6], v[7]);
        hi.x = pack_x2<TO>(v[8], v[9]); hi.y = pack_x2<TO>(v[10], v[11]); hi.z = 0u; hi.w = 0u;
        uint4* dst = reinterpret_cast<uint4*>(P + idx * 16);
        dst[0] = lo; dst[1] = hi;
    }
}

// w[co][c][k][k] (k = 7, pad 3 or k = 3, pad 1; stride 2) -> virtual OIHW w2[co][64][4][1]:
//   w2[co, q*16 + (py*2+px)*3 + c, r] = w[co, c, ky, kx],  ky = 2r+py-1-(3-pad), kx = 2q+px-1-(3-pad)   (0 outside)
__global__ void stem_s2d_weight_kernel(const float* __restrict__ w, int cout, int k, int pad, float* __restrict__ w2, int grad,
                                       float* __restrict__ gw) {
    const int total = cout * 64 * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i & 3, ci = (i >> 2) & 63, co = i >> 8;
        const int q = ci >> 4, e = ci & 15;
        if (e >= 12) { if (!grad) w2[i] = 0.f; continue; }
        const int par = e / 3, c = e - par * 3, py = par >> 1, px = par & 1;
        const int ky = 2 * r + py - 1 - (3 - pad), kx = 2 * q + px - 1 - (3 - pad);
        const bool ok = ky >= 0 && ky < k && kx >= 0 && kx < k;
        if (grad) {
            if (ok) gw[((co * 3 + c) * k + ky) * k + kx] += w2[i];        // the map is one-to-one: no atomics needed
        } else {
            w2[i] = ok ? w[((co * 3 + c) * k + ky) * k + kx] : 0.f;
        }
    }
}

}  // namespace rtsds

extern "C" int rtsds_stem_s2d_pack_ex(const void* x, int x_is_u8, const float* scale3, const float* bias3, int n, int h, int w,
                                      int p_dtype, void* P, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && P && n > 0 && h > 0 && w > 0, "stem_s2d_pack: bad argument");
    RTSDS_REQUIRE(p_dtype == RTSDS_BF16 || p_dtype == RTSDS_F16, "stem_s2d_pack: P must be bf16 or fp16");
    const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
    const long long total = static_cast<long long>(n) * (oh + 3) * (ow + 3);
    long long g = rtsds::cdiv(total, 256);
    const long long cap = 16LL * rtsds::num_sms();
    if (g > cap) g = cap;
    const float a0 = scale3 ? scale3[0] : 1.f, a1 = scale3 ? scale3[1] : 1.f, a2 = scale3 ? scale3[2] : 1.f;
    const float b0 = bias3 ? bias3[0] : 0.f, b1 = bias3 ? bias3[1] : 0.f, b2 = bias3 ? bias3[2] : 0.f;
    cudaStream_t st = rtsds::as_stream(s);
    const int gi = static_cast<int>(g);
#define S2D_PACK(TI, TO) rtsds::stem_s2d_pack_kernel<TI, TO><<<gi, 256, 0, st>>>(reinterpret_cast<const TI*>(x), n, h, w, oh + 3, ow + 3, a0, a1, a2, b0, b1, b2, reinterpret_cast<TO*>(P))
    if (x_is_u8 && p_dtype == RTSDS_F16) S2D_PACK(uint8_t, __half);
    else if (x_is_u8) S2D_PACK(uint8_t, __nv_bfloat16);
    else if (p_dtype == RTSDS_F16) S2D_PACK(float, __half);
    else S2D_PACK(float, __nv_bfloat16);
#undef S2D_PACK
    rtsds::count_launch();
    return rtsds::check_launch("stem_s2d_pack_kernel");
}

extern "C" int rtsds_stem_s2d_pack(const float* x, int n, int h, int w, void* P, rtsds_stream_t s) {
    return rtsds_stem_s2d_pack_ex(x, 0, nullptr, nullptr, n, h, w, RTSDS_BF16, P, s);
}

extern "C" int rtsds_stem_s2d_weight(const float* w_oihw, int cout, int k, int pad, float* w2, rtsds_stream_t s) {
    RTSDS_REQUIRE(w_oihw && w2 && cout > 0 && ((k == 7 && pad == 3) || (k == 3 && pad == 1)), "stem_s2d_weight: 7x7 p3 or 3x3 p1 only");
    rtsds::stem_s2d_weight_kernel<<<static_cast<int>(rtsds::cdiv(cout * 256, 256)), 256, 0, rtsds::as_stream(s)>>>(w_oihw, cout, k, pad, w2, 0, nullptr);
    rtsds::count_launch();
    return rtsds::check_launch("stem_s2d_weight_kernel");
}

extern "C" int rtsds_stem_s2d_weight_grad(const float* g2, int cout, int k, int pad, float* grad_oihw, rtsds_stream_t s) {
    RTSDS_REQUIRE(g2 && grad_oihw && cout > 0 && ((k == 7 && pad == 3) || (k == 3 && pad == 1)), "stem_s2d_weight_grad: 7x7 p3 or 3x3 p1 only");
    rtsds::stem_s2d_weight_kernel<<<static_cast<int>(rtsds::cdiv(cout * 256, 256)), 256, 0, rtsds::as_stream(s)>>>(nullptr, cout, k, pad, const_cast<float*>(g2), 1, grad_oihw);
    rtsds::count_launch();
    return rtsds::check_launch("stem_s2d_weight_kernel");
}
