// "Taps as N": a k x k stride-1 convolution with FEW output channels (the FeatureFusionModule's 3x3 1024 -> 19 conv,
// models/bisenet/build_bisenet.py:64,74) restated so that the wide input is read ONCE instead of once per filter tap.
//
//   forward   T[p, t*C + co] = sum_ci x[p, ci] * W[co, ci, t]                    one 1x1 GEMM, N = taps*C (171)
//             y[p, co]       = sum_t T[p + off(t), t*C + co]                     tapn_gather (this file), zero outside
//   backward  G[q, t*C + co] = dy[q - off(t), co]                                tapn_scatter (this file)
//             dx[q, ci]      = sum_{t,co} G[q, t*C+co] * W[co, ci, t]            one 1x1 GEMM, K = taps*C
//             dW[co, ci, t]  = sum_q G[q, t*C+co] * x[q, ci]                     one 1x1 wgrad GEMM
// The GEMMs are the tcgen05 kernels of conv_tc.cu on "virtual" 1x1 weights built by tapn_weights; the implicit-GEMM
// form re-read the 1024-channel input 9 times through L2 (189 MB L2->SM per 512x1024 frame for a 2.9 GFLOP layer).
// off(t) = (r*dil - pad, s*dil - pad) for t = r*k + s.
#include "common.cuh"

namespace rtsds {

// w[co][ci][taps] -> w_fwd[(t*C+co)][ci] (virtual OIHW [taps*C, cin, 1, 1]) and w_bwd[ci][(t*C+co) padded to kpad]
// (virtual OIHW [cin, kpad, 1, 1]); either output may be NULL.
__global__ void tapn_weights_kernel(const float* __restrict__ w, int c, int cin, int taps, int kpad,
                                    float* __restrict__ w_fwd, float* __restrict__ w_bwd) {
    const long long total = static_cast<long long>(cin) * kpad;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int n = static_cast<int>(i % kpad);
        const int ci = static_cast<int>(i / kpad);
        float v = 0.f;
        if (n < taps * c) {
            const int t = n / c, co = n - t * c;
            v = w[(static_cast<long long>(co) * cin + ci) * taps + t];
            if (w_fwd) w_fwd[static_cast<long long>(n) * cin + ci] = v;
        }
        if (w_bwd) w_bwd[i] = v;
    }
}

// grad[co][ci][t] += dw2[(t*C+co)][ci]   (dw2: the 1x1 wgrad result in OIHW form [taps*C, cin])
__global__ void tapn_weight_grad_kernel(const float* __restrict__ dw2, int c, int cin, int taps, float* __restrict__ grad) {
    const long long total = static_cast<long long>(c) * cin * taps;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % taps);
        const long long r = i / taps;
        const int ci = static_cast<int>(r % cin);
        const int co = static_cast<int>(r / cin);
        grad[i] += dw2[(static_cast<long long>(t) * c + co) * cin + ci];
    }
}

// y[p, co] = act(scale*sum_t T[p+off(t), t*C+co] + shift); optional per-channel sum / sum of squares of the raw value.
// One warp per output pixel (lane = channel, C <= 32): each (pixel, tap) read is C consecutive floats; the K*K loads
// of a pixel are independent and issued together.
template <int K>
__global__ void __launch_bounds__(256)
tapn_gather_kernel(const float* __restrict__ t_buf, int t_ld, int n, int h, int w, int c, int k_rt, int pad, int dil,
                   const float* __restrict__ scale, const float* __restrict__ shift, int act, float* stats,
                   unsigned long long* det, float* __restrict__ y, int y_ld, float* gap_out, float gap_scale) {
    __shared__ float s_stats[2][8][32];
    pdl_wait();
    const int k = K > 0 ? K : k_rt;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    // blockIdx.y = image: the fused global average pool (gap_out) sums per image
    const long long pix0 = static_cast<long long>(blockIdx.y) * h * w;
    const long long npix = pix0 + static_cast<long long>(h) * w;
    float gs = 0.f;
    const float sc = (scale && lane < c) ? scale[lane] : 1.f, sh = (shift && lane < c) ? shift[lane] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    for (long long p = pix0 + static_cast<long long>(blockIdx.x) * 8 + wrp; p < npix; p += static_cast<long long>(gridDim.x) * 8) {
        const int x = static_cast<int>(p % w);
        const int yy = static_cast<int>((p / w) % h);
        float acc = 0.f;
        if (lane < c) {
            if (K > 0) {
                float v[K > 0 ? K * K : 1];
#pragma unroll
                for (int r = 0; r < K; ++r)
#pragma unroll
                    for (int q = 0; q < K; ++q) {
                        const int iy = yy + r * dil - pad, ix = x + q * dil - pad;
                        const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < w;
                        const long long src = p + static_cast<long long>(iy - yy) * w + (ix - x);
                        v[r * K + q] = ok ? __ldg(t_buf + src * t_ld + (r * K + q) * c + lane) : 0.f;
                    }
#pragma unroll
                for (int t = 0; t < K * K; ++t) acc += v[t];
            } else {
                for (int r = 0; r < k; ++r) {
                    const int iy = yy + r * dil - pad;
                    if (iy < 0 || iy >= h) continue;
                    for (int q = 0; q < k; ++q) {
                        const int ix = x + q * dil - pad;
                        if (ix < 0 || ix >= w) continue;
                        const long long src = p + static_cast<long long>(iy - yy) * w + (ix - x);
                        acc += __ldg(t_buf + src * t_ld + (r * k + q) * c + lane);
                    }
                }
            }
            s1 += acc; s2 += acc * acc;
            float v = acc * sc + sh;
            if (act == RTSDS_ACT_RELU) v = fmaxf(v, 0.f);
            y[p * y_ld + lane] = v;
            gs += v;
        }
    }
    if (gap_out) {                  // AdaptiveAvgPool2d(1) of the FFM feature (build_bisenet.py:75) fused into its producer
        __syncthreads();
        s_stats[0][wrp][lane] = gs;
        __syncthreads();
        if (wrp == 0 && lane < c) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) a += s_stats[0][i][lane];
            // one partial mean per block, written (not accumulated): the consumer adds them in block order — deterministic
            gap_out[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * c + lane] = a * gap_scale;
        }
        __syncthreads();
    }
    if (stats) {
        s_stats[0][wrp][lane] = s1; s_stats[1][wrp][lane] = s2;
        __syncthreads();
        if (wrp == 0 && lane < c) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { a += s_stats[0][i][lane]; b += s_stats[1][i][lane]; }
            if (det) {            // deterministic mode: the block's sums (fixed warp order) go into exact accumulators
                det_add(det + 2 * lane, a);
                det_add(det + 2 * (c + lane), b);
            } else {
                atomicAdd(&stats[lane], a);
                atomicAdd(&stats[c + lane], b);
            }
        }
    }
}

// G[q, t*C+co] = dy[q - off(t), co] (0 outside the image; columns taps*C .. g_ld-1 = 0).  A block assembles the G rows
// of SC_PIX consecutive pixels in shared memory (one thread per (pixel, tap) copies C values) and writes them out as
// one contiguous run of 16-byte vectors.
constexpr int SC_PIX = 32;
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
tapn_scatter_kernel(const TI* __restrict__ dy, int dy_ld, int n, int h, int w, int c, int k, int pad, int dil,
                    TO* __restrict__ g, int g_ld) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    TO* s_g = reinterpret_cast<TO*>(s_raw);                  // [SC_PIX][g_ld]
    const int taps = k * k;
    const long long npix = static_cast<long long>(n) * h * w;
    const long long q0 = static_cast<long long>(blockIdx.x) * SC_PIX;
    const int np = static_cast<int>(min(static_cast<long long>(SC_PIX), npix - q0));
    for (int i = threadIdx.x; i < np * g_ld; i += blockDim.x) s_g[i] = from_f32<TO>(0.f);
    __syncthreads();
    for (int it = threadIdx.x; it < np * taps; it += blockDim.x) {
        const int pl = it / taps, t = it - pl * taps;
        const long long q = q0 + pl;
        const int r = t / k, s = t - r * k;
        const int x = static_cast<int>(q % w);
        const int yy = static_cast<int>((q / w) % h);
        const int iy = yy - (r * dil - pad), ix = x - (s * dil - pad);
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
            const TI* src = dy + (q + static_cast<long long>(iy - yy) * w + (ix - x)) * dy_ld;
            TO* dst = s_g + pl * g_ld + t * c;
            for (int j = 0; j < c; ++j) dst[j] = from_f32<TO>(to_f32(src[j]));
        }
    }
    __syncthreads();
    const int nvec = np * g_ld * static_cast<int>(sizeof(TO)) / 16;
    uint4* out = reinterpret_cast<uint4*>(g + q0 * g_ld);
    const uint4* in = reinterpret_cast<const uint4*>(s_g);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) out[i] = in[i];
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_tapn_weights(const float* w_oihw, int c, int cin, int k, int kpad, float* w_fwd, float* w_bwd,
                                  rtsds_stream_t s) {
    RTSDS_REQUIRE(w_oihw && (w_fwd || w_bwd) && c > 0 && c <= 32 && cin > 0 && k > 0 && kpad >= k * k * c, "tapn_weights: bad argument");
    const long long total = static_cast<long long>(cin) * kpad;
    long long g = cdiv(total, 256);
    if (g > 2048) g = 2048;
    tapn_weights_kernel<<<static_cast<int>(g), 256, 0, as_stream(s)>>>(w_oihw, c, cin, k * k, kpad, w_fwd, w_bwd);
    count_launch();
    return check_launch("tapn_weights_kernel");
}

extern "C" int rtsds_tapn_weight_grad(const float* dw2, int c, int cin, int k, float* grad_oihw, rtsds_stream_t s) {
    RTSDS_REQUIRE(dw2 && grad_oihw && c > 0 && cin > 0 && k > 0, "tapn_weight_grad: bad argument");
    const long long total = static_cast<long long>(c) * cin * k * k;
    long long g = cdiv(total, 256);
    if (g > 2048) g = 2048;
    tapn_weight_grad_kernel<<<static_cast<int>(g), 256, 0, as_stream(s)>>>(dw2, c, cin, k * k, grad_oihw);
    count_launch();
    return check_launch("tapn_weight_grad_kernel");
}

// blocks per image; with the fused pool one block = one partial, so the grid is kept to about one wave
static long long tapn_gather_blocks(int n, int h, int w, bool gap) {
    const long long npix = static_cast<long long>(h) * w;
    long long g = cdiv(npix, 8 * 4);
    const long long cap = cdiv((gap ? 1LL : 8LL) * num_sms(), n);
    if (g > cap) g = cap;
    return g < 1 ? 1 : g;
}
extern "C" int rtsds_tapn_gather_parts(int n, int h, int w) { return static_cast<int>(tapn_gather_blocks(n, h, w, true)); }

extern "C" int rtsds_tapn_gather(const float* t_buf, int t_ld, int n, int h, int w, int c, int k, int pad, int dil,
                                 const float* scale, const float* shift, int act, float* stats, float* y, int y_ld,
                                 float* gap_out, rtsds_stream_t s) {
    RTSDS_REQUIRE(t_buf && y && n > 0 && h > 0 && w > 0 && c > 0 && c <= 32 && k > 0 && t_ld >= k * k * c && y_ld >= c, "tapn_gather: bad argument");
    RTSDS_REQUIRE(n <= 65535, "tapn_gather: batch too large");
    const long long npix = static_cast<long long>(h) * w;
    const long long g = tapn_gather_blocks(n, h, w, gap_out != nullptr);
    const dim3 grid(static_cast<unsigned>(g), static_cast<unsigned>(n));
    const float gsc = 1.0f / static_cast<float>(npix);
    unsigned long long* det = nullptr;
    if (stats && det_mode()) {
        det = det_scratch(as_stream(s), 2 * static_cast<size_t>(c));
        if (!det) return RTSDS_ECUDA;
    }
    if (k == 3)
        launch_pdl(tapn_gather_kernel<3>, grid, dim3(256), 0, as_stream(s), t_buf, t_ld, n, h, w, c, k, pad, dil, scale, shift, act, stats, det, y, y_ld, gap_out, gsc);
    else
        launch_pdl(tapn_gather_kernel<0>, grid, dim3(256), 0, as_stream(s), t_buf, t_ld, n, h, w, c, k, pad, dil, scale, shift, act, stats, det, y, y_ld, gap_out, gsc);
    count_launch();
    int rc = check_launch("tapn_gather_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, stats, 2 * static_cast<size_t>(c), true, as_stream(s));
    return rc;
}

extern "C" int rtsds_tapn_scatter(const void* dy, int dy_ld, int dy_dtype, int n, int h, int w, int c, int k, int pad, int dil,
                                  void* g, int g_ld, int g_dtype, rtsds_stream_t s) {
    RTSDS_REQUIRE(dy && g && n > 0 && h > 0 && w > 0 && c > 0 && c <= 32 && k > 0 && dy_ld >= c && g_ld >= k * k * c, "tapn_scatter: bad argument");
    const size_t esz = g_dtype == RTSDS_BF16 ? 2 : 4;
    RTSDS_REQUIRE((g_ld * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "tapn_scatter: G rows must be multiples of 16 bytes");
    const long long grid = cdiv(static_cast<long long>(n) * h * w, SC_PIX);
    RTSDS_REQUIRE(grid < (1LL << 31), "tapn_scatter: too many pixels");
    const int gi = static_cast<int>(grid);
    const size_t smem = static_cast<size_t>(SC_PIX) * g_ld * esz;
    RTSDS_REQUIRE(smem <= 48 * 1024, "tapn_scatter: row too wide");
#define TAPN_SC(TI, TO) tapn_scatter_kernel<TI, TO><<<gi, 256, smem, as_stream(s)>>>(reinterpret_cast<const TI*>(dy), dy_ld, n, h, w, c, k, pad, dil, reinterpret_cast<TO*>(g), g_ld)
    if (dy_dtype == RTSDS_BF16 && g_dtype == RTSDS_BF16) TAPN_SC(__nv_bfloat16, __nv_bfloat16);
    else if (dy_dtype == RTSDS_F32 && g_dtype == RTSDS_BF16) TAPN_SC(float, __nv_bfloat16);
    else if (dy_dtype == RTSDS_F32 && g_dtype == RTSDS_F32) TAPN_SC(float, float);
    else { set_error("tapn_scatter: unsupported dtype combination"); return RTSDS_EINVAL; }
#undef TAPN_SC
    count_launch();
    return check_launch("tapn_scatter_kernel");
}
