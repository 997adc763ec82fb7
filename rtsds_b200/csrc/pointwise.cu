// Bandwidth-bound glue kernels of the BiSeNet forward (models/bisenet/build_bisenet.py):
// BatchNorm folding / finalisation, scale-shift-activation, global average
// pooling, the ARM channel gate, gated bilinear resize into the concat buffer,
// the FFM attention + final 1x1 conv, and the bilinear resize to NCHW logits.
// All are vectorised, coalesced, warp-shuffle / shared-memory reductions; none
// is GEMM-shaped.
#include "common.cuh"

namespace rtsds {

// ---------------------------------------------------------------- BatchNorm
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var,
                               const float* conv_bias, float eps, int c, float* scale, float* shift) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    float sc = (gamma ? gamma[i] : 1.f) / sqrtf(var[i] + eps);
    float sh = (beta ? beta[i] : 0.f) - mean[i] * sc;
    if (conv_bias) sh += conv_bias[i] * sc;
    scale[i] = sc; shift[i] = sh;
}

__global__ void bn_finalize_kernel(const float* stats, double count, const float* gamma, const float* beta,
                                   float eps, float momentum, int c, float* running_mean, float* running_var,
                                   float* scale, float* shift, float* save_mean, float* save_invstd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const double mean = static_cast<double>(stats[i]) / count;
    double var = static_cast<double>(stats[c + i]) / count - mean * mean;     // biased
    if (var < 0.0) var = 0.0;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
    const float sc = g * invstd;
    scale[i] = sc;
    shift[i] = b - static_cast<float>(mean) * sc;
    if (save_mean) save_mean[i] = static_cast<float>(mean);
    if (save_invstd) save_invstd[i] = invstd;
    if (running_mean) running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * static_cast<float>(mean);
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[i] = (1.f - momentum) * running_var[i] + momentum * static_cast<float>(unbiased);
    }
}

// y = act(scale*x + shift + residual); 8 channels per thread.  VEC: 16-byte accesses (c and all pitches
// multiples of 8); otherwise a scalar tail-safe path.
template <typename T> struct V8io;
template <> struct V8io<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};
template <> struct V8io<__half> {
    static __device__ __forceinline__ void ld(const __half* p, float (&v)[8]) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        float2 a = unpack_f16x2(u.x), b = unpack_f16x2(u.y), c = unpack_f16x2(u.z), d = unpack_f16x2(u.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
    static __device__ __forceinline__ void st(__half* p, const float (&v)[8]) {
        uint4 u;
        u.x = pack_f16x2(v[0], v[1]); u.y = pack_f16x2(v[2], v[3]); u.z = pack_f16x2(v[4], v[5]); u.w = pack_f16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};
template <> struct V8io<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <typename TX, typename TY, bool VEC>
__global__ void __launch_bounds__(256)
scale_shift_act_kernel(const TX* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                       const TY* residual, long long n_pix, int c, int x_ld, int res_ld, int y_ld,
                       int act, float slope, TY* y) {
    const int cg = (c + 7) / 8;
    if (VEC && blockDim.x % cg == 0) {
        // every thread keeps ONE 8-channel group: its scale / shift live in registers, two pixels in flight per iteration
        const int c0 = (threadIdx.x % cg) * 8;
        const long long prows = blockDim.x / cg, stride = prows * gridDim.x;
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale ? __ldg(scale + c0 + j) : 1.f; sh[j] = shift ? __ldg(shift + c0 + j) : 0.f; }
        for (long long p0 = static_cast<long long>(blockIdx.x) * prows + threadIdx.x / cg; p0 < n_pix; p0 += 2 * stride) {
            const long long p1 = p0 + stride;
            const bool two = p1 < n_pix;
            float v0[8], v1[8], r0[8], r1[8];
            V8io<TX>::ld(x + p0 * x_ld + c0, v0);
            if (two) V8io<TX>::ld(x + p1 * x_ld + c0, v1);
            if (residual) {
                V8io<TY>::ld(residual + p0 * res_ld + c0, r0);
                if (two) V8io<TY>::ld(residual + p1 * res_ld + c0, r1);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t0 = v0[j] * sc[j] + sh[j], t1 = v1[j] * sc[j] + sh[j];
                if (residual) { t0 += r0[j]; t1 += r1[j]; }
                v0[j] = apply_act(t0, act, slope);
                v1[j] = apply_act(t1, act, slope);
            }
            V8io<TY>::st(y + p0 * y_ld + c0, v0);
            if (two) V8io<TY>::st(y + p1 * y_ld + c0, v1);
        }
        return;
    }
    const long long total = n_pix * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pix = i / cg;
        const int c0 = static_cast<int>(i - pix * cg) * 8;
        if (VEC) {
            float v[8], r[8];
            V8io<TX>::ld(x + pix * x_ld + c0, v);
            if (residual) V8io<TY>::ld(residual + pix * res_ld + c0, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = v[j] * (scale ? __ldg(scale + c0 + j) : 1.f) + (shift ? __ldg(shift + c0 + j) : 0.f);
                if (residual) t += r[j];
                v[j] = apply_act(t, act, slope);
            }
            V8io<TY>::st(y + pix * y_ld + c0, v);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = c0 + j;
                if (ch < c) {
                    float v = to_f32(x[pix * x_ld + ch]) * (scale ? scale[ch] : 1.f) + (shift ? shift[ch] : 0.f);
                    if (residual) v += to_f32(residual[pix * res_ld + ch]);
                    y[pix * y_ld + ch] = from_f32<TY>(apply_act(v, act, slope));
                }
            }
        }
    }
}

// Train-mode BatchNorm in ONE launch: rtsds_bn_finalize + rtsds_scale_shift_act.  Every block derives scale / shift of all
// channels from the batch sums once (one channel per thread, exactly bn_finalize_kernel's arithmetic, so the value applied,
// the value written out for the backward pass and the two-launch form agree bit for bit) into shared memory; block 0 also
// writes scale / shift / save_mean / save_invstd and updates the running statistics.  (Per-THREAD derivation of its 8
// channels was measured 0.4-0.7 % slower per step than the two launches: 8 fp64 divide + sqrt chains per thread.)  Removes one dependent ~5-8 us launch per BatchNorm layer (24 per BiSeNet-R18 step, 104 per DeepLabV2 step).
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
bn_train_apply_kernel(const TX* __restrict__ x, const float* __restrict__ stats, double count, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float momentum, float* running_mean, float* running_var,
                      float* scale_out, float* shift_out, float* save_mean, float* save_invstd, const TY* residual,
                      long long n_pix, int c, int x_ld, int res_ld, int y_ld, int act, float slope, TY* y) {
    extern __shared__ float s_ss[];                     // [2*c]: scale | shift, computed once per block (one channel per thread)
    const int cg = c / 8;
    const int c0 = (threadIdx.x % cg) * 8;
    const long long prows = blockDim.x / cg, stride = prows * gridDim.x;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const double mean = static_cast<double>(stats[i]) / count;
        double var = static_cast<double>(stats[c + i]) / count - mean * mean;     // biased
        if (var < 0.0) var = 0.0;
        const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
        const float sc1 = g * invstd;
        const float sh1 = b - static_cast<float>(mean) * sc1;
        s_ss[i] = sc1; s_ss[c + i] = sh1;
        if (blockIdx.x == 0) {
            scale_out[i] = sc1;
            shift_out[i] = sh1;
            if (save_mean) save_mean[i] = static_cast<float>(mean);
            if (save_invstd) save_invstd[i] = invstd;
            if (running_mean) running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * static_cast<float>(mean);
            if (running_var) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_var[i] = (1.f - momentum) * running_var[i] + momentum * static_cast<float>(unbiased);
            }
        }
    }
    __syncthreads();
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = s_ss[c0 + j]; sh[j] = s_ss[c + c0 + j]; }
    for (long long p0 = static_cast<long long>(blockIdx.x) * prows + threadIdx.x / cg; p0 < n_pix; p0 += 2 * stride) {
        const long long p1 = p0 + stride;
        const bool two = p1 < n_pix;
        float v0[8], v1[8], r0[8], r1[8];
        V8io<TX>::ld(x + p0 * x_ld + c0, v0);
        if (two) V8io<TX>::ld(x + p1 * x_ld + c0, v1);
        if (residual) {
            V8io<TY>::ld(residual + p0 * res_ld + c0, r0);
            if (two) V8io<TY>::ld(residual + p1 * res_ld + c0, r1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t0 = v0[j] * sc[j] + sh[j], t1 = v1[j] * sc[j] + sh[j];
            if (residual) { t0 += r0[j]; t1 += r1[j]; }
            v0[j] = apply_act(t0, act, slope);
            v1[j] = apply_act(t1, act, slope);
        }
        V8io<TY>::st(y + p0 * y_ld + c0, v0);
        if (two) V8io<TY>::st(y + p1 * y_ld + c0, v1);
    }
}

// ---------------------------------------------------------------- global average pool
// grid (c/32, psplit, n), block 256 = 32 channels x 8 pixel lanes; out pre-zeroed.
template <typename T>
__global__ void __launch_bounds__(256)
gap_kernel(const T* __restrict__ x, long long hw, int c, int ld, float inv_hw, float* __restrict__ out, unsigned long long* det) {
    __shared__ float s[8][33];
    const int ch = blockIdx.x * 32 + (threadIdx.x & 31);
    const int pl = threadIdx.x >> 5;
    const int img = blockIdx.z;
    const long long per = (hw + gridDim.y - 1) / gridDim.y;
    const long long p0 = per * blockIdx.y;
    const long long p1 = (p0 + per < hw) ? p0 + per : hw;
    float acc = 0.f;
    if (ch < c) {
        const T* base = x + static_cast<long long>(img) * hw * ld + ch;
        long long p = p0 + pl;
        for (; p + 24 < p1; p += 32) {
            float a0 = to_f32(base[p * ld]), a1 = to_f32(base[(p + 8) * ld]);
            float a2 = to_f32(base[(p + 16) * ld]), a3 = to_f32(base[(p + 24) * ld]);
            acc += (a0 + a1) + (a2 + a3);
        }
        for (; p < p1; p += 8) acc += to_f32(base[p * ld]);
    }
    s[pl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (pl == 0 && ch < c) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x];
        if (det) det_add(det + 2 * (static_cast<long long>(img) * c + ch), t * inv_hw);      // deterministic mode (common.cuh)
        else atomicAdd(&out[static_cast<long long>(img) * c + ch], t * inv_hw);
    }
}

// ---------------------------------------------------------------- ARM gate
// One warp per output channel; loops over the batch.  n <= ARM_MAX_N in train mode.
constexpr int ARM_MAX_N = 128;
__global__ void __launch_bounds__(256)
arm_gate_kernel(const float* __restrict__ pooled, const float* __restrict__ w, const float* __restrict__ b,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                float* running_var, float eps, float momentum, int train, int n, int c,
                const float* __restrict__ mul, float* __restrict__ gate, float* lin_out, float* xhat_out) {
    __shared__ float s_lin[8][ARM_MAX_N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co = blockIdx.x * 8 + warp;
    if (co >= c) return;
    const float* wr = w + static_cast<long long>(co) * c;
    const float bias = b ? b[co] : 0.f;
    const float g = gamma[co], be = beta[co];
    if (!train) {
        const float sc = g / sqrtf(running_var[co] + eps);
        const float sh = be - running_mean[co] * sc;
        for (int i = 0; i < n; ++i) {
            float acc = 0.f;
            for (int k = lane; k < c; k += 32) acc = fmaf(wr[k], pooled[static_cast<long long>(i) * c + k], acc);
            acc = warp_sum(acc) + bias;
            if (lane == 0) {
                float v = 1.f / (1.f + expf(-(acc * sc + sh)));
                if (mul) v *= mul[static_cast<long long>(i) * c + co];
                gate[static_cast<long long>(i) * c + co] = v;
                if (lin_out) lin_out[static_cast<long long>(i) * c + co] = acc;
            }
        }
        return;
    }
    for (int i = 0; i < n; ++i) {
        float acc = 0.f;
        for (int k = lane; k < c; k += 32) acc = fmaf(wr[k], pooled[static_cast<long long>(i) * c + k], acc);
        acc = warp_sum(acc) + bias;
        if (lane == 0) s_lin[warp][i] = acc;
    }
    __syncwarp();
    float m = 0.f;
    for (int i = lane; i < n; i += 32) m += s_lin[warp][i];
    m = warp_sum(m) / n;
    float v = 0.f;
    for (int i = lane; i < n; i += 32) { float d = s_lin[warp][i] - m; v += d * d; }
    v = warp_sum(v) / n;                                 // biased variance
    const float invstd = rsqrtf(v + eps);
    for (int i = lane; i < n; i += 32) {
        const float lin = s_lin[warp][i];
        const float xh = (lin - m) * invstd;
        float o = 1.f / (1.f + expf(-(xh * g + be)));
        if (mul) o *= mul[static_cast<long long>(i) * c + co];
        gate[static_cast<long long>(i) * c + co] = o;
        if (lin_out) lin_out[static_cast<long long>(i) * c + co] = lin;
        if (xhat_out) xhat_out[static_cast<long long>(i) * c + co] = xh;
    }
    if (lane == 0) {
        if (running_mean) running_mean[co] = (1.f - momentum) * running_mean[co] + momentum * m;
        if (running_var) running_var[co] = (1.f - momentum) * running_var[co] + momentum * (v * n / (n - 1));
    }
}

// ---------------------------------------------------------------- gated bilinear resize NHWC -> NHWC slot
struct V8 { float v[8]; };
__device__ __forceinline__ V8 ld8(const __nv_bfloat16* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    V8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ V8 ld8(const __half* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    V8 r;
    float2 a = unpack_f16x2(u.x), b = unpack_f16x2(u.y), c = unpack_f16x2(u.z), d = unpack_f16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ void st8(__half* p, const V8& r) {
    uint4 u;
    u.x = pack_f16x2(r.v[0], r.v[1]); u.y = pack_f16x2(r.v[2], r.v[3]);
    u.z = pack_f16x2(r.v[4], r.v[5]); u.w = pack_f16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ V8 ld8(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    V8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const V8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8(float* p, const V8& r) {
    reinterpret_cast<float4*>(p)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// one thread = one destination pixel x 8 channels (16-byte loads/stores for bf16)
template <typename T>
__global__ void __launch_bounds__(256)
gate_resize_kernel(const T* __restrict__ src, int n, int h, int w, int c, int src_ld, const float* __restrict__ gate,
                   float gate_scale, int oh, int ow, float rh, float rw, T* __restrict__ dst, int dst_ld, int dst_coff,
                   FastDiv dcg, FastDiv dow, FastDiv doh) {
    const int cg = c / 8;
    const long long total = static_cast<long long>(n) * oh * ow * cg;          // < 2^31 (host check)
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned ug, ux, uy;
        unsigned r = fdivmod(static_cast<unsigned>(i), dcg, &ug);
        r = fdivmod(r, dow, &ux);
        const int img = static_cast<int>(fdivmod(r, doh, &uy));
        const int g8 = static_cast<int>(ug), ox = static_cast<int>(ux), oy = static_cast<int>(uy);
        const Lerp ly = lerp_src(oy, rh, h), lx = lerp_src(ox, rw, w);
        const T* b = src + static_cast<long long>(img) * h * w * src_ld + g8 * 8;
        const V8 p00 = ld8(b + (static_cast<long long>(ly.i0) * w + lx.i0) * src_ld);
        const V8 p01 = ld8(b + (static_cast<long long>(ly.i0) * w + lx.i1) * src_ld);
        const V8 p10 = ld8(b + (static_cast<long long>(ly.i1) * w + lx.i0) * src_ld);
        const V8 p11 = ld8(b + (static_cast<long long>(ly.i1) * w + lx.i1) * src_ld);
        V8 gv;
        if (gate) gv = ld8(gate + static_cast<long long>(img) * c + g8 * 8);
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = ly.l0 * (lx.l0 * p00.v[j] + lx.l1 * p01.v[j]) + ly.l1 * (lx.l0 * p10.v[j] + lx.l1 * p11.v[j]);
            if (gate) v *= gv.v[j];
            o.v[j] = v * gate_scale;
        }
        st8(dst + ((static_cast<long long>(img) * oh + oy) * ow + ox) * dst_ld + dst_coff + g8 * 8, o);
    }
}

// ---------------------------------------------------------------- FFM attention + final 1x1 conv
constexpr int FFM_MAX_C = 32, FFM_THREADS = 128, FFM_PX = 32;
// One block = tiles of 32 pixels: features are staged (coalesced) in shared memory with the attention applied, four
// threads share a pixel and take every 4th output channel of the final 1x1 conv, results leave coalesced.
template <typename T>
__global__ void __launch_bounds__(FFM_THREADS)
ffm_head_kernel(const T* __restrict__ f, int f_ld, const float* __restrict__ pooled, long long hw, int c,
                const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                const float* __restrict__ b2, const float* __restrict__ wc, const float* __restrict__ bc,
                float* attn_out, float* __restrict__ z, int z_ld) {
    __shared__ float s_h[FFM_MAX_C], s_a[FFM_MAX_C];
    __shared__ float s_w[FFM_MAX_C * FFM_MAX_C], s_b[FFM_MAX_C];
    __shared__ float s_g[FFM_PX][FFM_MAX_C + 1], s_o[FFM_PX][FFM_MAX_C + 1];
    const int img = blockIdx.y;
    const int t = threadIdx.x;
    const float* pp = pooled + static_cast<long long>(img) * c;
    if (t < c) {
        float acc = b1[t];
        for (int k = 0; k < c; ++k) acc = fmaf(w1[t * c + k], pp[k], acc);
        s_h[t] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (t < c) {
        float acc = b2[t];
        for (int k = 0; k < c; ++k) acc = fmaf(w2[t * c + k], s_h[k], acc);
        const float a = 1.f / (1.f + expf(-acc));
        s_a[t] = a;
        if (attn_out && blockIdx.x == 0) attn_out[static_cast<long long>(img) * c + t] = a;
    }
    // g = f*a + f is folded into the operand of the 1x1 conv
    for (int i = t; i < c * c; i += FFM_THREADS) s_w[i] = wc ? wc[i] : 0.f;
    if (t < c) s_b[t] = (wc && bc) ? bc[t] : 0.f;
    __syncthreads();
    const long long n_tiles = (hw + FFM_PX - 1) / FFM_PX;
    const int px = t >> 2, oq = t & 3;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p0 = tile * FFM_PX;
        for (int i = t; i < FFM_PX * FFM_MAX_C; i += FFM_THREADS) {
            const int q = i >> 5, k = i & 31;
            float v = 0.f;
            if (k < c && p0 + q < hw) v = to_f32(f[(static_cast<long long>(img) * hw + p0 + q) * f_ld + k]);
            s_g[q][k] = k < c ? fmaf(v, s_a[k], v) : 0.f;
        }
        __syncthreads();
        if (wc) {
            for (int o = oq; o < c; o += 4) {
                float acc = s_b[o];
                for (int k = 0; k < c; ++k) acc = fmaf(s_w[o * c + k], s_g[px][k], acc);
                s_o[px][o] = acc;
            }
        } else {
            for (int o = oq; o < c; o += 4) s_o[px][o] = s_g[px][o];
        }
        __syncthreads();
        for (int i = t; i < FFM_PX * FFM_MAX_C; i += FFM_THREADS) {
            const int q = i >> 5, k = i & 31;
            if (k < c && p0 + q < hw) z[(static_cast<long long>(img) * hw + p0 + q) * z_ld + k] = s_o[q][k];
        }
    }
}

static unsigned long long* g_arm_trace = nullptr;

// ---------------------------------------------------------------- ARM gate + gated bilinear resize, both ARMs, one launch
// Eval mode (folded BatchNorm).  A block owns 32 channels of one ARM and a chunk of destination pixels: it first evaluates
// ITS 32 gates  g[n,c] = sigmoid(BN(W[c,:] . pooled[n,:] + b[c])) (* pooled[n,c] for the `cx2 * tail` of
// build_bisenet.py:149) (* out_scale) — a warp per 4 channels, 32 dot products of length C, redundant across the pixel
// chunks but 16 K MACs — and then streams its pixels: bilinear gather from the 1/16 or 1/32 feature map, times the gate,
// 16-byte stores into the concat buffer slot.  Replaces arm_gate x2 + gate_resize x2 (4 launches, ~26 us at b=1).
struct ArmSide {
    const void* src; const float* pooled; const float* w; const float* b; const float* gamma; const float* beta;
    const float* mean; const float* var; float eps, out_scale; int h, w_, c, coff, mul_pooled, parts;
};
template <typename T>
__global__ void __launch_bounds__(256)
arm_gate_resize_kernel(ArmSide a0, ArmSide a1, int blocks0, int chunks, int n, int oh, int ow, T* __restrict__ dst, int dst_ld,
                       unsigned long long* trace) {
    __shared__ float s_gate[32];
    pdl_wait();
    if (trace && threadIdx.x == 0) {
        trace[blockIdx.x * 8 + 0] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        trace[blockIdx.x * 8 + 4] = gt;
    }
    extern __shared__ float s_pool[];                                     // [c]: pooled vector of the current image
    const bool second = static_cast<int>(blockIdx.x) >= blocks0;
    const ArmSide& a = second ? a1 : a0;
    const int local = second ? blockIdx.x - blocks0 : blockIdx.x;
    const int cg = local / chunks, chunk = local - cg * chunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npix = oh * ow;
    const int per = (npix + chunks - 1) / chunks;
    const int p_begin = chunk * per, p_end = min(p_begin + per, npix);
    const float rh = static_cast<float>(a.h) / static_cast<float>(oh), rw = static_cast<float>(a.w_) / static_cast<float>(ow);
    const T* src = reinterpret_cast<const T*>(a.src);
    for (int img = 0; img < n; ++img) {
        // pooled[n][parts][c]: per-CTA partial means of the producing conv's epilogue, added in part order (deterministic)
        // (loads batched eight at a time: a plain loop is one L2 round trip per part, 16-32 of them back to back)
        for (int ch = threadIdx.x; ch < a.c; ch += 256) {
            const float* pq = a.pooled + static_cast<long long>(img) * a.parts * a.c + ch;
            float acc = 0.f;
            int q = 0;
            for (; q + 8 <= a.parts; q += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(pq + static_cast<long long>(q + j) * a.c);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += v[j];
            }
            for (; q < a.parts; ++q) acc += __ldg(pq + static_cast<long long>(q) * a.c);
            s_pool[ch] = acc;
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[blockIdx.x * 8 + 1] = clock64();
        const float* pp = s_pool;
        {   // 32 gates, 8 threads each: every thread sums a stride-8 slice of its dot product, shuffles combine the slices
            const int gi = threadIdx.x >> 3, sub = threadIdx.x & 7;
            const int co = cg * 32 + gi;
            const float* wr = a.w + static_cast<long long>(co) * a.c;
            float acc0 = 0.f, acc1 = 0.f;
            for (int i0 = sub; i0 < a.c; i0 += 128) {                        // 16 weights in flight per thread (c % 32 == 0)
                float wv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) wv[j] = (i0 + 8 * j < a.c) ? __ldg(wr + i0 + 8 * j) : 0.f;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    if (i0 + 8 * j < a.c) {                               // pairs stay together: c is a multiple of 16
                        acc0 = fmaf(wv[j], pp[i0 + 8 * j], acc0);
                        acc1 = fmaf(wv[j + 1], pp[i0 + 8 * j + 8], acc1);
                    }
                }
            }
            float acc = acc0 + acc1;
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (sub == 0) {
                acc += a.b ? a.b[co] : 0.f;
                const float sc = a.gamma[co] / sqrtf(a.var[co] + a.eps);
                const float sh = a.beta[co] - a.mean[co] * sc;
                float v = 1.f / (1.f + expf(-(acc * sc + sh)));
                if (a.mul_pooled) v *= pp[co];
                s_gate[gi] = v * a.out_scale;
            }
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[blockIdx.x * 8 + 2] = clock64();
        const int g8 = threadIdx.x & 3;                                   // 8-channel group inside the block's 32 channels
        float gv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[j] = s_gate[g8 * 8 + j];
        const T* sb = src + static_cast<long long>(img) * a.h * a.w_ * a.c + cg * 32 + g8 * 8;
        T* db = dst + static_cast<long long>(img) * npix * dst_ld + a.coff + cg * 32 + g8 * 8;
        // two pixels per thread and trip: 8 independent 16-byte gathers in flight instead of 4
        for (int p = p_begin + (threadIdx.x >> 2); p < p_end; p += 128) {
            const int pb = p + 64;
            const bool two = pb < p_end;
            const int oy = p / ow, ox = p - oy * ow;
            const int oyb = two ? pb / ow : oy, oxb = two ? pb - oyb * ow : ox;
            const Lerp ly = lerp_src(oy, rh, a.h), lx = lerp_src(ox, rw, a.w_);
            const Lerp lyb = lerp_src(oyb, rh, a.h), lxb = lerp_src(oxb, rw, a.w_);
            const V8 p00 = ld8(sb + (static_cast<long long>(ly.i0) * a.w_ + lx.i0) * a.c);
            const V8 p01 = ld8(sb + (static_cast<long long>(ly.i0) * a.w_ + lx.i1) * a.c);
            const V8 p10 = ld8(sb + (static_cast<long long>(ly.i1) * a.w_ + lx.i0) * a.c);
            const V8 p11 = ld8(sb + (static_cast<long long>(ly.i1) * a.w_ + lx.i1) * a.c);
            const V8 q00 = ld8(sb + (static_cast<long long>(lyb.i0) * a.w_ + lxb.i0) * a.c);
            const V8 q01 = ld8(sb + (static_cast<long long>(lyb.i0) * a.w_ + lxb.i1) * a.c);
            const V8 q10 = ld8(sb + (static_cast<long long>(lyb.i1) * a.w_ + lxb.i0) * a.c);
            const V8 q11 = ld8(sb + (static_cast<long long>(lyb.i1) * a.w_ + lxb.i1) * a.c);
            V8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o.v[j] = (ly.l0 * (lx.l0 * p00.v[j] + lx.l1 * p01.v[j]) + ly.l1 * (lx.l0 * p10.v[j] + lx.l1 * p11.v[j])) * gv[j];
            st8(db + static_cast<long long>(p) * dst_ld, o);
            if (two) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o.v[j] = (lyb.l0 * (lxb.l0 * q00.v[j] + lxb.l1 * q01.v[j]) + lyb.l1 * (lxb.l0 * q10.v[j] + lxb.l1 * q11.v[j])) * gv[j];
                st8(db + static_cast<long long>(pb) * dst_ld, o);
            }
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) {
            trace[blockIdx.x * 8 + 3] = clock64();
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            trace[blockIdx.x * 8 + 5] = gt;
        }
    }
}

// ---------------------------------------------------------------- FFM attention + final 1x1 conv + x8 resize -> NCHW logits
// out[n,o,oy,ox] = bc[o] + sum_k Wc[o,k] (1 + a[n,k]) * bilinear(f)[k]:  the 1x1 conv commutes with the bilinear resize, so z
// is evaluated at feature resolution — here only for the source rows a block's output rows touch, in shared memory — and
// the block then writes its rows of all classes.  Replaces ffm_head + resize_nchw (and the z round trip through L2).
constexpr int FHR_ROWS = 4, FHR_THREADS = 256;
__global__ void __launch_bounds__(FHR_THREADS)
ffm_head_resize_kernel(const float* __restrict__ f, int f_ld, const float* __restrict__ pooled, int h, int w, int c,
                       const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                       const float* __restrict__ b2, const float* __restrict__ wc, const float* __restrict__ bc,
                       float* attn_out, int oh, int ow, float rh, float rw, int max_rows, int parts, float* __restrict__ out,
                       unsigned long long* trace) {
    extern __shared__ float s_z[];                                        // [rows][w][c]
    if (trace && threadIdx.x == 0) trace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 0] = clock64();
    __shared__ float s_h[FFM_MAX_C], s_a[FFM_MAX_C], s_b[FFM_MAX_C], s_p[8][FFM_MAX_C], s_pool[FFM_MAX_C];
    __shared__ float s_w[FFM_MAX_C * FFM_MAX_C], s_w1[FFM_MAX_C * FFM_MAX_C], s_w2[FFM_MAX_C * FFM_MAX_C];
    __shared__ float s_b1[FFM_MAX_C], s_b2[FFM_MAX_C];
    const int img = blockIdx.y, t = threadIdx.x;
    {   // every block repeats this prologue, and it is a chain of dependent global-memory round trips if done naively
        // (2 x c serial weight loads per thread): all operands come in with ONE round trip, spread over the block
        for (int i = t; i < c * c; i += FHR_THREADS) { s_w1[i] = __ldg(w1 + i); s_w2[i] = __ldg(w2 + i); s_w[i] = __ldg(wc + i); }
        if (t < c) { s_b1[t] = __ldg(b1 + t); s_b2[t] = __ldg(b2 + t); s_b[t] = bc ? __ldg(bc + t) : 0.f; }
        pdl_wait();                          // the weights above are parameters; everything below reads the predecessor's output
        // pooled[n][parts][c]: partial means of the producer (one per block of the gather), added in a fixed order
        const int ch = t & 31, stripe = t >> 5;
        float acc = 0.f;
        if (ch < c) {
            const float* pq = pooled + static_cast<long long>(img) * parts * c + ch;
            int q = stripe;
            for (; q + 24 < parts; q += 32) {                                  // four independent loads in flight
                const float v0 = __ldg(pq + static_cast<long long>(q) * c), v1 = __ldg(pq + static_cast<long long>(q + 8) * c);
                const float v2 = __ldg(pq + static_cast<long long>(q + 16) * c), v3 = __ldg(pq + static_cast<long long>(q + 24) * c);
                acc += v0; acc += v1; acc += v2; acc += v3;
            }
            for (; q < parts; q += 8) acc += __ldg(pq + static_cast<long long>(q) * c);
        }
        s_p[stripe][ch] = acc;
        __syncthreads();
        if (t < c) s_pool[t] = ((s_p[0][t] + s_p[1][t]) + (s_p[2][t] + s_p[3][t])) + ((s_p[4][t] + s_p[5][t]) + (s_p[6][t] + s_p[7][t]));
        __syncthreads();
    }
    const float* pp = s_pool;
    if (t < c) {
        float acc = s_b1[t];
        for (int k = 0; k < c; ++k) acc = fmaf(s_w1[t * c + k], pp[k], acc);
        s_h[t] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (t < c) {
        float acc = s_b2[t];
        for (int k = 0; k < c; ++k) acc = fmaf(s_w2[t * c + k], s_h[k], acc);
        const float a = 1.f / (1.f + expf(-acc));
        s_a[t] = a;
        if (attn_out && blockIdx.x == 0) attn_out[static_cast<long long>(img) * c + t] = a;
    }
    __syncthreads();
    for (int i = t; i < c * c; i += FHR_THREADS) s_w[i] *= 1.f + s_a[i % c];                 // f*a + f folded into the weights
    if (trace && threadIdx.x == 0) trace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 1] = clock64();
    const int oy0 = blockIdx.x * FHR_ROWS, oy1 = min(oy0 + FHR_ROWS, oh);
    const int ys = lerp_src(oy0, rh, h).i0, ye = lerp_src(oy1 - 1, rh, h).i1;
    const int rows = ye - ys + 1;                                          // <= max_rows (host)
    __syncthreads();
    for (int i = t; i < rows * w; i += FHR_THREADS) {
        const int r = i / w, x = i - r * w;
        const float* fp = f + ((static_cast<long long>(img) * h + ys + r) * w + x) * f_ld;
        float fv[FFM_MAX_C];
#pragma unroll
        for (int k = 0; k < FFM_MAX_C; k += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(fp + k));
            fv[k] = q.x; fv[k + 1] = q.y; fv[k + 2] = q.z; fv[k + 3] = q.w;
        }
        float* zp = s_z + static_cast<long long>(i) * c;
        for (int o = 0; o < c; ++o) {
            float acc = s_b[o];
#pragma unroll
            for (int k = 0; k < FFM_MAX_C; ++k)
                if (k < c) acc = fmaf(s_w[o * c + k], fv[k], acc);
            zp[o] = acc;
        }
    }
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 2] = clock64();
    const long long plane = static_cast<long long>(oh) * ow;
    // Row geometry of the block (shared by every thread): for the x8 head the FHR_ROWS output rows of a block all lie
    // between the same two source rows, and the 4 output pixels of a thread between the same two source columns; then one
    // class costs 4 shared-memory reads for 16 outputs instead of 64 (the write phase was bound by those reads).
    Lerp ly[FHR_ROWS];
    bool rows_uniform = (oy1 - oy0) == FHR_ROWS;
#pragma unroll
    for (int r = 0; r < FHR_ROWS; ++r) {
        ly[r] = lerp_src(min(oy0 + r, oh - 1), rh, h);
        rows_uniform = rows_uniform && ly[r].i0 == ly[0].i0 && ly[r].i1 == ly[0].i1;
    }
    for (int ox = t * 4; ox < ow; ox += FHR_THREADS * 4) {
        Lerp lx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) lx[j] = lerp_src(min(ox + j, ow - 1), rw, w);
        const bool vec = (ox + 4 <= ow) && ((ow & 3) == 0);
        const bool uniform = rows_uniform && vec && lx[1].i0 == lx[0].i0 && lx[2].i0 == lx[0].i0 && lx[3].i0 == lx[0].i0 &&
                             lx[1].i1 == lx[0].i1 && lx[2].i1 == lx[0].i1 && lx[3].i1 == lx[0].i1;
        if (uniform) {
            const float* z00 = s_z + (static_cast<long long>(ly[0].i0 - ys) * w + lx[0].i0) * c;
            const float* z01 = s_z + (static_cast<long long>(ly[0].i0 - ys) * w + lx[0].i1) * c;
            const float* z10 = s_z + (static_cast<long long>(ly[0].i1 - ys) * w + lx[0].i0) * c;
            const float* z11 = s_z + (static_cast<long long>(ly[0].i1 - ys) * w + lx[0].i1) * c;
            float* op = out + static_cast<long long>(img) * c * plane + static_cast<long long>(oy0) * ow + ox;
            for (int ch = 0; ch < c; ++ch) {
                const float a00 = z00[ch], a01 = z01[ch], a10 = z10[ch], a11 = z11[ch];
                float top[4], bot[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    top[j] = lx[j].l0 * a00 + lx[j].l1 * a01;
                    bot[j] = lx[j].l0 * a10 + lx[j].l1 * a11;
                }
#pragma unroll
                for (int r = 0; r < FHR_ROWS; ++r)
                    __stcs(reinterpret_cast<float4*>(op + ch * plane + static_cast<long long>(r) * ow),
                           make_float4(ly[r].l0 * top[0] + ly[r].l1 * bot[0], ly[r].l0 * top[1] + ly[r].l1 * bot[1],
                                       ly[r].l0 * top[2] + ly[r].l1 * bot[2], ly[r].l0 * top[3] + ly[r].l1 * bot[3]));
            }
            continue;
        }
        for (int oy = oy0; oy < oy1; ++oy) {
            const Lerp lyy = lerp_src(oy, rh, h);
            const float* r0 = s_z + static_cast<long long>(lyy.i0 - ys) * w * c;
            const float* r1 = s_z + static_cast<long long>(lyy.i1 - ys) * w * c;
            float* op = out + static_cast<long long>(img) * c * plane + static_cast<long long>(oy) * ow + ox;
            for (int ch = 0; ch < c; ++ch) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[j] = lyy.l0 * (lx[j].l0 * r0[lx[j].i0 * c + ch] + lx[j].l1 * r0[lx[j].i1 * c + ch]) +
                           lyy.l1 * (lx[j].l0 * r1[lx[j].i0 * c + ch] + lx[j].l1 * r1[lx[j].i1 * c + ch]);
                if (vec) {
                    __stcs(reinterpret_cast<float4*>(op + ch * plane), make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (ox + j < ow) op[ch * plane + j] = v[j];
                }
            }
        }
    }
    if (trace && threadIdx.x == 0) trace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 3] = clock64();
}

// ---------------------------------------------------------------- bilinear resize NHWC fp32 -> NCHW fp32
constexpr int RS_THREADS = 256, RS_PX = 4, RS_TILE = RS_THREADS * RS_PX;
__global__ void __launch_bounds__(RS_THREADS)
resize_nchw_kernel(const float* __restrict__ z, int h, int w, int c, int z_ld, int oh, int ow, float rh, float rw,
                   float* __restrict__ out) {
    extern __shared__ float sm[];     // [2][ncols][c]
    const int img = blockIdx.z, oy = blockIdx.y;
    const int ox0 = blockIdx.x * RS_TILE;
    const int ox_last = min(ox0 + RS_TILE, ow) - 1;
    const Lerp ly = lerp_src(oy, rh, h);
    const int xs = lerp_src(ox0, rw, w).i0;
    const int xe = lerp_src(ox_last, rw, w).i1;
    const int ncols = xe - xs + 1;
    float* s0 = sm;
    float* s1 = sm + ncols * c;
    const float* r0 = z + (static_cast<long long>(img) * h + ly.i0) * w * z_ld;
    const float* r1 = z + (static_cast<long long>(img) * h + ly.i1) * w * z_ld;
    for (int i = threadIdx.x; i < ncols * c; i += RS_THREADS) {
        const int col = i / c, ch = i - col * c;
        s0[i] = __ldg(r0 + static_cast<long long>(xs + col) * z_ld + ch);
        s1[i] = __ldg(r1 + static_cast<long long>(xs + col) * z_ld + ch);
    }
    __syncthreads();
    const int ox = ox0 + threadIdx.x * RS_PX;
    if (ox >= ow) return;
    Lerp lx[RS_PX];
#pragma unroll
    for (int j = 0; j < RS_PX; ++j) {
        lx[j] = lerp_src(min(ox + j, ow - 1), rw, w);
        lx[j].i0 -= xs; lx[j].i1 -= xs;
    }
    const long long plane = static_cast<long long>(oh) * ow;
    float* op = out + static_cast<long long>(img) * c * plane + static_cast<long long>(oy) * ow + ox;
    const bool vec = (ox + RS_PX <= ow) && ((ow & 3) == 0);
    for (int ch = 0; ch < c; ++ch) {
        float v[RS_PX];
#pragma unroll
        for (int j = 0; j < RS_PX; ++j) {
            v[j] = ly.l0 * (lx[j].l0 * s0[lx[j].i0 * c + ch] + lx[j].l1 * s0[lx[j].i1 * c + ch]) +
                   ly.l1 * (lx[j].l0 * s1[lx[j].i0 * c + ch] + lx[j].l1 * s1[lx[j].i1 * c + ch]);
        }
        if (vec) {
            __stcs(reinterpret_cast<float4*>(op + ch * plane), make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int j = 0; j < RS_PX; ++j)
                if (ox + j < ow) op[ch * plane + j] = v[j];
        }
    }
}

static int grid_for(long long total, int threads, int waves = 8) {
    long long want = cdiv(total, threads);
    long long cap = static_cast<long long>(waves) * num_sms();
    return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

// ---------------------------------------------------------------- layout converters (standalone sub-module forwards)
// NCHW fp32 [n,c,hw] <-> NHWC [n,hw,ld] (channel offset c_off) of type T, 32x32 tiles through shared memory so that both
// sides are coalesced.  block (32, 8); grid (ceil(hw/32), ceil(c/32), n).
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ x, int c, long long hw, T* __restrict__ y, int ld, int c_off) {
    __shared__ float s[32][33];
    const long long p0 = static_cast<long long>(blockIdx.x) * 32;
    const int c0 = blockIdx.y * 32, img = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = c0 + ty + 8 * j;
        const long long px = p0 + tx;
        s[ty + 8 * j][tx] = (ch < c && px < hw) ? __ldg(x + (static_cast<long long>(img) * c + ch) * hw + px) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long px = p0 + ty + 8 * j;
        const int ch = c0 + tx;
        if (ch < c && px < hw) y[(static_cast<long long>(img) * hw + px) * ld + c_off + ch] = from_f32<T>(s[tx][ty + 8 * j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ x, int ld, int c_off, int c, long long hw, float* __restrict__ y) {
    __shared__ float s[32][33];
    const long long p0 = static_cast<long long>(blockIdx.x) * 32;
    const int c0 = blockIdx.y * 32, img = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long px = p0 + ty + 8 * j;
        const int ch = c0 + tx;
        s[ty + 8 * j][tx] = (ch < c && px < hw) ? to_f32(x[(static_cast<long long>(img) * hw + px) * ld + c_off + ch]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = c0 + ty + 8 * j;
        const long long px = p0 + tx;
        if (ch < c && px < hw) y[(static_cast<long long>(img) * c + ch) * hw + px] = s[tx][ty + 8 * j];
    }
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                             const float* conv_bias, float eps, int c, float* scale, float* shift,
                             rtsds_stream_t s) {
    RTSDS_REQUIRE(mean && var && scale && shift && c > 0, "bn_fold: bad argument");
    bn_fold_kernel<<<static_cast<int>(cdiv(c, 128)), 128, 0, as_stream(s)>>>(gamma, beta, mean, var, conv_bias, eps, c, scale, shift);
    count_launch();
    return check_launch("bn_fold_kernel");
}

extern "C" int rtsds_bn_finalize(const float* stats, double count, const float* gamma, const float* beta, float eps,
                                 float momentum, int c, float* running_mean, float* running_var, float* scale,
                                 float* shift, float* save_mean, float* save_invstd, rtsds_stream_t s) {
    RTSDS_REQUIRE(stats && scale && shift && c > 0 && count > 0, "bn_finalize: bad argument");
    bn_finalize_kernel<<<static_cast<int>(cdiv(c, 128)), 128, 0, as_stream(s)>>>(stats, count, gamma, beta, eps, momentum, c,
                                                                             running_mean, running_var, scale, shift,
                                                                             save_mean, save_invstd);
    count_launch();
    return check_launch("bn_finalize_kernel");
}

extern "C" int rtsds_scale_shift_act(const void* x, const float* scale, const float* shift, const void* residual,
                                     int64_t n_pix, int c, int x_ld, int res_ld, int y_ld, int act, float slope,
                                     int x_dtype, int y_dtype, void* y, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && y && n_pix >= 0 && c > 0, "scale_shift_act: bad argument");
    if (n_pix == 0) return RTSDS_OK;
    const int grid = grid_for(n_pix * ((c + 7) / 8), 256);
    cudaStream_t st = as_stream(s);
    const size_t ex = x_dtype == RTSDS_BF16 ? 2 : 4, ey = y_dtype == RTSDS_BF16 ? 2 : 4;
    const bool vec = c % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && (!residual || res_ld % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(x) % (8 * ex) == 0) && (reinterpret_cast<uintptr_t>(y) % (8 * ey) == 0) &&
                     (!residual || reinterpret_cast<uintptr_t>(residual) % (8 * ey) == 0);
#define SSA(TX, TY)                                                                                                        \
    do {                                                                                                                   \
        if (vec) scale_shift_act_kernel<TX, TY, true><<<grid, 256, 0, st>>>(reinterpret_cast<const TX*>(x), scale, shift,     \
                     reinterpret_cast<const TY*>(residual), n_pix, c, x_ld, res_ld, y_ld, act, slope, reinterpret_cast<TY*>(y)); \
        else scale_shift_act_kernel<TX, TY, false><<<grid, 256, 0, st>>>(reinterpret_cast<const TX*>(x), scale, shift,        \
                     reinterpret_cast<const TY*>(residual), n_pix, c, x_ld, res_ld, y_ld, act, slope, reinterpret_cast<TY*>(y)); \
    } while (0)
    if (x_dtype == RTSDS_BF16 && y_dtype == RTSDS_BF16) SSA(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == RTSDS_F32 && y_dtype == RTSDS_F32) SSA(float, float);
    else if (x_dtype == RTSDS_F32 && y_dtype == RTSDS_BF16) SSA(float, __nv_bfloat16);
    else if (x_dtype == RTSDS_BF16 && y_dtype == RTSDS_F32) SSA(__nv_bfloat16, float);
    else { set_error("scale_shift_act: bad dtype"); return RTSDS_EINVAL; }
#undef SSA
    count_launch();
    return check_launch("scale_shift_act_kernel");
}

extern "C" int rtsds_bn_finalize_apply(const float* stats, double count, const float* gamma, const float* beta, float eps,
                                       float momentum, int c, float* running_mean, float* running_var, float* scale,
                                       float* shift, float* save_mean, float* save_invstd, const void* x,
                                       const void* residual, int64_t n_pix, int x_ld, int res_ld, int y_ld, int act,
                                       float slope, int x_dtype, int y_dtype, void* y, rtsds_stream_t s) {
    RTSDS_REQUIRE(stats && scale && shift && x && y && c > 0 && count > 0 && n_pix >= 0, "bn_finalize_apply: bad argument");
    const size_t ex = x_dtype == RTSDS_BF16 ? 2 : 4, ey = y_dtype == RTSDS_BF16 ? 2 : 4;
    const bool vec = c % 8 == 0 && 256 % (c / 8) == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && (!residual || res_ld % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(x) % (8 * ex) == 0) && (reinterpret_cast<uintptr_t>(y) % (8 * ey) == 0) &&
                     (!residual || reinterpret_cast<uintptr_t>(residual) % (8 * ey) == 0);
    static const bool off = getenv("RTSDS_NO_BN_FUSED_APPLY") != nullptr;
    if (!vec || n_pix == 0 || off) {        // odd channel counts / pitches: the two-launch form
        int rc = rtsds_bn_finalize(stats, count, gamma, beta, eps, momentum, c, running_mean, running_var, scale, shift, save_mean,
                                   save_invstd, s);
        if (rc != RTSDS_OK) return rc;
        return rtsds_scale_shift_act(x, scale, shift, residual, n_pix, c, x_ld, res_ld, y_ld, act, slope, x_dtype, y_dtype, y, s);
    }
    const int grid = grid_for(n_pix * (c / 8), 256);
    cudaStream_t st = as_stream(s);
#define BTA(TX, TY)                                                                                                          \
    bn_train_apply_kernel<TX, TY><<<grid, 256, 2 * c * sizeof(float), st>>>(reinterpret_cast<const TX*>(x), stats, count, gamma, beta, eps, momentum, \
        running_mean, running_var, scale, shift, save_mean, save_invstd, reinterpret_cast<const TY*>(residual), n_pix, c, x_ld,  \
        res_ld, y_ld, act, slope, reinterpret_cast<TY*>(y))
    if (x_dtype == RTSDS_BF16 && y_dtype == RTSDS_BF16) BTA(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == RTSDS_F32 && y_dtype == RTSDS_F32) BTA(float, float);
    else if (x_dtype == RTSDS_F32 && y_dtype == RTSDS_BF16) BTA(float, __nv_bfloat16);
    else if (x_dtype == RTSDS_BF16 && y_dtype == RTSDS_F32) BTA(__nv_bfloat16, float);
    else { set_error("bn_finalize_apply: bad dtype"); return RTSDS_EINVAL; }
#undef BTA
    count_launch();
    return check_launch("bn_train_apply_kernel");
}

extern "C" int rtsds_global_avgpool(const void* x, int n, int64_t hw, int c, int ld, int dtype, float* out,
                                    rtsds_stream_t s) {
    RTSDS_REQUIRE(x && out && n > 0 && hw > 0 && c > 0 && ld >= c, "global_avgpool: bad argument");
    cudaStream_t st = as_stream(s);
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * n * c, st);
    if (e != cudaSuccess) { set_error("global_avgpool: memset: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
    const int cb = static_cast<int>(cdiv(c, 32));
    long long psplit = cdiv(hw, 8 * 16);          // ~16 pixels per thread
    const long long cap = cdiv(4LL * num_sms(), static_cast<long long>(cb) * n);
    if (psplit > cap) psplit = cap;
    if (psplit < 1) psplit = 1;
    dim3 grid(cb, static_cast<unsigned>(psplit), n);
    const float inv = 1.0f / static_cast<float>(hw);
    unsigned long long* det = nullptr;
    if (psplit > 1 && det_mode()) {
        det = det_scratch(st, static_cast<size_t>(n) * c);
        if (!det) return RTSDS_ECUDA;
    }
    if (dtype == RTSDS_BF16)
        gap_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), hw, c, ld, inv, out, det);
    else if (dtype == RTSDS_F16)
        gap_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x), hw, c, ld, inv, out, det);
    else if (dtype == RTSDS_F32)
        gap_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), hw, c, ld, inv, out, det);
    else { set_error("global_avgpool: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    int rc = check_launch("gap_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, out, static_cast<size_t>(n) * c, false, st);
    return rc;
}

extern "C" int rtsds_arm_gate(const float* pooled, const float* w, const float* b, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float eps, float momentum,
                              int train, int n, int c, const float* mul, float* gate, float* lin_out,
                              float* xhat_out, rtsds_stream_t s) {
    RTSDS_REQUIRE(pooled && w && gamma && beta && gate && n > 0 && c > 0, "arm_gate: bad argument");
    RTSDS_REQUIRE(running_mean && running_var, "arm_gate: running stats required");
    if (train) {
        // nn.BatchNorm2d raises "Expected more than 1 value per channel when training" 
        RTSDS_REQUIRE(n >= 2, "arm_gate: train-mode BatchNorm over [N,C,1,1] needs N >= 2 (got %d)", n);
        RTSDS_REQUIRE(n <= ARM_MAX_N, "arm_gate: train-mode batch %d > %d", n, ARM_MAX_N);
    }
    arm_gate_kernel<<<static_cast<int>(cdiv(c, 8)), 256, 0, as_stream(s)>>>(pooled, w, b, gamma, beta, running_mean, running_var,
                                                                        eps, momentum, train, n, c, mul, gate, lin_out, xhat_out);
    count_launch();
    return check_launch("arm_gate_kernel");
}

extern "C" int rtsds_gate_resize_nhwc(const void* src, int n, int h, int w, int c, int src_ld, const float* gate,
                                      float gate_scale, int oh, int ow, void* dst, int dst_ld, int dst_coff, int dtype,
                                      rtsds_stream_t s) {
    RTSDS_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "gate_resize_nhwc: bad argument");
    RTSDS_REQUIRE(c > 0 && c % 8 == 0 && src_ld >= c && dst_ld >= dst_coff + c, "gate_resize_nhwc: bad channel layout");
    RTSDS_REQUIRE(src_ld % 8 == 0 && dst_ld % 8 == 0 && dst_coff % 8 == 0, "gate_resize_nhwc: pitches/offset must be multiples of 8");
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    RTSDS_REQUIRE(static_cast<long long>(n) * oh * ow * (c / 8) < (1LL << 31), "gate_resize_nhwc: tensor too large");
    const int grid = grid_for(static_cast<long long>(n) * oh * ow * (c / 8), 256);
    if (dtype == RTSDS_BF16)
        gate_resize_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(src), n, h, w, c,
                                                                          src_ld, gate, gate_scale, oh, ow, rh, rw,
                                                                          reinterpret_cast<__nv_bfloat16*>(dst), dst_ld, dst_coff,
                                                                          make_fastdiv(c / 8), make_fastdiv(ow), make_fastdiv(oh));
    else if (dtype == RTSDS_F16)
        gate_resize_kernel<__half><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const __half*>(src), n, h, w, c, src_ld, gate, gate_scale, oh, ow, rh, rw,
                                                                   reinterpret_cast<__half*>(dst), dst_ld, dst_coff,
                                                                   make_fastdiv(c / 8), make_fastdiv(ow), make_fastdiv(oh));
    else if (dtype == RTSDS_F32)
        gate_resize_kernel<float><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const float*>(src), n, h, w, c, src_ld, gate, gate_scale,
                                                                  oh, ow, rh, rw, reinterpret_cast<float*>(dst), dst_ld, dst_coff, make_fastdiv(c / 8), make_fastdiv(ow), make_fastdiv(oh));
    else { set_error("gate_resize_nhwc: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    return check_launch("gate_resize_kernel");
}

extern "C" int rtsds_ffm_head(const void* f, int f_dtype, int f_ld, const float* pooled, int n, int64_t hw, int c,
                              const float* w1, const float* b1, const float* w2, const float* b2, const float* wc,
                              const float* bc, float* attn_out, float* z, int z_ld, rtsds_stream_t s) {
    RTSDS_REQUIRE(f && pooled && w1 && b1 && w2 && b2 && z, "ffm_head: NULL argument");
    RTSDS_REQUIRE(n > 0 && hw > 0 && c > 0 && c <= FFM_MAX_C && f_ld >= c && z_ld >= c, "ffm_head: bad shape");
    long long bx = cdiv(hw, FFM_PX);
    const long long cap = cdiv(8LL * num_sms(), n);
    if (bx > cap) bx = cap;
    dim3 grid(static_cast<unsigned>(bx), n);
    if (f_dtype == RTSDS_BF16)
        ffm_head_kernel<__nv_bfloat16><<<grid, FFM_THREADS, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(f), f_ld, pooled, hw,
                                                                       c, w1, b1, w2, b2, wc, bc, attn_out, z, z_ld);
    else if (f_dtype == RTSDS_F32)
        ffm_head_kernel<float><<<grid, FFM_THREADS, 0, as_stream(s)>>>(reinterpret_cast<const float*>(f), f_ld, pooled, hw, c, w1, b1, w2,
                                                               b2, wc, bc, attn_out, z, z_ld);
    else { set_error("ffm_head: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    return check_launch("ffm_head_kernel");
}

extern "C" int rtsds_resize_to_nchw(const float* z, int n, int h, int w, int c, int z_ld, int oh, int ow, float* out,
                                    rtsds_stream_t s) {
    RTSDS_REQUIRE(z && out && n > 0 && h > 0 && w > 0 && c > 0 && z_ld >= c && oh > 0 && ow > 0, "resize_to_nchw: bad argument");
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    // source columns touched by one 1024-pixel tile
    const int max_cols = static_cast<int>(static_cast<double>(RS_TILE) * w / ow) + 4;
    size_t smem = sizeof(float) * 2 * static_cast<size_t>(max_cols < w ? max_cols : w) * c;
    RTSDS_REQUIRE(smem <= 200 * 1024, "resize_to_nchw: tile needs %zu bytes of shared memory", smem);
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(resize_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        done = true;
    }
    dim3 grid(static_cast<unsigned>(cdiv(ow, RS_TILE)), oh, n);
    resize_nchw_kernel<<<grid, RS_THREADS, smem, as_stream(s)>>>(z, h, w, c, z_ld, oh, ow, rh, rw, out);
    count_launch();
    return check_launch("resize_nchw_kernel");
}

// debug: per-block clock64 stamps of arm_gate_resize_kernel (start, pooled vector ready, gates ready, pixels streamed) and of
// ffm_head_resize_kernel (start, attention + weights ready, z rows ready, rows written)
extern "C" void rtsds_debug_arm_trace(void* buf) { g_arm_trace = reinterpret_cast<unsigned long long*>(buf); }

extern "C" int rtsds_arm_gate_resize(const RtsdsArmSide* a3, const RtsdsArmSide* a4, int dtype, int n, int oh, int ow, void* dst,
                                     int dst_ld, rtsds_stream_t s) {
    RTSDS_REQUIRE(a3 && a4 && dst && n > 0 && oh > 0 && ow > 0, "arm_gate_resize: bad argument");
    const RtsdsArmSide* as[2] = {a3, a4};
    ArmSide k[2];
    for (int i = 0; i < 2; ++i) {
        const RtsdsArmSide& a = *as[i];
        RTSDS_REQUIRE(a.src && a.pooled && a.w && a.gamma && a.beta && a.running_mean && a.running_var, "arm_gate_resize: NULL pointer");
        RTSDS_REQUIRE(a.c > 0 && a.c % 32 == 0 && a.h > 0 && a.w_in > 0 && a.dst_coff % 8 == 0 && a.dst_coff + a.c <= dst_ld,
                      "arm_gate_resize: channels must be a multiple of 32 and fit the destination slot");
        RTSDS_REQUIRE(a.pooled_parts >= 1, "arm_gate_resize: pooled_parts must be >= 1");
        k[i] = ArmSide{a.src, a.pooled, a.w, a.b, a.gamma, a.beta, a.running_mean, a.running_var, a.eps, a.out_scale,
                       a.h, a.w_in, a.c, a.dst_coff, a.mul_pooled, a.pooled_parts};
    }
    RTSDS_REQUIRE(dst_ld % 8 == 0, "arm_gate_resize: dst_ld must be a multiple of 8");
    const int groups = a3->c / 32 + a4->c / 32;
    const size_t sm = sizeof(float) * static_cast<size_t>(a3->c > a4->c ? a3->c : a4->c);
    RTSDS_REQUIRE(sm <= 40 * 1024, "arm_gate_resize: too many channels");
    // exactly ONE wave: every block costs ~10 us of mostly latency (pooled vector, gates, gathers), so a handful of blocks
    // left over for a second wave doubles the kernel (measured: 312 blocks on 296 slots = 23 us, per-block time 10 us)
    static int occ[3] = {0, 0, 0};
    const int oi = dtype == RTSDS_F16 ? 0 : (dtype == RTSDS_BF16 ? 1 : 2);
    if (!occ[oi]) {
        int o = 0;
        if (oi == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, arm_gate_resize_kernel<__half>, 256, 40 * 1024);
        else if (oi == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, arm_gate_resize_kernel<__nv_bfloat16>, 256, 40 * 1024);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, arm_gate_resize_kernel<float>, 256, 40 * 1024);
        occ[oi] = o > 0 ? o : 1;
    }
    int chunks = static_cast<int>(static_cast<long long>(occ[oi]) * num_sms() / groups);
    if (chunks < 1) chunks = 1;
    const int npix = oh * ow;
    if (chunks > npix / 64) chunks = npix / 64 > 0 ? npix / 64 : 1;
    const int blocks0 = (a3->c / 32) * chunks, blocks = groups * chunks;
    cudaStream_t st = as_stream(s);
    if (dtype == RTSDS_F16)
        launch_pdl(arm_gate_resize_kernel<__half>, dim3(blocks), dim3(256), sm, st, k[0], k[1], blocks0, chunks, n, oh, ow, reinterpret_cast<__half*>(dst), dst_ld, g_arm_trace);
    else if (dtype == RTSDS_BF16)
        launch_pdl(arm_gate_resize_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), sm, st, k[0], k[1], blocks0, chunks, n, oh, ow, reinterpret_cast<__nv_bfloat16*>(dst), dst_ld, g_arm_trace);
    else if (dtype == RTSDS_F32)
        launch_pdl(arm_gate_resize_kernel<float>, dim3(blocks), dim3(256), sm, st, k[0], k[1], blocks0, chunks, n, oh, ow, reinterpret_cast<float*>(dst), dst_ld, g_arm_trace);
    else { set_error("arm_gate_resize: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    return check_launch("arm_gate_resize_kernel");
}

extern "C" int rtsds_ffm_head_resize(const float* f, int f_ld, const float* pooled, int n, int h, int w, int c, const float* w1,
                                     const float* b1, const float* w2, const float* b2, const float* wc, const float* bc,
                                     float* attn_out, int pooled_parts, int oh, int ow, float* out, rtsds_stream_t s) {
    RTSDS_REQUIRE(f && pooled && w1 && b1 && w2 && b2 && wc && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && pooled_parts >= 1,
                  "ffm_head_resize: bad argument");
    RTSDS_REQUIRE(c > 0 && c <= FFM_MAX_C && f_ld >= FFM_MAX_C && f_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0,
                  "ffm_head_resize: c <= %d, feature pitch >= %d floats and 16-byte aligned", FFM_MAX_C, FFM_MAX_C);
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    const int max_rows = static_cast<int>(static_cast<double>(FHR_ROWS) * h / oh) + 3;
    const size_t smem = sizeof(float) * static_cast<size_t>(max_rows) * w * c;
    RTSDS_REQUIRE(smem <= 160 * 1024, "ffm_head_resize: %zu bytes of shared memory needed for %d source rows", smem, max_rows);
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(ffm_head_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); done = true; }
    dim3 grid(static_cast<unsigned>(cdiv(oh, FHR_ROWS)), n);
    launch_pdl(ffm_head_resize_kernel, grid, dim3(FHR_THREADS), smem, as_stream(s), f, f_ld, pooled, h, w, c, w1, b1, w2, b2, wc, bc, attn_out,
               oh, ow, rh, rw, max_rows, pooled_parts, out, g_arm_trace);
    count_launch();
    return check_launch("ffm_head_resize_kernel");
}

extern "C" int rtsds_nchw_to_nhwc(const float* x, int n, int c, int64_t hw, int dtype, void* y, int ld, int c_off, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && y && n > 0 && c > 0 && hw > 0 && ld >= c_off + c && c_off >= 0 && n <= 65535, "nchw_to_nhwc: bad argument");
    dim3 grid(static_cast<unsigned>(cdiv(hw, 32)), static_cast<unsigned>(cdiv(c, 32)), n), block(32, 8);
    if (dtype == RTSDS_F16) nchw_to_nhwc_kernel<__half><<<grid, block, 0, as_stream(s)>>>(x, c, hw, reinterpret_cast<__half*>(y), ld, c_off);
    else if (dtype == RTSDS_BF16) nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, as_stream(s)>>>(x, c, hw, reinterpret_cast<__nv_bfloat16*>(y), ld, c_off);
    else if (dtype == RTSDS_F32) nchw_to_nhwc_kernel<float><<<grid, block, 0, as_stream(s)>>>(x, c, hw, reinterpret_cast<float*>(y), ld, c_off);
    else { set_error("nchw_to_nhwc: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int rtsds_nhwc_to_nchw(const void* x, int dtype, int ld, int c_off, int n, int c, int64_t hw, float* y, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && y && n > 0 && c > 0 && hw > 0 && ld >= c_off + c && c_off >= 0 && n <= 65535, "nhwc_to_nchw: bad argument");
    dim3 grid(static_cast<unsigned>(cdiv(hw, 32)), static_cast<unsigned>(cdiv(c, 32)), n), block(32, 8);
    if (dtype == RTSDS_F16) nhwc_to_nchw_kernel<__half><<<grid, block, 0, as_stream(s)>>>(reinterpret_cast<const __half*>(x), ld, c_off, c, hw, y);
    else if (dtype == RTSDS_BF16) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, c_off, c, hw, y);
    else if (dtype == RTSDS_F32) nhwc_to_nchw_kernel<float><<<grid, block, 0, as_stream(s)>>>(reinterpret_cast<const float*>(x), ld, c_off, c, hw, y);
    else { set_error("nhwc_to_nchw: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    return check_launch("nhwc_to_nchw_kernel");
}
