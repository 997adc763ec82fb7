// Discriminator-specific kernels of the adversarial step (reference
// models/domain_shift/adversarial/model.py:30-83 driven by train.py:218-263).
//
//   * s2d_fwd / s2d_bwd: the discriminator's conv1 is a 4x4 stride-2 pad-1 conv over the
//     num_classes-channel probability map [N,19,H,W] fp32 NCHW (train.py:225 F.softmax).  It is
//     rewritten as a 2x2 stride-1 conv over the space-to-depth form of the padded input:
//         xs[n, i, j, (ph*2+pw)*32 + c] = x[n, c, 2i+ph-1, 2j+pw-1]      (zero outside the image)
//     an NHWC tensor with 4 parity planes x 32 channels = 128 channels, which is exactly two 64-channel
//     K tiles of the tcgen05 implicit-GEMM conv (csrc/conv_tc.cu) with dense stride-1 taps.  The
//     converter fuses the softmax over the class channel (so the probabilities never have to be
//     materialised in NCHW) and the bf16 cast; the adjoint converter fuses the softmax backward.
//   * act_bwd: LeakyReLU backward + bias gradient in one pass.
//   * disc_cls_*: classifier conv (Cout = 1, 4x4 s2 p1, bias) followed by AdaptiveAvgPool2d(1)
//     (model.py:49,52,58-60 / :73,76,80-81).  mean_pixels(conv(x)) is linear in x, so it collapses to
//     16 per-tap sums of x:  out[n] = b + 1/P * sum_{t,c} w[c,t] * S[n,t,c],
//     S[n,(r,s),c] = sum over the input pixels tap (r,s) visits.  One read of x forward; backward is
//     dW = g/P * S and dx = g/P * Weff(h,w,c) written together with the LeakyReLU mask of the
//     preceding layer.  HBM-bound, no tensor cores.
//   * bce_logits: nn.BCEWithLogitsLoss (main.py:132-134) against a constant target on the N logits.
#include "common.cuh"

namespace rtsds {

constexpr int S2D_CP = 32;            // channels per parity plane
constexpr int S2D_C = 4 * S2D_CP;     // channels of the space-to-depth tensor
constexpr int S2D_TJ = 32;            // s2d pixels per block along W (= 64 image columns)

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

// x: fp32 NCHW [n,c,h,w]; out: NHWC [n,hs,ws,128]
template <typename T>
__global__ void __launch_bounds__(128)
s2d_fwd_kernel(const float* __restrict__ x, int c, int h, int w, int hs, int ws, int softmax, T* __restrict__ out) {
    __shared__ __align__(16) T s_out[S2D_TJ * S2D_C];
    const int i = blockIdx.y, img = blockIdx.z, j0 = blockIdx.x * S2D_TJ;
    const int rr = threadIdx.x >> 6, cc = threadIdx.x & 63;
    const int hh = 2 * i - 1 + rr, ww = 2 * j0 - 1 + cc;
    const bool inb = hh >= 0 && hh < h && ww >= 0 && ww < w;
    const long long plane = static_cast<long long>(h) * w;
    const float* xp = x + static_cast<long long>(img) * c * plane + static_cast<long long>(hh) * w + ww;
    float v[S2D_CP];
#pragma unroll
    for (int ch = 0; ch < S2D_CP; ++ch) v[ch] = (inb && ch < c) ? __ldg(xp + ch * plane) : 0.0f;
    if (softmax && inb) {
        float m = -INFINITY;
#pragma unroll
        for (int ch = 0; ch < S2D_CP; ++ch) if (ch < c) m = fmaxf(m, v[ch]);
        float sum = 0.0f;
#pragma unroll
        for (int ch = 0; ch < S2D_CP; ++ch) {
            v[ch] = ch < c ? (sizeof(T) == 4 ? expf(v[ch] - m) : __expf(v[ch] - m)) : 0.0f;
            sum += v[ch];
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int ch = 0; ch < S2D_CP; ++ch) v[ch] *= inv;
    }
    T* dst = s_out + ((cc >> 1) * 4 + rr * 2 + (cc & 1)) * S2D_CP;
#pragma unroll
    for (int ch = 0; ch < S2D_CP; ++ch) dst[ch] = from_f32<T>(v[ch]);
    __syncthreads();
    const int nj = min(S2D_TJ, ws - j0);
    constexpr int VN = Vec16<T>::N;
    const int nvec = nj * S2D_C / VN;
    uint4* gdst = reinterpret_cast<uint4*>(out + ((static_cast<long long>(img) * hs + i) * ws + j0) * S2D_C);
    const uint4* ssrc = reinterpret_cast<const uint4*>(s_out);
    for (int k = threadIdx.x; k < nvec; k += 128) gdst[k] = ssrc[k];
}

// g: fp32 NHWC [n,hs,ws,128] gradient w.r.t. xs (kept in fp32: the softmax backward subtracts nearly equal
// numbers); p: the forward's xs (probabilities) when the softmax was fused, else NULL.
// dx: fp32 NCHW [n,c,h,w]:  softmax ? p*(g - sum_k g_k p_k) : g.
template <typename T>
__global__ void __launch_bounds__(128)
s2d_bwd_kernel(const float* __restrict__ g, const T* __restrict__ p, int c, int h, int w, int hs, int ws,
               float* __restrict__ dx) {
    __shared__ __align__(16) float s_g[S2D_TJ * S2D_C];
    __shared__ __align__(16) T s_p[S2D_TJ * S2D_C];
    const int i = blockIdx.y, img = blockIdx.z, j0 = blockIdx.x * S2D_TJ;
    const int nj = min(S2D_TJ, ws - j0);
    const long long base = ((static_cast<long long>(img) * hs + i) * ws + j0) * S2D_C;
    {
        const uint4* gs = reinterpret_cast<const uint4*>(g + base);
        uint4* sd = reinterpret_cast<uint4*>(s_g);
        for (int k = threadIdx.x; k < nj * S2D_C / 4; k += 128) sd[k] = __ldg(gs + k);
        if (p) {
            constexpr int VN = Vec16<T>::N;
            const uint4* ps = reinterpret_cast<const uint4*>(p + base);
            uint4* pd = reinterpret_cast<uint4*>(s_p);
            for (int k = threadIdx.x; k < nj * S2D_C / VN; k += 128) pd[k] = __ldg(ps + k);
        }
    }
    __syncthreads();
    const int rr = threadIdx.x >> 6, cc = threadIdx.x & 63;
    const int hh = 2 * i - 1 + rr, ww = 2 * j0 - 1 + cc;
    if (!(hh >= 0 && hh < h && ww >= 0 && ww < w)) return;
    const int slot = ((cc >> 1) * 4 + rr * 2 + (cc & 1)) * S2D_CP;
    float gv[S2D_CP];
#pragma unroll
    for (int ch = 0; ch < S2D_CP; ++ch) gv[ch] = ch < c ? s_g[slot + ch] : 0.0f;
    if (p) {
        float pv[S2D_CP];
        float dot = 0.0f;
#pragma unroll
        for (int ch = 0; ch < S2D_CP; ++ch) {
            pv[ch] = ch < c ? to_f32(s_p[slot + ch]) : 0.0f;
            dot += gv[ch] * pv[ch];
        }
#pragma unroll
        for (int ch = 0; ch < S2D_CP; ++ch) gv[ch] = pv[ch] * (gv[ch] - dot);
    }
    const long long plane = static_cast<long long>(h) * w;
    float* dp = dx + static_cast<long long>(img) * c * plane + static_cast<long long>(hh) * w + ww;
#pragma unroll
    for (int ch = 0; ch < S2D_CP; ++ch) if (ch < c) dp[ch * plane] = gv[ch];
}

// w[co][c][4][4] -> w2[co][128][2][2]:  w2[co, (ph*2+pw)*32+c, a, b] = w[co, c, 2a+ph, 2b+pw]
__global__ void s2d_weight_kernel(const float* __restrict__ w, int cout, int c, float* __restrict__ w2) {
    const int total = cout * S2D_C * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ab = i & 3, a = ab >> 1, b = ab & 1;
        const int k = (i >> 2) % S2D_C, co = (i >> 2) / S2D_C;
        const int par = k / S2D_CP, ch = k % S2D_CP, ph = par >> 1, pw = par & 1;
        w2[i] = ch < c ? w[((co * c + ch) * 4 + 2 * a + ph) * 4 + 2 * b + pw] : 0.0f;
    }
}
// gw[co][c][r][s] += gw2[co, ((r&1)*2+(s&1))*32+c, r>>1, s>>1]
__global__ void s2d_weight_grad_kernel(const float* __restrict__ gw2, int cout, int c, float* __restrict__ gw) {
    const int total = cout * c * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int s = i & 3, r = (i >> 2) & 3;
        const int ch = (i >> 4) % c, co = (i >> 4) / c;
        const int k = ((r & 1) * 2 + (s & 1)) * S2D_CP + ch;
        gw[i] += gw2[((co * S2D_C + k) * 2 + (r >> 1)) * 2 + (s >> 1)];
    }
}

// ---- 8-wide helpers (as in backward.cu) ----
struct G8 { float v[8]; };
__device__ __forceinline__ G8 ldg8(const __nv_bfloat16* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    G8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ G8 ldg8(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    G8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void stg8(__nv_bfloat16* p, const G8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void stg8(float* p, const G8& r) {
    reinterpret_cast<float4*>(p)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// d_raw = dy * (y > 0 ? 1 : slope)  (act: RTSDS_ACT_RELU -> slope 0);  dbias[c] += sum d_raw
template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* dy, int dy_ld, const T* __restrict__ y, int y_ld, long long n_pix, int c,
               float slope, T* d_raw, int d_raw_ld, float* dbias, unsigned long long* det) {
    extern __shared__ float sh[];             // [c]
    for (int i = threadIdx.x; i < c; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int cg = c / 8;
    const int g8 = threadIdx.x % cg;
    const int prow = threadIdx.x / cg, prows = blockDim.x / cg;
    float s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = 0.f;
    if (prow < prows) {
        for (long long p = static_cast<long long>(blockIdx.x) * prows + prow; p < n_pix; p += static_cast<long long>(gridDim.x) * prows) {
            const G8 d = ldg8(dy + p * dy_ld + g8 * 8);
            const G8 yy = ldg8(y + p * y_ld + g8 * 8);
            G8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o.v[j] = yy.v[j] > 0.f ? d.v[j] : d.v[j] * slope;
                s1[j] += o.v[j];
            }
            stg8(d_raw + p * d_raw_ld + g8 * 8, o);
        }
        if (dbias && det) {          // deterministic mode (common.cuh): this thread's sums into the exact accumulators
#pragma unroll
            for (int j = 0; j < 8; ++j) det_add(det + 2 * (g8 * 8 + j), s1[j]);
        } else if (dbias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&sh[g8 * 8 + j], s1[j]);
        }
    }
    if (dbias && !det) {
        __syncthreads();
        for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&dbias[i], sh[i]);
    }
}

// ---- classifier (Cout = 1, 4x4 s2 p1) + global average pool ----
// tapsum[n][r*4+s][c] += sum of x[n,hh,ww,c] over the pixels tap (r,s) visits; one block per image row.
template <typename T>
__global__ void __launch_bounds__(256)
disc_cls_tapsum_kernel(const T* __restrict__ x, int ld, int h, int w, int c, int oh, int ow, float* __restrict__ tapsum,
                       unsigned long long* det) {
    extern __shared__ float sh[];             // [4][c]
    const int hh = blockIdx.x, img = blockIdx.y;
    for (int i = threadIdx.x; i < 4 * c; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int cg = c / 8;
    const int g8 = threadIdx.x % cg;
    const int lane = threadIdx.x / cg, lanes = blockDim.x / cg;
    float a0[8], a1[8], a2[8], a3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; }
    if (lane < lanes) {
        const T* row = x + (static_cast<long long>(img) * h + hh) * w * ld + g8 * 8;
        for (int ww = lane; ww < w; ww += lanes) {
            const G8 v = ldg8(row + static_cast<long long>(ww) * ld);
            const int s0 = (ww + 1) & 1;
            const int oa = (ww + 1 - s0) >> 1;           // output column of tap s0; tap s0+2 visits oa-1
            const bool va = oa < ow, vb = oa >= 1 && oa - 1 < ow;
            if (s0 == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { if (va) a0[j] += v.v[j]; if (vb) a2[j] += v.v[j]; }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) { if (va) a1[j] += v.v[j]; if (vb) a3[j] += v.v[j]; }
            }
        }
        if (det) {
            // deterministic mode: this thread's four column-tap sums straight into the exact accumulators of the (up to two)
            // row taps this input row feeds; nothing goes through the block's shared-memory atomics
            const int r0d = (hh + 1) & 1;
            const int oad = (hh + 1 - r0d) >> 1;
            const bool vad = oad < oh, vbd = oad >= 1 && oad - 1 < oh;
            unsigned long long* dd = det + 2 * (static_cast<long long>(img) * 16 * c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = g8 * 8 + j;
                const float av[4] = {a0[j], a1[j], a2[j], a3[j]};
#pragma unroll
                for (int sx = 0; sx < 4; ++sx) {
                    if (vad) det_add(dd + 2 * ((r0d * 4 + sx) * c + ch), av[sx]);
                    if (vbd) det_add(dd + 2 * (((r0d + 2) * 4 + sx) * c + ch), av[sx]);
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sh[0 * c + g8 * 8 + j], a0[j]);
            atomicAdd(&sh[1 * c + g8 * 8 + j], a1[j]);
            atomicAdd(&sh[2 * c + g8 * 8 + j], a2[j]);
            atomicAdd(&sh[3 * c + g8 * 8 + j], a3[j]);
        }
        }
    }
    if (det) return;
    __syncthreads();
    const int r0 = (hh + 1) & 1;
    const int oa = (hh + 1 - r0) >> 1;
    const bool va = oa < oh, vb = oa >= 1 && oa - 1 < oh;
    float* dst = tapsum + static_cast<long long>(img) * 16 * c;
    for (int i = threadIdx.x; i < 4 * c; i += blockDim.x) {
        const int s = i / c, ch = i - s * c;
        const float v = sh[i];
        if (va) atomicAdd(&dst[(r0 * 4 + s) * c + ch], v);
        if (vb) atomicAdd(&dst[((r0 + 2) * 4 + s) * c + ch], v);
    }
}

// out[n] = b + inv_p * sum_{t,c} w[c*16+t] * tapsum[n][t][c]
__global__ void __launch_bounds__(256)
disc_cls_finish_kernel(const float* __restrict__ tapsum, const float* __restrict__ w, const float* __restrict__ bias,
                       int c, float inv_p, float* __restrict__ out) {
    __shared__ float red[8];
    const int img = blockIdx.x;
    const float* ts = tapsum + static_cast<long long>(img) * 16 * c;
    float acc = 0.f;
    for (int i = threadIdx.x; i < 16 * c; i += blockDim.x) {
        const int t = i / c, ch = i - t * c;
        acc += ts[i] * __ldg(w + ch * 16 + t);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k];
        out[img] = t * inv_p + (bias ? bias[0] : 0.f);
    }
}

// dW[c*16+t] += inv_p * sum_n g[n] * tapsum[n][t][c];  dbias += sum_n g[n]
__global__ void disc_cls_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ tapsum, int n, int c,
                                      float inv_p, float g_scale, float* dw, float* dbias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 16 * c) {
        const int t = i / c, ch = i - t * c;
        float acc = 0.f;
        for (int k = 0; k < n; ++k) acc += g[k] * tapsum[(static_cast<long long>(k) * 16 + t) * c + ch];
        if (dw) dw[ch * 16 + t] += acc * inv_p * g_scale;
    }
    if (i == 0 && dbias) {
        float acc = 0.f;
        for (int k = 0; k < n; ++k) acc += g[k];
        dbias[0] += acc * g_scale;
    }
}

// dx[n,hh,ww,c] = g[n]*inv_p * sum over the taps that visit (hh,ww) of w[c,r,s]; fused with the preceding
// LeakyReLU: d_raw = dx * (x > 0 ? 1 : slope) (x is that layer's activation output); dbias_prev += sum d_raw.
template <typename T>
__global__ void __launch_bounds__(256)
disc_cls_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ wgt, const T* __restrict__ x, int ld, int h,
                      int w, int c, int oh, int ow, float inv_p, float g_scale, int masked, float slope,
                      T* __restrict__ d_raw, int d_ld, float* dbias_prev, unsigned long long* det) {
    extern __shared__ float sh[];             // [4][c] row-combined weights, [c] bias sums
    float* s_w = sh;
    float* s_b = sh + 4 * c;
    const int hh = blockIdx.x, img = blockIdx.y;
    const int r0 = (hh + 1) & 1;
    const int oa_r = (hh + 1 - r0) >> 1;
    const bool ra = oa_r < oh, rb = oa_r >= 1 && oa_r - 1 < oh;
    for (int i = threadIdx.x; i < 4 * c; i += blockDim.x) {
        const int s = i / c, ch = i - s * c;
        float v = 0.f;
        if (ra) v += __ldg(wgt + ch * 16 + r0 * 4 + s);
        if (rb) v += __ldg(wgt + ch * 16 + (r0 + 2) * 4 + s);
        s_w[i] = v;
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) s_b[i] = 0.f;
    __syncthreads();
    const float gs = g[img] * inv_p * g_scale;
    const int cg = c / 8;
    const int g8 = threadIdx.x % cg;
    const int lane = threadIdx.x / cg, lanes = blockDim.x / cg;
    float bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bs[j] = 0.f;
    if (lane < lanes) {
        const long long rowoff = (static_cast<long long>(img) * h + hh) * w;
        for (int ww = lane; ww < w; ww += lanes) {
            const int s0 = (ww + 1) & 1;
            const int oa = (ww + 1 - s0) >> 1;
            const bool va = oa < ow, vb = oa >= 1 && oa - 1 < ow;
            G8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float wv = 0.f;
                if (va) wv += s_w[s0 * c + g8 * 8 + j];
                if (vb) wv += s_w[(s0 + 2) * c + g8 * 8 + j];
                o.v[j] = wv * gs;
            }
            if (masked) {
                const G8 xv = ldg8(x + (rowoff + ww) * ld + g8 * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] = xv.v[j] > 0.f ? o.v[j] : o.v[j] * slope;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[j] += o.v[j];
            stg8(d_raw + (rowoff + ww) * d_ld + g8 * 8, o);
        }
        if (dbias_prev && det) {     // deterministic mode (common.cuh)
#pragma unroll
            for (int j = 0; j < 8; ++j) det_add(det + 2 * (g8 * 8 + j), bs[j]);
        } else if (dbias_prev) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&s_b[g8 * 8 + j], bs[j]);
        }
    }
    if (dbias_prev && !det) {
        __syncthreads();
        for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&dbias_prev[i], s_b[i]);
    }
}

// loss = scale * mean_n( max(x,0) - x*t + log1p(exp(-|x|)) );  dlogit[n] = scale * (sigmoid(x) - t) / n
__global__ void bce_logits_kernel(const float* __restrict__ logit, int n, float target, float scale, float* loss,
                                  float* dlogit) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = logit[i];
        acc += fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x)));
        if (dlogit) dlogit[i] = scale * (1.0f / (1.0f + expf(-x)) - target) / static_cast<float>(n);
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float t = 0.f;
        for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) t += red[k];
        loss[0] = scale * t / static_cast<float>(n);
    }
}

static int cls_threads(int c) {
    // threads = (c/8) channel groups x pixel lanes, at most 256
    const int cg = c / 8;
    int lanes = 256 / cg;
    if (lanes < 1) lanes = 1;
    return cg * lanes;
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_s2d_out_size(int in_size) { return in_size / 2 + 1; }

extern "C" int rtsds_s2d_fwd(const float* x, int n, int c, int h, int w, int softmax, int dtype, void* out,
                             rtsds_stream_t s) {
    RTSDS_REQUIRE(x && out && n > 0 && c > 0 && c <= S2D_CP && h >= 2 && w >= 2, "s2d_fwd: bad argument (c <= 32)");
    RTSDS_REQUIRE(dtype == RTSDS_BF16 || dtype == RTSDS_F32, "s2d_fwd: bad dtype");
    RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "s2d_fwd: alignment");
    const int hs = h / 2 + 1, ws = w / 2 + 1;
    dim3 grid(static_cast<unsigned>(cdiv(ws, S2D_TJ)), hs, n);
    if (dtype == RTSDS_BF16)
        s2d_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(s)>>>(x, c, h, w, hs, ws, softmax, reinterpret_cast<__nv_bfloat16*>(out));
    else
        s2d_fwd_kernel<float><<<grid, 128, 0, as_stream(s)>>>(x, c, h, w, hs, ws, softmax, reinterpret_cast<float*>(out));
    count_launch();
    return check_launch("s2d_fwd_kernel");
}

extern "C" int rtsds_s2d_bwd(const float* g, const void* p, int n, int c, int h, int w, int dtype, float* dx,
                             rtsds_stream_t s) {
    RTSDS_REQUIRE(g && dx && n > 0 && c > 0 && c <= S2D_CP && h >= 2 && w >= 2, "s2d_bwd: bad argument (c <= 32)");
    RTSDS_REQUIRE(dtype == RTSDS_BF16 || dtype == RTSDS_F32, "s2d_bwd: bad dtype (of p)");
    const int hs = h / 2 + 1, ws = w / 2 + 1;
    dim3 grid(static_cast<unsigned>(cdiv(ws, S2D_TJ)), hs, n);
    if (dtype == RTSDS_BF16)
        s2d_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(s)>>>(g, reinterpret_cast<const __nv_bfloat16*>(p), c, h, w, hs, ws, dx);
    else
        s2d_bwd_kernel<float><<<grid, 128, 0, as_stream(s)>>>(g, reinterpret_cast<const float*>(p), c, h, w, hs, ws, dx);
    count_launch();
    return check_launch("s2d_bwd_kernel");
}

extern "C" int rtsds_s2d_weight(const float* w_oihw, int cout, int c, float* w2, rtsds_stream_t s) {
    RTSDS_REQUIRE(w_oihw && w2 && cout > 0 && c > 0 && c <= S2D_CP, "s2d_weight: bad argument");
    const int total = cout * S2D_C * 4;
    s2d_weight_kernel<<<static_cast<int>(cdiv(total, 256)), 256, 0, as_stream(s)>>>(w_oihw, cout, c, w2);
    count_launch();
    return check_launch("s2d_weight_kernel");
}

extern "C" int rtsds_s2d_weight_grad(const float* gw2, int cout, int c, float* gw_oihw, rtsds_stream_t s) {
    RTSDS_REQUIRE(gw2 && gw_oihw && cout > 0 && c > 0 && c <= S2D_CP, "s2d_weight_grad: bad argument");
    const int total = cout * c * 16;
    s2d_weight_grad_kernel<<<static_cast<int>(cdiv(total, 256)), 256, 0, as_stream(s)>>>(gw2, cout, c, gw_oihw);
    count_launch();
    return check_launch("s2d_weight_grad_kernel");
}

extern "C" int rtsds_act_bwd(const void* dy, int dy_ld, const void* y, int y_ld, int64_t n_pix, int c, int act,
                             float slope, int dtype, void* d_raw, int d_raw_ld, float* dbias, rtsds_stream_t s) {
    RTSDS_REQUIRE(dy && y && d_raw && n_pix > 0 && c > 0 && c % 8 == 0 && c <= 2048, "act_bwd: bad argument (c % 8 == 0)");
    RTSDS_REQUIRE(dy_ld >= c && y_ld >= c && d_raw_ld >= c && dy_ld % 8 == 0 && y_ld % 8 == 0 && d_raw_ld % 8 == 0, "act_bwd: pitches");
    RTSDS_REQUIRE(act == RTSDS_ACT_RELU || act == RTSDS_ACT_LRELU, "act_bwd: activation");
    if (act == RTSDS_ACT_RELU) slope = 0.f;
    const int threads = cls_threads(c);
    const int prows = threads / (c / 8);
    long long want = cdiv(n_pix, static_cast<long long>(prows) * 4);
    const long long cap = 8LL * num_sms();
    const int grid = static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
    const size_t sm = sizeof(float) * c;
    unsigned long long* det = nullptr;
    if (dbias && det_mode()) {
        det = det_scratch(as_stream(s), static_cast<size_t>(c));
        if (!det) return RTSDS_ECUDA;
    }
    if (dtype == RTSDS_BF16)
        act_bwd_kernel<__nv_bfloat16><<<grid, threads, sm, as_stream(s)>>>(
            reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, reinterpret_cast<const __nv_bfloat16*>(y), y_ld, n_pix, c, slope,
            reinterpret_cast<__nv_bfloat16*>(d_raw), d_raw_ld, dbias, det);
    else if (dtype == RTSDS_F32)
        act_bwd_kernel<float><<<grid, threads, sm, as_stream(s)>>>(reinterpret_cast<const float*>(dy), dy_ld,
                                                                  reinterpret_cast<const float*>(y), y_ld, n_pix, c, slope,
                                                                  reinterpret_cast<float*>(d_raw), d_raw_ld, dbias, det);
    else { set_error("act_bwd: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    int rc = check_launch("act_bwd_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, dbias, static_cast<size_t>(c), true, as_stream(s));
    return rc;
}

extern "C" int rtsds_disc_cls_fwd(const void* x, int ld, int dtype, int n, int h, int w, int c, const float* w_oihw,
                                  const float* bias, float* tapsum, float* out, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && w_oihw && tapsum && out && n > 0 && h >= 2 && w >= 2, "disc_cls_fwd: bad argument");
    RTSDS_REQUIRE(c > 0 && c % 8 == 0 && c <= 2048 && ld >= c && ld % 8 == 0, "disc_cls_fwd: c must be a multiple of 8, <= 2048");
    const int oh = h / 2, ow = w / 2;          // 4x4, stride 2, pad 1
    cudaError_t e = cudaMemsetAsync(tapsum, 0, sizeof(float) * n * 16 * c, as_stream(s));
    if (e != cudaSuccess) { set_error("disc_cls_fwd: memset: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
    const int threads = cls_threads(c);
    dim3 grid(h, n);
    const size_t sm = sizeof(float) * 4 * c;
    unsigned long long* det = nullptr;
    const size_t n_ts = static_cast<size_t>(n) * 16 * c;
    if (det_mode()) {
        det = det_scratch(as_stream(s), n_ts);
        if (!det) return RTSDS_ECUDA;
    }
    if (dtype == RTSDS_BF16)
        disc_cls_tapsum_kernel<__nv_bfloat16><<<grid, threads, sm, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, h, w, c, oh, ow, tapsum, det);
    else if (dtype == RTSDS_F32)
        disc_cls_tapsum_kernel<float><<<grid, threads, sm, as_stream(s)>>>(reinterpret_cast<const float*>(x), ld, h, w, c, oh, ow, tapsum, det);
    else { set_error("disc_cls_fwd: bad dtype"); return RTSDS_EINVAL; }
    count_launch();
    int rc = check_launch("disc_cls_tapsum_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, tapsum, n_ts, true, as_stream(s));
    if (rc != RTSDS_OK) return rc;
    disc_cls_finish_kernel<<<n, 256, 0, as_stream(s)>>>(tapsum, w_oihw, bias, c, 1.0f / (static_cast<float>(oh) * ow), out);
    count_launch();
    return check_launch("disc_cls_finish_kernel");
}

extern "C" int rtsds_disc_cls_bwd(const float* g, float g_scale, const float* tapsum, const float* w_oihw, const void* x,
                                  int ld, int dtype, int n, int h, int w, int c, int masked, float slope, void* d_raw,
                                  int d_ld, float* dbias_prev, float* dw, float* dbias, rtsds_stream_t s) {
    RTSDS_REQUIRE(g && tapsum && w_oihw && n > 0 && h >= 2 && w >= 2, "disc_cls_bwd: bad argument");
    RTSDS_REQUIRE(c > 0 && c % 8 == 0 && c <= 2048, "disc_cls_bwd: c must be a multiple of 8, <= 2048");
    const int oh = h / 2, ow = w / 2;
    const float inv_p = 1.0f / (static_cast<float>(oh) * ow);
    if (dw || dbias) {
        disc_cls_wgrad_kernel<<<static_cast<int>(cdiv(16 * c, 256)), 256, 0, as_stream(s)>>>(g, tapsum, n, c, inv_p, g_scale, dw, dbias);
        count_launch();
        int rc = check_launch("disc_cls_wgrad_kernel");
        if (rc != RTSDS_OK) return rc;
    }
    if (d_raw) {
        RTSDS_REQUIRE(d_ld >= c && d_ld % 8 == 0 && (!masked || (x && ld >= c && ld % 8 == 0)), "disc_cls_bwd: pitches");
        const int threads = cls_threads(c);
        dim3 grid(h, n);
        const size_t sm = sizeof(float) * 5 * c;
        unsigned long long* det = nullptr;
        if (dbias_prev && det_mode()) {
            det = det_scratch(as_stream(s), static_cast<size_t>(c));
            if (!det) return RTSDS_ECUDA;
        }
        if (dtype == RTSDS_BF16)
            disc_cls_dgrad_kernel<__nv_bfloat16><<<grid, threads, sm, as_stream(s)>>>(
                g, w_oihw, reinterpret_cast<const __nv_bfloat16*>(x), ld, h, w, c, oh, ow, inv_p, g_scale, masked, slope,
                reinterpret_cast<__nv_bfloat16*>(d_raw), d_ld, dbias_prev, det);
        else if (dtype == RTSDS_F32)
            disc_cls_dgrad_kernel<float><<<grid, threads, sm, as_stream(s)>>>(g, w_oihw, reinterpret_cast<const float*>(x), ld, h, w, c,
                                                                            oh, ow, inv_p, g_scale, masked, slope,
                                                                            reinterpret_cast<float*>(d_raw), d_ld, dbias_prev, det);
        else { set_error("disc_cls_bwd: bad dtype"); return RTSDS_EINVAL; }
        count_launch();
        int rc = check_launch("disc_cls_dgrad_kernel");
        if (rc == RTSDS_OK && det) rc = det_finish(det, dbias_prev, static_cast<size_t>(c), true, as_stream(s));
        return rc;
    }
    return RTSDS_OK;
}

extern "C" int rtsds_bce_logits(const float* logit, int n, float target, float scale, float* loss, float* dlogit,
                                rtsds_stream_t s) {
    RTSDS_REQUIRE(logit && n > 0 && (loss || dlogit), "bce_logits: bad argument");
    bce_logits_kernel<<<1, 256, 0, as_stream(s)>>>(logit, n, target, scale, loss, dlogit);
    count_launch();
    return check_launch("bce_logits_kernel");
}
