// CUDA-core implicit-GEMM convolution with fp32 accumulation.
//
// This is the "fp32 check mode" of BASELINE.json (logits within 1e-4 of the
// reference) and, instantiated for bf16 tensors, an independent cross-check of
// the tcgen05 kernel (same operands, different machinery).  It accepts any
// cin / filter / stride / dilation.  Not the production path for bf16.
//
// Tiling: 64 output pixels x 64 output channels per 256-thread block, 4x4
// register tile per thread, K walked as (tap, 16-channel chunk) through shared
// memory.
#include "common.cuh"

namespace rtsds {

constexpr int SC_TM = 64, SC_TN = 64, SC_TK = 16, SC_THREADS = 256;

struct SimtParams {
    int n, h, w, cin, in_ld;
    int cout, cout_pad, out_ld, res_ld;
    int kh, kw, stride, pad, dil, oh, ow;
    int act;
    float slope;
    int out_dtype;
    long long m_total;
};

template <typename T>
__global__ void __launch_bounds__(SC_THREADS)
conv_simt_kernel(const T* __restrict__ x, const T* __restrict__ wgt, const float* __restrict__ scale,
                 const float* __restrict__ shift, const void* __restrict__ residual, float* stats, void* y,
                 SimtParams p) {
    __shared__ float sA[SC_TK][SC_TM + 4];
    __shared__ float sB[SC_TK][SC_TN + 4];
    __shared__ int s_img[SC_TM], s_ih0[SC_TM], s_iw0[SC_TM];
    __shared__ float s_sum[SC_TN], s_sq[SC_TN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = static_cast<long long>(blockIdx.x) * SC_TM;
    const int n0 = blockIdx.y * SC_TN;

    if (tid < SC_TM) {
        long long m = m0 + tid;
        if (m < p.m_total) {
            int img = static_cast<int>(m / (static_cast<long long>(p.oh) * p.ow));
            long long r = m - static_cast<long long>(img) * p.oh * p.ow;
            int oy = static_cast<int>(r / p.ow), ox = static_cast<int>(r - static_cast<long long>(oy) * p.ow);
            s_img[tid] = img; s_ih0[tid] = oy * p.stride - p.pad; s_iw0[tid] = ox * p.stride - p.pad;
        } else {
            s_img[tid] = -1; s_ih0[tid] = 0; s_iw0[tid] = 0;
        }
        s_sum[tid] = 0.f; s_sq[tid] = 0.f;
    }
    __syncthreads();

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = p.kh * p.kw;
    const long long wrow = static_cast<long long>(taps) * p.cin;   // elements per output channel
    const int a_pix = tid >> 2, a_k = (tid & 3) * 4;                // A loader: pixel, 4 channels
    const int b_co = tid >> 2, b_k = (tid & 3) * 4;                 // B loader: out channel, 4 k

    for (int t = 0; t < taps; ++t) {
        const int r = t / p.kw, q = t - r * p.kw;
        const int img = s_img[a_pix];
        const int ih = s_ih0[a_pix] + r * p.dil, iw = s_iw0[a_pix] + q * p.dil;
        const bool inb = img >= 0 && ih >= 0 && ih < p.h && iw >= 0 && iw < p.w;
        const T* xp = inb ? x + ((static_cast<long long>(img) * p.h + ih) * p.w + iw) * p.in_ld : nullptr;
        const int co = n0 + b_co;
        const T* wp = (co < p.cout_pad) ? wgt + co * wrow + static_cast<long long>(t) * p.cin : nullptr;
        for (int c0 = 0; c0 < p.cin; c0 += SC_TK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + a_k + j;
                sA[a_k + j][a_pix] = (xp && c < p.cin) ? to_f32(xp[c]) : 0.f;
                const int cb = c0 + b_k + j;
                sB[b_k + j][b_co] = (wp && cb < p.cin) ? to_f32(wp[cb]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SC_TK; ++k) {
                float4 a = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
                float4 b = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
                float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // epilogue
    float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= p.m_total) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= p.cout) continue;
            float raw = acc[i][j];
            cs[j] += raw; cq[j] += raw * raw;
            float v = raw * (scale ? scale[co] : 1.f) + (shift ? shift[co] : 0.f);
            if (p.out_dtype == RTSDS_BF16) {
                if (residual) v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(residual)[m * p.res_ld + co]);
                reinterpret_cast<__nv_bfloat16*>(y)[m * p.out_ld + co] = __float2bfloat16_rn(apply_act(v, p.act, p.slope));
            } else {
                if (residual) v += reinterpret_cast<const float*>(residual)[m * p.res_ld + co];
                reinterpret_cast<float*>(y)[m * p.out_ld + co] = apply_act(v, p.act, p.slope);
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&s_sum[tx * 4 + j], cs[j]);
            atomicAdd(&s_sq[tx * 4 + j], cq[j]);
        }
        __syncthreads();
        if (tid < SC_TN && n0 + tid < p.cout) {
            atomicAdd(&stats[n0 + tid], s_sum[tid]);
            atomicAdd(&stats[p.cout + n0 + tid], s_sq[tid]);
        }
    }
}

// OIHW fp32 -> [cout_pad][kh*kw][cin]
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int taps, int cout_pad,
                                   T* __restrict__ out) {
    const long long total = static_cast<long long>(cout_pad) * taps * cin;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % cin);
        const long long r = i / cin;
        const int t = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        float v = (co < cout) ? w[(static_cast<long long>(co) * cin + c) * taps + t] : 0.f;
        out[i] = from_f32<T>(v);
    }
}

int conv_cout_pad(int cout);

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_conv2d_simt_fwd(const RtsdsConvDesc* d, const void* x, const void* w, const float* scale,
                                     const float* shift, const void* residual, float* stats, void* y,
                                     rtsds_stream_t s) {
    RTSDS_REQUIRE(d && x && w && y, "conv2d_simt_fwd: NULL argument");
    RTSDS_REQUIRE(d->in_dtype == RTSDS_BF16 || d->in_dtype == RTSDS_F32, "conv2d_simt_fwd: bad in_dtype");
    RTSDS_REQUIRE(d->out_dtype == RTSDS_BF16 || d->out_dtype == RTSDS_F32, "conv2d_simt_fwd: bad out_dtype");
    RTSDS_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "conv2d_simt_fwd: empty tensor");
    RTSDS_REQUIRE(d->stride >= 1 && d->dil >= 1 && d->pad >= 0 && d->kh >= 1 && d->kw >= 1, "conv2d_simt_fwd: bad geometry");
    const int exp_oh = (d->h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
    const int exp_ow = (d->w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
    RTSDS_REQUIRE(d->oh == exp_oh && d->ow == exp_ow, "conv2d_simt_fwd: oh/ow (%d,%d) != expected (%d,%d)", d->oh, d->ow, exp_oh, exp_ow);
    RTSDS_REQUIRE(d->in_ld >= d->cin && d->out_ld >= d->cout, "conv2d_simt_fwd: pitch smaller than channel count");
    if (residual) RTSDS_REQUIRE(d->res_ld >= d->cout, "conv2d_simt_fwd: res_ld < cout");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    SimtParams p;
    p.n = d->n; p.h = d->h; p.w = d->w; p.cin = d->cin; p.in_ld = d->in_ld;
    p.cout = d->cout; p.cout_pad = conv_cout_pad(d->cout); p.out_ld = d->out_ld; p.res_ld = d->res_ld;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil; p.oh = d->oh; p.ow = d->ow;
    p.act = d->act; p.slope = d->slope; p.out_dtype = d->out_dtype;
    p.m_total = static_cast<long long>(d->n) * d->oh * d->ow;
    dim3 grid(static_cast<unsigned>(cdiv(p.m_total, SC_TM)), static_cast<unsigned>(cdiv(d->cout, SC_TN)));
    if (d->in_dtype == RTSDS_BF16)
        conv_simt_kernel<__nv_bfloat16><<<grid, SC_THREADS, 0, as_stream(s)>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w), scale, shift,
            residual, stats, y, p);
    else
        conv_simt_kernel<float><<<grid, SC_THREADS, 0, as_stream(s)>>>(
            reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(w), scale, shift, residual, stats, y, p);
    count_launch();
    return check_launch("conv_simt_kernel");
}

extern "C" int rtsds_pack_conv_weight(const float* w_oihw, int cout, int cin, int kh, int kw, int cout_pad,
                                      int dtype, void* w_packed, rtsds_stream_t s) {
    RTSDS_REQUIRE(w_oihw && w_packed, "pack_conv_weight: NULL argument");
    RTSDS_REQUIRE(cout > 0 && cin > 0 && kh > 0 && kw > 0 && cout_pad >= cout, "pack_conv_weight: bad shape");
    RTSDS_REQUIRE(dtype == RTSDS_BF16 || dtype == RTSDS_F32, "pack_conv_weight: bad dtype");
    const long long total = static_cast<long long>(cout_pad) * kh * kw * cin;
    int grid = static_cast<int>(cdiv(total, 256) > 2048 ? 2048 : cdiv(total, 256));
    if (dtype == RTSDS_BF16)
        pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(w_oihw, cout, cin, kh * kw, cout_pad,
                                                                         reinterpret_cast<__nv_bfloat16*>(w_packed));
    else
        pack_weight_kernel<float><<<grid, 256, 0, as_stream(s)>>>(w_oihw, cout, cin, kh * kw, cout_pad,
                                                                 reinterpret_cast<float*>(w_packed));
    count_launch();
    return check_launch("pack_weight_kernel");
}
