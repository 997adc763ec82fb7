// Generic "tap GEMM" description shared by the tensor-core and CUDA-core convolution kernels:
//   y[n, oy, ox, co] = sum_t sum_c  view_{map[t]}[n, oy + dh[t], ox + dw[t], c] * w[co][kb[t]*kalign + c]
// A forward conv (stride 1 or 2: one dense view per input parity) and each parity class of a
// dgrad are instances; out-of-range view coordinates read as zero (= padding).
#pragma once
#include "common.cuh"
#include <cstring>

namespace rtsds {

constexpr int TAP_MAX = 16;

struct TapView { const void* base; int wd, hd; long long sw, sh, sn; bool used; };
struct TapProblem {
    TapView view[4];
    int ck;                       // reduction channels per tap (multiple of 64 on the tensor-core path)
    int c_extent;                 // channels that really exist in the views (<= ck; the rest reads as zero)
    int n_img, oh, ow;            // output grid
    int n_taps;
    signed char dh[TAP_MAX], dw[TAP_MAX], map[TAP_MAX];
    int kb[TAP_MAX];              // weight K offset of tap t, in units of kalign channels
    const void* w; long long w_ktot;
    int cout;                     // GEMM N (real), padded via conv_cout_pad
    long long out_sn, out_sh, out_sw, res_sn, res_sh, res_sw;
    const float* scale; const float* shift; const void* residual; float* stats; void* y;
    unsigned long long* det;      // deterministic mode: exact accumulators [2*cout] replacing the atomics on stats (set by the launcher only)
    int out_dtype, act; float slope;
    int split_req;
    int in_f16;                   // operands are IEEE half (RTSDS_F16) instead of bf16
    float* gap_out;               // [n_img][cout] fp32: += mean over the output pixels of the final value (tensor-core path)
};


static inline int fwd_problem(const RtsdsConvDesc* d, const void* x, const void* w, int kalign, int in_esize, TapProblem* t) {
    memset(t, 0, sizeof(*t));
    RTSDS_REQUIRE(d->cin > 0 && d->cin % kalign == 0, "conv2d fwd: cin=%d must be a multiple of %d", d->cin, kalign);
    RTSDS_REQUIRE(d->stride == 1 || d->stride == 2, "conv2d fwd: stride %d unsupported", d->stride);
    RTSDS_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= TAP_MAX, "conv2d_tc_fwd: %dx%d filter unsupported", d->kh, d->kw);
    RTSDS_REQUIRE(d->dil >= 1 && d->pad >= 0, "conv2d_tc_fwd: bad dil/pad");
    RTSDS_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->cout > 0, "conv2d_tc_fwd: empty tensor");
    const int exp_oh = (d->h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
    const int exp_ow = (d->w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
    RTSDS_REQUIRE(d->oh == exp_oh && d->ow == exp_ow, "conv2d_tc_fwd: oh/ow (%d,%d) != expected (%d,%d)", d->oh, d->ow, exp_oh, exp_ow);
    RTSDS_REQUIRE(d->in_ld >= d->cin && (kalign == 1 || d->in_ld % 8 == 0), "conv2d fwd: in_ld=%d must be >= cin (and a multiple of 8 on the tensor-core path)", d->in_ld);
    t->ck = d->cin; t->c_extent = d->cin; t->n_img = d->n; t->oh = d->oh; t->ow = d->ow;
    t->n_taps = d->kh * d->kw;
    t->w = w; t->w_ktot = static_cast<long long>(t->n_taps) * d->cin; t->cout = d->cout;
    const char* xb = reinterpret_cast<const char*>(x);
    const long long ld = d->in_ld;
    const int st = d->stride;
    for (int r = 0; r < d->kh; ++r) {
        for (int q = 0; q < d->kw; ++q) {
            const int i = r * d->kw + q;
            const int a = r * d->dil - d->pad, b = q * d->dil - d->pad;
            const int hp = ((a % st) + st) % st, wp = ((b % st) + st) % st;
            const int dh = (a - hp) / st, dw = (b - wp) / st;
            RTSDS_REQUIRE(dh >= -128 && dh <= 127 && dw >= -128 && dw <= 127, "conv2d_tc_fwd: tap offset out of range");
            const int mi = hp * 2 + wp;
            t->dh[i] = static_cast<signed char>(dh); t->dw[i] = static_cast<signed char>(dw);
            t->map[i] = static_cast<signed char>(mi); t->kb[i] = i * (d->cin / kalign);
            if (!t->view[mi].used) {
                const int hd = (d->h - hp + st - 1) / st, wd = (d->w - wp + st - 1) / st;
                RTSDS_REQUIRE(hd > 0 && wd > 0, "conv2d_tc_fwd: degenerate parity view");
                t->view[mi] = TapView{xb + (static_cast<long long>(hp) * d->w + wp) * ld * in_esize, wd, hd, st * ld,
                                      static_cast<long long>(st) * d->w * ld, static_cast<long long>(d->h) * d->w * ld, true};
            }
        }
    }
    t->out_sw = d->out_ld; t->out_sh = static_cast<long long>(d->ow) * d->out_ld; t->out_sn = t->out_sh * d->oh;
    t->res_sw = d->res_ld; t->res_sh = static_cast<long long>(d->ow) * d->res_ld; t->res_sn = t->res_sh * d->oh;
    t->out_dtype = d->out_dtype; t->act = d->act; t->slope = d->slope; t->split_req = d->split_k;
    t->in_f16 = d->in_dtype == RTSDS_F16;
    return RTSDS_OK;
}


// ---- dgrad: dx[n,h,w,ci] = sum_{r,s,co} dy[n,(h+pad-r*dil)/st,(w+pad-s*dil)/st,co] * W[co,ci,r,s] (+ residual) ----
// `d` is the FORWARD geometry.  dy: NHWC bf16 with pitch d->out_ld >= ck = roundup(cout,64), channels
// cout..ck-1 zero.  w_dgrad: [cin_pad][kh*kw][ck] (rtsds_pack_conv_weight_dgrad).  dx: NHWC with pitch d->in_ld.
// Stride 2 runs one tap-GEMM per input-pixel parity class (dense taps, strided output rows).
static inline int dgrad_problem(const RtsdsConvDesc* d, const void* dy, const void* w, const void* residual, void* dx,
                                int out_dtype, int ph, int pw, int kalign, int in_esize, TapProblem* t) {
    memset(t, 0, sizeof(*t));
    const int st = d->stride;
    const int ck = static_cast<int>(cdiv(d->cout, kalign) * kalign);
    t->ck = ck; t->c_extent = d->cout; t->n_img = d->n;
    t->oh = (d->h - ph + st - 1) / st; t->ow = (d->w - pw + st - 1) / st;
    t->w = w; t->cout = d->cin;
    t->w_ktot = static_cast<long long>(d->kh) * d->kw * ck;
    int nt = 0;
    for (int r = 0; r < d->kh; ++r) {
        const int a = ph + d->pad - r * d->dil;
        if (((a % st) + st) % st) continue;
        for (int q = 0; q < d->kw; ++q) {
            const int b = pw + d->pad - q * d->dil;
            if (((b % st) + st) % st) continue;
            const int dh = a / st, dw = b / st;      // exact division
            RTSDS_REQUIRE(dh >= -128 && dh <= 127 && dw >= -128 && dw <= 127, "conv2d_tc_dgrad: tap offset out of range");
            t->dh[nt] = static_cast<signed char>(dh); t->dw[nt] = static_cast<signed char>(dw); t->map[nt] = 0;
            t->kb[nt] = (r * d->kw + q) * (ck / kalign);
            ++nt;
        }
    }
    t->n_taps = nt;
    const long long ldy = d->out_ld;
    t->view[0] = TapView{dy, d->ow, d->oh, ldy, static_cast<long long>(d->ow) * ldy, static_cast<long long>(d->oh) * d->ow * ldy, true};
    const long long ldx = d->in_ld;
    const size_t es = out_dtype == RTSDS_BF16 ? 2 : 4;
    const long long off = (static_cast<long long>(ph) * d->w + pw) * ldx;
    t->y = reinterpret_cast<char*>(dx) + off * es;
    t->out_sw = st * ldx; t->out_sh = static_cast<long long>(st) * d->w * ldx; t->out_sn = static_cast<long long>(d->h) * d->w * ldx;
    if (residual) {
        const long long ldr = d->res_ld ? d->res_ld : ldx;
        t->residual = reinterpret_cast<const char*>(residual) + (static_cast<long long>(ph) * d->w + pw) * ldr * es;
        t->res_sw = st * ldr; t->res_sh = static_cast<long long>(st) * d->w * ldr; t->res_sn = static_cast<long long>(d->h) * d->w * ldr;
    }
    t->out_dtype = out_dtype; t->act = RTSDS_ACT_NONE; t->split_req = d->split_k;
    return RTSDS_OK;
}


}  // namespace rtsds
