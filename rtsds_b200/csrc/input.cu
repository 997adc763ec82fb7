// Device-side input pipeline: what the reference does on the CPU inside its Dataset / transforms
// (main.py:60-108, datasets/cityscapes.py:66-72, datasets/gta5.py:68-82, utils.py:67-75) for every sample —
//     read_image(...).float()  ->  transforms.Resize(size, antialias=True)  ->  transforms.Normalize(mean, std)
//     read_image(...).long()   ->  transforms.Resize(size, antialias=True)  ->  IntRangeTransformer(0, num_classes)
// — evaluated on the GPU from the raw uint8 image / label planes, so that a frame crosses PCIe as 3 bytes per pixel instead
// of 12 (and a label as 1 byte instead of 8).  Memory-bound, one pass each.
//
// Resize semantics are torchvision's for tensors: F.interpolate(mode="bilinear", align_corners=False, antialias=True),
// i.e. ATen's _upsample_bilinear2d_aa: a separable triangle filter whose support grows with the down-scaling factor,
//     scale = in/out, support = max(scale, 1), center = scale*(i+0.5), xmin = max(int(center-support+0.5), 0),
//     xsize = min(int(center+support+0.5), in) - xmin, w_j = tri((j + xmin - center + 0.5) / max(scale,1)) / sum_j
// (for equal sizes this is the identity).  Integer tensors (labels) go through float, are ROUNDED (torch.round: half to
// even) and cast back (torchvision _cast_squeeze_out), then clamped to [lo, hi] (IntRangeTransformer) and widened to int64.
#include "common.cuh"

namespace rtsds {

struct AaAxis { float scale, support, invscale; };
static inline AaAxis make_axis(int in, int out) {
    AaAxis a;
    a.scale = static_cast<float>(in) / static_cast<float>(out);      // area_pixel_compute_scale, align_corners=False
    a.support = a.scale >= 1.f ? a.scale : 1.f;                       // interp_size/2 * scale, interp_size = 2
    a.invscale = a.scale >= 1.f ? 1.f / a.scale : 1.f;
    return a;
}
__device__ __forceinline__ void aa_window(int i, const AaAxis& a, int in, int* xmin, int* xsize, float* center) {
    // ATen rounds the centre to float BEFORE it is used (center = scale * (i + 0.5)); __fmul_rn keeps nvcc from contracting
    // the product into the subtractions below, which would use the unrounded centre: half an ulp of a coordinate near 2000
    // is 6e-5 pixel — 1e-2 grey levels after the filter, harmless but not what the reference computes
    const float c = __fmul_rn(a.scale, static_cast<float>(i) + 0.5f);
    int lo = static_cast<int>(c - a.support + 0.5f);
    if (lo < 0) lo = 0;
    int hi = static_cast<int>(c + a.support + 0.5f);
    if (hi > in) hi = in;
    *xmin = lo; *xsize = hi - lo; *center = c;
}
__device__ __forceinline__ float aa_tri(float x) { x = fabsf(x); return x < 1.f ? 1.f - x : 0.f; }

constexpr int AA_MAX_TAPS = 24;        // window per axis; covers down-scaling factors up to ~11

// value at output (oy, ox) of plane `src` [h][w]; LOAD converts one source element to float
template <typename S>
__device__ __forceinline__ float aa_sample(const S* __restrict__ src, int h, int w, int oy, int ox, const AaAxis& ay, const AaAxis& ax) {
    int y0, ny, x0, nx;
    float cy, cx;
    aa_window(oy, ay, h, &y0, &ny, &cy);
    aa_window(ox, ax, w, &x0, &nx, &cx);
    float wx[AA_MAX_TAPS];
    float sx = 0.f;
#pragma unroll 4
    for (int j = 0; j < nx; ++j) { wx[j] = aa_tri((static_cast<float>(j + x0) - cx + 0.5f) * ax.invscale); sx += wx[j]; }
    const float inv_sx = sx != 0.f ? 1.f / sx : 0.f;
    float acc = 0.f, sy = 0.f;
    for (int i = 0; i < ny; ++i) {
        const float wy = aa_tri((static_cast<float>(i + y0) - cy + 0.5f) * ay.invscale);
        sy += wy;
        const S* row = src + static_cast<long long>(y0 + i) * w + x0;
        float r = 0.f;
#pragma unroll 4
        for (int j = 0; j < nx; ++j) r = fmaf(wx[j], static_cast<float>(row[j]), r);
        acc = fmaf(wy, r * inv_sx, acc);           // horizontal pass normalised first, as the separable ATen kernel does
    }
    return sy != 0.f ? acc / sy : 0.f;
}

// image: uint8 NCHW [n,3,h,w] -> fp32 NCHW [n,3,oh,ow], out = resize(float(src)) * scale[c] + bias[c]
__global__ void __launch_bounds__(256)
image_u8_kernel(const uint8_t* __restrict__ src, int n, int c, int h, int w, int oh, int ow, AaAxis ay, AaAxis ax, int identity,
                float s0, float s1, float s2, float b0, float b1, float b2, float* __restrict__ dst) {
    const long long total = static_cast<long long>(n) * c * oh * ow;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % ow);
        const long long r = i / ow;
        const int oy = static_cast<int>(r % oh);
        const long long pl = r / oh;                  // plane index n*c + ch
        const int ch = static_cast<int>(pl % c);
        const uint8_t* p = src + pl * h * w;
        const float v = identity ? static_cast<float>(p[static_cast<long long>(oy) * w + ox]) : aa_sample(p, h, w, oy, ox, ay, ax);
        const float sc = ch == 0 ? s0 : (ch == 1 ? s1 : s2), bi = ch == 0 ? b0 : (ch == 1 ? b1 : b2);
        dst[i] = fmaf(v, sc, bi);
    }
}

// same-size case (a camera frame already at network resolution): pure convert + normalise, 16 pixels per thread
// (one 16-byte load, four 16-byte stores).  plane = h*w must be a multiple of 16 and src 16-byte aligned (host check).
__global__ void __launch_bounds__(256)
image_u8_identity_kernel(const uint8_t* __restrict__ src, long long total16, int plane16, int c, float s0, float s1, float s2,
                         float b0, float b1, float b2, float* __restrict__ dst) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total16;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>((i / plane16) % c);
        const float sc = ch == 0 ? s0 : (ch == 1 ? s1 : s2), bi = ch == 0 ? b0 : (ch == 1 ? b1 : b2);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
        float4* o = reinterpret_cast<float4*>(dst) + i * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = make_float4(fmaf(static_cast<float>(wds[k] & 0xffu), sc, bi), fmaf(static_cast<float>((wds[k] >> 8) & 0xffu), sc, bi),
                               fmaf(static_cast<float>((wds[k] >> 16) & 0xffu), sc, bi), fmaf(static_cast<float>(wds[k] >> 24), sc, bi));
    }
}

// labels: S (uint8 or int64) [n,h,w] -> int64 [n,oh,ow] = clamp(round(resize(float(src))), lo, hi)
template <typename S>
__global__ void __launch_bounds__(256)
label_kernel(const S* __restrict__ src, int n, int h, int w, int oh, int ow, AaAxis ay, AaAxis ax, int identity, int clamp,
             long long lo, long long hi, long long* __restrict__ dst) {
    const long long total = static_cast<long long>(n) * oh * ow;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % ow);
        const long long r = i / ow;
        const int oy = static_cast<int>(r % oh);
        const S* p = src + (r / oh) * h * w;
        long long v;
        if (identity) v = static_cast<long long>(p[static_cast<long long>(oy) * w + ox]);
        else v = static_cast<long long>(rintf(aa_sample(p, h, w, oy, ox, ay, ax)));      // torch.round: half to even
        if (clamp) v = v < lo ? lo : (v > hi ? hi : v);
        dst[i] = v;
    }
}

}  // namespace rtsds

using namespace rtsds;

static int aa_check(int h, int w, int oh, int ow, const char* who) {
    RTSDS_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0, "%s: empty tensor", who);
    const float sy = static_cast<float>(h) / oh, sx = static_cast<float>(w) / ow;
    RTSDS_REQUIRE(2.f * (sy > 1.f ? sy : 1.f) + 2.f <= AA_MAX_TAPS && 2.f * (sx > 1.f ? sx : 1.f) + 2.f <= AA_MAX_TAPS,
                  "%s: down-scaling factor (%g, %g) beyond the supported filter window", who, sy, sx);
    return RTSDS_OK;
}

extern "C" int rtsds_image_u8_to_f32(const uint8_t* src, int n, int c, int h, int w, int oh, int ow, const float* scale3,
                                     const float* bias3, float* dst, rtsds_stream_t s) {
    RTSDS_REQUIRE(src && dst && scale3 && bias3 && n > 0 && c >= 1 && c <= 3, "image_u8_to_f32: bad argument (c must be 1..3)");
    int rc = aa_check(h, w, oh, ow, "image_u8_to_f32");
    if (rc != RTSDS_OK) return rc;
    const long long total = static_cast<long long>(n) * c * oh * ow;
    const long long plane = static_cast<long long>(h) * w;
    if (h == oh && w == ow && plane % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const long long t16 = total / 16;
        const int g16 = static_cast<int>(cdiv(t16, 256) > 16LL * num_sms() ? 16LL * num_sms() : cdiv(t16, 256));
        image_u8_identity_kernel<<<g16, 256, 0, as_stream(s)>>>(src, t16, static_cast<int>(plane / 16), c, scale3[0], scale3[c > 1 ? 1 : 0],
                                                                scale3[c > 2 ? 2 : 0], bias3[0], bias3[c > 1 ? 1 : 0], bias3[c > 2 ? 2 : 0], dst);
        count_launch();
        return check_launch("image_u8_identity_kernel");
    }
    const int grid = static_cast<int>(cdiv(total, 256) > 32LL * num_sms() ? 32LL * num_sms() : cdiv(total, 256));
    image_u8_kernel<<<grid, 256, 0, as_stream(s)>>>(src, n, c, h, w, oh, ow, make_axis(h, oh), make_axis(w, ow), h == oh && w == ow,
                                                    scale3[0], scale3[c > 1 ? 1 : 0], scale3[c > 2 ? 2 : 0], bias3[0],
                                                    bias3[c > 1 ? 1 : 0], bias3[c > 2 ? 2 : 0], dst);
    count_launch();
    return check_launch("image_u8_kernel");
}

extern "C" int rtsds_label_resize_clamp(const void* src, int src_is_u8, int n, int h, int w, int oh, int ow, int clamp,
                                        int64_t lo, int64_t hi, int64_t* dst, rtsds_stream_t s) {
    RTSDS_REQUIRE(src && dst && n > 0, "label_resize_clamp: bad argument");
    int rc = aa_check(h, w, oh, ow, "label_resize_clamp");
    if (rc != RTSDS_OK) return rc;
    const long long total = static_cast<long long>(n) * oh * ow;
    const int grid = static_cast<int>(cdiv(total, 256) > 32LL * num_sms() ? 32LL * num_sms() : cdiv(total, 256));
    const int ident = h == oh && w == ow;
    if (src_is_u8)
        label_kernel<uint8_t><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const uint8_t*>(src), n, h, w, oh, ow, make_axis(h, oh),
                                                              make_axis(w, ow), ident, clamp, lo, hi, reinterpret_cast<long long*>(dst));
    else
        label_kernel<long long><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const long long*>(src), n, h, w, oh, ow, make_axis(h, oh),
                                                                make_axis(w, ow), ident, clamp, lo, hi, reinterpret_cast<long long*>(dst));
    count_launch();
    return check_launch("label_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// F.adaptive_avg_pool2d on NCHW fp32 logits (train.py:410,438,445, adversarial_train_2: the generator's prediction is
// pooled to the target label size before softmax -> discriminator).  Window of output o along an axis:
// [floor(o*in/out), ceil((o+1)*in/out)) — ATen's start_index / end_index.  Backward: every input pixel collects dy/area
// from the (at most two per axis) windows that contain it — a gather, no atomics, deterministic.
namespace rtsds {
__device__ __forceinline__ int ap_start(int o, int in, int out) { return static_cast<int>((static_cast<long long>(o) * in) / out); }
__device__ __forceinline__ int ap_end(int o, int in, int out) { return static_cast<int>((static_cast<long long>(o + 1) * in + out - 1) / out); }

__global__ void __launch_bounds__(256)
adaptive_avgpool_fwd_kernel(const float* __restrict__ x, long long planes, int h, int w, int oh, int ow, float* __restrict__ y) {
    const long long total = planes * oh * ow;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % ow);
        const long long r = i / ow;
        const int oy = static_cast<int>(r % oh);
        const float* p = x + (r / oh) * h * w;
        const int y0 = ap_start(oy, h, oh), y1 = ap_end(oy, h, oh), x0 = ap_start(ox, w, ow), x1 = ap_end(ox, w, ow);
        float acc = 0.f;
        for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) acc += p[static_cast<long long>(yy) * w + xx];
        y[i] = acc / static_cast<float>((y1 - y0) * (x1 - x0));
    }
}

__global__ void __launch_bounds__(256)
adaptive_avgpool_bwd_kernel(const float* __restrict__ dy, long long planes, int h, int w, int oh, int ow, float* __restrict__ dx) {
    const long long total = planes * h * w;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ix = static_cast<int>(i % w);
        const long long r = i / w;
        const int iy = static_cast<int>(r % h);
        const float* g = dy + (r / h) * oh * ow;
        // candidate outputs: every o whose window [floor(o*in/out), ceil((o+1)*in/out)) can contain i lies in
        // [floor(i*out/in) - 1, floor((i+1)*out/in) + 1]
        const int cy0 = static_cast<int>((static_cast<long long>(iy) * oh) / h) - 1, cy1 = static_cast<int>((static_cast<long long>(iy + 1) * oh) / h) + 1;
        const int cx0 = static_cast<int>((static_cast<long long>(ix) * ow) / w) - 1, cx1 = static_cast<int>((static_cast<long long>(ix + 1) * ow) / w) + 1;
        float acc = 0.f;
        for (int oy = max(cy0, 0); oy <= min(cy1, oh - 1); ++oy) {
            const int y0 = ap_start(oy, h, oh), y1 = ap_end(oy, h, oh);
            if (iy < y0 || iy >= y1) continue;
            for (int ox = max(cx0, 0); ox <= min(cx1, ow - 1); ++ox) {
                const int x0 = ap_start(ox, w, ow), x1 = ap_end(ox, w, ow);
                if (ix < x0 || ix >= x1) continue;
                acc += g[static_cast<long long>(oy) * ow + ox] / static_cast<float>((y1 - y0) * (x1 - x0));
            }
        }
        dx[i] = acc;
    }
}
}  // namespace rtsds

extern "C" int rtsds_adaptive_avgpool_nchw_fwd(const float* x, int64_t planes, int h, int w, int oh, int ow, float* y, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && y && planes > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "adaptive_avgpool_fwd: bad argument");
    const long long total = planes * oh * ow;
    const int grid = static_cast<int>(cdiv(total, 256) > 32LL * num_sms() ? 32LL * num_sms() : cdiv(total, 256));
    adaptive_avgpool_fwd_kernel<<<grid, 256, 0, as_stream(s)>>>(x, planes, h, w, oh, ow, y);
    count_launch();
    return check_launch("adaptive_avgpool_fwd_kernel");
}

extern "C" int rtsds_adaptive_avgpool_nchw_bwd(const float* dy, int64_t planes, int h, int w, int oh, int ow, float* dx, rtsds_stream_t s) {
    RTSDS_REQUIRE(dy && dx && planes > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "adaptive_avgpool_bwd: bad argument");
    const long long total = planes * h * w;
    const int grid = static_cast<int>(cdiv(total, 256) > 32LL * num_sms() ? 32LL * num_sms() : cdiv(total, 256));
    adaptive_avgpool_bwd_kernel<<<grid, 256, 0, as_stream(s)>>>(dy, planes, h, w, oh, ow, dx);
    count_launch();
    return check_launch("adaptive_avgpool_bwd_kernel");
}
