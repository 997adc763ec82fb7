// Shared helpers for librtsds_b200 (sm_100a only).
#pragma once
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/rtsds_b200.h"

namespace rtsds {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  check_launch(const char* what);          // cudaGetLastError -> RTSDS_ECUDA

#define RTSDS_REQUIRE(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            ::rtsds::set_error(__VA_ARGS__);                      \
            return RTSDS_EINVAL;                                  \
        }                                                         \
    } while (0)

static inline cudaStream_t as_stream(rtsds_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch for the small kernels between the tensor-core convs: the grid may be scheduled while its
// predecessor in the stream drains, and every thread calls pdl_wait() before it touches memory the predecessor wrote
// (pdl_wait() returns once the predecessor has completed and flushed).  Opt-in with RTSDS_PDL_GLUE=1: measured on the
// batch-1 frame it is 2 % SLOWER than plain launches (0.343-0.345 vs 0.336-0.337 ms, same box, alternating runs) — the
// early-resident waiting blocks take SM slots from the draining conv — so the default launches these kernels serialised.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static int use_pdl = -1;
    if (use_pdl < 0) {
        const char* e = getenv("RTSDS_NO_PDL");
        const char* g = getenv("RTSDS_PDL_GLUE");
        use_pdl = (g && g[0] == '1' && !(e && e[0] == '1')) ? 1 : 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
int num_sms();

// ---- deterministic (order-independent) accumulation ------------------------------------------------------------------
// rtsds_set_deterministic(1): every cross-CTA floating-point reduction of the training path that normally goes through
// fp32 atomics (train-mode BatchNorm sums, weight-gradient partials, BatchNorm-backward sums) is accumulated EXACTLY
// instead, in 64.64 two's-complement fixed point with integer atomics, and rounded to fp32 once at the end.  Integer
// addition is associative, so the result does not depend on the order in which CTAs arrive: two runs are bit-identical.
// Range +-9.2e18, resolution 5.4e-20 per contribution (an fp32 contribution below that is dropped: it could not have
// changed an fp32 sum of ordinary magnitude either).  An accumulator is two 64-bit words {low, high}; the scratch that
// holds them belongs to the library (one region per stream), the caller-visible buffers stay fp32.
__device__ __forceinline__ void det_add(unsigned long long* acc, float v) {
    if (v == 0.0f || !(fabsf(v) < 9.0e18f)) return;            // zeros, NaN/inf and out-of-range values are skipped
    const double d = static_cast<double>(v);
    const double fl = floor(d);
    const long long hi = static_cast<long long>(fl);
    const unsigned long long lo = __double2ull_rz((d - fl) * 18446744073709551616.0);     // (d - fl) in [0,1): < 2^64
    const unsigned long long old = atomicAdd(acc, lo);
    const unsigned long long carry = (old + lo < old) ? 1ull : 0ull;
    const unsigned long long h = static_cast<unsigned long long>(hi) + carry;
    if (h) atomicAdd(acc + 1, h);
}
// out-of-line pair (sum, sum of squares) for the hot epilogues: keeps the fp64 conversion code out of their register budget
static __device__ __noinline__ void det_add2(unsigned long long* a, float va, unsigned long long* b, float vb) {
    det_add(a, va);
    det_add(b, vb);
}
static __device__ __noinline__ void det_add1(unsigned long long* a, float va) { det_add(a, va); }
__device__ __forceinline__ double det_value(const unsigned long long* acc) {
    return static_cast<double>(static_cast<long long>(acc[1])) + static_cast<double>(acc[0]) * 5.421010862427522e-20;   // 2^-64
}
bool det_mode();
// zeroed scratch of n accumulators (16 bytes each) for this stream; NULL + error set when the allocation fails
unsigned long long* det_scratch(cudaStream_t st, size_t n);
// dst[i] = (accumulate ? dst[i] : 0) + fp32(acc[i]) for i < n, on the stream
int det_finish(const unsigned long long* acc, float* dst, size_t n, bool accumulate, cudaStream_t st);

// ---- dtype helpers ---------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
// fp16 conversions SATURATE to the largest finite half (cvt.rn.satfinite): an out-of-range activation degrades to a
// clipped value instead of poisoning everything downstream with inf / NaN
template <> __device__ __forceinline__ __half from_f32<__half>(float v) {
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return __ushort_as_half(h);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
}
// 16-bit storage selected at run time (warp-uniform flag): fp16 (RTSDS_F16) or bf16 (RTSDS_BF16)
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi, bool f16) { return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
__device__ __forceinline__ float2 unpack_16x2(uint32_t u, bool f16) { return f16 ? unpack_f16x2(u) : unpack_bf16x2(u); }
__device__ __forceinline__ float ld_16(const void* p, long long i, bool f16) {
    return f16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_16(void* p, long long i, float v, bool f16) {
    if (f16) reinterpret_cast<__half*>(p)[i] = from_f32<__half>(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
// typed forms for kernels templated on the storage type
template <typename T> __device__ __forceinline__ uint32_t pack_x2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack_x2<__nv_bfloat16>(float lo, float hi) { return pack_bf16x2(lo, hi); }
template <> __device__ __forceinline__ uint32_t pack_x2<__half>(float lo, float hi) { return pack_f16x2(lo, hi); }
template <typename T> __device__ __forceinline__ float2 unpack_x2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack_x2<__nv_bfloat16>(uint32_t u) { return unpack_bf16x2(u); }
template <> __device__ __forceinline__ float2 unpack_x2<__half>(uint32_t u) { return unpack_f16x2(u); }
static inline bool is_16bit(int dtype) { return dtype == RTSDS_BF16 || dtype == RTSDS_F16; }
static inline size_t dtype_size(int dtype) { return dtype == RTSDS_F32 ? 4 : 2; }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == RTSDS_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == RTSDS_ACT_LRELU) return v > 0.0f ? v : v * slope;
    return v;
}

// ---- warp / block reductions ----------------------------------------------
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Bilinear source index, align_corners=False, as ATen's
// area_pixel_compute_source_index (negative source clamped to 0).
// Division by a runtime constant as multiply + shift (exact for n < 2^31): the index decompositions of the grid-stride
// kernels otherwise cost several 64-bit divisions per 16 bytes moved.
struct FastDiv { unsigned d, mul, shr; };
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d;
    unsigned l = 0;
    while ((1ull << l) < d) ++l;
    f.shr = 31 + l;
    f.mul = static_cast<unsigned>(((1ull << f.shr) + d - 1) / d);
    if (d == 1) { f.mul = 0; f.shr = 0; }
    return f;
}
__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv& f) {
    return f.d == 1 ? n : static_cast<unsigned>((static_cast<unsigned long long>(n) * f.mul) >> f.shr);
}
struct Div3 { FastDiv a, b, c; };
// n -> (n / d, n % d)
__device__ __forceinline__ unsigned fdivmod(unsigned n, const FastDiv& f, unsigned* rem) {
    const unsigned q = fdiv(n, f);
    *rem = n - q * f.d;
    return q;
}

struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_src(int dst, float rscale, int in_size) {
    float r = rscale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    r = r < 0.0f ? 0.0f : r;
    int i0 = static_cast<int>(r);
    if (i0 > in_size - 1) i0 = in_size - 1;
    int p = (i0 < in_size - 1) ? 1 : 0;
    Lerp o;
    o.i0 = i0; o.i1 = i0 + p;
    o.l1 = r - static_cast<float>(i0);
    o.l0 = 1.0f - o.l1;
    return o;
}
// scale used by F.interpolate(size=...) in fp32: (float)in / out
static inline float resize_scale(int in_size, int out_size) {
    return static_cast<float>(in_size) / static_cast<float>(out_size);
}

}  // namespace rtsds
