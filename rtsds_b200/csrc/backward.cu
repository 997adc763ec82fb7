// Backward kernels of the bandwidth-bound parts of the training step (autograd of
// models/bisenet/build_bisenet.py driven by loss.backward(), train.py:95):
// BatchNorm(+ReLU) backward, max-pool backward, adjoint bilinear resizes, ARM and
// FFM attention backward, per-channel sums (bias grads), stem weight gradient.
// All HBM-bound: vectorised 16-byte accesses, register/shuffle/shared-memory
// reductions, one global atomic per channel per block.
#include "common.cuh"
#include "ptx.cuh"

namespace rtsds {

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    F8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ F8 ld8(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8(float* p, const F8& r) {
    reinterpret_cast<float4*>(p)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

static int grid_for(long long total, int threads, int waves = 8) {
    long long want = cdiv(total, threads);
    long long cap = static_cast<long long>(waves) * num_sms();
    return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

// ---------------------------------------------------------------- BatchNorm(+ReLU) backward
// g = dy * (relu ? y > 0 : 1);  sums[c] += sum g;  sums[C+c] += sum g * xhat,  xhat = (raw-mean)*invstd
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ raw,
                     int raw_ld, const float* __restrict__ mean, const float* __restrict__ invstd,
                     const float* __restrict__ fsc, const float* __restrict__ fsh, long long n_pix, int c,
                     int relu, float* sums, unsigned long long* det) {
    extern __shared__ float sh[];             // [2*cp]
    const int cg = (c + 7) / 8, cp = cg * 8;
    for (int i = threadIdx.x; i < 2 * cp; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int g8 = threadIdx.x % cg;
    const int prow = threadIdx.x / cg, prows = blockDim.x / cg;
    // ReLU mask: from the stored activation y, or (y == NULL) recomputed from raw with the forward's own
    // scale / shift — the same fp32 fma the forward evaluated, so the mask is bit-identical and y is not re-read
    const bool rawmask = relu && y == nullptr;
    float s1[8], s2[8], mu[8], is[8], fa[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s1[j] = 0.f; s2[j] = 0.f;
        const int ch = g8 * 8 + j;
        mu[j] = ch < c ? mean[ch] : 0.f; is[j] = ch < c ? invstd[ch] : 0.f;
        fa[j] = (rawmask && ch < c) ? fsc[ch] : 0.f; fb[j] = (rawmask && ch < c) ? fsh[ch] : 0.f;
    }
    if (prow < prows) {
        // two pixels per iteration: all loads are issued before the first use (memory-level parallelism)
        const long long stride = static_cast<long long>(gridDim.x) * prows;
        for (long long p = static_cast<long long>(blockIdx.x) * prows + prow; p < n_pix; p += 2 * stride) {
            const long long q = p + stride;
            const bool two = q < n_pix;
            const F8 d0 = ld8(dy + p * dy_ld + g8 * 8);
            const F8 r0 = ld8(raw + p * raw_ld + g8 * 8);
            F8 d1, r1, y0, y1;
            if (two) { d1 = ld8(dy + q * dy_ld + g8 * 8); r1 = ld8(raw + q * raw_ld + g8 * 8); }
            if (relu && !rawmask) {
                y0 = ld8(y + p * y_ld + g8 * 8);
                if (two) y1 = ld8(y + q * y_ld + g8 * 8);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float m0 = rawmask ? fmaf(r0.v[j], fa[j], fb[j]) : (relu ? y0.v[j] : 1.f);
                const float g0 = (relu && !(m0 > 0.f)) ? 0.f : d0.v[j];
                s1[j] += g0;
                s2[j] += g0 * (r0.v[j] - mu[j]) * is[j];
                if (two) {
                    const float m1 = rawmask ? fmaf(r1.v[j], fa[j], fb[j]) : (relu ? y1.v[j] : 1.f);
                    const float g1 = (relu && !(m1 > 0.f)) ? 0.f : d1.v[j];
                    s1[j] += g1;
                    s2[j] += g1 * (r1.v[j] - mu[j]) * is[j];
                }
            }
        }
        if (det) {            // deterministic mode: this thread's partial sums go straight into the exact accumulators
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (g8 * 8 + j < c) {
                    det_add(det + 2 * (g8 * 8 + j), s1[j]);
                    det_add(det + 2 * (c + g8 * 8 + j), s2[j]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&sh[g8 * 8 + j], s1[j]);
                atomicAdd(&sh[cp + g8 * 8 + j], s2[j]);
            }
        }
    }
    __syncthreads();
    if (det) return;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        atomicAdd(&sums[i], sh[i]);
        atomicAdd(&sums[c + i], sh[cp + i]);
    }
}

// d_raw = gamma*invstd*(g - sum_g/M - xhat*sum_gx/M); optional g_out = g (gradient of the residual branch).
// Parameter gradients: dgamma += sum_gx, dbeta += sum_g (block 0).
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ raw,
                    int raw_ld, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ sums, const float* __restrict__ fsc,
                    const float* __restrict__ fsh, long long n_pix, int c, int relu,
                    TO* __restrict__ d_raw, int d_raw_ld, T* __restrict__ g_out, int g_ld, float* dgamma, float* dbeta) {
    const int cg = (c + 7) / 8;
    const bool rawmask = relu && y == nullptr;
    const float inv_m = 1.0f / static_cast<float>(n_pix);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < c; i += blockDim.x) {
            if (dgamma) dgamma[i] += sums[c + i];
            if (dbeta) dbeta[i] += sums[i];
        }
    }
    if (blockDim.x % cg == 0) {
        // one 8-channel group per thread: d_raw = ca*g + cb*raw + cc with per-channel coefficients in registers
        //   ca = gamma*invstd,  cb = -gamma*invstd^2*sum_gx/M,  cc = -ca*sum_g/M - cb*mean
        const int g8 = threadIdx.x % cg;
        const long long prows = blockDim.x / cg, stride = prows * gridDim.x;
        float ca[8], cb[8], cc[8], fa[8], fb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = g8 * 8 + j;
            if (ch < c) {
                const float is = invstd[ch], gm = gamma ? gamma[ch] : 1.f;
                ca[j] = gm * is;
                cb[j] = -gm * is * is * sums[c + ch] * inv_m;
                cc[j] = -ca[j] * sums[ch] * inv_m - cb[j] * mean[ch];
                fa[j] = rawmask ? fsc[ch] : 0.f; fb[j] = rawmask ? fsh[ch] : 0.f;
            } else { ca[j] = cb[j] = cc[j] = fa[j] = fb[j] = 0.f; }
        }
        for (long long p = static_cast<long long>(blockIdx.x) * prows + threadIdx.x / cg; p < n_pix; p += stride) {
            const F8 d = ld8(dy + p * dy_ld + g8 * 8);
            const F8 r = ld8(raw + p * raw_ld + g8 * 8);
            F8 yy;
            if (rawmask) {
#pragma unroll
                for (int j = 0; j < 8; ++j) yy.v[j] = fmaf(r.v[j], fa[j], fb[j]);
            } else if (relu) {
                yy = ld8(y + p * y_ld + g8 * 8);
            }
            F8 o, gg;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g = (relu && !(yy.v[j] > 0.f)) ? 0.f : d.v[j];
                gg.v[j] = (g8 * 8 + j < c) ? g : 0.f;
                o.v[j] = fmaf(ca[j], g, fmaf(cb[j], r.v[j], cc[j]));
            }
            st8(d_raw + p * d_raw_ld + g8 * 8, o);
            if (g_out) st8(g_out + p * g_ld + g8 * 8, gg);
        }
        return;
    }
    const long long total = n_pix * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long p = i / cg;
        const int g8 = static_cast<int>(i - p * cg);
        const F8 d = ld8(dy + p * dy_ld + g8 * 8);
        const F8 r = ld8(raw + p * raw_ld + g8 * 8);
        F8 yy;
        if (rawmask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = g8 * 8 + j;
                yy.v[j] = ch < c ? fmaf(r.v[j], fsc[ch], fsh[ch]) : 0.f;
            }
        } else if (relu) {
            yy = ld8(y + p * y_ld + g8 * 8);
        }
        F8 o, gg;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = g8 * 8 + j;
            if (ch < c) {
                const float g = (relu && !(yy.v[j] > 0.f)) ? 0.f : d.v[j];
                const float is = invstd[ch];
                const float xh = (r.v[j] - mean[ch]) * is;
                o.v[j] = (gamma ? gamma[ch] : 1.f) * is * (g - sums[ch] * inv_m - xh * sums[c + ch] * inv_m);
                gg.v[j] = g;
            } else { o.v[j] = 0.f; gg.v[j] = 0.f; }
        }
        st8(d_raw + p * d_raw_ld + g8 * 8, o);
        if (g_out) st8(g_out + p * g_ld + g8 * 8, gg);
    }
}

// ---- streamed forms (bf16, contiguous [n_pix, c], c = 8 * 2^k) -----------------------------------------------
// The register form above keeps at most 4-6 16-byte loads per thread in flight at ~120 registers, i.e. ~32 KB per
// SM, and stalls near 3 TB/s (latency bound).  Here one thread streams 8 KB tiles of every input into a 4-stage
// shared-memory ring with cp.async.bulk (128+ KB in flight per SM, independent of registers); the 256 threads read
// the tiles back conflict-free, one 16-byte chunk = one 8-channel group of one pixel each.
constexpr int BNS_TILE = 8192;   // bytes per tensor per stage
constexpr int bns_stages(int nt) { return 12 / nt; }   // 96 KB ring per block, two blocks per SM
constexpr int BNS_CHUNKS = BNS_TILE / 16 / 256;   // 16-byte chunks per thread per tile (2)

__device__ __forceinline__ F8 cvt8(const uint4& u) {
    F8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}

template <int NT>
struct BnStream {
    static constexpr int BNS_STAGES = bns_stages(NT);
    uint8_t* ring;
    uint64_t* full;
    const __nv_bfloat16* src[NT];
    long long n_pix, n_tiles;
    int c, tile_px;
    __device__ __forceinline__ void issue(int s, long long tile) const {
        const long long px0 = tile * tile_px;
        const long long left = n_pix - px0;
        const uint32_t bytes = static_cast<uint32_t>((left < tile_px ? left : tile_px) * c * 2);
        ptx::mbar_expect_tx(&full[s], NT * bytes);
#pragma unroll
        for (int t = 0; t < NT; ++t)
            ptx::bulk_load_1d(ring + (s * NT + t) * BNS_TILE, src[t] + px0 * c, bytes, &full[s]);
    }
    __device__ __forceinline__ void start(uint8_t* smem) {
        ring = smem;
        full = reinterpret_cast<uint64_t*>(smem + BNS_STAGES * NT * BNS_TILE);
        tile_px = BNS_TILE / (c * 2);
        n_tiles = (n_pix + tile_px - 1) / tile_px;
        if (threadIdx.x == 0) {
            for (int s = 0; s < BNS_STAGES; ++s) ptx::mbar_init(&full[s], 1);
            ptx::fence_barrier_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int s = 0; s < BNS_STAGES; ++s) {
                const long long tile = blockIdx.x + static_cast<long long>(s) * gridDim.x;
                if (tile < n_tiles) issue(s, tile);
            }
        }
    }
    __device__ __forceinline__ const uint4* stage(int s, int t) const {
        return reinterpret_cast<const uint4*>(ring + (s * NT + t) * BNS_TILE);
    }
};

static size_t bns_smem(int nt, int floats = 0) { return static_cast<size_t>(bns_stages(nt)) * nt * BNS_TILE + 64 + sizeof(float) * floats; }

template <bool WITH_Y>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_stream_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                            const __nv_bfloat16* __restrict__ raw, const float* __restrict__ mean,
                            const float* __restrict__ invstd, const float* __restrict__ fsc,
                            const float* __restrict__ fsh, long long n_pix, int c, int relu, float* sums) {
    extern __shared__ __align__(128) uint8_t bns_smem_buf[];
    constexpr int NT = WITH_Y ? 3 : 2;
    constexpr int BNS_STAGES = bns_stages(NT);
    float* sh = reinterpret_cast<float*>(bns_smem_buf + BNS_STAGES * NT * BNS_TILE + 64);   // [2*c]
    BnStream<NT> st;
    st.src[0] = dy; st.src[1] = raw;
    if (WITH_Y) st.src[NT - 1] = y;
    st.n_pix = n_pix; st.c = c;
    for (int i = threadIdx.x; i < 2 * c; i += 256) sh[i] = 0.f;
    st.start(bns_smem_buf);
    const int cg = c >> 3, g8 = threadIdx.x % cg, prow = threadIdx.x / cg, prows = 256 / cg;
    const bool rawmask = relu && !WITH_Y;
    float s1[8], s2[8], mu[8], fa[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = g8 * 8 + j;
        s1[j] = 0.f; s2[j] = 0.f; mu[j] = mean[ch];
        fa[j] = rawmask ? fsc[ch] : 0.f; fb[j] = rawmask ? fsh[ch] : 0.f;
    }
    int it = 0;
    for (long long tile = blockIdx.x; tile < st.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % BNS_STAGES;
        ptx::mbar_wait(&st.full[s], (it / BNS_STAGES) & 1);
        uint4 ud[BNS_CHUNKS], ur[BNS_CHUNKS], uy[BNS_CHUNKS];
#pragma unroll
        for (int k = 0; k < BNS_CHUNKS; ++k) {
            ud[k] = st.stage(s, 0)[threadIdx.x + k * 256];
            ur[k] = st.stage(s, 1)[threadIdx.x + k * 256];
            if (WITH_Y) uy[k] = st.stage(s, NT - 1)[threadIdx.x + k * 256];
        }
        __syncthreads();   // every thread has read stage s: refill it
        if (threadIdx.x == 0) {
            const long long next = tile + static_cast<long long>(BNS_STAGES) * gridDim.x;
            if (next < st.n_tiles) st.issue(s, next);
        }
        const long long px0 = tile * st.tile_px;
#pragma unroll
        for (int k = 0; k < BNS_CHUNKS; ++k) {
            if (px0 + prow + k * prows >= n_pix) continue;
            const F8 d = cvt8(ud[k]), r = cvt8(ur[k]);
            F8 yy;
            if (WITH_Y) yy = cvt8(uy[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float m = WITH_Y ? yy.v[j] : fmaf(r.v[j], fa[j], fb[j]);
                const float g = (relu && !(m > 0.f)) ? 0.f : d.v[j];
                s1[j] += g;
                s2[j] = fmaf(g, r.v[j] - mu[j], s2[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        atomicAdd(&sh[g8 * 8 + j], s1[j]);
        atomicAdd(&sh[c + g8 * 8 + j], s2[j] * invstd[g8 * 8 + j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * c; i += 256) atomicAdd(&sums[i], sh[i]);
}

template <bool WITH_Y>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_stream_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                           const __nv_bfloat16* __restrict__ raw, const float* __restrict__ mean,
                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                           const float* __restrict__ sums, const float* __restrict__ fsc, const float* __restrict__ fsh,
                           long long n_pix, int c, int relu, __nv_bfloat16* __restrict__ d_raw,
                           __nv_bfloat16* __restrict__ g_out, float* dgamma, float* dbeta) {
    extern __shared__ __align__(128) uint8_t bns_smem_buf[];
    constexpr int NT = WITH_Y ? 3 : 2;
    constexpr int BNS_STAGES = bns_stages(NT);
    BnStream<NT> st;
    st.src[0] = dy; st.src[1] = raw;
    if (WITH_Y) st.src[NT - 1] = y;
    st.n_pix = n_pix; st.c = c;
    st.start(bns_smem_buf);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < c; i += 256) {
            if (dgamma) dgamma[i] += sums[c + i];
            if (dbeta) dbeta[i] += sums[i];
        }
    }
    const int cg = c >> 3, g8 = threadIdx.x % cg, prow = threadIdx.x / cg, prows = 256 / cg;
    const bool rawmask = relu && !WITH_Y;
    const float inv_m = 1.0f / static_cast<float>(n_pix);
    float ca[8], cb[8], cc[8], fa[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = g8 * 8 + j;
        const float is = invstd[ch], gm = gamma ? gamma[ch] : 1.f;
        ca[j] = gm * is;
        cb[j] = -gm * is * is * sums[c + ch] * inv_m;
        cc[j] = -ca[j] * sums[ch] * inv_m - cb[j] * mean[ch];
        fa[j] = rawmask ? fsc[ch] : 0.f; fb[j] = rawmask ? fsh[ch] : 0.f;
    }
    int it = 0;
    for (long long tile = blockIdx.x; tile < st.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % BNS_STAGES;
        ptx::mbar_wait(&st.full[s], (it / BNS_STAGES) & 1);
        uint4 ud[BNS_CHUNKS], ur[BNS_CHUNKS], uy[BNS_CHUNKS];
#pragma unroll
        for (int k = 0; k < BNS_CHUNKS; ++k) {
            ud[k] = st.stage(s, 0)[threadIdx.x + k * 256];
            ur[k] = st.stage(s, 1)[threadIdx.x + k * 256];
            if (WITH_Y) uy[k] = st.stage(s, NT - 1)[threadIdx.x + k * 256];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long next = tile + static_cast<long long>(BNS_STAGES) * gridDim.x;
            if (next < st.n_tiles) st.issue(s, next);
        }
        const long long px0 = tile * st.tile_px;
#pragma unroll
        for (int k = 0; k < BNS_CHUNKS; ++k) {
            const long long p = px0 + prow + k * prows;
            if (p >= n_pix) continue;
            const F8 d = cvt8(ud[k]), r = cvt8(ur[k]);
            F8 yy, o, gg;
            if (WITH_Y) yy = cvt8(uy[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float m = WITH_Y ? yy.v[j] : fmaf(r.v[j], fa[j], fb[j]);
                const float g = (relu && !(m > 0.f)) ? 0.f : d.v[j];
                gg.v[j] = g;
                o.v[j] = fmaf(ca[j], g, fmaf(cb[j], r.v[j], cc[j]));
            }
            st8(d_raw + p * c + g8 * 8, o);
            if (g_out) st8(g_out + p * c + g8 * 8, gg);
        }
    }
}

// streamed forms apply to contiguous bf16 tensors whose 8-channel groups divide the 256-thread block
static bool bns_ok(int c, int dtype, int dy_ld, int raw_ld, const void* y, int y_ld, int64_t n_pix) {
    static const bool off = getenv("RTSDS_NO_BN_STREAM") != nullptr;
    if (off || dtype != RTSDS_BF16 || c < 8 || c > 2048 || (c & (c - 1)) != 0) return false;
    if (dy_ld != c || raw_ld != c || (y && y_ld != c)) return false;
    return n_pix * c >= (1 << 20);   // small maps: the register form has less fixed cost
}

// out[c] (+)= sum over pixels of x[p][c]   (bias gradients)
template <typename T>
__global__ void __launch_bounds__(256)
channel_sum_kernel(const T* __restrict__ x, int ld, long long n_pix, int c, float* out, unsigned long long* det) {
    __shared__ float sh[8][33];
    const int ch = blockIdx.x * 32 + (threadIdx.x & 31);
    const int pl = threadIdx.x >> 5;
    float acc = 0.f;
    if (ch < c)
        for (long long p = static_cast<long long>(blockIdx.y) * 8 + pl; p < n_pix; p += static_cast<long long>(gridDim.y) * 8)
            acc += to_f32(x[p * ld + ch]);
    sh[pl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (pl == 0 && ch < c) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
        if (det) det_add(det + 2 * ch, t);               // deterministic mode (common.cuh)
        else atomicAdd(&out[ch], t);
    }
}

// ---------------------------------------------------------------- max-pool backward (3x3, s2, p1)
// nn.MaxPool2d routes each window's gradient to its FIRST maximum (row-major scan, as ATen's max_pool2d
// records it).  Two stages per block, both through shared memory:
//   1. for a tile of (MP_TH+1) x (MP_TW+1) windows x 64 channels: load the 9 inputs once, find the position
//      (0..8) of the first maximum per channel (8 nibbles per 8-channel group) and stage the window's dy;
//   2. every input pixel the block owns (2*MP_TH x 2*MP_TW) gathers from the <= 4 windows that contain it.
// ~2.7 x-loads + 1.2 dy-loads per input pixel instead of ~35 in a per-pixel recomputation.
constexpr int MP_TH = 8, MP_TW = 16, MP_WIN = (MP_TH + 1) * (MP_TW + 1);
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, int h, int w, int c, int oh, int ow, int tiles_x,
                   T* __restrict__ dx) {
    __shared__ uint32_t s_idx[MP_WIN * 8];
    __shared__ F8 s_dy[MP_WIN * 8];
    const int tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int oy0 = ty * MP_TH, ox0 = tx * MP_TW;
    const int img = blockIdx.z;
    const int cg0 = blockIdx.y * 8;
    const int ncg = min(8, c / 8 - cg0);
    const T* xin = x + static_cast<long long>(img) * h * w * c + cg0 * 8;
    const T* dyin = dy + static_cast<long long>(img) * oh * ow * c + cg0 * 8;
    for (int it = threadIdx.x; it < MP_WIN * 8; it += 256) {
        const int cg = it & 7, wi = it >> 3;
        const int a = wi / (MP_TW + 1), b = wi - a * (MP_TW + 1);
        const int oy = oy0 + a, ox = ox0 + b;
        uint32_t packed = 0xffffffffu;
        F8 g;
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] = 0.f;
        if (cg < ncg && oy < oh && ox < ow) {
            float best[8];
            int pos[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; pos[j] = 15; }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int yy = oy * 2 - 1 + r;
                if (yy < 0 || yy >= h) continue;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int xx = ox * 2 - 1 + q;
                    if (xx < 0 || xx >= w) continue;
                    const F8 v = ld8(xin + (static_cast<long long>(yy) * w + xx) * c + cg * 8);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (v.v[j] > best[j] || pos[j] == 15 || v.v[j] != v.v[j]) { best[j] = v.v[j]; pos[j] = r * 3 + q; }
                }
            }
            packed = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) packed |= static_cast<uint32_t>(pos[j]) << (4 * j);
            g = ld8(dyin + (static_cast<long long>(oy) * ow + ox) * c + cg * 8);
        }
        s_idx[it] = packed;
        s_dy[it] = g;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < 2 * MP_TH * 2 * MP_TW * 8; it += 256) {
        const int cg = it & 7, pi = it >> 3;
        const int ry = pi / (2 * MP_TW), rx = pi - ry * (2 * MP_TW);
        const int iy = 2 * oy0 + ry, ix = 2 * ox0 + rx;
        if (cg >= ncg || iy >= h || ix >= w) continue;
        F8 acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
        const int a_lo = iy / 2, a_hi = (iy + 1) / 2, b_lo = ix / 2, b_hi = (ix + 1) / 2;
        for (int oy = a_lo; oy <= a_hi; ++oy) {
            for (int ox = b_lo; ox <= b_hi; ++ox) {
                const int wi = ((oy - oy0) * (MP_TW + 1) + (ox - ox0)) * 8 + cg;
                const uint32_t pk = s_idx[wi];
                const uint32_t p = static_cast<uint32_t>((iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)));
                const F8& g = s_dy[wi];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (((pk >> (4 * j)) & 15u) == p) acc.v[j] += g.v[j];
            }
        }
        st8(dx + (static_cast<long long>(img) * h * w + static_cast<long long>(iy) * w + ix) * c + (cg0 + cg) * 8, acc);
    }
}

// Backward from the recorded first-maximum positions (maxpool_idx_kernel): one thread per (input pixel, 8-channel
// group) gathers from the <= 4 windows that contain it; x is not read at all (4 B of index per 16 B of dy instead).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_idx_kernel(const uint32_t* __restrict__ idx, const T* __restrict__ dy, int n, int h, int w, int c, int oh, int ow,
                       T* __restrict__ dx, Div3 dv) {
    const int cg = c / 8;
    const long long total = static_cast<long long>(n) * h * w * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned ug, ux, uy;
        unsigned r = fdivmod(static_cast<unsigned>(i), dv.a, &ug);
        r = fdivmod(r, dv.b, &ux);
        const int img = static_cast<int>(fdivmod(r, dv.c, &uy));
        const int g = static_cast<int>(ug), ix = static_cast<int>(ux), iy = static_cast<int>(uy);
        F8 acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
        // the (up to) 2 x 2 windows containing the pixel; all 8 loads are issued before the first use
        const int oy0 = iy / 2, ox0 = ix / 2;
        uint32_t pk[4];
        F8 gd[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int oy = oy0 + (k >> 1), ox = ox0 + (k & 1);
            ok[k] = oy < oh && ox < ow && ((k >> 1) == 0 || (iy & 1)) && ((k & 1) == 0 || (ix & 1));
            const long long wi = ((static_cast<long long>(img) * oh + (ok[k] ? oy : oy0)) * ow + (ok[k] ? ox : ox0)) * cg + g;
            pk[k] = ok[k] ? __ldg(idx + wi) : 0xffffffffu;
            gd[k] = ld8(dy + wi * 8);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int oy = oy0 + (k >> 1), ox = ox0 + (k & 1);
            const uint32_t p = static_cast<uint32_t>((iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)));
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (ok[k] && ((pk[k] >> (4 * j)) & 15u) == p) acc.v[j] += gd[k].v[j];
        }
        st8(dx + i * 8, acc);
    }
}

// ---------------------------------------------------------------- adjoint bilinear resize helpers
// candidate destination range [lo, hi] that can reference source index s (checked exactly by the caller)
__device__ __forceinline__ void dst_range(int s, float rscale, int out_size, int* lo, int* hi) {
    const float inv = 1.0f / rscale;
    int a = static_cast<int>(floorf((static_cast<float>(s) - 1.0f + 0.5f) * inv - 0.5f)) - 1;
    int b = static_cast<int>(ceilf((static_cast<float>(s) + 1.0f + 0.5f) * inv - 0.5f)) + 1;
    *lo = a < 0 ? 0 : a;
    *hi = b > out_size - 1 ? out_size - 1 : b;
}

// d_src[n,y,x,c] = sum over dst pixels of bilinear weight * d_dst[n,oy,ox,coff+c]   (adjoint of gate_resize
// before the gate multiply); dgate[n,c] += sum_pix d_src * src.  8 channels per thread.
template <typename T>
__global__ void __launch_bounds__(256)
resize_bwd_nhwc_kernel(const T* __restrict__ d_dst, int dst_ld, int dst_coff, int n, int h, int w, int c, int oh, int ow,
                       float rh, float rw, const T* __restrict__ src, float* __restrict__ d_src, float* dgate,
                       unsigned long long* det) {
    extern __shared__ float sh[];     // [c] per-image partial dgate (block works on one image)
    const int cg = c / 8;
    const int img = blockIdx.y;
    for (int i = threadIdx.x; i < c; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const long long total = static_cast<long long>(h) * w * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g8 = static_cast<int>(i % cg);
        const long long r = i / cg;
        const int x = static_cast<int>(r % w), y = static_cast<int>(r / w);
        int ylo, yhi, xlo, xhi;
        dst_range(y, rh, oh, &ylo, &yhi);
        dst_range(x, rw, ow, &xlo, &xhi);
        F8 acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
        // column weights once per thread; per contributing destination row the loads go out four at a time
        constexpr int RB_MAXC = 12;
        float wxs[RB_MAXC];
#pragma unroll
        for (int k = 0; k < RB_MAXC; ++k) {
            const int ox = xlo + k;
            float wx = 0.f;
            if (ox <= xhi) {
                const Lerp lx = lerp_src(ox, rw, w);
                wx = (lx.i0 == x ? lx.l0 : 0.f) + (lx.i1 == x ? lx.l1 : 0.f);
            }
            wxs[k] = wx;
        }
        for (int oy = ylo; oy <= yhi; ++oy) {
            const Lerp ly = lerp_src(oy, rh, h);
            const float wy = (ly.i0 == y ? ly.l0 : 0.f) + (ly.i1 == y ? ly.l1 : 0.f);
            if (wy == 0.f) continue;
            const T* rowp = d_dst + ((static_cast<long long>(img) * oh + oy) * ow + xlo) * dst_ld + dst_coff + g8 * 8;
#pragma unroll
            for (int k0 = 0; k0 < RB_MAXC; k0 += 4) {
                F8 g[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (wxs[k0 + k] != 0.f) g[k] = ld8(rowp + static_cast<long long>(k0 + k) * dst_ld);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (wxs[k0 + k] != 0.f) {
                        const float wgt = wy * wxs[k0 + k];
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(wgt, g[k].v[j], acc.v[j]);
                    }
                }
            }
            for (int ox = xlo + RB_MAXC; ox <= xhi; ++ox) {      // upsampling factors above ~4.5: the rest of the range
                const Lerp lx = lerp_src(ox, rw, w);
                const float wx = (lx.i0 == x ? lx.l0 : 0.f) + (lx.i1 == x ? lx.l1 : 0.f);
                if (wx == 0.f) continue;
                const F8 g = ld8(d_dst + ((static_cast<long long>(img) * oh + oy) * ow + ox) * dst_ld + dst_coff + g8 * 8);
                const float wgt = wy * wx;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(wgt, g.v[j], acc.v[j]);
            }
        }
        const long long sp = ((static_cast<long long>(img) * h + y) * w + x) * c + g8 * 8;
        st8(d_src + sp, acc);
        if (dgate) {
            const F8 s = ld8(src + sp);
#pragma unroll
            if (det) {            // deterministic mode: per-thread products straight into the exact accumulators of dgate
#pragma unroll
                for (int j = 0; j < 8; ++j) det_add(det + 2 * (static_cast<long long>(img) * c + g8 * 8 + j), acc.v[j] * s.v[j]);
            } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&sh[g8 * 8 + j], acc.v[j] * s.v[j]);
            }
        }
    }
    if (dgate && !det) {
        __syncthreads();
        for (int i = threadIdx.x; i < c; i += blockDim.x)
            if (sh[i] != 0.f) atomicAdd(&dgate[static_cast<long long>(img) * c + i], sh[i]);
    }
}

// dx[n,p,c] = d_gated[n,p,c] * gate[n,c] + add[n,c] * add_scale     (ARM: x*gate and the GAP branch)
template <typename TO>
__global__ void __launch_bounds__(256)
gate_bwd_finish_kernel(const float* __restrict__ d_gated, const float* __restrict__ gate, const float* __restrict__ add,
                       float add_scale, int n, long long hw, int c, TO* __restrict__ dx) {
    const int cg = c / 8;
    const long long total = static_cast<long long>(n) * hw * cg;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g8 = static_cast<int>(i % cg);
        const long long p = i / cg;
        const int img = static_cast<int>(p / hw);
        const F8 d = ld8(d_gated + p * c + g8 * 8);
        const F8 g = ld8(gate + static_cast<long long>(img) * c + g8 * 8);
        F8 a;
        if (add) a = ld8(add + static_cast<long long>(img) * c + g8 * 8);
        F8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = d.v[j] * g.v[j] + (add ? a.v[j] * add_scale : 0.f);
        st8(dx + p * c + g8 * 8, o);
    }
}

// ---------------------------------------------------------------- ARM backward (tiny, [N,C])
// gate = sigmoid(gamma*xhat+beta) * mul,  xhat = BN_over_batch(lin),  lin = W pooled + b
// phase 1 (one warp per output channel co): dgamma, dbeta, db, dlin[n,co]; d(mul) -> dpooled_direct
// phase 2 (one thread per (n, ci)): dpooled[n,ci] = sum_co dlin[n,co] W[co,ci] (+ direct); dW[co,ci] += sum_n dlin*pooled
constexpr int ARMB_MAX_N = 128;
__global__ void __launch_bounds__(256)
arm_bwd_phase1_kernel(const float* __restrict__ dgate, const float* __restrict__ lin, const float* __restrict__ xhat,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mul,
                      float eps, int n, int c, float* dlin, float* dmul, float* dgamma, float* dbeta, float* dbias) {
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    if (co >= c) return;
    const float g = gamma[co], be = beta[co];
    float mean = 0.f;
    for (int i = 0; i < n; ++i) mean += lin[static_cast<long long>(i) * c + co];
    mean /= n;
    float var = 0.f;
    for (int i = 0; i < n; ++i) { const float d = lin[static_cast<long long>(i) * c + co] - mean; var += d * d; }
    var /= n;
    const float invstd = rsqrtf(var + eps);
    float sdx = 0.f, sdxx = 0.f, sg = 0.f, sb = 0.f;
    for (int i = 0; i < n; ++i) {
        const long long k = static_cast<long long>(i) * c + co;
        const float xh = xhat[k];
        const float s = 1.f / (1.f + expf(-(xh * g + be)));
        const float m = mul ? mul[k] : 1.f;
        const float dg = dgate[k];
        if (dmul) dmul[k] = dg * s;
        const float du = dg * m * s * (1.f - s);
        sg += du * xh; sb += du;
        const float dxh = du * g;
        sdx += dxh; sdxx += dxh * xh;
    }
    float sl = 0.f;
    for (int i = 0; i < n; ++i) {
        const long long k = static_cast<long long>(i) * c + co;
        const float xh = xhat[k];
        const float s = 1.f / (1.f + expf(-(xh * g + be)));
        const float m = mul ? mul[k] : 1.f;
        const float dxh = dgate[k] * m * s * (1.f - s) * g;
        const float dl = invstd * (dxh - sdx / n - xh * sdxx / n);
        dlin[k] = dl;
        sl += dl;
    }
    if (dgamma) dgamma[co] += sg;
    if (dbeta) dbeta[co] += sb;
    if (dbias) dbias[co] += sl;
}

__global__ void __launch_bounds__(256)
arm_bwd_phase2_kernel(const float* __restrict__ dlin, const float* __restrict__ w, const float* __restrict__ pooled,
                      const float* __restrict__ dmul, int n, int c, float* dpooled, float* dw) {
    // blockIdx.y = 0..n-1: dpooled rows;  blockIdx.y = n: dW (one thread per (co, ci) strip)
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= c) return;
    if (static_cast<int>(blockIdx.y) < n) {
        const int i = blockIdx.y;
        float acc = dmul ? dmul[static_cast<long long>(i) * c + ci] : 0.f;
        for (int co = 0; co < c; ++co) acc = fmaf(dlin[static_cast<long long>(i) * c + co], w[static_cast<long long>(co) * c + ci], acc);
        dpooled[static_cast<long long>(i) * c + ci] = acc;
    }
}

// dW[co][ci] += sum_n dlin[n][co] * pooled[n][ci]; one thread per (co, ci)
__global__ void __launch_bounds__(256)
arm_bwd_dw_kernel(const float* __restrict__ dlin, const float* __restrict__ pooled, int n, int c, float* dw) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(c) * c) return;
    const int co = static_cast<int>(i / c), ci = static_cast<int>(i - static_cast<long long>(co) * c);
    float acc = 0.f;
    for (int k = 0; k < n; ++k) acc = fmaf(dlin[static_cast<long long>(k) * c + co], pooled[static_cast<long long>(k) * c + ci], acc);
    dw[i] += acc;
}

// ---------------------------------------------------------------- FFM head backward
// forward: a = sigmoid(W2 relu(W1 p + b1) + b2), g = f*(1+a), z = Wc g + bc (Wc NULL: z = g)
// pass 1: dg = Wc^T dz;  da_raw[n,c] += dg*f;  dWc += dz (x) g;  dbc += dz.
// Pixels are staged through shared memory in chunks of 128.  Thread t owns pixel t of the chunk for
// da_raw (registers); for the dWc outer product lane o (< c) of every warp owns row o: it walks the
// warp's 32 pixels keeping dWc[o][0..c) and dbc[o] in registers (broadcast shared loads, no atomics).
constexpr int FB_MAXC = 32;
constexpr int FB_CHUNK = 128;
__global__ void __launch_bounds__(FB_CHUNK)
ffm_head_bwd_pass1_kernel(const float* __restrict__ dz, int dz_ld, const float* __restrict__ f, int f_ld,
                          const float* __restrict__ attn, const float* __restrict__ wc, long long hw, int c,
                          float* da_raw, float* dwc, float* dbc, unsigned long long* det) {
    __shared__ float s_w[FB_MAXC * FB_MAXC];
    __shared__ float s_dz[FB_CHUNK][FB_MAXC + 1], s_g[FB_CHUNK][FB_MAXC + 1];
    __shared__ float s_da[FB_MAXC], s_a1[FB_MAXC];
    const int img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < c * c; i += blockDim.x) s_w[i] = wc ? wc[i] : 0.f;
    if (tid < FB_MAXC) { s_da[tid] = 0.f; s_a1[tid] = tid < c ? 1.f + attn[static_cast<long long>(img) * c + tid] : 0.f; }
    __syncthreads();
    float da[FB_MAXC], dw_row[FB_MAXC];
    float db = 0.f;
#pragma unroll
    for (int k = 0; k < FB_MAXC; ++k) { da[k] = 0.f; dw_row[k] = 0.f; }
    for (long long p0 = static_cast<long long>(blockIdx.x) * FB_CHUNK; p0 < hw; p0 += static_cast<long long>(gridDim.x) * FB_CHUNK) {
        const long long p = p0 + tid;
        const bool ok = p < hw;
        const float* dzp = dz + (static_cast<long long>(img) * hw + (ok ? p : 0)) * dz_ld;
        const float* fp = f + (static_cast<long long>(img) * hw + (ok ? p : 0)) * f_ld;
        float dzv[FB_MAXC], fv[FB_MAXC];
#pragma unroll
        for (int k = 0; k < FB_MAXC; ++k) {
            dzv[k] = (ok && k < c) ? dzp[k] : 0.f;
            fv[k] = (ok && k < c) ? fp[k] : 0.f;
            s_dz[tid][k] = dzv[k];
            s_g[tid][k] = fv[k] * s_a1[k];
        }
#pragma unroll
        for (int k = 0; k < FB_MAXC; ++k) {
            if (k < c) {
                float dg = 0.f;
                if (wc) { for (int o = 0; o < c; ++o) dg = fmaf(s_w[o * c + k], dzv[o], dg); } else dg = dzv[k];
                da[k] += dg * fv[k];
            }
        }
        __syncwarp();
        if (wc && lane < c) {
            for (int q = 0; q < 32; ++q) {
                const float d = s_dz[warp * 32 + q][lane];
                db += d;
#pragma unroll
                for (int k = 0; k < FB_MAXC; ++k)
                    if (k < c) dw_row[k] = fmaf(d, s_g[warp * 32 + q][k], dw_row[k]);
            }
        }
        __syncwarp();
    }
    if (det) {
        // deterministic mode: exact accumulators laid out as [da_raw: n*c | dwc: c*c | dbc: c]; every thread adds its own
        // (fixed-order) partial sums, nothing goes through the block's shared-memory atomics
        const long long n_img = gridDim.y;
#pragma unroll
        for (int k = 0; k < FB_MAXC; ++k)          // (compile-time indices: a runtime one would push da[] into local memory)
            if (k < c) det_add(det + 2 * (static_cast<long long>(img) * c + k), da[k]);
        if (wc && lane < c) {
            if (dbc) det_add(det + 2 * (n_img * c + static_cast<long long>(c) * c + lane), db);
            if (dwc) {
#pragma unroll
                for (int k = 0; k < FB_MAXC; ++k)
                    if (k < c) det_add(det + 2 * (n_img * c + lane * c + k), dw_row[k]);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < FB_MAXC; ++k)
        if (k < c) atomicAdd(&s_da[k], da[k]);
    if (wc && lane < c) {
        if (dbc) atomicAdd(&dbc[lane], db);
        if (dwc) {
#pragma unroll
            for (int k = 0; k < FB_MAXC; ++k)
                if (k < c) atomicAdd(&dwc[lane * c + k], dw_row[k]);
        }
    }
    __syncthreads();
    for (int i = tid; i < c; i += blockDim.x) atomicAdd(&da_raw[static_cast<long long>(img) * c + i], s_da[i]);
}

// pass 2 (one block per image, tiny): attention MLP backward -> dpooled[n,c] and parameter grads
__global__ void __launch_bounds__(64)
ffm_head_bwd_pass2_kernel(const float* __restrict__ da_raw, const float* __restrict__ pooled, const float* __restrict__ attn,
                          const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2, int c,
                          float* dpooled, float* dw1, float* db1, float* dw2, float* db2, unsigned long long* det) {
    __shared__ float s_h[FB_MAXC], s_dv[FB_MAXC], s_dh[FB_MAXC];
    const int img = blockIdx.x, t = threadIdx.x;
    const float* pp = pooled + static_cast<long long>(img) * c;
    if (t < c) {
        float acc = b1[t];
        for (int k = 0; k < c; ++k) acc = fmaf(w1[t * c + k], pp[k], acc);
        s_h[t] = acc;                                        // pre-ReLU
        const float a = attn[static_cast<long long>(img) * c + t];
        s_dv[t] = da_raw[static_cast<long long>(img) * c + t] * a * (1.f - a);    // grad at conv2 output
    }
    __syncthreads();
    if (t < c) {
        float acc = 0.f;
        for (int o = 0; o < c; ++o) acc = fmaf(w2[o * c + t], s_dv[o], acc);
        s_dh[t] = s_h[t] > 0.f ? acc : 0.f;                  // grad at conv1 output (through ReLU)
        if (det) {            // deterministic mode: [db2: c | dw2: c*c | db1: c | dw1: c*c]
            det_add(det + 2 * t, s_dv[t]);
            for (int k = 0; k < c; ++k) det_add(det + 2 * (c + t * c + k), s_dv[t] * fmaxf(s_h[k], 0.f));
        } else {
        atomicAdd(&db2[t], s_dv[t]);
        for (int k = 0; k < c; ++k) atomicAdd(&dw2[t * c + k], s_dv[t] * fmaxf(s_h[k], 0.f));
        }
    }
    __syncthreads();
    if (t < c) {
        float acc = 0.f;
        for (int o = 0; o < c; ++o) acc = fmaf(w1[o * c + t], s_dh[o], acc);
        dpooled[static_cast<long long>(img) * c + t] = acc;
        if (det) {
            det_add(det + 2 * (c + c * c + t), s_dh[t]);
            for (int k = 0; k < c; ++k) det_add(det + 2 * (2 * c + c * c + t * c + k), s_dh[t] * pp[k]);
        } else {
        atomicAdd(&db1[t], s_dh[t]);
        for (int k = 0; k < c; ++k) atomicAdd(&dw1[t * c + k], s_dh[t] * pp[k]);
        }
    }
}

// pass 3 (per pixel): df = (Wc^T dz)*(1+a) + dpooled/hw
__global__ void __launch_bounds__(128)
ffm_head_bwd_pass3_kernel(const float* __restrict__ dz, int dz_ld, const float* __restrict__ attn,
                          const float* __restrict__ wc, const float* __restrict__ dpooled, long long hw, int c,
                          float* __restrict__ df, int df_ld) {
    __shared__ float s_w[FB_MAXC * FB_MAXC], s_a[FB_MAXC], s_dp[FB_MAXC];
    const int img = blockIdx.y;
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) s_w[i] = wc ? wc[i] : 0.f;
    if (threadIdx.x < c) {
        s_a[threadIdx.x] = 1.f + attn[static_cast<long long>(img) * c + threadIdx.x];
        s_dp[threadIdx.x] = dpooled[static_cast<long long>(img) * c + threadIdx.x] / static_cast<float>(hw);
    }
    __syncthreads();
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < hw;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float* dzp = dz + (static_cast<long long>(img) * hw + p) * dz_ld;
        float dzv[FB_MAXC];
#pragma unroll
        for (int k = 0; k < FB_MAXC; ++k) dzv[k] = k < c ? dzp[k] : 0.f;
        float* o = df + (static_cast<long long>(img) * hw + p) * df_ld;
#pragma unroll
        for (int k = 0; k < FB_MAXC; ++k) {
            if (k < c) {
                float dg = 0.f;
                if (wc) { for (int q = 0; q < c; ++q) dg = fmaf(s_w[q * c + k], dzv[q], dg); } else dg = dzv[k];
                o[k] = dg * s_a[k] + s_dp[k];
            }
        }
    }
}

// ---------------------------------------------------------------- adjoint of resize_to_nchw
// dz[n,y,x,c] = sum over output pixels of bilinear weight * dout[n,c,oy,ox]; one thread per (n,y,x,c)
__global__ void __launch_bounds__(256)
resize_nchw_bwd_kernel(const float* __restrict__ dout, int n, int c, int oh, int ow, int h, int w, float rh, float rw,
                       float* __restrict__ dz, int z_ld) {
    const long long total = static_cast<long long>(n) * h * w * c;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // x fastest so that neighbouring threads read neighbouring dout columns
        const int x = static_cast<int>(i % w);
        long long r = i / w;
        const int y = static_cast<int>(r % h); r /= h;
        const int ch = static_cast<int>(r % c);
        const int img = static_cast<int>(r / c);
        int ylo, yhi, xlo, xhi;
        dst_range(y, rh, oh, &ylo, &yhi);
        dst_range(x, rw, ow, &xlo, &xhi);
        const float* plane = dout + (static_cast<long long>(img) * c + ch) * oh * ow;
        // column weights once per thread (the interpolation coordinates cost ~20 instructions each: recomputing them
        // for every (row, column) candidate made this kernel instruction bound), then rows x columns of load + fma
        constexpr int RN_MAXC = 20;                    // 2*scale + 4 candidates: upsampling factors up to 8
        float wxs[RN_MAXC];
#pragma unroll
        for (int k = 0; k < RN_MAXC; ++k) {
            const int ox = xlo + k;
            float wx = 0.f;
            if (ox <= xhi) {
                const Lerp lx = lerp_src(ox, rw, w);
                wx = (lx.i0 == x ? lx.l0 : 0.f) + (lx.i1 == x ? lx.l1 : 0.f);
            }
            wxs[k] = wx;
        }
        float acc = 0.f;
        for (int oy = ylo; oy <= yhi; ++oy) {
            const Lerp ly = lerp_src(oy, rh, h);
            const float wy = (ly.i0 == y ? ly.l0 : 0.f) + (ly.i1 == y ? ly.l1 : 0.f);
            if (wy == 0.f) continue;
            const float* rowp = plane + static_cast<long long>(oy) * ow + xlo;
            float rowacc = 0.f;
#pragma unroll
            for (int k = 0; k < RN_MAXC; ++k)
                if (wxs[k] != 0.f) rowacc = fmaf(wxs[k], __ldg(rowp + k), rowacc);
            for (int ox = xlo + RN_MAXC; ox <= xhi; ++ox) {
                const Lerp lx = lerp_src(ox, rw, w);
                const float wx = (lx.i0 == x ? lx.l0 : 0.f) + (lx.i1 == x ? lx.l1 : 0.f);
                if (wx != 0.f) rowacc = fmaf(wx, __ldg(plane + static_cast<long long>(oy) * ow + ox), rowacc);
            }
            acc = fmaf(wy, rowacc, acc);
        }
        dz[((static_cast<long long>(img) * h + y) * w + x) * z_ld + ch] = acc;
    }
}


// ---------------------------------------------------------------- stem weight gradient
// dw[co][ci][r][s] += sum_{n,oy,ox} d_raw[n,oy,ox,co] * x[n,ci,oy*2-pad+r,ox*2-pad+s]   (x NCHW fp32, d_raw NHWC)
// block = one (ci, r) pair and a slab of output rows; thread = (co 0..63, s-group); shared-memory staged rows.
template <typename T, int K>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ d_raw, int n, int cin, int h, int w, int oh, int ow,
                  int pad, int rows_per_block, float* __restrict__ dw) {
    constexpr int TW = 64;                       // output columns per smem tile
    __shared__ float s_x[TW * 2 + K];            // input row segment
    __shared__ float s_d[TW][65];                // d_raw tile [ox][co]
    const int ci = blockIdx.y / K, r = blockIdx.y % K;
    const int img = blockIdx.z;
    const int oy0 = blockIdx.x * rows_per_block;
    const int co = threadIdx.x & 63, sg = threadIdx.x >> 6;      // 4 s-groups
    float acc[(K + 3) / 4];
#pragma unroll
    for (int i = 0; i < (K + 3) / 4; ++i) acc[i] = 0.f;
    for (int oy = oy0; oy < min(oy0 + rows_per_block, oh); ++oy) {
        const int iy = oy * 2 - pad + r;
        if (iy < 0 || iy >= h) continue;         // uniform across the block
        const float* xrow = x + ((static_cast<long long>(img) * cin + ci) * h + iy) * w;
        for (int ox0 = 0; ox0 < ow; ox0 += TW) {
            __syncthreads();
            for (int i = threadIdx.x; i < TW * 2 + K; i += 256) {
                const int ix = ox0 * 2 - pad + i;
                s_x[i] = (ix >= 0 && ix < w) ? xrow[ix] : 0.f;
            }
            for (int i = threadIdx.x; i < TW * 64; i += 256) {
                const int o = i >> 6, c = i & 63;
                s_d[o][c] = (ox0 + o < ow) ? to_f32(d_raw[((static_cast<long long>(img) * oh + oy) * ow + ox0 + o) * 64 + c]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < (K + 3) / 4; ++i) {
                const int s = sg + i * 4;
                if (s < K) {
                    float a = 0.f;
#pragma unroll 8
                    for (int o = 0; o < TW; ++o) a = fmaf(s_d[o][co], s_x[o * 2 + s], a);
                    acc[i] += a;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < (K + 3) / 4; ++i) {
        const int s = sg + i * 4;
        if (s < K && acc[i] != 0.f) atomicAdd(&dw[((static_cast<long long>(co) * cin + ci) * K + r) * K + s], acc[i]);
    }
}

}  // namespace rtsds

using namespace rtsds;

#define DISPATCH_T(dtype, CALL_BF16, CALL_F32, name)                       \
    do {                                                                   \
        if ((dtype) == RTSDS_BF16) { CALL_BF16; }                          \
        else if ((dtype) == RTSDS_F32) { CALL_F32; }                       \
        else { set_error(name ": bad dtype"); return RTSDS_EINVAL; }       \
    } while (0)

static int bn_bwd_reduce_impl(const void* dy, int dy_ld, const void* y, int y_ld, const void* raw, int raw_ld,
                              const float* mean, const float* invstd, const float* fsc, const float* fsh, int64_t n_pix, int c,
                              int relu, int dtype, float* sums, rtsds_stream_t s) {
    RTSDS_REQUIRE(dy && raw && mean && invstd && sums && n_pix > 0 && c > 0, "bn_bwd_reduce: bad argument");
    RTSDS_REQUIRE(!relu || y || (fsc && fsh), "bn_bwd_reduce: the ReLU mask needs y or the forward scale/shift");
    RTSDS_REQUIRE(dy_ld % 8 == 0 && raw_ld % 8 == 0 && (!relu || !y || y_ld % 8 == 0), "bn_bwd_reduce: pitches must be multiples of 8");
    const int cg = (c + 7) / 8;
    RTSDS_REQUIRE(cg <= 256 && dy_ld >= cg * 8 && raw_ld >= cg * 8, "bn_bwd_reduce: channel count / pitch");
    cudaStream_t st = as_stream(s);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, st);
    if (e != cudaSuccess) { set_error("bn_bwd_reduce: memset: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
    unsigned long long* det = nullptr;
    if (det_mode()) {             // deterministic mode: register-form kernel + exact accumulators, rounded into sums afterwards
        det = det_scratch(st, 2 * static_cast<size_t>(c));
        if (!det) return RTSDS_ECUDA;
    }
    if (!det && bns_ok(c, dtype, dy_ld, raw_ld, relu ? y : nullptr, y_ld, n_pix)) {
        const bool with_y = relu && y;
        const int nt = with_y ? 3 : 2;
        const long long n_tiles = cdiv(n_pix, static_cast<int64_t>(BNS_TILE / (c * 2)));
        const long long cap = 2LL * num_sms();
        const int grid = static_cast<int>(n_tiles < cap ? n_tiles : cap);
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(bn_bwd_reduce_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bns_smem(3, 4096)));
            cudaFuncSetAttribute(bn_bwd_reduce_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bns_smem(2, 4096)));
            attr = true;
        }
        const __nv_bfloat16 *pdy = reinterpret_cast<const __nv_bfloat16*>(dy), *py = reinterpret_cast<const __nv_bfloat16*>(y),
                            *praw = reinterpret_cast<const __nv_bfloat16*>(raw);
        if (with_y)
            bn_bwd_reduce_stream_kernel<true><<<grid, 256, bns_smem(nt, 2 * c), st>>>(pdy, py, praw, mean, invstd, fsc, fsh, n_pix, c, relu, sums);
        else
            bn_bwd_reduce_stream_kernel<false><<<grid, 256, bns_smem(nt, 2 * c), st>>>(pdy, nullptr, praw, mean, invstd, fsc, fsh, n_pix, c, relu, sums);
        count_launch();
        return check_launch("bn_bwd_reduce_stream_kernel");
    }
    const int prows = 256 / cg;
    const int grid = grid_for(cdiv(n_pix, 2), prows, 8);
    const size_t sm = sizeof(float) * 2 * cg * 8;
    DISPATCH_T(dtype,
               (bn_bwd_reduce_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>(
                   reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, reinterpret_cast<const __nv_bfloat16*>(y), y_ld,
                   reinterpret_cast<const __nv_bfloat16*>(raw), raw_ld, mean, invstd, fsc, fsh, n_pix, c, relu, sums, det)),
               (bn_bwd_reduce_kernel<float><<<grid, 256, sm, st>>>(
                   reinterpret_cast<const float*>(dy), dy_ld, reinterpret_cast<const float*>(y), y_ld,
                   reinterpret_cast<const float*>(raw), raw_ld, mean, invstd, fsc, fsh, n_pix, c, relu, sums, det)),
               "bn_bwd_reduce");
    count_launch();
    int rc = check_launch("bn_bwd_reduce_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, sums, 2 * static_cast<size_t>(c), false, st);
    return rc;
}

extern "C" int rtsds_bn_bwd_reduce(const void* dy, int dy_ld, const void* y, int y_ld, const void* raw, int raw_ld,
                                   const float* mean, const float* invstd, int64_t n_pix, int c, int relu, int dtype,
                                   float* sums, rtsds_stream_t s) {
    RTSDS_REQUIRE(!relu || y, "bn_bwd_reduce: y required for the ReLU mask");
    return bn_bwd_reduce_impl(dy, dy_ld, y, y_ld, raw, raw_ld, mean, invstd, nullptr, nullptr, n_pix, c, relu, dtype, sums, s);
}

extern "C" int rtsds_bn_bwd_reduce_rawmask(const void* dy, int dy_ld, const void* raw, int raw_ld, const float* mean,
                                           const float* invstd, const float* fwd_scale, const float* fwd_shift, int64_t n_pix,
                                           int c, int dtype, float* sums, rtsds_stream_t s) {
    return bn_bwd_reduce_impl(dy, dy_ld, nullptr, 0, raw, raw_ld, mean, invstd, fwd_scale, fwd_shift, n_pix, c, 1, dtype, sums, s);
}

static int bn_bwd_apply_impl(const void* dy, int dy_ld, const void* y, int y_ld, const void* raw, int raw_ld,
                             const float* mean, const float* invstd, const float* gamma, const float* sums, const float* fsc,
                             const float* fsh, int64_t n_pix, int c, int relu, int dtype, void* d_raw, int d_raw_ld,
                             int d_raw_dtype, void* g_out, int g_ld, float* dgamma, float* dbeta, rtsds_stream_t s) {
    RTSDS_REQUIRE(dy && raw && mean && invstd && sums && d_raw && n_pix > 0 && c > 0, "bn_bwd_apply: bad argument");
    RTSDS_REQUIRE(!relu || y || (fsc && fsh), "bn_bwd_apply: the ReLU mask needs y or the forward scale/shift");
    const int cg = (c + 7) / 8;
    RTSDS_REQUIRE(dy_ld % 8 == 0 && raw_ld % 8 == 0 && d_raw_ld % 8 == 0 && d_raw_ld >= cg * 8, "bn_bwd_apply: pitches must be multiples of 8");
    RTSDS_REQUIRE(!g_out || (g_ld % 8 == 0 && g_ld >= cg * 8), "bn_bwd_apply: g_out pitch");
    cudaStream_t st = as_stream(s);
    if (d_raw_dtype == RTSDS_BF16 && d_raw_ld == c && (!g_out || g_ld == c) &&
        bns_ok(c, dtype, dy_ld, raw_ld, relu ? y : nullptr, y_ld, n_pix)) {
        const bool with_y = relu && y;
        const int nt = with_y ? 3 : 2;
        const long long n_tiles = cdiv(n_pix, static_cast<int64_t>(BNS_TILE / (c * 2)));
        const long long cap = 2LL * num_sms();
        const int sgrid = static_cast<int>(n_tiles < cap ? n_tiles : cap);
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(bn_bwd_apply_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bns_smem(3)));
            cudaFuncSetAttribute(bn_bwd_apply_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bns_smem(2)));
            attr = true;
        }
        const __nv_bfloat16 *pdy = reinterpret_cast<const __nv_bfloat16*>(dy), *py = reinterpret_cast<const __nv_bfloat16*>(y),
                            *praw = reinterpret_cast<const __nv_bfloat16*>(raw);
        __nv_bfloat16 *pdr = reinterpret_cast<__nv_bfloat16*>(d_raw), *pg = reinterpret_cast<__nv_bfloat16*>(g_out);
        if (with_y)
            bn_bwd_apply_stream_kernel<true><<<sgrid, 256, bns_smem(nt), st>>>(pdy, py, praw, mean, invstd, gamma, sums, fsc, fsh,
                                                                               n_pix, c, relu, pdr, pg, dgamma, dbeta);
        else
            bn_bwd_apply_stream_kernel<false><<<sgrid, 256, bns_smem(nt), st>>>(pdy, nullptr, praw, mean, invstd, gamma, sums, fsc,
                                                                                fsh, n_pix, c, relu, pdr, pg, dgamma, dbeta);
        count_launch();
        return check_launch("bn_bwd_apply_stream_kernel");
    }
    const int grid = grid_for(n_pix * cg, 256);
#define BN_APPLY(T, TO)                                                                                                  \
    bn_bwd_apply_kernel<T, TO><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(dy), dy_ld, reinterpret_cast<const T*>(y), \
                                                     y_ld, reinterpret_cast<const T*>(raw), raw_ld, mean, invstd, gamma,   \
                                                     sums, fsc, fsh, n_pix, c, relu, reinterpret_cast<TO*>(d_raw), d_raw_ld, \
                                                     reinterpret_cast<T*>(g_out), g_ld, dgamma, dbeta)
    if (dtype == RTSDS_BF16 && d_raw_dtype == RTSDS_BF16) BN_APPLY(__nv_bfloat16, __nv_bfloat16);
    else if (dtype == RTSDS_F32 && d_raw_dtype == RTSDS_F32) BN_APPLY(float, float);
    else if (dtype == RTSDS_F32 && d_raw_dtype == RTSDS_BF16) BN_APPLY(float, __nv_bfloat16);
    else { set_error("bn_bwd_apply: unsupported dtype combination"); return RTSDS_EINVAL; }
#undef BN_APPLY
    count_launch();
    return check_launch("bn_bwd_apply_kernel");
}

extern "C" int rtsds_bn_bwd_apply(const void* dy, int dy_ld, const void* y, int y_ld, const void* raw, int raw_ld,
                                  const float* mean, const float* invstd, const float* gamma, const float* sums,
                                  int64_t n_pix, int c, int relu, int dtype, void* d_raw, int d_raw_ld, int d_raw_dtype,
                                  void* g_out, int g_ld, float* dgamma, float* dbeta, rtsds_stream_t s) {
    RTSDS_REQUIRE(!relu || y, "bn_bwd_apply: y required for the ReLU mask");
    return bn_bwd_apply_impl(dy, dy_ld, y, y_ld, raw, raw_ld, mean, invstd, gamma, sums, nullptr, nullptr, n_pix, c, relu, dtype,
                             d_raw, d_raw_ld, d_raw_dtype, g_out, g_ld, dgamma, dbeta, s);
}

extern "C" int rtsds_bn_bwd_apply_rawmask(const void* dy, int dy_ld, const void* raw, int raw_ld, const float* mean,
                                          const float* invstd, const float* gamma, const float* sums, const float* fwd_scale,
                                          const float* fwd_shift, int64_t n_pix, int c, int dtype, void* d_raw, int d_raw_ld,
                                          int d_raw_dtype, void* g_out, int g_ld, float* dgamma, float* dbeta, rtsds_stream_t s) {
    return bn_bwd_apply_impl(dy, dy_ld, nullptr, 0, raw, raw_ld, mean, invstd, gamma, sums, fwd_scale, fwd_shift, n_pix, c, 1, dtype,
                             d_raw, d_raw_ld, d_raw_dtype, g_out, g_ld, dgamma, dbeta, s);
}

extern "C" int rtsds_channel_sum(const void* x, int ld, int64_t n_pix, int c, int dtype, float* out, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && out && n_pix > 0 && c > 0 && ld >= c, "channel_sum: bad argument");
    long long py = cdiv(n_pix, 8 * 32);
    const long long cap = cdiv(4LL * num_sms(), cdiv(c, 32));
    if (py > cap) py = cap;
    if (py < 1) py = 1;
    dim3 grid(static_cast<unsigned>(cdiv(c, 32)), static_cast<unsigned>(py));
    unsigned long long* det = nullptr;
    if (py > 1 && det_mode()) {
        det = det_scratch(as_stream(s), static_cast<size_t>(c));
        if (!det) return RTSDS_ECUDA;
    }
    DISPATCH_T(dtype, (channel_sum_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, n_pix, c, out, det)),
               (channel_sum_kernel<float><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const float*>(x), ld, n_pix, c, out, det)), "channel_sum");
    count_launch();
    int rc = check_launch("channel_sum_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, out, static_cast<size_t>(c), true, as_stream(s));
    return rc;
}

extern "C" int rtsds_maxpool3x3s2_bwd(const void* x, const void* dy, int n, int h, int w, int c, int dtype, int ceil_mode,
                                      void* dx, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool_bwd: bad argument");
    auto osz = [&](int in) {
        int o = ceil_mode ? in / 2 + 1 : (in - 1) / 2 + 1;
        if (ceil_mode && (o - 1) * 2 >= in + 1) --o;
        return o;
    };
    const int oh = osz(h), ow = osz(w);
    // every input pixel belongs to a tile: tiles cover ceil(h/2) x ceil(w/2) window positions
    const int tiles_x = static_cast<int>(cdiv(cdiv(w, 2), MP_TW)), tiles_y = static_cast<int>(cdiv(cdiv(h, 2), MP_TH));
    RTSDS_REQUIRE(n <= 65535 && c / 8 <= 8 * 65535, "maxpool_bwd: shape out of range");
    dim3 grid(static_cast<unsigned>(tiles_x * tiles_y), static_cast<unsigned>(cdiv(c / 8, 8)), n);
    DISPATCH_T(dtype,
               (maxpool_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), h, w, c, oh, ow, tiles_x, reinterpret_cast<__nv_bfloat16*>(dx))),
               (maxpool_bwd_kernel<float><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(dy), h, w, c, oh, ow, tiles_x, reinterpret_cast<float*>(dx))),
               "maxpool_bwd");
    count_launch();
    return check_launch("maxpool_bwd_kernel");
}

extern "C" int rtsds_maxpool3x3s2_bwd_idx(const uint32_t* idx, const void* dy, int n, int h, int w, int c, int dtype,
                                          int ceil_mode, void* dx, rtsds_stream_t s) {
    RTSDS_REQUIRE(idx && dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool_bwd_idx: bad argument");
    auto osz = [&](int in) {
        int o = ceil_mode ? (in + 2 - 3 + 1) / 2 + 1 : (in + 2 - 3) / 2 + 1;
        if (ceil_mode && (o - 1) * 2 >= in + 1) --o;
        return o;
    };
    const int oh = osz(h), ow = osz(w);
    const long long total = static_cast<long long>(n) * h * w * (c / 8);
    RTSDS_REQUIRE(total < (1LL << 31), "maxpool_bwd_idx: tensor too large");
    const int grid = grid_for(total, 256, 16);
    const Div3 dv = {make_fastdiv(c / 8), make_fastdiv(w), make_fastdiv(h)};
    DISPATCH_T(dtype,
               (maxpool_bwd_idx_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(idx, reinterpret_cast<const __nv_bfloat16*>(dy), n, h, w, c, oh, ow, reinterpret_cast<__nv_bfloat16*>(dx), dv)),
               (maxpool_bwd_idx_kernel<float><<<grid, 256, 0, as_stream(s)>>>(idx, reinterpret_cast<const float*>(dy), n, h, w, c, oh, ow, reinterpret_cast<float*>(dx), dv)),
               "maxpool_bwd_idx");
    count_launch();
    return check_launch("maxpool_bwd_idx_kernel");
}

extern "C" int rtsds_resize_bwd_nhwc(const void* d_dst, int dst_ld, int dst_coff, int n, int h, int w, int c, int oh, int ow,
                                     const void* src, int dtype, float* d_src, float* dgate, rtsds_stream_t s) {
    RTSDS_REQUIRE(d_dst && d_src && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && oh > 0 && ow > 0, "resize_bwd_nhwc: bad argument");
    RTSDS_REQUIRE(dst_ld % 8 == 0 && dst_coff % 8 == 0 && (!dgate || src), "resize_bwd_nhwc: layout");
    cudaStream_t st = as_stream(s);
    if (dgate) {
        cudaError_t e = cudaMemsetAsync(dgate, 0, sizeof(float) * n * c, st);
        if (e != cudaSuccess) { set_error("resize_bwd_nhwc: memset: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
    }
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    long long bx = cdiv(static_cast<long long>(h) * w * (c / 8), 256);
    const long long cap = cdiv(8LL * num_sms(), n);
    if (bx > cap) bx = cap;
    dim3 grid(static_cast<unsigned>(bx), n);
    const size_t sm = sizeof(float) * c;
    unsigned long long* det = nullptr;
    if (dgate && det_mode()) {
        det = det_scratch(st, static_cast<size_t>(n) * c);
        if (!det) return RTSDS_ECUDA;
    }
    DISPATCH_T(dtype,
               (resize_bwd_nhwc_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_dst), dst_ld, dst_coff, n, h, w, c, oh, ow, rh, rw, reinterpret_cast<const __nv_bfloat16*>(src), d_src, dgate, det)),
               (resize_bwd_nhwc_kernel<float><<<grid, 256, sm, st>>>(reinterpret_cast<const float*>(d_dst), dst_ld, dst_coff, n, h, w, c, oh, ow, rh, rw, reinterpret_cast<const float*>(src), d_src, dgate, det)),
               "resize_bwd_nhwc");
    count_launch();
    int rc = check_launch("resize_bwd_nhwc_kernel");
    if (rc == RTSDS_OK && det) rc = det_finish(det, dgate, static_cast<size_t>(n) * c, true, st);
    return rc;
}

extern "C" int rtsds_gate_bwd_finish(const float* d_gated, const float* gate, const float* add, float add_scale, int n,
                                     int64_t hw, int c, int out_dtype, void* dx, rtsds_stream_t s) {
    RTSDS_REQUIRE(d_gated && gate && dx && n > 0 && hw > 0 && c > 0 && c % 8 == 0, "gate_bwd_finish: bad argument");
    const int grid = grid_for(static_cast<long long>(n) * hw * (c / 8), 256);
    DISPATCH_T(out_dtype,
               (gate_bwd_finish_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(s)>>>(d_gated, gate, add, add_scale, n, hw, c, reinterpret_cast<__nv_bfloat16*>(dx))),
               (gate_bwd_finish_kernel<float><<<grid, 256, 0, as_stream(s)>>>(d_gated, gate, add, add_scale, n, hw, c, reinterpret_cast<float*>(dx))),
               "gate_bwd_finish");
    count_launch();
    return check_launch("gate_bwd_finish_kernel");
}

extern "C" int rtsds_arm_gate_bwd(const float* dgate, const float* pooled, const float* lin, const float* xhat,
                                  const float* w, const float* gamma, const float* beta, const float* mul, float eps, int n,
                                  int c, float* dlin_ws, float* dmul_ws, float* dpooled, float* dw, float* dbias,
                                  float* dgamma, float* dbeta, rtsds_stream_t s) {
    RTSDS_REQUIRE(dgate && pooled && lin && xhat && w && gamma && beta && dlin_ws && dpooled, "arm_gate_bwd: NULL argument");
    RTSDS_REQUIRE(n >= 2 && c > 0, "arm_gate_bwd: needs N >= 2");
    RTSDS_REQUIRE(!mul || dmul_ws, "arm_gate_bwd: dmul workspace required when mul is given");
    cudaStream_t st = as_stream(s);
    arm_bwd_phase1_kernel<<<static_cast<int>(cdiv(c, 128)), 128, 0, st>>>(dgate, lin, xhat, gamma, beta, mul, eps, n, c, dlin_ws,
                                                                      mul ? dmul_ws : nullptr, dgamma, dbeta, dbias);
    dim3 grid(static_cast<unsigned>(cdiv(c, 128)), n);
    arm_bwd_phase2_kernel<<<grid, 128, 0, st>>>(dlin_ws, w, pooled, mul ? dmul_ws : nullptr, n, c, dpooled, dw);
    if (dw) arm_bwd_dw_kernel<<<static_cast<int>(cdiv(static_cast<long long>(c) * c, 256)), 256, 0, st>>>(dlin_ws, pooled, n, c, dw);
    count_launch(dw ? 3 : 2);
    return check_launch("arm_bwd kernels");
}

extern "C" int rtsds_ffm_head_bwd(const float* dz, int dz_ld, const float* f, int f_ld, const float* pooled,
                                  const float* attn, int n, int64_t hw, int c, const float* w1, const float* b1,
                                  const float* w2, const float* wc, float* da_ws, float* dpooled_ws, float* df, int df_ld,
                                  float* dw1, float* db1, float* dw2, float* db2, float* dwc, float* dbc, rtsds_stream_t s) {
    RTSDS_REQUIRE(dz && f && pooled && attn && w1 && b1 && w2 && da_ws && dpooled_ws && df, "ffm_head_bwd: NULL argument");
    RTSDS_REQUIRE(dw1 && db1 && dw2 && db2, "ffm_head_bwd: gradient buffers required");
    RTSDS_REQUIRE(n > 0 && hw > 0 && c > 0 && c <= FB_MAXC && dz_ld >= c && f_ld >= c && df_ld >= c, "ffm_head_bwd: bad shape");
    cudaStream_t st = as_stream(s);
    cudaError_t e = cudaMemsetAsync(da_ws, 0, sizeof(float) * n * c, st);
    if (e != cudaSuccess) { set_error("ffm_head_bwd: memset: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
    long long bx = cdiv(hw, FB_CHUNK);
    const long long cap = cdiv(2LL * num_sms(), n);
    if (bx > cap) bx = cap;
    dim3 grid(static_cast<unsigned>(bx), n);
    const bool det_on = det_mode();
    const size_t nc = static_cast<size_t>(n) * c, cc = static_cast<size_t>(c) * c;
    unsigned long long* det = nullptr;
    if (det_on) {
        det = det_scratch(st, nc + cc + c);
        if (!det) return RTSDS_ECUDA;
    }
    ffm_head_bwd_pass1_kernel<<<grid, FB_CHUNK, 0, st>>>(dz, dz_ld, f, f_ld, attn, wc, hw, c, da_ws, dwc, dbc, det);
    if (det) {
        int rc = det_finish(det, da_ws, nc, true, st);
        if (rc == RTSDS_OK && wc && dwc) rc = det_finish(det + 2 * nc, dwc, cc, true, st);
        if (rc == RTSDS_OK && wc && dbc) rc = det_finish(det + 2 * (nc + cc), dbc, static_cast<size_t>(c), true, st);
        if (rc != RTSDS_OK) return rc;
        det = det_scratch(st, 2 * c + 2 * cc);          // pass 2: [db2 | dw2 | db1 | dw1]
        if (!det) return RTSDS_ECUDA;
    }
    ffm_head_bwd_pass2_kernel<<<n, 64, 0, st>>>(da_ws, pooled, attn, w1, b1, w2, c, dpooled_ws, dw1, db1, dw2, db2, det);
    if (det) {
        int rc = det_finish(det, db2, static_cast<size_t>(c), true, st);
        if (rc == RTSDS_OK) rc = det_finish(det + 2 * c, dw2, cc, true, st);
        if (rc == RTSDS_OK) rc = det_finish(det + 2 * (c + cc), db1, static_cast<size_t>(c), true, st);
        if (rc == RTSDS_OK) rc = det_finish(det + 2 * (2 * c + cc), dw1, cc, true, st);
        if (rc != RTSDS_OK) return rc;
    }
    long long bx3 = cdiv(hw, 128);
    const long long cap3 = cdiv(8LL * num_sms(), n);
    if (bx3 > cap3) bx3 = cap3;
    dim3 grid3(static_cast<unsigned>(bx3), n);
    ffm_head_bwd_pass3_kernel<<<grid3, 128, 0, st>>>(dz, dz_ld, attn, wc, dpooled_ws, hw, c, df, df_ld);
    count_launch(3);
    return check_launch("ffm_head_bwd kernels");
}

extern "C" int rtsds_resize_to_nchw_bwd(const float* dout, int n, int c, int oh, int ow, int h, int w, float* dz, int z_ld,
                                        rtsds_stream_t s) {
    RTSDS_REQUIRE(dout && dz && n > 0 && c > 0 && oh > 0 && ow > 0 && h > 0 && w > 0 && z_ld >= c, "resize_to_nchw_bwd: bad argument");
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    const int grid = grid_for(static_cast<long long>(n) * h * w * c, 256);
    resize_nchw_bwd_kernel<<<grid, 256, 0, as_stream(s)>>>(dout, n, c, oh, ow, h, w, rh, rw, dz, z_ld);
    count_launch();
    return check_launch("resize_nchw_bwd_kernel");
}

extern "C" int rtsds_stem_conv_wgrad(const float* x, const void* d_raw, int d_dtype, int n, int cin, int h, int w, int cout,
                                     int k, int stride, int pad, float* dw_oihw, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && d_raw && dw_oihw, "stem_conv_wgrad: NULL argument");
    RTSDS_REQUIRE(cout == 64 && stride == 2 && cin >= 1 && cin <= 32, "stem_conv_wgrad: unsupported geometry");
    const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
    int rpb = static_cast<int>(cdiv(static_cast<long long>(oh) * n * cin * k, 4LL * num_sms()));
    if (rpb < 1) rpb = 1;
    dim3 grid(static_cast<unsigned>(cdiv(oh, rpb)), cin * k, n);
    cudaStream_t st = as_stream(s);
#define STEM_WG(T, KK) stem_wgrad_kernel<T, KK><<<grid, 256, 0, st>>>(x, reinterpret_cast<const T*>(d_raw), n, cin, h, w, oh, ow, pad, rpb, dw_oihw)
    if (d_dtype == RTSDS_BF16) {
        if (k == 3) STEM_WG(__nv_bfloat16, 3); else if (k == 4) STEM_WG(__nv_bfloat16, 4); else if (k == 7) STEM_WG(__nv_bfloat16, 7);
        else { set_error("stem_conv_wgrad: k=%d unsupported", k); return RTSDS_EUNSUP; }
    } else if (d_dtype == RTSDS_F32) {
        if (k == 3) STEM_WG(float, 3); else if (k == 4) STEM_WG(float, 4); else if (k == 7) STEM_WG(float, 7);
        else { set_error("stem_conv_wgrad: k=%d unsupported", k); return RTSDS_EUNSUP; }
    } else { set_error("stem_conv_wgrad: bad dtype"); return RTSDS_EINVAL; }
#undef STEM_WG
    count_launch();
    return check_launch("stem_wgrad_kernel");
}
