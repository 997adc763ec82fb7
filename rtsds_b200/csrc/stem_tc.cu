// Tensor-core stems: the two first convolutions of BiSeNet read the same image
//   context path  conv7x7 s2 p3 3->64 (torchvision resnet conv1, build_contextpath.py:19)
//   spatial path  conv3x3 s2 p1 3->64 (build_bisenet.py:24)
// and have the same output grid, and the 3x3 window is the centre of the 7x7 window.  They are fused
// into ONE implicit GEMM:  D[128 pixels, 128 = 64 cp + 64 sp channels] = A[128, K] * B[128, K]^T,
// K = 3*7*7 = 147 (padded to 192), the 3x3 weights embedded (zero-padded) in the 7x7 K layout.
//
// There is no TMA path for an im2col of a 3-channel NCHW fp32 image, so the A tile is built by
// 128 gather threads (one output pixel each).  The input patch of the tile ((2*TH+5) x (2*TW+5) x 3
// floats) is first staged in shared memory with coalesced loads -- prefetched into registers one tile
// ahead, so its L2/HBM latency hides behind the build of the current tile -- and split by column parity
// so that the stride-2 window reads are bank-conflict free; the 147 window values of a pixel are then
// shared-memory loads at compile-time offsets, converted to bf16 and written with 16-byte stores straight
// into the 128-byte-swizzled K-major UMMA layout.  (Gathering straight from global memory cost one L2
// round trip per 8 values: 5.4 us per tile, 5x the HBM time of the kernel.)
// Persistent CTAs, warp-specialised:
//   warps 0-3  gather (double-buffered A), warp 4 MMA issuer (tcgen05, 2 TMEM accumulators),
//   warps 5-8  epilogue (tcgen05.ld -> folded BN + ReLU -> bf16 NHWC to both outputs; train mode:
//              raw outputs + per-channel sum / sum of squares).
// The same gathered tile is also exactly the MN-major B operand of the weight-gradient GEMM
//   dW[128 co, K] = sum_pixels d_raw[pix, co] * im2col[pix, K]
// (stem_wgrad_tc_kernel): d_raw tiles arrive by TMA, the im2col tile is re-gathered.
#include "common.cuh"
#include "ptx.cuh"
#include <cstring>

namespace rtsds {

constexpr int SK = 192;                         // padded K (3 swizzle atoms of 64 bf16)
constexpr int SK_REAL = 147;
constexpr int S_ATOM_BYTES = 128 * 128;         // 128 rows x 128 B
constexpr int S_A_BYTES = 3 * S_ATOM_BYTES;     // one gathered tile: 48 KiB
constexpr int S_THREADS = 9 * 32;

struct StemTcParams {
    const float* x;            // NCHW fp32 [n,3,h,w]
    int n, h, w, oh, ow;
    int tile_w, tile_h, tiles_w, tiles_h, tiles_total;
    // forward
    const __nv_bfloat16* wpk;  // [128][192] bf16, K order (c, r, s) of the 7x7 window
    const float* scale;        // [128] or NULL
    const float* shift;        // [128] or NULL
    int relu;
    float* stats_cp;           // [2*64] or NULL (train)
    float* stats_sp;
    __nv_bfloat16* y_cp;       // NHWC [n,oh,ow,64]
    __nv_bfloat16* y_sp;
    // eval: MaxPool2d(3, 2, 1) of the context-path map fused into the epilogue.  pool: NHWC [n,ph,pw,64], ZERO on entry; every
    // tile max-reduces its (post-ReLU, >= 0) partial windows into it; y_cp itself is then never written.
    void* pool;
    int ph, pw;
    // wgrad
    float* dw;                 // [128][192] fp32, accumulated
    // deterministic mode (common.cuh: det_add): exact accumulators instead of the fp32 atomics — forward: [2][2*64] in the
    // layout of stats_cp then stats_sp; wgrad: [128][192] in the layout of dw.  NULL otherwise.
    unsigned long long* det;
};

// ---- input patch staging ----------------------------------------------------------------------------------
template <int TW>
struct Patch {
    static constexpr int TH = 128 / TW;
    static constexpr int PH = 2 * TH + 5, PW = 2 * TW + 5;
    // row pitch / parity-plane offset (floats): lanes of one warp span 32/TW tile rows (2 patch rows apart);
    // the pitch places those rows in disjoint bank groups
    static constexpr int RP = TW == 32 ? 72 : (TW == 16 ? 40 : 24);
    static constexpr int PO = RP / 2;
    static constexpr int N_EL = 3 * PH * PW;
    static constexpr int PRE = (N_EL + 127) / 128;       // prefetch registers per gather thread (<= 22)
    static_assert(3 * PH * RP <= 3072 && PO >= TW + 3, "patch buffer");
};
constexpr int S_PATCH_FLOATS = 3072;                 // 12 KiB per staged patch
constexpr int S_PATCHES = 3;                         // ring: the patch of tile i+2 is in flight while tile i is built

// asynchronous (cp.async) copy of tile t's patch into shared memory, zero-filled outside the image
template <int TW>
__device__ __forceinline__ void patch_issue(const StemTcParams& p, int t, int tid, float* patch) {
    using P = Patch<TW>;
    const int per = p.tiles_w * p.tiles_h;
    const int img = t / per;
    const int r = t - img * per;
    const int th = r / p.tiles_w;
    const int iy_base = th * P::TH * 2 - 3, ix_base = (r - th * p.tiles_w) * TW * 2 - 3;
    const float* xi = p.x + static_cast<long long>(img) * 3 * p.h * p.w;
#pragma unroll
    for (int k = 0; k < P::PRE; ++k) {
        const int idx = tid + k * 128;
        const int c = idx / (P::PH * P::PW), rem = idx - c * (P::PH * P::PW);
        const int prow = rem / P::PW, pcol = rem - prow * P::PW;
        const int iy = iy_base + prow, ix = ix_base + pcol;
        const bool ok = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        if (idx < P::N_EL)
            ptx::cp_async_4(patch + (c * P::PH + prow) * P::RP + (pcol & 1) * P::PO + (pcol >> 1),
                            ok ? xi + (static_cast<long long>(c) * p.h + iy) * p.w + ix : p.x, ok);
    }
}

// packed 16-bit pair maximum / 16-byte max-reduction to global memory (REDG.MAX.F16x8 / BF16x8)
template <bool F16>
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    if (F16) asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
template <bool F16>
__device__ __forceinline__ void red_max_16x8(void* dst, uint4 v) {
    if (F16) asm volatile("red.global.v4.f16x2.max.noftz [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("red.global.v4.bf16x2.max.noftz [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// one thread builds row `m` of the swizzled tile from the staged patch.  Fully unrolled: every
// k = (c*7 + r)*7 + s is a compile-time constant, so are all shared-memory offsets.
template <int TW, bool F16>
__device__ __forceinline__ void gather_row(const float* patch, uint8_t* tile, int m) {
    using P = Patch<TW>;
    const int my = m / TW, mx = m - my * TW;
    const float* pb = patch + 2 * my * P::RP + mx;
#pragma unroll
    for (int chunk = 0; chunk < SK / 8; ++chunk) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = chunk * 8 + e;
            if (k < SK_REAL) {
                const int c = k / 49, r = (k % 49) / 7, sx = k % 7;
                v[e] = pb[(c * P::PH + r) * P::RP + (sx & 1) * P::PO + (sx >> 1)];
            } else {
                v[e] = 0.f;
            }
        }
        uint4 u;
        if (F16) {
            u.x = pack_f16x2(v[0], v[1]); u.y = pack_f16x2(v[2], v[3]);
            u.z = pack_f16x2(v[4], v[5]); u.w = pack_f16x2(v[6], v[7]);
        } else {
            u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        }
        const int atom = chunk >> 3, j = chunk & 7;
        *reinterpret_cast<uint4*>(tile + atom * S_ATOM_BYTES + m * 128 + ((j ^ (m & 7)) << 4)) = u;
    }
}

// the gather warps' loop: keep two patches in flight, build the current tile, publish it
#define STEM_GATHER_LOOP(FULL_BAR, EMPTY_BAR, TILE_BASE, EXTRA)                                              \
    {                                                                                                        \
        const int m = threadIdx.x;                                                                           \
        const int gstep = static_cast<int>(gridDim.x);                                                       \
        for (int k = 0; k < 2; ++k) {                                                                        \
            if (static_cast<int>(blockIdx.x) + k * gstep < p.tiles_total)                                    \
                patch_issue<TW>(p, blockIdx.x + k * gstep, m, s_patch + k * S_PATCH_FLOATS);                 \
            ptx::cp_async_commit();                                                                          \
        }                                                                                                    \
        int it = 0;                                                                                          \
        for (int t = blockIdx.x; t < p.tiles_total; t += gstep, ++it) {                                      \
            const int buf = it & 1;                                                                          \
            const uint32_t ph = (it >> 1) & 1;                                                               \
            ptx::cp_async_wait<1>();            /* this thread's part of tile `it`'s patch has landed */     \
            asm volatile("bar.sync 2, 128;" ::: "memory");   /* everyone's has; tile it-1 is fully built */  \
            if (t + 2 * gstep < p.tiles_total)                                                               \
                patch_issue<TW>(p, t + 2 * gstep, m, s_patch + ((it + 2) % S_PATCHES) * S_PATCH_FLOATS);     \
            ptx::cp_async_commit();                                                                          \
            ptx::mbar_wait(&EMPTY_BAR[buf], ph ^ 1);                                                         \
            EXTRA                                                                                            \
            gather_row<TW, STEM_F16>(s_patch + (it % S_PATCHES) * S_PATCH_FLOATS, TILE_BASE + buf * S_A_BYTES, m);     \
            ptx::fence_proxy_async();                                                                        \
            ptx::mbar_arrive(&FULL_BAR[buf]);                                                                \
        }                                                                                                    \
        ptx::cp_async_wait<0>();                                                                             \
    }

__device__ __forceinline__ void tile_coords(const StemTcParams& p, int t, int m, int* img, int* oy, int* ox) {
    const int per = p.tiles_w * p.tiles_h;
    *img = t / per;
    const int r = t - *img * per;
    const int th = r / p.tiles_w;
    const int ml = m / p.tile_w;
    *oy = th * p.tile_h + ml;
    *ox = (r - th * p.tiles_w) * p.tile_w + (m - ml * p.tile_w);
}

__device__ __forceinline__ float warp_tsum(float (&v)[32], int lane) {
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

struct StemWgMaps { CUtensorMap d_cp, d_sp; };      // [64, ow, oh, n] bf16 each: d_raw (wgrad loads) / y (forward stores)

constexpr int SF_THREADS = 13 * 32;                 // 4 gather + 1 MMA + 8 epilogue warps
template <int TW, bool F16>
__global__ void __launch_bounds__(SF_THREADS, 1)
stem_fwd_tc_kernel(const __grid_constant__ StemWgMaps maps, const StemTcParams p) {
    constexpr bool STEM_F16 = F16;                        // operand / output type: IEEE half (eval-mode inference) or bf16
    constexpr uint32_t IDESC = ptx::umma_idesc_16(128, 128, F16);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_a = smem;                                   // 2 x 48 KiB
    uint8_t* s_b = smem + 2 * S_A_BYTES;                   // 48 KiB: weights [128 rows][192 k], K-major swizzled
    uint8_t* s_out = s_b + S_A_BYTES;                      // 2 x 16 KiB: bf16 output tiles [128 px][64 ch] (TMA store source)
    float* s_scale = reinterpret_cast<float*>(s_out + 2 * S_ATOM_BYTES);
    float* s_shift = s_scale + 128;
    float* s_stats = s_shift + 128;                        // [2*128]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(s_stats + 256);      // [2]
    uint64_t* a_empty = a_full + 2;                                     // [2]
    uint64_t* d_full = a_empty + 2;                                     // [2]
    uint64_t* d_empty = d_full + 2;                                     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
    float* s_patch = reinterpret_cast<float*>(tmem_slot + 4);          // 3 x 12 KiB staged input patches

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // weights -> smem (swizzled), scale/shift, barriers
    for (int i = threadIdx.x; i < 128 * (SK / 8); i += SF_THREADS) {
        const int row = i / (SK / 8), chunk = i - row * (SK / 8);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.wpk + row * SK + chunk * 8));
        const int atom = chunk >> 3, j = chunk & 7;
        *reinterpret_cast<uint4*>(s_b + atom * S_ATOM_BYTES + row * 128 + ((j ^ (row & 7)) << 4)) = u;
    }
    for (int i = threadIdx.x; i < 128; i += SF_THREADS) {
        s_scale[i] = p.scale ? p.scale[i] : 1.f;
        s_shift[i] = p.shift ? p.shift[i] : 0.f;
        s_stats[i] = 0.f; s_stats[128 + i] = 0.f;
    }
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&maps.d_cp); ptx::prefetch_tmap(&maps.d_sp);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&a_full[i], 128); ptx::mbar_init(&a_empty[i], 1);
            ptx::mbar_init(&d_full[i], 1); ptx::mbar_init(&d_empty[i], 256);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 4) { ptx::tmem_alloc(tmem_slot, 256); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();                              // weight tile written by generic stores, read by UMMA
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ================= gather =================
        STEM_GATHER_LOOP(a_full, a_empty, s_a, )
    } else if (warp == 4) {
        // ================= MMA =================
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                ptx::mbar_wait(&d_empty[buf], ph ^ 1);     // epilogue drained this accumulator
                ptx::mbar_wait(&a_full[buf], ph);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(s_a + buf * S_A_BYTES), sb = ptx::smem_u32(s_b);
#pragma unroll
                for (int k = 0; k < SK / 16; ++k) {
                    const uint32_t off = (k >> 2) * S_ATOM_BYTES + (k & 3) * 32;
                    ptx::umma_bf16(tmem_base + buf * 128, ptx::umma_desc_k_sw128(sa + off), ptx::umma_desc_k_sw128(sb + off),
                                   IDESC, k > 0 ? 1u : 0u);
                }
                ptx::umma_commit(&a_empty[buf]);
                ptx::umma_commit(&d_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        // Two sets of four warps: set 0 drains the context-path half of the accumulator (columns 0..63), set 1 the
        // spatial-path half.  Each thread owns one pixel; the bf16 tile is staged in shared memory (swizzled, so
        // the 16-byte stores are conflict free) and leaves with ONE TMA store per set -- per-thread global
        // stores of 128-byte-strided rows cost 32 LSU passes per instruction.
        const int set = (warp - 5) >> 2;                   // 0: context path, 1: spatial path
        const int q = warp & 3;                            // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;
        const bool issuer = q == 0 && lane == 0;
        const bool affine = p.scale != nullptr || p.shift != nullptr || p.relu;
        uint8_t* stage = s_out + set * S_ATOM_BYTES;
        const CUtensorMap* omap = set == 0 ? &maps.d_cp : &maps.d_sp;
        const int bar_id = 3 + set;
        int it = 0;
        for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            int img, oy, ox;
            tile_coords(p, t, m, &img, &oy, &ox);
            const bool valid = oy < p.oh && ox < p.ow;
            ptx::mbar_wait(&d_full[buf], ph);
            ptx::tc_fence_after();
            uint4 packed[8];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int c0 = set * 64 + half * 32;
                uint32_t r[32];
                ptx::tmem_ld_32x32(tmem_base + buf * 128 + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
                ptx::tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (p.stats_cp) {
                    float tt[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) tt[j] = valid ? v[j] : 0.f;
                    const float s1 = warp_tsum(tt, lane);
#pragma unroll
                    for (int j = 0; j < 32; ++j) tt[j] = valid ? v[j] * v[j] : 0.f;
                    const float s2 = warp_tsum(tt, lane);
                    if (p.det) {
                        const int ch = c0 + lane;                      // 0..63 context path, 64..127 spatial path
                        unsigned long long* d = p.det + 2 * ((ch >> 6) * 128 + (ch & 63));
                        det_add(d, s1);
                        det_add(d + 2 * 64, s2);
                    } else {
                        atomicAdd(&s_stats[c0 + lane], s1);
                        atomicAdd(&s_stats[128 + c0 + lane], s2);
                    }
                }
                if (affine) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float xv = fmaf(v[j], s_scale[c0 + j], s_shift[c0 + j]);
                        v[j] = p.relu ? fmaxf(xv, 0.f) : xv;
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    if (F16) {
                        u.x = pack_f16x2(v[g * 8 + 0], v[g * 8 + 1]); u.y = pack_f16x2(v[g * 8 + 2], v[g * 8 + 3]);
                        u.z = pack_f16x2(v[g * 8 + 4], v[g * 8 + 5]); u.w = pack_f16x2(v[g * 8 + 6], v[g * 8 + 7]);
                    } else {
                        u.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); u.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                        u.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); u.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                    }
                    packed[half * 4 + g] = u;
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&d_empty[buf]);               // accumulator is in registers: the MMA warp may reuse it
            if (issuer) ptx::bulk_wait_read0();            // previous tile's TMA store has read the staging tile
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(stage + m * 128 + ((j ^ (m & 7)) << 4)) = packed[j];
            ptx::fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (set == 0 && p.pool) {
                // ---- fused 3x3 stride-2 max-pool (build_contextpath.py:21, nn.MaxPool2d(3, 2, 1)): the staged tile covers rows
                // [y0, y0+th) x columns [x0, x0+tw) of the stem output; it holds part of the windows of pooled rows
                // y0/2 .. (y0+th)/2 and columns x0/2 .. (x0+tw)/2.  Each (pooled pixel, 8-channel group) takes the maximum
                // over the window pixels inside this tile and max-reduces it into the (zeroed) pooled map: values are
                // post-ReLU (>= 0) and max is exact, so the result is bit-identical to pooling the stored map.
                int i0, y0, x0;
                tile_coords(p, t, 0, &i0, &y0, &x0);
                const int th = p.tile_h, tw = p.tile_w;
                const int npy = th / 2 + 1, npx = tw / 2 + 1;
                const int items = npy * npx * 8;
                for (int it2 = m; it2 < items; it2 += 128) {
                    const int g = it2 & 7;
                    const int pp = it2 >> 3;
                    const int ly = pp / npx, lx = pp - ly * npx;
                    const int py = y0 / 2 + ly, px = x0 / 2 + lx;
                    if (py >= p.ph || px >= p.pw) continue;
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
                    bool any = false;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int ry = 2 * py + dy - y0;
                        if (ry < 0 || ry >= th || 2 * py + dy >= p.oh) continue;
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int rx = 2 * px + dx - x0;
                            if (rx < 0 || rx >= tw || 2 * px + dx >= p.ow) continue;
                            const int mm = ry * tw + rx;
                            const uint4 u = *reinterpret_cast<const uint4*>(stage + mm * 128 + ((g ^ (mm & 7)) << 4));
                            acc.x = max16x2<F16>(acc.x, u.x); acc.y = max16x2<F16>(acc.y, u.y);
                            acc.z = max16x2<F16>(acc.z, u.z); acc.w = max16x2<F16>(acc.w, u.w);
                            any = true;
                        }
                    }
                    if (any)
                        red_max_16x8<F16>(reinterpret_cast<uint8_t*>(p.pool) +
                                          (((static_cast<long long>(i0) * p.ph + py) * p.pw + px) * 64 + g * 8) * 2, acc);
                }
            } else if (issuer) {
                int i0, y0, x0;
                tile_coords(p, t, 0, &i0, &y0, &x0);
                ptx::tma_store_4d(omap, stage, 0, x0, y0, i0);     // clipped at the image border
                ptx::bulk_commit();
            }
        }
        if (issuer) ptx::bulk_wait0();
        if (p.stats_cp && !p.det) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int i = threadIdx.x - 160;              // 0..255
            if (i < 128) {
                float* dst = i < 64 ? p.stats_cp : p.stats_sp;
                atomicAdd(&dst[i & 63], s_stats[i]);
                atomicAdd(&dst[64 + (i & 63)], s_stats[128 + i]);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 256); }
}

// ---- weight gradient: dW[128 co][192 k] += sum_pixels d_raw[pix][co] * im2col[pix][k] --------------------
template <int TW>
__global__ void __launch_bounds__(S_THREADS, 1)
stem_wgrad_tc_kernel(const __grid_constant__ StemWgMaps maps, const StemTcParams p) {
    constexpr bool STEM_F16 = false;                      // training operands are bf16
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(128, SK, 1, 1);     // both operands MN-major, N = 192
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_b = smem;                                   // 2 x 48 KiB gathered im2col tiles
    uint8_t* s_a = smem + 2 * S_A_BYTES;                   // 2 x 32 KiB d_raw tiles (two 64-channel boxes)
    uint64_t* b_full = reinterpret_cast<uint64_t*>(s_a + 2 * 2 * S_ATOM_BYTES);
    uint64_t* a_full = b_full + 2;
    uint64_t* empty = a_full + 2;
    uint64_t* d_full = empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
    float* s_patch = reinterpret_cast<float*>(tmem_slot + 2);          // 3 x 12 KiB staged input patches
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&maps.d_cp); ptx::prefetch_tmap(&maps.d_sp);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&b_full[i], 128); ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&empty[i], 1); }
        ptx::mbar_init(d_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 4) { ptx::tmem_alloc(tmem_slot, 256); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool have = static_cast<int>(blockIdx.x) < p.tiles_total;

    if (warp < 4) {
        STEM_GATHER_LOOP(b_full, empty, s_b,
            if (m == 0) {       /* d_raw tiles by TMA (out-of-range pixels are zero-filled: they contribute nothing) */
                int img; int oy; int ox;
                tile_coords(p, t, 0, &img, &oy, &ox);
                ptx::mbar_expect_tx(&a_full[buf], 2 * S_ATOM_BYTES);
                ptx::tma_load_4d(s_a + buf * 2 * S_ATOM_BYTES, &maps.d_cp, &a_full[buf], 0, ox, oy, img);
                ptx::tma_load_4d(s_a + buf * 2 * S_ATOM_BYTES + S_ATOM_BYTES, &maps.d_sp, &a_full[buf], 0, ox, oy, img);
            })
    } else if (warp == 4) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                ptx::mbar_wait(&a_full[buf], ph);
                ptx::mbar_wait(&b_full[buf], ph);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(s_a + buf * 2 * S_ATOM_BYTES), sb = ptx::smem_u32(s_b + buf * S_A_BYTES);
#pragma unroll
                for (int k = 0; k < 128 / 16; ++k)       // 16 pixels per MMA
                    ptx::umma_bf16(tmem_base, ptx::umma_desc_mn_sw128(sa + k * 2048, S_ATOM_BYTES),
                                   ptx::umma_desc_mn_sw128(sb + k * 2048, S_ATOM_BYTES), IDESC, (it > 0 || k > 0) ? 1u : 0u);
                ptx::umma_commit(&empty[buf]);
            }
            if (have) ptx::umma_commit(d_full);
        }
        __syncwarp();
    } else if (have) {
        const int q = warp & 3;
        const int co = q * 32 + lane;
        ptx::mbar_wait(d_full, 0);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < SK; c0 += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < SK_REAL) {
                    if (p.det) det_add(p.det + 2 * (co * SK + c0 + j), __uint_as_float(r[j]));
                    else atomicAdd(&p.dw[co * SK + c0 + j], __uint_as_float(r[j]));
                }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 256); }
}

// pack both stems' OIHW fp32 weights into the fused [128][192] bf16 operand
template <typename T>
__global__ void stem_pack_kernel(const float* __restrict__ w7, const float* __restrict__ w3, T* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 128 * SK) return;
    const int row = i / SK, k = i - row * SK;
    float v = 0.f;
    if (k < SK_REAL) {
        const int c = k / 49, r = (k % 49) / 7, s = k % 7;
        if (row < 64) v = w7[(row * 3 + c) * 49 + r * 7 + s];
        else if (r >= 2 && r <= 4 && s >= 2 && s <= 4) v = w3[((row - 64) * 3 + c) * 9 + (r - 2) * 3 + (s - 2)];
    }
    out[i] = from_f32<T>(v);
}

// scatter the fused fp32 gradient [128][192] back to the two OIHW gradients (accumulate), and clear it
__global__ void stem_unpack_kernel(float* __restrict__ dw, float* __restrict__ g7, float* __restrict__ g3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 128 * SK) return;
    const int row = i / SK, k = i - row * SK;
    const float v = dw[i];
    dw[i] = 0.f;
    if (k >= SK_REAL) return;
    const int c = k / 49, r = (k % 49) / 7, s = k % 7;
    if (row < 64) { if (g7) g7[(row * 3 + c) * 49 + r * 7 + s] += v; }
    else if (g3 && r >= 2 && r <= 4 && s >= 2 && s <= 4) g3[((row - 64) * 3 + c) * 9 + (r - 2) * 3 + (s - 2)] += v;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int stem_geometry(int n, int h, int w, StemTcParams* p) {
    p->n = n; p->h = h; p->w = w;
    p->oh = (h - 1) / 2 + 1; p->ow = (w - 1) / 2 + 1;
    p->tile_w = p->ow >= 32 ? 32 : (p->ow >= 16 ? 16 : 8);
    p->tile_h = 128 / p->tile_w;
    p->tiles_w = static_cast<int>(cdiv(p->ow, p->tile_w));
    p->tiles_h = static_cast<int>(cdiv(p->oh, p->tile_h));
    const long long tt = static_cast<long long>(n) * p->tiles_w * p->tiles_h;
    if (tt >= (1LL << 30)) { set_error("stem_tc: too many tiles"); return RTSDS_EINVAL; }
    p->tiles_total = static_cast<int>(tt);
    return RTSDS_OK;
}

// tensor maps over two NHWC bf16 [n, oh, ow, 64] tensors, one box = one 128-pixel tile (128-byte swizzle)
static int stem_make_maps(const StemTcParams& p, const void* base_cp, const void* base_sp, StemWgMaps* maps, const char* who) {
    static EncodeTiledFn2 enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            set_error("%s: cuTensorMapEncodeTiled unavailable", who);
            return RTSDS_ECUDA;
        }
        enc = reinterpret_cast<EncodeTiledFn2>(fp);
    }
    memset(maps, 0, sizeof(*maps));
    const void* bases[2] = {base_cp, base_sp};
    CUtensorMap* ms[2] = {&maps->d_cp, &maps->d_sp};
    for (int i = 0; i < 2; ++i) {
        cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(p.ow), static_cast<cuuint64_t>(p.oh), static_cast<cuuint64_t>(p.n)};
        cuuint64_t strides[3] = {64 * 2, static_cast<cuuint64_t>(p.ow) * 64 * 2, static_cast<cuuint64_t>(p.oh) * p.ow * 64 * 2};
        cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.tile_w), static_cast<cuuint32_t>(p.tile_h), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(ms[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(bases[i]), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed: %d", who, static_cast<int>(r)); return RTSDS_ECUDA; }
    }
    return RTSDS_OK;
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_stem_pack_weights(const float* w7_oihw, const float* w3_oihw, int dtype, void* wpk, rtsds_stream_t s) {
    RTSDS_REQUIRE(w7_oihw && w3_oihw && wpk, "stem_pack_weights: NULL argument");
    RTSDS_REQUIRE(is_16bit(dtype), "stem_pack_weights: dtype must be bf16 or fp16");
    const int g = static_cast<int>(cdiv(128 * SK, 256));
    if (dtype == RTSDS_F16) stem_pack_kernel<__half><<<g, 256, 0, as_stream(s)>>>(w7_oihw, w3_oihw, reinterpret_cast<__half*>(wpk));
    else stem_pack_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(s)>>>(w7_oihw, w3_oihw, reinterpret_cast<__nv_bfloat16*>(wpk));
    count_launch();
    return check_launch("stem_pack_kernel");
}

// Fused tensor-core stems.  x: NCHW fp32 [n,3,h,w]; wpk from rtsds_stem_pack_weights; scale/shift: fp32 [128]
// (context-path BN in 0..63, spatial-path BN in 64..127) or NULL; stats_*: fp32 [2*64] train-mode sums or NULL.
static int stem_pair_fwd_impl(const float* x, int n, int h, int w, const void* wpk, const float* scale,
                              const float* shift, int relu, float* stats_cp, float* stats_sp, int dtype, void* y_cp,
                              void* y_sp, void* pool, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && wpk && (y_cp || pool) && y_sp && n > 0 && h > 0 && w > 0, "stem_pair_tc_fwd: bad argument");
    RTSDS_REQUIRE(!pool || (relu && !stats_cp), "stem_pair_tc_fwd_pool: the fused max-pool needs the ReLU epilogue (eval mode)");
    RTSDS_REQUIRE(is_16bit(dtype), "stem_pair_tc_fwd: dtype must be bf16 or fp16");
    RTSDS_REQUIRE((stats_cp == nullptr) == (stats_sp == nullptr), "stem_pair_tc_fwd: stats go together");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    StemTcParams p;
    memset(&p, 0, sizeof(p));
    rc = stem_geometry(n, h, w, &p);
    if (rc != RTSDS_OK) return rc;
    p.x = x; p.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk); p.scale = scale; p.shift = shift; p.relu = relu;
    p.stats_cp = stats_cp; p.stats_sp = stats_sp;
    p.y_cp = reinterpret_cast<__nv_bfloat16*>(y_cp); p.y_sp = reinterpret_cast<__nv_bfloat16*>(y_sp);
    p.pool = pool; p.ph = (p.oh + 2 - 3) / 2 + 1; p.pw = (p.ow + 2 - 3) / 2 + 1;
    RTSDS_REQUIRE(!pool || (p.tile_h % 2 == 0 && p.tile_w % 2 == 0 && (reinterpret_cast<uintptr_t>(pool) & 15) == 0),
                  "stem_pair_tc_fwd_pool: tile geometry / alignment");
    StemWgMaps maps;
    rc = stem_make_maps(p, y_cp ? y_cp : y_sp, y_sp, &maps, "stem_pair_tc_fwd");
    if (rc != RTSDS_OK) return rc;
    const size_t smem = 1024 + 3 * S_A_BYTES + 2 * S_ATOM_BYTES + 512 * 4 + 8 * 8 + 16 + S_PATCHES * S_PATCH_FLOATS * 4;
    static bool done = false;
    if (!done) {
        const void* fns[6] = {(const void*)stem_fwd_tc_kernel<32, false>, (const void*)stem_fwd_tc_kernel<16, false>,
                              (const void*)stem_fwd_tc_kernel<8, false>, (const void*)stem_fwd_tc_kernel<32, true>,
                              (const void*)stem_fwd_tc_kernel<16, true>, (const void*)stem_fwd_tc_kernel<8, true>};
        for (const void* f : fns) {
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("stem_pair_tc_fwd: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
        }
        done = true;
    }
    const int grid = p.tiles_total < num_sms() ? p.tiles_total : num_sms();
    const bool f16 = dtype == RTSDS_F16;
    if (stats_cp && det_mode()) {
        p.det = det_scratch(as_stream(s), 256);
        if (!p.det) return RTSDS_ECUDA;
    }
#define STEM_FWD(TWv)                                                                                         \
    do {                                                                                                      \
        if (f16) stem_fwd_tc_kernel<TWv, true><<<grid, SF_THREADS, smem, as_stream(s)>>>(maps, p);            \
        else stem_fwd_tc_kernel<TWv, false><<<grid, SF_THREADS, smem, as_stream(s)>>>(maps, p);               \
    } while (0)
    if (p.tile_w == 32) STEM_FWD(32);
    else if (p.tile_w == 16) STEM_FWD(16);
    else STEM_FWD(8);
#undef STEM_FWD
    count_launch();
    rc = check_launch("stem_fwd_tc_kernel");
    if (rc == RTSDS_OK && p.det) rc = det_finish(p.det, stats_cp, 128, true, as_stream(s));
    if (rc == RTSDS_OK && p.det) rc = det_finish(p.det + 2 * 128, stats_sp, 128, true, as_stream(s));
    return rc;
}

extern "C" int rtsds_stem_pair_tc_fwd(const float* x, int n, int h, int w, const void* wpk, const float* scale,
                                      const float* shift, int relu, float* stats_cp, float* stats_sp, int dtype, void* y_cp,
                                      void* y_sp, rtsds_stream_t s) {
    RTSDS_REQUIRE(y_cp, "stem_pair_tc_fwd: NULL y_cp");
    return stem_pair_fwd_impl(x, n, h, w, wpk, scale, shift, relu, stats_cp, stats_sp, dtype, y_cp, y_sp, nullptr, s);
}

// Eval-mode form with nn.MaxPool2d(3, 2, 1) of the context-path stem fused in: the 1/2-resolution map is never written,
// `pool` (NHWC [n, (oh-1)/2+1, (ow-1)/2+1, 64], 16-bit) must be ZERO on entry and receives the pooled map.
extern "C" int rtsds_stem_pair_tc_fwd_pool(const float* x, int n, int h, int w, const void* wpk, const float* scale,
                                           const float* shift, int dtype, void* pool, void* y_sp, rtsds_stream_t s) {
    RTSDS_REQUIRE(pool, "stem_pair_tc_fwd_pool: NULL pool");
    return stem_pair_fwd_impl(x, n, h, w, wpk, scale, shift, 1, nullptr, nullptr, dtype, nullptr, y_sp, pool, s);
}

// Fused weight gradient of both stems.  d_raw_*: NHWC bf16 [n,oh,ow,64]; dw_ws: fp32 [128*192] scratch that is
// zero on entry and zero again on return; g7/g3: OIHW fp32 gradients, accumulated (NULL = frozen).
extern "C" int rtsds_stem_pair_tc_wgrad(const float* x, int n, int h, int w, const void* d_raw_cp, const void* d_raw_sp,
                                        float* dw_ws, float* g7_oihw, float* g3_oihw, rtsds_stream_t s) {
    RTSDS_REQUIRE(x && d_raw_cp && d_raw_sp && dw_ws && n > 0 && h > 0 && w > 0, "stem_pair_tc_wgrad: bad argument");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    StemTcParams p;
    memset(&p, 0, sizeof(p));
    rc = stem_geometry(n, h, w, &p);
    if (rc != RTSDS_OK) return rc;
    p.x = x; p.dw = dw_ws;
    StemWgMaps maps;
    rc = stem_make_maps(p, d_raw_cp, d_raw_sp, &maps, "stem_pair_tc_wgrad");
    if (rc != RTSDS_OK) return rc;
    const size_t smem = 1024 + 2 * S_A_BYTES + 4 * S_ATOM_BYTES + 7 * 8 + 16 + S_PATCHES * S_PATCH_FLOATS * 4;
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(stem_wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_wgrad_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_wgrad_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("stem_pair_tc_wgrad: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
        done = true;
    }
    const int grid = p.tiles_total < num_sms() ? p.tiles_total : num_sms();
    cudaStream_t st = as_stream(s);
    if (det_mode()) {
        p.det = det_scratch(st, 128 * SK);
        if (!p.det) return RTSDS_ECUDA;
    }
    if (p.tile_w == 32) stem_wgrad_tc_kernel<32><<<grid, S_THREADS, smem, st>>>(maps, p);
    else if (p.tile_w == 16) stem_wgrad_tc_kernel<16><<<grid, S_THREADS, smem, st>>>(maps, p);
    else stem_wgrad_tc_kernel<8><<<grid, S_THREADS, smem, st>>>(maps, p);
    if (p.det) {                                   // dw_ws (zero on entry) receives the rounded exact sums, then unpacks as usual
        rc = det_finish(p.det, dw_ws, 128 * SK, true, st);
        if (rc != RTSDS_OK) return rc;
    }
    stem_unpack_kernel<<<static_cast<int>(cdiv(128 * SK, 256)), 256, 0, st>>>(dw_ws, g7_oihw, g3_oihw);
    count_launch(2);
    return check_launch("stem_wgrad_tc kernels");
}
