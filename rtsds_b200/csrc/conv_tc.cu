// tcgen05 / TMEM implicit-GEMM convolution fed by TMA (sm_100a).
//
// Replaces nn.Conv2d (+ folded BatchNorm / bias + residual + ReLU/LeakyReLU)
// for every conv with cin % 64 == 0 on the hot path:
//   models/bisenet/build_bisenet.py:11-18 (ConvBlock), :64 (FFM conv),
//   torchvision BasicBlock convs via models/bisenet/build_contextpath.py:22-25,
//   models/deeplabv2/deeplabv2.py:13-24 (Bottleneck), :59-61 (ASPP),
//   models/domain_shift/adversarial/model.py:46-48 (discriminator conv2-4).
//
// GEMM view:  D[M = 128 output pixels, N = BLOCK_N output channels]
//             += A[M, K = 64 input channels of one filter tap] * B[N, K]^T
// looped over (filter tap, 64-channel chunk).
//   * A tile: ONE 4-D TMA box [64 ch, tile_w, tile_h, 1 image] of the NHWC bf16
//     activation tensor, shifted by the tap offset; out-of-bounds pixels are
//     zero-filled by TMA (= the conv padding).  The box lands in shared memory
//     as 128 rows x 128 B with the 128-byte swizzle: exactly the canonical
//     K-major SWIZZLE_128B UMMA operand, no register staging.
//     Stride-2 convs use one TMA map per input parity (h%2, w%2), so every tap
//     is again a dense box.
//   * B tile: 2-D TMA box [64 k, BLOCK_N rows] of the packed weights
//     [cout_pad][tap][cin] (K-major, SWIZZLE_128B).
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread issues
//     tcgen05.mma, accumulators in TMEM), warps 2-5 = epilogue
//     (tcgen05.ld -> scale/shift/residual/activation -> bf16/fp32 NHWC store,
//     optional per-channel sum / sum-of-squares for train-mode BatchNorm).
//   * multi-stage smem ring with full/empty mbarriers; tcgen05.commit releases
//     a stage back to the producer and finally signals the epilogue.
//   * split_k > 1: each z-slice of the grid reduces a share of the k-blocks
//     and writes raw fp32 partials; a finishing kernel sums the slices in a
//     fixed order (deterministic) and applies the epilogue.
#include "common.cuh"
#include "ptx.cuh"
#include "tap_problem.cuh"
#include <mutex>
#include <cstring>
#include <cstdlib>

namespace rtsds {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;              // bf16 elements = 128 B = one swizzle atom
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KiB
constexpr int TC_THREADS = 192;

struct TcMaps {
    CUtensorMap a[4];   // activation views (one per input parity for stride 2)
    CUtensorMap b;      // packed weights
};

struct TcParams {
    int n_img, oh, ow;            // output grid the M tiles cover
    int tile_w, tile_h, tiles_w, tiles_h;
    int cout, cout_pad;
    int n_taps, kchunks;          // k-blocks = n_taps * kchunks
    int stages;
    int split_k;
    int cluster_reduce;           // split-K slices of one tile form a thread-block cluster and reduce through DSMEM
    int mc;                       // 2: pairs of N tiles (cluster y) share every A tile — each CTA loads half of it and TMA multicasts
    int mc_dw, mc_dh;             // box-coordinate step (w, h) of the second half of the A tile
    int res_pre;                  // the 16-bit residual row may be prefetched whole (Cout is a multiple of the N tile)
    int rbuf_bytes;               // cluster split-K: receive buffer [split][128/split rows][BLOCK_N+4] fp32 after the TMA ring
    int b_early;                  // weight tiles of the first ring pass are fetched BEFORE the programmatic-dependency wait
    unsigned long long* trace;    // debug: 16 time stamps per CTA (rtsds_debug_conv_trace), NULL in production
    int n_tiles_n, total_tiles, tiles_per_cta;      // persistent kernel: tile id = n_tile * m_tiles + m_tile
    int halo_d, halo_rows, a_stage_bytes, halo_baseoff;   // halo mode: dilation, rows of the halo tile, bytes per A stage
    signed char tap_dh[TAP_MAX], tap_dw[TAP_MAX], tap_map[TAP_MAX];
    short tap_kb[TAP_MAX];            // first weight k-block of each tap
    long long out_sn, out_sh, out_sw;     // element strides of y
    long long res_sn, res_sh, res_sw;     // element strides of residual
    const float* scale;
    const float* shift;
    const void* residual;                 // same dtype as y
    float* stats;
    int dbg;                              // RTSDS_TC_DBG bit mask, persistent kernel only (bound-finding experiments; results are
                                          // WRONG when set): 1 no global stores, 2 no tcgen05.ld, 4 no MMA, 8 no A-tile TMA
    unsigned long long* det;              // deterministic mode: [2*cout] exact accumulators (common.cuh: det_add) instead of stats
    void* y;
    float* partial;                       // split-K slices [split][M_total][cout_pad]
    int out_dtype, act;
    float slope;
    int f16;                              // operands (and 16-bit outputs / residuals) are IEEE half instead of bf16
    float* gap_out;                       // [n_img][cout] fp32, += gap_scale * sum over pixels of the final output (or NULL)
    float gap_scale;
};

// Reduce 32 per-thread values across the 32 lanes of a warp so that lane j
// ends up with the sum of element j (31 shuffles).
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
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

constexpr int TR_PITCH = 36;                                  // floats per row of the per-warp transpose scratch (16-byte aligned rows)
constexpr int TR_BYTES = 4 * 32 * TR_PITCH * 4;               // four epilogue warps

// Epilogue of one 32-column chunk of an accumulator row: optional train-mode BatchNorm statistics of the raw value,
// scale/shift (folded BatchNorm or bias), residual add, activation, 16-byte stores.  Shared by both conv kernels.
// rpre != NULL: the 16-bit residual of this row (all BLOCK_N channels, 16-byte pieces) was fetched into registers while the
// main loop ran, instead of paying its latency here.
// MODE 0 compiles the train-mode BatchNorm sums (and the deterministic-mode call) out: the eval-mode instantiation of
// the one-tile kernel keeps the register allocation and schedule it had before those paths existed (measured: their mere
// presence cost the batch-1 frame 1.3 %).
// MODE: 0 = lean (no BatchNorm sums, no debug paths), 1 = train (BatchNorm sums through fp32 atomics, no debug paths),
// 2 = full (sums incl. the deterministic mode, time stamps, A-tile multicast, stage-removal switches).
template <int BLOCK_N, int MODE = 2>
__device__ __forceinline__ void tc_epilogue_chunk(const TcParams& p, float (&v)[32], int c0, int n0, bool valid, long long out_off,
                                                  long long res_off, const float* s_scale, const float* s_shift, float* s_stats,
                                                  int lane, const uint4* rpre = nullptr, float* s_tr = nullptr) {
    constexpr bool STATS = MODE >= 1, FULL = MODE == 2;
    const int co0 = n0 + c0;
            if (STATS && p.stats) {
            float s1, s2;
            if (s_tr) {
                // column sums of this warp's 32 x 32 block through a shared-memory transpose (s_tr: this warp's own
                // [32][TR_PITCH] floats): 8 row stores + 32 conflict-free column loads per thread instead of two
                // 31-shuffle butterflies with their selects (measured 0.65 us of a 1.45 us chunk, epi_timeline.py)
                float4* rowp = reinterpret_cast<float4*>(s_tr + lane * TR_PITCH);
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    rowp[g] = valid ? make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
                float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float x0 = s_tr[i * TR_PITCH + lane], x1 = s_tr[(i + 1) * TR_PITCH + lane];
                    a0 += x0; a1 += x1;
                    b0 = fmaf(x0, x0, b0); b1 = fmaf(x1, x1, b1);
                }
                __syncwarp();
                s1 = a0 + a1; s2 = b0 + b1;
            } else {
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = valid ? v[j] : 0.0f;
            s1 = warp_transpose_sum(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = valid ? v[j] * v[j] : 0.0f;
            s2 = warp_transpose_sum(t, lane);
            }
            if (FULL && p.det) {          // order-independent: this warp's 32-row sums go straight into the exact accumulators
                if (co0 + lane < p.cout) det_add2(p.det + 2 * (co0 + lane), s1, p.det + 2 * (p.cout + co0 + lane), s2);
            } else {
                atomicAdd(&s_stats[c0 + lane], s1);
                atomicAdd(&s_stats[BLOCK_N + c0 + lane], s2);
            }
        }

        const bool full = (co0 + 32 <= p.cout);
        const bool out16 = p.out_dtype != RTSDS_F32;
        const bool f16 = p.out_dtype == RTSDS_F16;
        if (valid) {
            // scale/shift (folded BatchNorm or bias) -> + residual -> activation, all in registers
            // (the four epilogue warps sit on four different schedulers, one warp each: nothing hides latency, so every
            // per-element branch or scalar shared load costs its full latency — uniform switches are hoisted out of the
            // element loops and the per-channel constants come in as 16-byte broadcast loads)
            // (train-mode forward writes the RAW conv output and dgrad has no affine at all: both skip this — measured
            // 0.34 us of a 0.8 us chunk on the training shapes, tools/debug/epi_timeline.py)
            if (p.scale != nullptr || p.shift != nullptr) {
                const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
                const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 a = sc4[g], b = sh4[g];
                    v[4 * g + 0] = fmaf(v[4 * g + 0], a.x, b.x); v[4 * g + 1] = fmaf(v[4 * g + 1], a.y, b.y);
                    v[4 * g + 2] = fmaf(v[4 * g + 2], a.z, b.z); v[4 * g + 3] = fmaf(v[4 * g + 3], a.w, b.w);
                }
            }
            if (FULL && p.trace && threadIdx.x == 64 && c0 == 0) p.trace[32ull * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) + 16] = clock64();
            if (p.residual) {
                if (out16 && full) {
                    uint4 rv[4];
                    if (rpre) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) rv[g] = rpre[c0 / 8 + g];
                    } else {
                        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + res_off + co0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) rv[g] = __ldg(rp + g);            // four independent loads in flight
                    }
                    if (f16) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float2 a = unpack_f16x2(rv[g].x), b = unpack_f16x2(rv[g].y), c = unpack_f16x2(rv[g].z), d = unpack_f16x2(rv[g].w);
                            v[g * 8 + 0] += a.x; v[g * 8 + 1] += a.y; v[g * 8 + 2] += b.x; v[g * 8 + 3] += b.y;
                            v[g * 8 + 4] += c.x; v[g * 8 + 5] += c.y; v[g * 8 + 6] += d.x; v[g * 8 + 7] += d.y;
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float2 a = unpack_bf16x2(rv[g].x), b = unpack_bf16x2(rv[g].y), c = unpack_bf16x2(rv[g].z), d = unpack_bf16x2(rv[g].w);
                            v[g * 8 + 0] += a.x; v[g * 8 + 1] += a.y; v[g * 8 + 2] += b.x; v[g * 8 + 3] += b.y;
                            v[g * 8 + 4] += c.x; v[g * 8 + 5] += c.y; v[g * 8 + 6] += d.x; v[g * 8 + 7] += d.y;
                        }
                    }
                } else if (out16) {
                    const uint16_t* res = reinterpret_cast<const uint16_t*>(p.residual) + res_off + co0;
                    {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (co0 + j < p.cout) v[j] += ld_16(res, j, f16);
                    }
                } else {
                    const float* res = reinterpret_cast<const float*>(p.residual) + res_off + co0;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 rv = __ldg(reinterpret_cast<const float4*>(res + j));
                            v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (co0 + j < p.cout) v[j] += res[j];
                    }
                }
            }
            if (p.act == RTSDS_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            } else if (p.act == RTSDS_ACT_LRELU) {
                const float slope = p.slope;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : v[j] * slope;
            }
            if (FULL && p.trace && threadIdx.x == 64 && c0 == 0) p.trace[32ull * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) + 17] = clock64();
        }
        if (p.gap_out) {
            // global average pool of THIS layer's output fused into its epilogue (ARM AdaptiveAvgPool2d(1) / context-path
            // tail, build_bisenet.py:46, build_contextpath.py:27-28): per-channel sums of the final fp32 values
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = valid ? v[j] : 0.0f;
            const float s1 = warp_transpose_sum(t, lane);
            s_stats[(threadIdx.x >> 5 & 3) * BLOCK_N + c0 + lane] += s1;      // this warp's own slot: fixed summation order
        }
        if (valid && !(FULL && (p.dbg & 1))) {
            if (out16 && full) {
                uint4 o[4];
                if (f16) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        o[g].x = pack_f16x2(v[g * 8 + 0], v[g * 8 + 1]); o[g].y = pack_f16x2(v[g * 8 + 2], v[g * 8 + 3]);
                        o[g].z = pack_f16x2(v[g * 8 + 4], v[g * 8 + 5]); o[g].w = pack_f16x2(v[g * 8 + 6], v[g * 8 + 7]);
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        o[g].x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); o[g].y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                        o[g].z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); o[g].w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                    }
                }
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y) + out_off + co0);
#pragma unroll
                for (int g = 0; g < 4; ++g) dst[g] = o[g];
            } else if (out16) {
                uint16_t* dst = reinterpret_cast<uint16_t*>(p.y) + out_off + co0;
                {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + j < p.cout) st_16(dst, j, v[j], f16);
                }
            } else {
                float* dst = reinterpret_cast<float*>(p.y) + out_off + co0;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + j < p.cout) dst[j] = v[j];
                }
            }
        }
    }

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    constexpr bool STATS = MODE >= 1, FULL = MODE == 2;       // (MODE: see tc_epilogue_chunk)
    constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
    constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    const uint32_t IDESC = ptx::umma_idesc_16(TC_BLOCK_M, BLOCK_N, p.f16 != 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + static_cast<size_t>(stages) * TC_A_BYTES;
    float* rbuf = reinterpret_cast<float*>(smem_b + static_cast<size_t>(stages) * B_BYTES);
    float* s_scale = reinterpret_cast<float*>(smem_b + static_cast<size_t>(stages) * B_BYTES + p.rbuf_bytes);
    float* s_shift = s_scale + BLOCK_N;
    float* s_stats = s_shift + BLOCK_N;                     // [2*BLOCK_N] BatchNorm sums, or [4 warps][BLOCK_N] fused-pool slots
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stats + 4 * BLOCK_N);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* tmem_full_bar = empty_bar + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- tile coordinates ----
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int th = trem / p.tiles_w;
    const int h0 = th * p.tile_h;
    const int w0 = (trem - th * p.tiles_w) * p.tile_w;
    const int n0 = blockIdx.y * BLOCK_N;
    const int kb_total = p.n_taps * p.kchunks;
    const int kb_begin = static_cast<int>((static_cast<long long>(kb_total) * blockIdx.z) / p.split_k);
    const int kb_end = static_cast<int>((static_cast<long long>(kb_total) * (blockIdx.z + 1)) / p.split_k);
    // the lean (STATS = false) instantiation is the production eval kernel: no time stamps, no A-tile multicast either
    const int mc_ = FULL ? p.mc : 1;
    unsigned long long* trace = (FULL && p.trace) ? p.trace + 32ull * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
    if (trace && threadIdx.x == 0) { trace[0] = ptx::globaltimer(); trace[1] = clock64(); }

    // cluster = (1, mc, split_k): rank = y + mc * z.  y pairs share A tiles (multicast), z slices share the output tile
    const bool clustered = p.cluster_reduce || mc_ > 1;
    const uint32_t crank = clustered ? ptx::cluster_ctarank() : 0u;
    const uint32_t yrank = mc_ > 1 ? crank % static_cast<uint32_t>(mc_) : 0u;
    const uint32_t zrank = mc_ > 1 ? crank / static_cast<uint32_t>(mc_) : crank;
    const uint16_t pair_mask = static_cast<uint16_t>(((1u << mc_) - 1u) << (zrank * mc_));

    // ---- one-time setup ----
    if (warp == 0 && lane == 0) {
        if (p.n_taps > 0) {
            ptx::prefetch_tmap(&maps.b);
            ptx::prefetch_tmap(&maps.a[0]);
        }
        for (int i = 0; i < stages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], mc_ > 1 ? mc_ : 1);     // multicast: every CTA of the pair must release the stage
        }
        ptx::mbar_init(tmem_full_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < 4 * BLOCK_N; i += TC_THREADS - 64) s_stats[i] = 0.0f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    // peers write into this CTA's shared memory (multicast TMA / pushed split-K partials): it must exist, i.e. the CTA must
    // have started.  Multicast needs that before the first load (full sync); the pushes only before the epilogue, so there
    // the arrive is here and every thread waits at the end of its main-loop role (long since complete by then).
    const bool started_barrier = p.cluster_reduce && mc_ <= 1;
    if (mc_ > 1) ptx::cluster_sync_all();
    else if (started_barrier) ptx::cluster_arrive();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the
    // tail of the previous kernel in the stream; global memory is only touched after the dependency has resolved.
    // The weights do not depend on the previous kernel: the producer thread requests the B tiles of the first ring pass
    // BEFORE waiting, so their L2/HBM latency overlaps the predecessor's tail as well.
    const int kb_pre = (p.b_early && mc_ <= 1) ? min(stages, kb_end - kb_begin) : 0;
    if (trace && threadIdx.x == 0) trace[2] = clock64();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kb_pre; ++i) {
            const int kb = kb_begin + i;
            const int tap = kb / p.kchunks;
            const int cc = kb - tap * p.kchunks;
            ptx::mbar_expect_tx(&full_bar[i], TC_A_BYTES + B_BYTES);
            ptx::tma_load_2d(smem_b + static_cast<size_t>(i) * B_BYTES, &maps.b, &full_bar[i], (p.tap_kb[tap] + cc) * TC_BLOCK_K, n0);
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (trace && threadIdx.x == 0) trace[3] = clock64();
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BLOCK_N; i += TC_THREADS - 64) {
            const int co = n0 + i;
            s_scale[i] = (p.scale && co < p.cout) ? p.scale[co] : 1.0f;
            s_shift[i] = (p.shift && co < p.cout) ? p.shift[co] : 0.0f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
    }

    if (warp == 0) {
        // =================== TMA producer ===================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // (tap, channel chunk) of the k-block advance incrementally: a division by the runtime chunk count per TMA
            // issue sat on the critical path of this one-thread loop
            int tap = kb_begin / p.kchunks;
            int cc = kb_begin - tap * p.kchunks;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const bool pre = kb - kb_begin < kb_pre;          // first ring pass: stage empty by construction, B already in flight
                if (!pre) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full_bar[stage], TC_A_BYTES + B_BYTES);
                }
                if (mc_ > 1) {
                    // this CTA fetches half yrank of the A tile (the maps' box is half a tile) and multicasts it to the pair;
                    // the other half arrives from the peer, on this same barrier
                    ptx::tma_load_4d_mc(smem_a + static_cast<size_t>(stage) * TC_A_BYTES + yrank * (TC_A_BYTES / 2),
                                        &maps.a[p.tap_map[tap]], &full_bar[stage], cc * TC_BLOCK_K,
                                        w0 + p.tap_dw[tap] + static_cast<int>(yrank) * p.mc_dw,
                                        h0 + p.tap_dh[tap] + static_cast<int>(yrank) * p.mc_dh, img, pair_mask);
                } else
                ptx::tma_load_4d(smem_a + static_cast<size_t>(stage) * TC_A_BYTES,
                                 &maps.a[p.tap_map[tap]], &full_bar[stage], cc * TC_BLOCK_K,
                                 w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], img);
                if (!pre)
                    ptx::tma_load_2d(smem_b + static_cast<size_t>(stage) * B_BYTES, &maps.b, &full_bar[stage],
                                     (p.tap_kb[tap] + cc) * TC_BLOCK_K, n0);
                if (++stage == stages) { stage = 0; phase ^= 1; }
                if (++cc == p.kchunks) { cc = 0; ++tap; }
            }
            if (trace) trace[4] = clock64();
        }
        __syncwarp();
        if (started_barrier) ptx::cluster_wait();
    } else if (warp == 1) {
        // =================== MMA issuer ===================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                if (trace && kb == kb_begin) trace[5] = clock64();
                const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_a + static_cast<size_t>(stage) * TC_A_BYTES));
                const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_b + static_cast<size_t>(stage) * B_BYTES));
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                    // +32 B per 16-element K step inside the 128-B swizzle atom
                    ptx::umma_bf16(tmem_base, da + 2 * k, db + 2 * k, IDESC,
                                   (kb > kb_begin || k > 0) ? 1u : 0u);
                }
                if (mc_ > 1) ptx::umma_commit_mc(&empty_bar[stage], pair_mask);   // releases the stage in BOTH CTAs of the pair
                else ptx::umma_commit(&empty_bar[stage]);      // frees the smem stage when the MMAs retire
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (kb_end > kb_begin) ptx::umma_commit(tmem_full_bar);   // accumulator complete
            if (trace) trace[6] = clock64();
        }
        __syncwarp();
        if (started_barrier) ptx::cluster_wait();
    } else {
        // =================== epilogue (4 warps, 128 TMEM lanes) ===================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // GEMM row = pixel inside the tile
        const int hl = row / p.tile_w;
        const int oh = h0 + hl;
        const int ow = w0 + (row - hl * p.tile_w);
        const bool valid = (oh < p.oh) && (ow < p.ow);
        const bool have_acc = kb_end > kb_begin;     // a tap-less problem (e.g. odd pixels of a 1x1 s2 dgrad) is all zeros
        // residual row -> registers while the main loop runs (narrow tiles only: 16 bytes x BLOCK_N/8 per thread)
        constexpr bool RES_PRE = BLOCK_N <= 64;
        uint4 rres[RES_PRE ? BLOCK_N / 8 : 1];
        const bool res_pre = RES_PRE && p.residual && p.res_pre && valid && p.split_k == 1;
        if (res_pre) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + img * p.res_sn + oh * p.res_sh +
                                                            ow * p.res_sw + n0);
#pragma unroll
            for (int g = 0; g < (RES_PRE ? BLOCK_N / 8 : 1); ++g) rres[g] = __ldg(rp + g);
        }
        if (have_acc) {
            ptx::mbar_wait(tmem_full_bar, 0);
            ptx::tc_fence_after();
        }
        if (started_barrier) ptx::cluster_wait();
        if (trace && threadIdx.x == 64) trace[7] = clock64();

        const long long m_total = static_cast<long long>(p.n_img) * p.oh * p.ow;
        const long long pix_lin = (static_cast<long long>(img) * p.oh + oh) * p.ow + ow;
        const long long out_off = img * p.out_sn + oh * p.out_sh + ow * p.out_sw;
        const long long res_off = img * p.res_sn + oh * p.res_sh + ow * p.res_sw;

        // One 32-column chunk of this thread's accumulator row.
        auto do_chunk = [&](uint32_t (&r)[32], int c0) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            const int co0 = n0 + c0;
            if (p.cluster_reduce) {
                // PUSH: this fp32 partial row goes straight into the receive buffer of the CTA that owns the row
                // (slot = this CTA's split rank).  Remote stores are fire-and-forget: no DSMEM load latency anywhere.
                const int rows_per = TC_BLOCK_M / p.split_k;
                const int owner = row / rows_per;
                const float* dst = rbuf + (static_cast<int>(zrank) * rows_per + (row - owner * rows_per)) * (BLOCK_N + 4) + c0;
                const uint32_t ra = ptx::mapa_u32(dst, yrank + static_cast<uint32_t>(owner * (mc_ > 1 ? mc_ : 1)));
#pragma unroll
                for (int j = 0; j < 32; j += 4) ptx::st_dsmem_f4(ra + j * 4, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            } else if (p.split_k > 1) {
                if (valid) {
                    float* dst = p.partial + (static_cast<long long>(blockIdx.z) * m_total + pix_lin) * p.cout_pad + co0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            } else {
                // (transpose scratch of the BatchNorm sums: the operand ring, idle once the accumulator is complete — not with
                // A-tile multicast, where a peer's TMA may still write into this CTA's ring)
                tc_epilogue_chunk<BLOCK_N, MODE>(p, v, c0, n0, valid, out_off, res_off, s_scale, s_shift, s_stats, lane, res_pre ? rres : nullptr,
                                                  (STATS && mc_ <= 1) ? reinterpret_cast<float*>(smem) + (warp & 3) * 32 * TR_PITCH : nullptr);
            }
        };
        // (two chunks per trip — both TMEM loads in flight, two interleaved instruction streams — was measured: 168
        // registers, small spills, no gain; the chunk time is dominated by the row-per-lane 16-byte global stores)
        constexpr int CW = 32;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += CW) {
            uint32_t ra[32], rb[32];
            if (have_acc) {
                ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, ra);
                if (CW == 64) ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0 + 32, rb);
                ptx::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) ra[j] = rb[j] = 0u;
            }
            if (trace && threadIdx.x == 64) trace[c0 == 0 ? 13 : 15] = clock64();
            do_chunk(ra, c0);
            if (CW == 64) do_chunk(rb, c0 + 32);
            if (trace && threadIdx.x == 64 && c0 == 0) trace[14] = clock64();
        }
        if (STATS && p.stats && !(FULL && p.det) && p.split_k == 1) {
            // all 4 epilogue warps have added their rows: named barrier 1, 128 threads
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) {
                const int co = n0 + i;
                if (co < p.cout) {
                    atomicAdd(&p.stats[co], s_stats[i]);
                    atomicAdd(&p.stats[p.cout + co], s_stats[BLOCK_N + i]);
                }
            }
        }
        if (p.gap_out && p.split_k == 1) {
            // this CTA's partial of the fused pool: [image][tile of the image][cout], summed by the consumer in tile order
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) {
                const int co = n0 + i;
                if (co < p.cout)
                    p.gap_out[(static_cast<long long>(img) * tiles_per_img + trem) * p.cout + co] =
                        ((s_stats[i] + s_stats[BLOCK_N + i]) + (s_stats[2 * BLOCK_N + i] + s_stats[3 * BLOCK_N + i])) * p.gap_scale;
            }
        }
    }

    if (trace && threadIdx.x == 64) trace[8] = clock64();
    if (p.cluster_reduce) {
        // ---- split-K reduction inside the cluster: every CTA has pushed its partial rows into the receive buffer of the
        // rank that owns them; CTA r now sums rows [r*128/split, (r+1)*128/split) over the `split` slots of its OWN shared
        // memory (slot order = deterministic) and runs the epilogue on them.  No workspace round trip through L2, no
        // finish kernel, no remote loads; nobody touches a peer's memory after this barrier, so CTAs exit independently.
        // (requesting the residual of the first reduction task BEFORE this barrier was measured: the four live 16-byte
        // registers across the barrier cost the lean kernel 24 registers and 1 % of the frame — not kept)
        ptx::cluster_sync_all();
        if (trace && threadIdx.x == 64) trace[11] = clock64();
        if (warp >= 2) {
            const int split = p.split_k;
            const int rows_per = TC_BLOCK_M / split;                   // 64, 32 or 16
            const int rank = static_cast<int>(zrank);
            const int ntasks = rows_per * (BLOCK_N / 32);
            for (int task = threadIdx.x - 64; task < ntasks; task += 128) {
                const int chunk = task / rows_per;
                const int lrow = task - chunk * rows_per;
                const int row = rank * rows_per + lrow;
                const int c0 = chunk * 32;
                const int hl = row / p.tile_w;
                const int oh = h0 + hl;
                const int ow = w0 + (row - hl * p.tile_w);
                const bool valid = (oh < p.oh) && (ow < p.ow);
                const long long out_off = img * p.out_sn + oh * p.out_sh + ow * p.out_sw;
                const long long res_off = img * p.res_sn + oh * p.res_sh + ow * p.res_sw;
                const float* src = rbuf + lrow * (BLOCK_N + 4) + c0;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
                for (int q = 0; q < split; ++q) {                      // slice order: deterministic sum
                    const float4* sq = reinterpret_cast<const float4*>(src + q * rows_per * (BLOCK_N + 4));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = sq[j];
                        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                    }
                }
                tc_epilogue_chunk<BLOCK_N, MODE>(p, v, c0, n0, valid, out_off, res_off, s_scale, s_shift, s_stats, lane, nullptr,
                                                  (STATS && mc_ <= 1) ? reinterpret_cast<float*>(smem) + (warp & 3) * 32 * TR_PITCH : nullptr);
            }
            if (STATS && p.stats && !(FULL && p.det)) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) {
                    const int co = n0 + i;
                    if (co < p.cout) {
                        atomicAdd(&p.stats[co], s_stats[i]);
                        atomicAdd(&p.stats[p.cout + co], s_stats[BLOCK_N + i]);
                    }
                }
            }
            if (p.gap_out) {        // partial index: (tile of the image, cluster rank) — every rank reduced its own rows
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int rank_ = static_cast<int>(zrank);
                for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) {
                    const int co = n0 + i;
                    if (co < p.cout)
                        p.gap_out[((static_cast<long long>(img) * tiles_per_img + trem) * p.split_k + rank_) * p.cout + co] =
                            ((s_stats[i] + s_stats[BLOCK_N + i]) + (s_stats[2 * BLOCK_N + i] + s_stats[3 * BLOCK_N + i])) * p.gap_scale;
                }
            }
        }
        if (trace && threadIdx.x == 64) trace[12] = clock64();
    } else if (mc_ > 1) {
        ptx::cluster_sync_all();                    // nobody leaves while the peer's MMA commits still arrive on its barriers
    }

    // ---- teardown ----
    ptx::tc_fence_before();
    __syncthreads();
    if (trace && threadIdx.x == 64) { trace[9] = clock64(); trace[10] = ptx::globaltimer(); }
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================
// Persistent variant for grids of many tiles (training batches, stems): one CTA per SM walks a contiguous range of
// output tiles.
//   * the weights of the CTA's N tile stay RESIDENT in shared memory when all their K blocks fit (3x3 convs with
//     <= 128 input channels, 1x1 convs, stems): only activation tiles stream through the TMA ring;
//   * two TMEM accumulators: the epilogue of tile i (tcgen05.ld, BatchNorm statistics, stores) overlaps the TMA / MMA
//     main loop of tile i+1;
//   * barrier init, TMEM allocation and descriptor prefetch are paid once per CTA instead of once per tile.
// =====================================================================================
// HALO (3x3 stride-1 convs, tile = 8 x 16 output pixels): instead of one TMA box per filter tap, ONE box of the
// (16+2d) x 16-pixel input neighbourhood per 64-channel chunk lands in shared memory (row pitch 16 pixels = 2048 B,
// a multiple of the 1024-byte swizzle pattern) and each tap's A operand is a WINDOW into it: UMMA descriptor start =
// tile + ((r*d)*16 + s*d) * 128 B, 8-row groups 2048 B apart.  4x fewer TMA rows than the tap-wise form — the TMA row
// rate (~128 B per 8 clk per SM), not HBM or the tensor pipe, is what bounds the tap-wise kernels on 64/128-channel layers.
//
// EPI = number of epilogue warp groups (4 warps each).  The epilogue of these kernels is LATENCY bound, not instruction
// bound: one warp per scheduler spends ~1200 cycles on the ~100 instructions of a 32-column chunk (tcgen05.ld + wait,
// convert, strided 16-byte stores; tools/debug/epi_timeline.py), which caps a CTA at ~1 output tile per 3-6 us while the MMA
// loop of the 64..256-channel training layers needs 1-2 us.  With EPI = 2 a second group drains the NEXT tile at the same
// time (accumulator ring of 2*EPI TMEM buffers; tile j of the CTA -> buffer j % (2*EPI), group j % EPI), two warps per
// scheduler hide each other's latency, and every group keeps its own scale/shift/statistics staging and named barrier.
// STATS = false: the instantiation for launches without BatchNorm sums and without the debug switches (dgrad, eval) — those
// paths are compiled out, as in conv_tc_kernel.
template <int BLOCK_N, bool B_RESIDENT, bool HALO, int EPI, int MODE>
__global__ void __launch_bounds__(64 + 128 * EPI, 1)
conv_tcp_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    constexpr bool STATS = MODE >= 1, FULL = MODE == 2;       // (MODE: see tc_epilogue_chunk)
    const int dbg_ = FULL ? p.dbg : 0;
    constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
    constexpr int NACC = 2 * EPI;
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;
    const uint32_t IDESC = ptx::umma_idesc_16(TC_BLOCK_M, BLOCK_N, p.f16 != 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    const int kb_total = p.n_taps * p.kchunks;
    const int a_stage = HALO ? p.a_stage_bytes : TC_A_BYTES;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + static_cast<size_t>(stages) * a_stage;
    const size_t b_slots = B_RESIDENT ? static_cast<size_t>(kb_total) : static_cast<size_t>(stages);
    float* s_epi = reinterpret_cast<float*>(smem_b + b_slots * B_BYTES);      // per group: scale | shift | stats [2] (BLOCK_N each)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_epi + EPI * 4 * BLOCK_N);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* acc_full = empty_bar + stages;                // [NACC]
    uint64_t* acc_empty = acc_full + NACC;                  // [NACC]
    uint64_t* b_full = acc_empty + NACC;
    uint64_t* b_free = b_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_free + 1);
    // transpose scratch of the BatchNorm sums (only allocated when p.stats: tcp_smem_bytes' `extra`)
    float* s_tr_all = (STATS && p.stats) ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~uintptr_t(15)) : nullptr;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = p.tiles_w * p.tiles_h * p.n_img;
    const int t_begin = blockIdx.x * p.tiles_per_cta;
    const int t_end = min(t_begin + p.tiles_per_cta, p.total_tiles);
    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&maps.b);
        ptx::prefetch_tmap(&maps.a[0]);
        for (int i = 0; i < stages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < NACC; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 128); }
        ptx::mbar_init(b_full, 1);
        ptx::mbar_init(b_free, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, TMEM_COLS); ptx::tmem_relinquish(); }
    if (warp >= 2)
        for (int i = threadIdx.x - 64; i < EPI * 4 * BLOCK_N; i += 128 * EPI) s_epi[i] = 0.0f;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // =================== TMA producer ===================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, bfree_phase = 0;
            int cur_n = -1;
            for (int tile = t_begin; tile < t_end; ++tile) {
                const int n_tile = tile / m_tiles, m_tile = tile - n_tile * m_tiles;
                const int img = m_tile / tiles_per_img;
                const int trem = m_tile - img * tiles_per_img;
                const int th = trem / p.tiles_w;
                const int h0 = th * p.tile_h, w0 = (trem - th * p.tiles_w) * p.tile_w;
                const int n0 = n_tile * BLOCK_N;
                if (B_RESIDENT && n_tile != cur_n) {
                    if (cur_n >= 0) { ptx::mbar_wait(b_free, bfree_phase); bfree_phase ^= 1; }   // MMAs of the old N tile retired
                    ptx::mbar_expect_tx(b_full, static_cast<uint32_t>(kb_total) * B_BYTES);
                    for (int kb = 0, tap = 0, cc = 0; kb < kb_total; ++kb) {
                        ptx::tma_load_2d(smem_b + static_cast<size_t>(kb) * B_BYTES, &maps.b, b_full, (p.tap_kb[tap] + cc) * TC_BLOCK_K, n0);
                        if (++cc == p.kchunks) { cc = 0; ++tap; }
                    }
                    cur_n = n_tile;
                }
                if (HALO) {
                    for (int cc = 0; cc < p.kchunks; ++cc) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (dbg_ & 8) { ptx::mbar_arrive(&full_bar[stage]); if (++stage == stages) { stage = 0; phase ^= 1; } continue; }
                        ptx::mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.halo_rows) * 16 * 128);
                        ptx::tma_load_4d(smem_a + static_cast<size_t>(stage) * a_stage, &maps.a[1], &full_bar[stage], cc * TC_BLOCK_K,
                                         w0 - p.halo_d, h0 - p.halo_d, img);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                } else {
                for (int kb = 0, tap = 0, cc = -1; kb < kb_total; ++kb) {
                    if (++cc == p.kchunks) { cc = 0; ++tap; }              // (tap, chunk) advance without a division per TMA issue
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if ((dbg_ & 8) && B_RESIDENT) { ptx::mbar_arrive(&full_bar[stage]); if (++stage == stages) { stage = 0; phase ^= 1; } continue; }
                    ptx::mbar_expect_tx(&full_bar[stage], TC_A_BYTES + (B_RESIDENT ? 0 : B_BYTES));
                    ptx::tma_load_4d(smem_a + static_cast<size_t>(stage) * TC_A_BYTES, &maps.a[p.tap_map[tap]], &full_bar[stage],
                                     cc * TC_BLOCK_K, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], img);
                    if (!B_RESIDENT)
                        ptx::tma_load_2d(smem_b + static_cast<size_t>(stage) * B_BYTES, &maps.b, &full_bar[stage],
                                         (p.tap_kb[tap] + cc) * TC_BLOCK_K, n0);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =================== MMA issuer ===================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, bfull_phase = 0;
            int cur_n = -1;
            for (int tile = t_begin; tile < t_end; ++tile) {
                const int n_tile = tile / m_tiles;
                if (B_RESIDENT && n_tile != cur_n) {
                    ptx::mbar_wait(b_full, bfull_phase);
                    bfull_phase ^= 1;
                    cur_n = n_tile;
                }
                const int j = tile - t_begin, acc = j % NACC;
                const uint32_t use_parity = static_cast<uint32_t>(j / NACC) & 1u;      // k-th use of this accumulator: parity k & 1
                ptx::mbar_wait(&acc_empty[acc], use_parity ^ 1);            // the epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * ACC_COLS;
                if (HALO) {
                    for (int cc = 0; cc < p.kchunks; ++cc) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t a_base = ptx::smem_u32(smem_a + static_cast<size_t>(stage) * a_stage);
                        for (int tap = 0; tap < p.n_taps; ++tap) {
                            // window of this tap: rows (dh + d) .. +16, columns (dw + d) .. +8 of the 16-pixel-wide halo tile
                            const uint32_t row_off = static_cast<uint32_t>((p.tap_dh[tap] + p.halo_d) * 16 + (p.tap_dw[tap] + p.halo_d));
                            const uint32_t a_addr = a_base + row_off * 128u;
                            const uint64_t da = ptx::umma_desc_k_sw128_ex(a_addr, 2048u, p.halo_baseoff ? (row_off & 7u) : 0u);
                            const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_b + static_cast<size_t>(tap * p.kchunks + cc) * B_BYTES));
                            if (dbg_ & 4) continue;
#pragma unroll
                            for (int k = 0; k < TC_BLOCK_K / 16; ++k)
                                ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, IDESC, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                } else {
                for (int kb = 0; kb < kb_total; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_a + static_cast<size_t>(stage) * TC_A_BYTES));
                    const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_b + static_cast<size_t>(B_RESIDENT ? kb : stage) * B_BYTES));
                    if (!(dbg_ & 4)) {
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / 16; ++k)
                        ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                }
                ptx::umma_commit(&acc_full[acc]);
                if (B_RESIDENT && tile + 1 < t_end && (tile + 1) / m_tiles != n_tile) ptx::umma_commit(b_free);
            }
        }
        __syncwarp();
    } else {
        // =================== epilogue (EPI groups of 4 warps, 128 TMEM lanes each) ===================
        const int q = warp & 3;                                 // TMEM lane quarter this warp may read
        const int grp = (warp - 2) >> 2;                        // epilogue group: tiles j = grp, grp + EPI, ...
        const int tig = static_cast<int>(threadIdx.x) - 64 - 128 * grp;      // thread in group, 0..127
        const int bar_id = 1 + grp;
        float* s_scale = s_epi + grp * 4 * BLOCK_N;
        float* s_shift = s_scale + BLOCK_N;
        float* s_stats = s_shift + BLOCK_N;                     // [2*BLOCK_N]
        float* s_tr = s_tr_all ? s_tr_all + (grp * 4 + q) * 32 * TR_PITCH : nullptr;
        const int row = q * 32 + lane;
        const int hl = row / p.tile_w, wl = row - hl * p.tile_w;
        int cur_n = -1;
        for (int tile = t_begin + grp; tile < t_end; tile += EPI) {
            const int n_tile = tile / m_tiles, m_tile = tile - n_tile * m_tiles;
            const int img = m_tile / tiles_per_img;
            const int trem = m_tile - img * tiles_per_img;
            const int th = trem / p.tiles_w;
            const int oh = th * p.tile_h + hl, ow = (trem - th * p.tiles_w) * p.tile_w + wl;
            const int n0 = n_tile * BLOCK_N;
            const bool valid = (oh < p.oh) && (ow < p.ow);
            if (n_tile != cur_n) {
                // new N tile: flush the statistics of the old one, load this one's scale / shift
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (STATS && p.stats && !(FULL && p.det) && cur_n >= 0) {
                    for (int i = tig; i < BLOCK_N; i += 128) {
                        const int co = cur_n * BLOCK_N + i;
                        if (co < p.cout) {
                            atomicAdd(&p.stats[co], s_stats[i]);
                            atomicAdd(&p.stats[p.cout + co], s_stats[BLOCK_N + i]);
                        }
                        s_stats[i] = 0.0f; s_stats[BLOCK_N + i] = 0.0f;
                    }
                }
                for (int i = tig; i < BLOCK_N; i += 128) {
                    const int co = n0 + i;
                    s_scale[i] = (p.scale && co < p.cout) ? p.scale[co] : 1.0f;
                    s_shift[i] = (p.shift && co < p.cout) ? p.shift[co] : 0.0f;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                cur_n = n_tile;
            }
            const int j = tile - t_begin, acc = j % NACC;
            ptx::mbar_wait(&acc_full[acc], static_cast<uint32_t>(j / NACC) & 1u);
            ptx::tc_fence_after();
            const long long out_off = img * p.out_sn + oh * p.out_sh + ow * p.out_sw;
            const long long res_off = img * p.res_sn + oh * p.res_sh + ow * p.res_sw;
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * ACC_COLS;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t r[32];
                if (dbg_ & 2) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) r[jj] = static_cast<uint32_t>(jj + c0);
                } else {
                    ptx::tmem_ld_32x32(t_addr + c0, r);
                    ptx::tmem_ld_wait();
                }
                if (c0 + 32 >= BLOCK_N) {                 // last read of this accumulator: hand it back to the MMA warp
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&acc_empty[acc]);
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                tc_epilogue_chunk<BLOCK_N, MODE>(p, v, c0, n0, valid, out_off, res_off, s_scale, s_shift, s_stats, lane, nullptr, s_tr);
            }
        }
        if (STATS && p.stats && !(FULL && p.det) && cur_n >= 0) {
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            for (int i = tig; i < BLOCK_N; i += 128) {
                const int co = cur_n * BLOCK_N + i;
                if (co < p.cout) {
                    atomicAdd(&p.stats[co], s_stats[i]);
                    atomicAdd(&p.stats[p.cout + co], s_stats[BLOCK_N + i]);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, TMEM_COLS); }
}

// Sum split-K slices in slice order and apply the epilogue.
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ partial, int split, long long m_total, int cout,
                     int cout_pad, int oh, int ow, TcParams p) {
    extern __shared__ float s_acc[];    // [2*cout] when stats
    if (p.stats) {
        for (int i = threadIdx.x; i < 2 * cout; i += blockDim.x) s_acc[i] = 0.0f;
        __syncthreads();
    }
    const int groups = cout_pad / 4;
    const long long total = m_total * groups;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pix = i / groups;
        const int co = static_cast<int>(i - pix * groups) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < split; ++s) {
            float4 t = __ldg(reinterpret_cast<const float4*>(partial + (static_cast<long long>(s) * m_total + pix) * cout_pad + co));
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        float v[4] = {a.x, a.y, a.z, a.w};
        const int img = static_cast<int>(pix / (static_cast<long long>(oh) * ow));
        const long long rem = pix - static_cast<long long>(img) * oh * ow;
        const int y = static_cast<int>(rem / ow);
        const int x = static_cast<int>(rem - static_cast<long long>(y) * ow);
        const long long out_off = img * p.out_sn + y * p.out_sh + x * p.out_sw;
        const long long res_off = img * p.res_sn + y * p.res_sh + x * p.res_sw;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = co + j;
            if (c >= cout) break;
            float raw = v[j];
            if (p.det) {
                det_add(p.det + 2 * c, raw);
                det_add(p.det + 2 * (cout + c), raw * raw);
            } else if (p.stats) {
                atomicAdd(&s_acc[c], raw);
                atomicAdd(&s_acc[cout + c], raw * raw);
            }
            float o = raw * (p.scale ? p.scale[c] : 1.0f) + (p.shift ? p.shift[c] : 0.0f);
            if (p.out_dtype != RTSDS_F32) {
                const bool f16 = p.out_dtype == RTSDS_F16;
                if (p.residual) o += ld_16(p.residual, res_off + c, f16);
                o = apply_act(o, p.act, p.slope);
                st_16(p.y, out_off + c, o, f16);
            } else {
                if (p.residual) o += reinterpret_cast<const float*>(p.residual)[res_off + c];
                o = apply_act(o, p.act, p.slope);
                reinterpret_cast<float*>(p.y)[out_off + c] = o;
            }
            (void)img;
        }
    }
    if (p.stats && !p.det) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * cout; i += blockDim.x)
            if (s_acc[i] != 0.0f) atomicAdd(&p.stats[i], s_acc[i]);
    }
}

// ---- host side -----------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 4-D bf16 activation view [C, W, H, N] with element strides (1, sw, sh, sn); box [64, bw, bh, 1].
static int make_act_map(CUtensorMap* m, const void* base, int c, int wd, int hd, int n, long long sw,
                        long long sh, long long sn, int bw, int bh) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable"); return RTSDS_ECUDA; }
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(wd), static_cast<cuuint64_t>(hd),
                          static_cast<cuuint64_t>(n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(sw * 2), static_cast<cuuint64_t>(sh * 2),
                             static_cast<cuuint64_t>(sn * 2)};
    cuuint32_t box[4] = {TC_BLOCK_K, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv_tc: cuTensorMapEncodeTiled(activation) failed: %d (dims %d,%d,%d,%d strides %lld,%lld,%lld box %d,%d)",
                  static_cast<int>(r), c, wd, hd, n, sw, sh, sn, bw, bh);
        return RTSDS_ECUDA;
    }
    return RTSDS_OK;
}

static int make_weight_map(CUtensorMap* m, const void* base, long long ktot, int rows, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable"); return RTSDS_ECUDA; }
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot * 2)};
    cuuint32_t box[2] = {TC_BLOCK_K, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv_tc: cuTensorMapEncodeTiled(weights) failed: %d (k %lld rows %d box %d)", static_cast<int>(r),
                  ktot, rows, box_rows);
        return RTSDS_ECUDA;
    }
    return RTSDS_OK;
}

int conv_cout_pad(int cout) {
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    return static_cast<int>(cdiv(cout, 128) * 128);
}

static void pick_tile(int oh, int ow, int* tw, int* th) {
    // tile_w * tile_h == 128; minimise padded area, prefer wide tiles (longer contiguous rows).
    const int cand_w[5] = {128, 64, 32, 16, 8};
    long long best = -1;
    for (int i = 0; i < 5; ++i) {
        int w_ = cand_w[i], h_ = 128 / w_;
        long long area = cdiv(ow, w_) * w_ * cdiv(oh, h_) * h_;
        if (best < 0 || area < best) { best = area; *tw = w_; *th = h_; }
    }
}

static size_t tc_smem_bytes(int block_n, int stages, size_t rbuf = 0) {
    return 1024 + static_cast<size_t>(stages) * (TC_A_BYTES + block_n * TC_BLOCK_K * 2) + rbuf + 6 * block_n * 4 +
           (2 * stages + 1) * 8 + 16;
}

template <int BLOCK_N>
static int launch_tc(const TcMaps& maps, const TcParams& p, dim3 grid, cudaStream_t st) {
    size_t smem = tc_smem_bytes(BLOCK_N, p.stages, p.rbuf_bytes);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
        attr_done = true;
    }
    static int use_pdl = -1;
    if (use_pdl < 0) { const char* e = getenv("RTSDS_NO_PDL"); use_pdl = (e && e[0] == '1') ? 0 : 1; }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (use_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (p.cluster_reduce || p.mc > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = static_cast<unsigned>(p.mc > 1 ? p.mc : 1);
        attr[na].val.clusterDim.z = static_cast<unsigned>(p.cluster_reduce ? p.split_k : 1);
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const bool full = p.trace || p.mc > 1 || p.dbg || p.det;
    cudaError_t le = full ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, 2>, maps, p)
                     : p.stats ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, 1>, maps, p)
                               : cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, 0>, maps, p);
    if (le != cudaSuccess) { set_error("conv_tc_kernel: launch: %s", cudaGetErrorString(le)); return RTSDS_ECUDA; }
    count_launch();
    return check_launch("conv_tc_kernel");
}

constexpr int TCP_EPI_MAX = 2;                                // epilogue groups of the persistent kernel (sizes its shared memory)
static size_t tcp_smem_bytes(int block_n, int stages, int b_slots, int a_stage = TC_A_BYTES, size_t extra = 0) {
    return 1024 + static_cast<size_t>(stages) * a_stage + static_cast<size_t>(b_slots) * block_n * TC_BLOCK_K * 2 +
           TCP_EPI_MAX * 4 * block_n * 4 + (2 * stages + 4 * TCP_EPI_MAX + 2) * 8 + 16 + extra;
}

template <int BLOCK_N, bool B_RESIDENT, bool HALO, int EPI, int MODE>
static int launch_tcp_epi(const TcMaps& maps, const TcParams& p, int grid, size_t smem, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcp_kernel<BLOCK_N, B_RESIDENT, HALO, EPI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("conv_tcp: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
        attr_done = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(64 + 128 * EPI);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tcp_kernel<BLOCK_N, B_RESIDENT, HALO, EPI, MODE>, maps, p);
    if (le != cudaSuccess) { set_error("conv_tcp_kernel: launch: %s", cudaGetErrorString(le)); return RTSDS_ECUDA; }
    count_launch();
    return check_launch("conv_tcp_kernel");
}

// RTSDS_TCP_EPI=1: one epilogue group (the round-1 form) instead of two — A/B switch
template <int BLOCK_N, bool B_RESIDENT, bool HALO = false>
static int launch_tcp(const TcMaps& maps, const TcParams& p, int grid, size_t smem, cudaStream_t st) {
    static int epi = -1;
    if (epi < 0) { const char* e = getenv("RTSDS_TCP_EPI"); epi = (e && e[0] == '1') ? 1 : 2; }
    if (epi == 1) return launch_tcp_epi<BLOCK_N, B_RESIDENT, HALO, 1, 2>(maps, p, grid, smem, st);
    if (p.dbg || p.det) return launch_tcp_epi<BLOCK_N, B_RESIDENT, HALO, 2, 2>(maps, p, grid, smem, st);
    if (p.stats) return launch_tcp_epi<BLOCK_N, B_RESIDENT, HALO, 2, 1>(maps, p, grid, smem, st);
    return launch_tcp_epi<BLOCK_N, B_RESIDENT, HALO, 2, 0>(maps, p, grid, smem, st);
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_conv_cout_pad(int cout) { return conv_cout_pad(cout); }

// Tuning knobs (process-wide, set from Python for experiments): block_n in {0=auto,32,64,128},
// stages 0=auto.
static int g_force_block_n = 0, g_force_stages = 0;
static unsigned long long* g_tc_trace = nullptr;
// Debug: subsequent non-persistent conv_tc launches write 16 stamps per CTA into `buf` (device memory, >= 16*8*CTAs bytes);
// NULL switches it off.  tools/conv_timeline.py reads them.
extern "C" void rtsds_debug_conv_trace(void* buf) { g_tc_trace = reinterpret_cast<unsigned long long*>(buf); }
extern "C" void rtsds_conv2d_tc_tune(int block_n, int stages) { g_force_block_n = block_n; g_force_stages = stages; }

static int tc_auto_split(long long ctas, int kb_total) {
    int split = 1;
    const int sms = num_sms();
    while (ctas * split * 2 <= sms && kb_total / (split * 2) >= 4 && split < 16) split *= 2;
    return split;
}

static int tc_pick_block_n(int cout_pad, long long m_tiles, int kb_total) {
    if (g_force_block_n && cout_pad % g_force_block_n == 0) return g_force_block_n;
    if (cout_pad <= 64) return cout_pad;
    const long long wide = m_tiles * (cout_pad / 128);
    if (wide >= num_sms()) return 128;
    // Few M tiles.  The K loop of these launches runs at the L2->SM fabric rate, and a 128-wide tile moves a third fewer
    // operand bytes per flop than a 64-wide one; the missing CTAs come back as split-K slices, which reduce inside a
    // cluster (<= 4 deep) — worth it while every slice keeps >= 8 k-blocks and the grid still covers most of the SMs
    // (tools/deep_conv_sweep.py: 256->256 32x64 13.3 -> 12.0 us, 128->128 64x128 12.3 -> 11.6 us; 128->256 s2 and the
    // 512-channel layers stay on 64-wide tiles).
    const int sp = tc_auto_split(wide, kb_total);
    if (sp >= 2 && sp <= 4 && kb_total / sp >= 8 && wide * sp * 5 >= 4LL * num_sms()) return 128;
    return 64;      // narrower N tiles give more CTAs
}

static void tp_plan(const TapProblem& t, int* block_n, int* split, int* tile_w, int* tile_h) {
    pick_tile(t.oh, t.ow, tile_w, tile_h);
    const long long m_tiles = static_cast<long long>(t.n_img) * cdiv(t.ow, *tile_w) * cdiv(t.oh, *tile_h);
    const int cp = conv_cout_pad(t.cout);
    const int kb_total = t.n_taps * (t.ck / TC_BLOCK_K);
    *block_n = tc_pick_block_n(cp, m_tiles, kb_total);
    int sp = t.split_req;
    if (sp <= 0) sp = tc_auto_split(m_tiles * (cp / *block_n), kb_total);
    if (sp > kb_total) sp = kb_total;
    if (sp > 16) sp = 16;
    if (sp < 1) sp = 1;
    if (t.gap_out) sp = sp >= 4 ? 4 : (sp >= 2 ? 2 : 1);      // the fused pool rides on the in-cluster split-K reduction
    *split = sp;
}

static size_t tp_workspace(const TapProblem& t) {
    int bn, sp, tw, th;
    tp_plan(t, &bn, &sp, &tw, &th);
    if (sp <= 1) return 0;
    return static_cast<size_t>(sp) * t.n_img * t.oh * t.ow * conv_cout_pad(t.cout) * sizeof(float);
}

static int tp_run_inner(const TapProblem& t, void* workspace, size_t ws_bytes, cudaStream_t stream, unsigned long long* det);

// Deterministic mode: the BatchNorm sums of this launch go into exact accumulators, rounded into t.stats afterwards.
static int tp_run(const TapProblem& t, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    if (!t.stats || !det_mode()) return tp_run_inner(t, workspace, ws_bytes, stream, nullptr);
    unsigned long long* det = det_scratch(stream, 2 * static_cast<size_t>(t.cout));
    if (!det) return RTSDS_ECUDA;
    int rc = tp_run_inner(t, workspace, ws_bytes, stream, det);
    if (rc != RTSDS_OK) return rc;
    return det_finish(det, t.stats, 2 * static_cast<size_t>(t.cout), true, stream);
}

static int tp_run_inner(const TapProblem& t, void* workspace, size_t ws_bytes, cudaStream_t stream, unsigned long long* det) {
    TcMaps maps;
    TcParams p;
    memset(&maps, 0, sizeof(maps));
    memset(&p, 0, sizeof(p));
    int block_n, split;
    tp_plan(t, &block_n, &split, &p.tile_w, &p.tile_h);
    // ---- halo mode: 3x3 stride-1 tap grids (forward or dgrad), weights resident, >= 2 waves of 8x16 tiles ----
    static int halo_mode = -1, halo_baseoff = 0;     // measured: the UMMA unit swizzles on absolute smem address bits, base offset stays 0
    if (halo_mode < 0) {
        const char* e = getenv("RTSDS_NO_HALO");
        halo_mode = (e && e[0] == '1') ? 0 : 1;
        const char* b = getenv("RTSDS_HALO_BASEOFF");
        if (b) halo_baseoff = atoi(b);
    }
    bool halo = false;
    int halo_d = 0;
    if (halo_mode && split == 1 && t.n_taps == 9 && t.view[0].used && !t.gap_out) {
        for (int i = 0; i < 9; ++i) halo_d = max(halo_d, max(abs(t.dh[i]), abs(t.dw[i])));
        halo = halo_d == 1 || halo_d == 2 || halo_d == 4;
        bool seen[9] = {false};
        for (int i = 0; i < 9 && halo; ++i) {
            if (t.map[i] != 0 || t.dh[i] % halo_d || t.dw[i] % halo_d) { halo = false; break; }
            const int r = t.dh[i] / halo_d + 1, q = t.dw[i] / halo_d + 1;
            if (r < 0 || r > 2 || q < 0 || q > 2 || seen[r * 3 + q]) { halo = false; break; }
            seen[r * 3 + q] = true;
            if (t.kb[i] != i * (t.ck / TC_BLOCK_K)) halo = false;
        }
        if (halo) {
            const int cp = conv_cout_pad(t.cout);
            const int a_stage = (16 + 2 * halo_d) * 16 * 128;
            const int kbt = 9 * (t.ck / TC_BLOCK_K);
            // the whole Cout must be ONE resident N tile (otherwise every N tile re-loads the halo, and streaming the
            // weights costs as many TMA rows as the taps did): in practice the 64-input-channel 3x3 layers
            int bn = 0;
            if ((cp == 64 || cp == 128) && tcp_smem_bytes(cp, 2, kbt, a_stage) <= 227 * 1024) bn = cp;
            const long long mt = static_cast<long long>(t.n_img) * cdiv(t.ow, 8) * cdiv(t.oh, 16);
            static long long halo_min_tiles = -1;
            if (halo_min_tiles < 0) { const char* e = getenv("RTSDS_HALO_MIN_TILES"); halo_min_tiles = e ? atoll(e) : 2LL * num_sms(); }
            if (bn == 0 || mt * (cp / bn) < halo_min_tiles) halo = false;
            else { block_n = bn; p.tile_w = 8; p.tile_h = 16; p.halo_d = halo_d; p.halo_rows = 16 + 2 * halo_d; p.a_stage_bytes = a_stage; p.halo_baseoff = halo_baseoff; }
        }
    }
    p.n_img = t.n_img; p.oh = t.oh; p.ow = t.ow;
    p.tiles_w = static_cast<int>(cdiv(t.ow, p.tile_w));
    p.tiles_h = static_cast<int>(cdiv(t.oh, p.tile_h));
    p.cout = t.cout; p.cout_pad = conv_cout_pad(t.cout);
    p.n_taps = t.n_taps; p.kchunks = t.ck / TC_BLOCK_K;
    // ---- A-tile multicast: the N tiles of one M tile read the same activations.  With an even number of N tiles and a
    // grid small enough that the non-persistent kernel runs it, pairs of N tiles form a cluster along y; each CTA fetches
    // half of every A tile and TMA multicasts it to both (L2 -> SM activation traffic halves).
    static int mc_mode = -1;
    if (mc_mode < 0) { const char* e = getenv("RTSDS_MC"); mc_mode = e ? atoi(e) : 0; }
    p.mc = 1;
    {
        const int n_tiles_ = p.cout_pad / block_n;
        const long long tot = static_cast<long long>(t.n_img) * p.tiles_w * p.tiles_h * n_tiles_;
        const int zdim = (split == 2 || split == 4 || split == 8) ? split : 1;       // upper bound of the split-K cluster depth
        if (mc_mode && !halo && t.n_taps > 0 && n_tiles_ % 2 == 0 && 2 * zdim <= 8 && tot * split < 2LL * num_sms()) {
            p.mc = 2;
            if (p.tile_h >= 2) { p.mc_dh = p.tile_h / 2; p.mc_dw = 0; }
            else { p.mc_dh = 0; p.mc_dw = p.tile_w / 2; }
        }
    }
    const int box_w = p.mc > 1 && p.tile_h < 2 ? p.tile_w / 2 : p.tile_w;
    const int box_h = p.mc > 1 && p.tile_h >= 2 ? p.tile_h / 2 : p.tile_h;
    p.out_sn = t.out_sn; p.out_sh = t.out_sh; p.out_sw = t.out_sw;
    p.res_sn = t.res_sn; p.res_sh = t.res_sh; p.res_sw = t.res_sw;
    p.scale = t.scale; p.shift = t.shift; p.residual = t.residual; p.stats = t.stats; p.det = det; p.y = t.y;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("RTSDS_TC_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
    p.out_dtype = t.out_dtype; p.act = t.act; p.slope = t.slope; p.f16 = t.in_f16;
    p.gap_out = t.gap_out; p.gap_scale = t.gap_out ? 1.0f / (static_cast<float>(t.oh) * static_cast<float>(t.ow)) : 0.f;
    if (t.gap_out && t.stats) { set_error("conv_tc: gap_out and stats are mutually exclusive"); return RTSDS_EINVAL; }
    int first_used = -1;
    for (int i = 0; i < 4; ++i) {
        if (!t.view[i].used) continue;
        int rc = make_act_map(&maps.a[i], t.view[i].base, t.c_extent, t.view[i].wd, t.view[i].hd, t.n_img, t.view[i].sw,
                              t.view[i].sh, t.view[i].sn, box_w, box_h);
        if (rc != RTSDS_OK) return rc;
        if (first_used < 0) first_used = i;
    }
    if (first_used < 0 && t.n_taps > 0) { set_error("conv_tc: no activation view"); return RTSDS_EINVAL; }
    for (int i = 0; i < 4; ++i)
        if (!t.view[i].used && first_used >= 0) maps.a[i] = maps.a[first_used];
    for (int i = 0; i < t.n_taps; ++i) { p.tap_dh[i] = t.dh[i]; p.tap_dw[i] = t.dw[i]; p.tap_map[i] = t.map[i]; p.tap_kb[i] = static_cast<short>(t.kb[i]); }
    if (halo) {      // one box of 16 x (16+2d) pixels per 64-channel chunk; out-of-image pixels are zero-filled = the padding
        int rc = make_act_map(&maps.a[1], t.view[0].base, t.c_extent, t.view[0].wd, t.view[0].hd, t.n_img, t.view[0].sw, t.view[0].sh,
                              t.view[0].sn, 16, p.halo_rows);
        if (rc != RTSDS_OK) return rc;
    }
    if (t.n_taps > 0) {
        int rc = make_weight_map(&maps.b, t.w, t.w_ktot, p.cout_pad, block_n);
        if (rc != RTSDS_OK) return rc;
    }
    const long long m_tiles = static_cast<long long>(t.n_img) * p.tiles_w * p.tiles_h;
    const int n_tiles = p.cout_pad / block_n;
    const int kb_total = p.n_taps * p.kchunks;
    p.split_k = split;
    const long long m_total = static_cast<long long>(t.n_img) * t.oh * t.ow;
    if (split > 1) {
        const size_t need = static_cast<size_t>(split) * m_total * p.cout_pad * sizeof(float);
        if (!workspace || ws_bytes < need) {
            set_error("conv_tc: split_k=%d needs %zu workspace bytes, got %zu", split, need, ws_bytes);
            return RTSDS_EWS;
        }
        RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "conv_tc: workspace alignment");
        p.partial = reinterpret_cast<float*>(workspace);
    }
    // stages: keep two CTAs resident per SM (epilogue of one overlaps the main loop of the other)
    int stages = g_force_stages ? g_force_stages : (block_n == 128 ? 3 : (block_n == 64 ? 4 : 5));
    const int kb_per = kb_total / split;
    // a grid that fits in one wave leaves one CTA per SM: use all of its shared memory as pipeline depth (the K loop of
    // these small problems is TMA-latency bound), instead of keeping room for a second resident CTA
    static int wave_stages = -1, b_early = -1;
    if (wave_stages < 0) { const char* e = getenv("RTSDS_TC_WAVE_STAGES"); wave_stages = e ? atoi(e) : 12; }
    if (b_early < 0) { const char* e = getenv("RTSDS_TC_B_EARLY"); b_early = e ? atoi(e) : 1; }
    p.b_early = b_early;
    p.trace = g_tc_trace;
    if (!g_force_stages && m_tiles * n_tiles * split <= num_sms()) stages = wave_stages;
    if (stages > kb_per) stages = kb_per < 2 ? 2 : kb_per;
    while (tc_smem_bytes(block_n, stages) > 227 * 1024) --stages;
    // split-K of 2 or 4: the slices of a tile run as one thread-block cluster and reduce through distributed shared memory
    static int cluster_mode = -1;
    if (cluster_mode < 0) { const char* e = getenv("RTSDS_NO_CLUSTER_SPLITK"); cluster_mode = (e && e[0] == '1') ? 0 : 1; }
    // (split 8: 16 rows per rank — a warp of the reduction then spans two 32-column chunks, which the per-warp BatchNorm /
    // pool sums do not allow; plain epilogues only)
    if (cluster_mode && (split == 2 || split == 4 || (split == 8 && !t.stats && !t.gap_out)) && block_n >= 32) {
        // receive buffer for the pushed partial rows (NOT an overlay of the ring: peers write it while this CTA still streams)
        const size_t need = static_cast<size_t>(TC_BLOCK_M) * (block_n + 4) * sizeof(float);
        while (stages > 2 && tc_smem_bytes(block_n, stages, need) > 227 * 1024) --stages;
        if (tc_smem_bytes(block_n, stages, need) <= 227 * 1024) { p.cluster_reduce = 1; p.rbuf_bytes = static_cast<int>(need); }
    }
    if (t.gap_out && split > 1 && !p.cluster_reduce) {
        set_error("conv_tc: the fused global pool needs split_k 1, 2 or 4 (thread-block-cluster reduction); got %d", split);
        return RTSDS_EUNSUP;
    }
    p.stages = stages;
    p.res_pre = t.residual && t.out_dtype != RTSDS_F32 && t.cout % block_n == 0;
    RTSDS_REQUIRE(m_tiles <= 0x7fffffffLL, "conv_tc: too many tiles");

    // many tiles, no split-K: persistent CTAs (resident weights when they fit, double-buffered accumulators)
    static int persist_mode = -1;
    if (persist_mode < 0) { const char* e = getenv("RTSDS_NO_PERSISTENT"); persist_mode = (e && e[0] == '1') ? 0 : 1; }
    const long long total_tiles = m_tiles * n_tiles;
    const size_t tr_extra = t.stats ? TCP_EPI_MAX * TR_BYTES + 16 : 0;      // per-warp transpose scratch of the BatchNorm sums
    // (only with resident weights: when they have to stream, two co-resident non-persistent CTAs per SM hide more latency)
    if (halo) {
        int st = 6;
        while (st > 2 && tcp_smem_bytes(block_n, st, kb_total, p.a_stage_bytes, tr_extra) > 227 * 1024) --st;
        if (st > p.kchunks * 4) st = max(2, p.kchunks * 4);
        p.stages = st;
        p.n_tiles_n = n_tiles;
        p.total_tiles = static_cast<int>(total_tiles);
        int ctas = num_sms();
        p.tiles_per_cta = static_cast<int>(cdiv(total_tiles, ctas));
        ctas = static_cast<int>(cdiv(total_tiles, p.tiles_per_cta));
        const size_t smem = tcp_smem_bytes(block_n, st, kb_total, p.a_stage_bytes, tr_extra);
        if (block_n == 128) return launch_tcp<128, true, true>(maps, p, ctas, smem, stream);
        if (block_n == 64) return launch_tcp<64, true, true>(maps, p, ctas, smem, stream);
        return launch_tcp<32, true, true>(maps, p, ctas, smem, stream);
    }
    const bool resident = kb_total >= 1 && tcp_smem_bytes(block_n, 3, kb_total, TC_A_BYTES, tr_extra) <= 227 * 1024;
    // Weights that do not fit stream through the ring with the A tiles.  Measured (tools/conv_micro.py, two epilogue groups):
    // faster than two co-resident one-tile CTAs up to K = 1152 (64->128 s2 with statistics 187 -> 124 us, 128->256 s2
    // 120 -> 105 us, 128->128 65 -> 62 us), slower beyond (256->256 54 -> 66 us, 512->512 d4 124 -> 146 us): the longer
    // the K loop, the more the second resident CTA's main loop is worth.  RTSDS_PERSIST_STREAM=<max k-blocks> overrides.
    static int persist_stream = -1;
    if (persist_stream < 0) { const char* e = getenv("RTSDS_PERSIST_STREAM"); persist_stream = e ? atoi(e) : 18; }
    if (persist_mode && split == 1 && (resident || kb_total <= persist_stream) && total_tiles >= 2LL * num_sms() && total_tiles < (1LL << 30) && !t.gap_out) {
        const size_t b_all = static_cast<size_t>(kb_total) * block_n * TC_BLOCK_K * 2;
        int st = resident ? 8 : (block_n == 128 ? 6 : 8);
        while (st > 2 && tcp_smem_bytes(block_n, st, resident ? kb_total : st, TC_A_BYTES, tr_extra) > 227 * 1024) --st;
        (void)b_all;
        p.stages = st;
        p.n_tiles_n = n_tiles;
        p.total_tiles = static_cast<int>(total_tiles);
        int ctas = num_sms();
        p.tiles_per_cta = static_cast<int>(cdiv(total_tiles, ctas));
        ctas = static_cast<int>(cdiv(total_tiles, p.tiles_per_cta));
        const size_t smem = tcp_smem_bytes(block_n, st, resident ? kb_total : st, TC_A_BYTES, tr_extra);
        if (block_n == 128) return resident ? launch_tcp<128, true>(maps, p, ctas, smem, stream) : launch_tcp<128, false>(maps, p, ctas, smem, stream);
        if (block_n == 64) return resident ? launch_tcp<64, true>(maps, p, ctas, smem, stream) : launch_tcp<64, false>(maps, p, ctas, smem, stream);
        if (block_n == 32) return resident ? launch_tcp<32, true>(maps, p, ctas, smem, stream) : launch_tcp<32, false>(maps, p, ctas, smem, stream);
    }
    dim3 grid(static_cast<unsigned>(m_tiles), static_cast<unsigned>(n_tiles), static_cast<unsigned>(split));
    int rc;
    if (block_n == 128) rc = launch_tc<128>(maps, p, grid, stream);
    else if (block_n == 64) rc = launch_tc<64>(maps, p, grid, stream);
    else if (block_n == 32) rc = launch_tc<32>(maps, p, grid, stream);
    else { set_error("conv_tc: block_n %d", block_n); return RTSDS_EUNSUP; }
    if (rc != RTSDS_OK) return rc;
    if (split > 1 && !p.cluster_reduce) {
        const long long total = m_total * (p.cout_pad / 4);
        long long want = cdiv(total, 256);
        int g = static_cast<int>(want > 4LL * num_sms() ? 4LL * num_sms() : want);
        size_t sm = t.stats ? 2 * static_cast<size_t>(t.cout) * sizeof(float) : 0;
        splitk_finish_kernel<<<g, 256, sm, stream>>>(p.partial, split, m_total, t.cout, p.cout_pad, t.oh, t.ow, p);
        count_launch();
        rc = check_launch("splitk_finish_kernel");
    }
    return rc;
}

extern "C" size_t rtsds_conv2d_tc_workspace_bytes(const RtsdsConvDesc* d) {
    if (!d) return 0;
    TapProblem t;
    if (fwd_problem(d, nullptr, nullptr, TC_BLOCK_K, 2, &t) != RTSDS_OK) return 0;
    return tp_workspace(t);
}

extern "C" int rtsds_conv2d_tc_fwd(const RtsdsConvDesc* d, const void* x, const void* w, const float* scale,
                                   const float* shift, const void* residual, float* stats, void* y,
                                   void* workspace, size_t ws_bytes, rtsds_stream_t s) {
    RTSDS_REQUIRE(d && x && w && y, "conv2d_tc_fwd: NULL argument");
    RTSDS_REQUIRE(is_16bit(d->in_dtype), "conv2d_tc_fwd: input must be bf16 or fp16");
    RTSDS_REQUIRE(d->out_dtype == d->in_dtype || d->out_dtype == RTSDS_F32, "conv2d_tc_fwd: out_dtype must be the input type or fp32");
    RTSDS_REQUIRE(d->out_ld >= d->cout, "conv2d_tc_fwd: out_ld < cout");
    RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv2d_tc_fwd: pointers must be 16-byte aligned");
    const int vec = d->out_dtype != RTSDS_F32 ? 8 : 4;
    RTSDS_REQUIRE(d->out_ld % vec == 0, "conv2d_tc_fwd: out_ld=%d must be a multiple of %d", d->out_ld, vec);
    if (residual) {
        RTSDS_REQUIRE(d->res_ld >= d->cout && d->res_ld % vec == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                      "conv2d_tc_fwd: residual pitch/alignment");
    }
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    TapProblem t;
    rc = fwd_problem(d, x, w, TC_BLOCK_K, 2, &t);
    if (rc != RTSDS_OK) return rc;
    t.scale = scale; t.shift = shift; t.residual = residual; t.stats = stats; t.y = y;
    return tp_run(t, workspace, ws_bytes, as_stream(s));
}

// Same, with the global average pool of the layer's OUTPUT fused into the epilogue — AdaptiveAvgPool2d(1) of
// build_bisenet.py:46 and the context-path tail of build_contextpath.py:27-28 without a separate pass over feature3 /
// feature4.  DETERMINISTIC: every CTA writes the (pre-scaled) partial sum of its own rows to
// gap_out[image][part][cout], parts = rtsds_conv2d_tc_gap_parts(d); the consumer adds the parts in index order
// (mean = sum over parts).  No atomics, nothing to zero.
extern "C" int rtsds_conv2d_tc_gap_parts(const RtsdsConvDesc* d) {
    if (!d) return 0;
    TapProblem t;
    if (fwd_problem(d, nullptr, nullptr, TC_BLOCK_K, 2, &t) != RTSDS_OK) return 0;
    float dummy;
    t.gap_out = &dummy;
    int bn, sp, tw, th;
    tp_plan(t, &bn, &sp, &tw, &th);
    return static_cast<int>(cdiv(t.ow, tw) * cdiv(t.oh, th)) * sp;
}

extern "C" int rtsds_conv2d_tc_fwd_gap(const RtsdsConvDesc* d, const void* x, const void* w, const float* scale,
                                       const float* shift, const void* residual, void* y, float* gap_out,
                                       void* workspace, size_t ws_bytes, rtsds_stream_t s) {
    RTSDS_REQUIRE(d && x && w && y && gap_out, "conv2d_tc_fwd_gap: NULL argument");
    RTSDS_REQUIRE(is_16bit(d->in_dtype), "conv2d_tc_fwd_gap: input must be bf16 or fp16");
    RTSDS_REQUIRE(d->out_dtype == d->in_dtype || d->out_dtype == RTSDS_F32, "conv2d_tc_fwd_gap: out_dtype must be the input type or fp32");
    RTSDS_REQUIRE(d->out_ld >= d->cout && d->out_ld % (d->out_dtype != RTSDS_F32 ? 8 : 4) == 0, "conv2d_tc_fwd_gap: out_ld");
    RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv2d_tc_fwd_gap: pointers must be 16-byte aligned");
    if (residual)
        RTSDS_REQUIRE(d->res_ld >= d->cout && (reinterpret_cast<uintptr_t>(residual) & 15) == 0, "conv2d_tc_fwd_gap: residual pitch/alignment");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    TapProblem t;
    rc = fwd_problem(d, x, w, TC_BLOCK_K, 2, &t);
    if (rc != RTSDS_OK) return rc;
    t.scale = scale; t.shift = shift; t.residual = residual; t.stats = nullptr; t.y = y; t.gap_out = gap_out;
    return tp_run(t, workspace, ws_bytes, as_stream(s));
}

extern "C" size_t rtsds_conv2d_tc_dgrad_workspace_bytes(const RtsdsConvDesc* d) {
    if (!d || (d->stride != 1 && d->stride != 2)) return 0;
    size_t m = 0;
    for (int ph = 0; ph < d->stride; ++ph)
        for (int pw = 0; pw < d->stride; ++pw) {
            TapProblem t;
            if (dgrad_problem(d, nullptr, nullptr, nullptr, nullptr, RTSDS_BF16, ph, pw, TC_BLOCK_K, 2, &t) != RTSDS_OK) return 0;
            if (t.oh <= 0 || t.ow <= 0 || t.n_taps == 0) continue;
            size_t b = tp_workspace(t);
            if (b > m) m = b;
        }
    return m;
}

extern "C" int rtsds_conv2d_tc_dgrad(const RtsdsConvDesc* d, const void* dy, const void* w_dgrad, const void* residual,
                                     void* dx, int dx_dtype, void* workspace, size_t ws_bytes, rtsds_stream_t s) {
    RTSDS_REQUIRE(d && dy && w_dgrad && dx, "conv2d_tc_dgrad: NULL argument");
    RTSDS_REQUIRE(d->stride == 1 || d->stride == 2, "conv2d_tc_dgrad: stride %d unsupported", d->stride);
    RTSDS_REQUIRE(d->kh * d->kw <= TAP_MAX, "conv2d_tc_dgrad: filter too large");
    const int ck = static_cast<int>(cdiv(d->cout, TC_BLOCK_K) * TC_BLOCK_K);
    RTSDS_REQUIRE(d->out_ld >= ck && d->out_ld % 8 == 0, "conv2d_tc_dgrad: dy pitch %d must be >= %d (cout rounded up to 64, zero padded)", d->out_ld, ck);
    RTSDS_REQUIRE(dx_dtype == RTSDS_BF16 || dx_dtype == RTSDS_F32, "conv2d_tc_dgrad: bad dx dtype");
    const int vec = dx_dtype == RTSDS_BF16 ? 8 : 4;
    RTSDS_REQUIRE(d->in_ld >= d->cin && d->in_ld % vec == 0, "conv2d_tc_dgrad: dx pitch");
    RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(w_dgrad) & 15) == 0, "conv2d_tc_dgrad: pointers must be 16-byte aligned");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    for (int ph = 0; ph < d->stride; ++ph)
        for (int pw = 0; pw < d->stride; ++pw) {
            TapProblem t;
            rc = dgrad_problem(d, dy, w_dgrad, residual, dx, dx_dtype, ph, pw, TC_BLOCK_K, 2, &t);
            if (rc != RTSDS_OK) return rc;
            if (t.oh <= 0 || t.ow <= 0) continue;
            rc = tp_run(t, workspace, ws_bytes, as_stream(s));
            if (rc != RTSDS_OK) return rc;
        }
    return RTSDS_OK;
}

// =====================================================================================
// wgrad on tensor cores: dW[co][tap][ci] = sum_pixels dy[pix][co] * x[pix shifted by tap][ci]
//
// GEMM view per (tap, 128-wide co tile, ci tile): D[M = co, N = ci] += A[M, K = 16 pixels] * B[N, K]^T with
// K running over the pixels of 128-pixel tiles.  Both operands are "MN-major": in NHWC the channel
// (M or N) index is the contiguous one.  A TMA box [64 ch, tile_w, tile_h, 1] lands as 128 pixel rows
// of 128 B with the 128-byte swizzle, which is exactly the canonical MN-major SWIZZLE_128B UMMA
// operand (8-pixel groups 1024 B apart = SBO; further 64-channel chunks one box apart = LBO).
// The pixel range is split over CTAs; partial results are accumulated into the fp32 dW with
// red.global.add (dW is a few MB at most).
// =====================================================================================
namespace rtsds {

constexpr int WG_BOX_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;    // [128 pixels][64 channels] bf16 = 16 KiB

struct WgMaps {
    CUtensorMap a;        // dy  [ck, ow, oh, n]
    CUtensorMap b[4];     // x parity views [cin, wd, hd, n]
};

struct WgParams {
    int n_img, oh, ow, tile_w, tile_h, tiles_w, tiles_h;
    int cout, cin, n_taps, ci_tiles;
    int tiles_total, tiles_per_split;
    int stages;
    signed char tap_dh[TAP_MAX], tap_dw[TAP_MAX], tap_map[TAP_MAX];
    float* dw;
    unsigned long long* det;      // deterministic mode: exact accumulators with dw's layout (common.cuh: det_add), else NULL
};

// Deterministic mode: this thread's accumulator row, one column at a time, into the exact accumulators (common.cuh: det_add)
// that mirror dw.  Out of line and column-wise so that the production epilogue's registers are untouched.
__device__ __noinline__ void wgrad_det_epilogue(const WgParams& p, uint32_t trow, int n_cols, int block_n, int co, int ci0, int tap0) {
    for (int cc = 0; cc < n_cols; ++cc) {
        const float v = __uint_as_float(ptx::tmem_ld_32x32_x1(trow + cc));       // warp-uniform: every lane loads its row
        ptx::tmem_ld_wait();
        const int tt = cc / block_n, ci = ci0 + cc - tt * block_n;
        if (co < p.cout && ci < p.cin)
            det_add(p.det + 2 * ((static_cast<long long>(co) * p.n_taps + tap0 + tt) * p.cin + ci), v);
    }
}

// TAPS filter taps per CTA share one dy tile (the A operand): each tap has its own x tile (B operand) and its own
// accumulator columns [tt*BLOCK_N, (tt+1)*BLOCK_N) in TMEM.  TAPS = 3 for the 64-input-channel 3x3 layers: the dy tile
// crosses L2 -> SM three times instead of nine.
template <int BLOCK_N, int TAPS>     // ci per CTA: 64, 128 or 256
__global__ void __launch_bounds__(TC_THREADS)
wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
    constexpr int A_BYTES = 2 * WG_BOX_BYTES;                  // 128 co = two 64-channel boxes
    constexpr int B1_BYTES = (BLOCK_N / 64) * WG_BOX_BYTES;    // one tap
    constexpr int B_BYTES = TAPS * B1_BYTES;
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(TC_BLOCK_M, BLOCK_N, 1, 1);
    constexpr uint32_t TMEM_COLS = TAPS * BLOCK_N <= 64 ? 64 : (TAPS * BLOCK_N <= 128 ? 128 : (TAPS * BLOCK_N <= 256 ? 256 : 512));

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + static_cast<size_t>(stages) * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + static_cast<size_t>(stages) * B_BYTES);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* tmem_full_bar = empty_bar + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap0 = (blockIdx.y / p.ci_tiles) * TAPS;
    const int ci0 = (blockIdx.y % p.ci_tiles) * BLOCK_N;
    const int co0 = blockIdx.z * TC_BLOCK_M;
    const int t_begin = blockIdx.x * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&maps.a);
        ptx::prefetch_tmap(&maps.b[p.tap_map[tap0]]);
        for (int i = 0; i < stages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        ptx::mbar_init(tmem_full_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            // rows 64..127 of the co tile do not exist when cout <= co0 + 64: their box is not loaded at all (the MMA
            // then reads stale shared memory for those rows, which only reaches accumulator rows nobody stores)
            const bool hi = co0 + 64 < p.cout;
            for (int t = t_begin; t < t_end; ++t) {
                const int img = t / tiles_per_img;
                const int r = t - img * tiles_per_img;
                const int th = r / p.tiles_w;
                const int h0 = th * p.tile_h, w0 = (r - th * p.tiles_w) * p.tile_w;
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_expect_tx(&full_bar[stage], (hi ? A_BYTES : WG_BOX_BYTES) + B_BYTES);
                uint8_t* sa = smem_a + static_cast<size_t>(stage) * A_BYTES;
                uint8_t* sb = smem_b + static_cast<size_t>(stage) * B_BYTES;
                ptx::tma_load_4d(sa, &maps.a, &full_bar[stage], co0, w0, h0, img);
                if (hi) ptx::tma_load_4d(sa + WG_BOX_BYTES, &maps.a, &full_bar[stage], co0 + 64, w0, h0, img);
#pragma unroll
                for (int tt = 0; tt < TAPS; ++tt) {
                    const int tap = tap0 + tt;
                    const CUtensorMap* bm = &maps.b[p.tap_map[tap]];
#pragma unroll
                    for (int j = 0; j < BLOCK_N / 64; ++j)
                        ptx::tma_load_4d(sb + tt * B1_BYTES + j * WG_BOX_BYTES, bm, &full_bar[stage], ci0 + j * 64,
                                         w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], img);
                }
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem_a + static_cast<size_t>(stage) * A_BYTES);
                const uint32_t sb = ptx::smem_u32(smem_b + static_cast<size_t>(stage) * B_BYTES);
#pragma unroll
                for (int tt = 0; tt < TAPS; ++tt) {
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_M / 16; ++k) {      // 16 pixels (rows) per MMA = 2048 B
                        const uint64_t da = ptx::umma_desc_mn_sw128(sa + k * 2048, WG_BOX_BYTES);
                        const uint64_t db = ptx::umma_desc_mn_sw128(sb + tt * B1_BYTES + k * 2048, WG_BOX_BYTES);
                        ptx::umma_bf16(tmem_base + tt * BLOCK_N, da, db, IDESC, (t > t_begin || k > 0) ? 1u : 0u);
                    }
                }
                ptx::umma_commit(&empty_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (t_end > t_begin) ptx::umma_commit(tmem_full_bar);
        }
        __syncwarp();
    } else if (t_end > t_begin) {
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        ptx::mbar_wait(tmem_full_bar, 0);
        ptx::tc_fence_after();
        if (p.det) wgrad_det_epilogue(p, tmem_base + (static_cast<uint32_t>(q * 32) << 16), TAPS * BLOCK_N, BLOCK_N, co, ci0, tap0);
#pragma unroll 1
        for (int cc = p.det ? TAPS * BLOCK_N : 0; cc < TAPS * BLOCK_N; cc += 32) {
            const int tt = cc / BLOCK_N, c0 = cc - tt * BLOCK_N;
            uint32_t r[32];
            ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cc, r);
            ptx::tmem_ld_wait();
            if (co < p.cout) {
                float* dst = p.dw + (static_cast<long long>(co) * p.n_taps + tap0 + tt) * p.cin + ci0 + c0;
                if (ci0 + c0 + 32 <= p.cin) {
                    // 16-byte vector reductions: 8 per thread instead of 32 scalar atomics (cin % 64 == 0 keeps dst aligned)
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                                     "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                                     : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (ci0 + c0 + j < p.cin) atomicAdd(dst + j, __uint_as_float(r[j]));
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int BLOCK_N, int TAPS>
static int launch_wgrad(const WgMaps& maps, const WgParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = 1024 + static_cast<size_t>(p.stages) * (2 + TAPS * (BLOCK_N / 64)) * WG_BOX_BYTES + (2 * p.stages + 1) * 8 + 16;
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RTSDS_ECUDA; }
        done = true;
    }
    wgrad_tc_kernel<BLOCK_N, TAPS><<<grid, TC_THREADS, smem, st>>>(maps, p);
    count_launch();
    return check_launch("wgrad_tc_kernel");
}

}  // namespace rtsds

// x: NHWC bf16 (pitch d->in_ld), dy: NHWC bf16 (pitch d->out_ld, a multiple of 8; channels >= cout read as zero
// via TMA bounds), dw_packed: fp32 [cout][kh*kw][cin], ACCUMULATED into (caller zeroes it).
// x views / taps come from a TapProblem (a regular conv or the sliding-window stem view); dy: NHWC bf16 [n,oh,ow,ldy].
static int wgrad_run(const TapProblem& t, int cin, int cout, const void* dy, long long ldy, float* dw_packed, cudaStream_t st) {
    WgMaps maps;
    WgParams p;
    memset(&maps, 0, sizeof(maps));
    memset(&p, 0, sizeof(p));
    pick_tile(t.oh, t.ow, &p.tile_w, &p.tile_h);
    p.n_img = t.n_img; p.oh = t.oh; p.ow = t.ow;
    p.tiles_w = static_cast<int>(cdiv(t.ow, p.tile_w));
    p.tiles_h = static_cast<int>(cdiv(t.oh, p.tile_h));
    p.cout = cout; p.cin = cin; p.n_taps = t.n_taps;
    const int block_n = cin >= 256 ? 256 : (cin >= 128 ? 128 : 64);
    p.ci_tiles = static_cast<int>(cdiv(cin, block_n));
    const long long tiles_total = static_cast<long long>(t.n_img) * p.tiles_w * p.tiles_h;
    RTSDS_REQUIRE(tiles_total < (1LL << 30), "conv2d_tc_wgrad: too many tiles");
    p.tiles_total = static_cast<int>(tiles_total);
    const int co_tiles = static_cast<int>(cdiv(cout, TC_BLOCK_M));
    // 64-input-channel layers with a multiple of 3 taps: 3 taps per CTA share the dy tile
    static int multi_tap = -1;
    if (multi_tap < 0) { const char* e = getenv("RTSDS_NO_WGRAD_MULTITAP"); multi_tap = (e && e[0] == '1') ? 0 : 1; }
    const int taps_per = (multi_tap && block_n == 64 && t.n_taps % 3 == 0) ? 3 : 1;
    const long long base = static_cast<long long>(t.n_taps / taps_per) * p.ci_tiles * co_tiles;
    // one CTA per SM is resident (shared memory): aim at whole waves, never a ragged extra one
    long long splits = (base >= num_sms() / 2 ? 1LL : 2LL) * num_sms() / base;
    if (splits > tiles_total) splits = tiles_total;
    if (splits < 1) splits = 1;
    p.tiles_per_split = static_cast<int>(cdiv(tiles_total, splits));
    splits = cdiv(tiles_total, p.tiles_per_split);
    for (int i = 0; i < t.n_taps; ++i) { p.tap_dh[i] = t.dh[i]; p.tap_dw[i] = t.dw[i]; p.tap_map[i] = t.map[i]; }
    p.dw = dw_packed;
    // dy view: channel extent = cout (TMA zero-fills channels >= cout), pitch ldy
    int rc = make_act_map(&maps.a, dy, cout, t.ow, t.oh, t.n_img, ldy, static_cast<long long>(t.ow) * ldy,
                          static_cast<long long>(t.oh) * t.ow * ldy, p.tile_w, p.tile_h);
    if (rc != RTSDS_OK) return rc;
    int first = -1;
    for (int i = 0; i < 4; ++i) {
        if (!t.view[i].used) continue;
        rc = make_act_map(&maps.b[i], t.view[i].base, cin, t.view[i].wd, t.view[i].hd, t.n_img, t.view[i].sw, t.view[i].sh,
                          t.view[i].sn, p.tile_w, p.tile_h);
        if (rc != RTSDS_OK) return rc;
        if (first < 0) first = i;
    }
    for (int i = 0; i < 4; ++i)
        if (!t.view[i].used) maps.b[i] = maps.b[first];
    int stages = taps_per == 3 ? 2 : (block_n == 256 ? 2 : (block_n == 128 ? 3 : 4));
    if (stages > p.tiles_per_split) stages = p.tiles_per_split < 2 ? 2 : p.tiles_per_split;
    p.stages = stages;
    dim3 grid(static_cast<unsigned>(splits), static_cast<unsigned>(t.n_taps / taps_per * p.ci_tiles), static_cast<unsigned>(co_tiles));
    const size_t n_dw = static_cast<size_t>(cout) * t.n_taps * cin;
    if (det_mode()) {
        p.det = det_scratch(st, n_dw);
        if (!p.det) return RTSDS_ECUDA;
    }
    if (taps_per == 3) rc = launch_wgrad<64, 3>(maps, p, grid, st);
    else if (block_n == 256) rc = launch_wgrad<256, 1>(maps, p, grid, st);
    else if (block_n == 128) rc = launch_wgrad<128, 1>(maps, p, grid, st);
    else rc = launch_wgrad<64, 1>(maps, p, grid, st);
    if (rc == RTSDS_OK && p.det) rc = det_finish(p.det, dw_packed, n_dw, true, st);
    return rc;
}

extern "C" int rtsds_conv2d_tc_wgrad(const RtsdsConvDesc* d, const void* x, const void* dy, float* dw_packed,
                                     rtsds_stream_t s) {
    RTSDS_REQUIRE(d && x && dy && dw_packed, "conv2d_tc_wgrad: NULL argument");
    RTSDS_REQUIRE(d->in_dtype == RTSDS_BF16, "conv2d_tc_wgrad: operands must be bf16");
    RTSDS_REQUIRE(d->out_ld % 8 == 0 && d->out_ld >= d->cout, "conv2d_tc_wgrad: dy pitch must be a multiple of 8");
    RTSDS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0, "conv2d_tc_wgrad: alignment");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    TapProblem t;
    rc = fwd_problem(d, x, nullptr, TC_BLOCK_K, 2, &t);
    if (rc != RTSDS_OK) return rc;
    return wgrad_run(t, d->cin, d->cout, dy, d->out_ld, dw_packed, as_stream(s));
}

// =====================================================================================
// Space-to-depth stems.  A k x k stride-2 conv on the 3-channel image (7x7 p3: torchvision resnet conv1 via
// models/bisenet/build_contextpath.py:19 and deeplabv2.py:72; 3x3 p1: build_bisenet.py:24) is a 4x4 stride-1 conv over
// the space-to-depth image (12 channels, stored padded to 16).  In the padded NHWC tensor
//     P[n, i, j, (py*2+px)*3 + c] = x[n, c, 2(i-2)+py, 2(j-2)+px]        [n, OH+3, OW+3, 16] bf16, zero outside
// the 4 pixels x 16 channels of one window ROW are 64 contiguous bf16 = one 128-byte UMMA K row, and windows of
// neighbouring output pixels overlap with a pixel pitch of 16 elements.  A TMA tensor map with that (overlapping)
// pixel stride makes the stem an ordinary 4-tap implicit GEMM (K = 4 x 64) for the kernels above — forward and
// weight gradient — instead of a thread-gathered im2col.
// =====================================================================================
static void stem_s2d_problem(const void* P, int n, int oh, int ow, TapProblem* t) {
    memset(t, 0, sizeof(*t));
    const long long rowp = static_cast<long long>(ow + 3) * 16;
    t->view[0] = TapView{P, ow, oh + 3, 16, rowp, rowp * (oh + 3), true};
    t->ck = 64; t->c_extent = 64; t->n_img = n; t->oh = oh; t->ow = ow;
    t->n_taps = 4;
    for (int r = 0; r < 4; ++r) { t->dh[r] = static_cast<signed char>(r); t->dw[r] = 0; t->map[r] = 0; t->kb[r] = r; }
    t->w_ktot = 256;
}

static int stem_s2d_conv_fwd_impl(const void* P, int p_dtype, int n, int oh, int ow, const void* w_packed, int cout,
                                  const float* scale, const float* shift, int act, float* stats, void* y,
                                  int out_ld, int out_dtype, rtsds_stream_t s) {
    RTSDS_REQUIRE(P && w_packed && y && n > 0 && oh > 0 && ow > 0 && cout > 0, "stem_s2d_conv_fwd: bad argument");
    RTSDS_REQUIRE(is_16bit(p_dtype) && (out_dtype == p_dtype || out_dtype == RTSDS_F32), "stem_s2d_conv_fwd: bad dtypes");
    RTSDS_REQUIRE(out_ld >= cout && out_ld % (out_dtype != RTSDS_F32 ? 8 : 4) == 0, "stem_s2d_conv_fwd: out_ld");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    TapProblem t;
    stem_s2d_problem(P, n, oh, ow, &t);
    t.w = w_packed; t.cout = cout;
    t.out_sw = out_ld; t.out_sh = static_cast<long long>(ow) * out_ld; t.out_sn = t.out_sh * oh;
    t.scale = scale; t.shift = shift; t.stats = stats; t.y = y;
    t.out_dtype = out_dtype; t.act = act; t.split_req = 1; t.in_f16 = p_dtype == RTSDS_F16;
    return tp_run(t, nullptr, 0, as_stream(s));
}

extern "C" int rtsds_stem_s2d_conv_fwd(const void* P, int n, int oh, int ow, const void* w_packed, int cout,
                                       const float* scale, const float* shift, int act, float* stats, void* y,
                                       int out_ld, int out_dtype, rtsds_stream_t s) {
    return stem_s2d_conv_fwd_impl(P, RTSDS_BF16, n, oh, ow, w_packed, cout, scale, shift, act, stats, y, out_ld, out_dtype, s);
}

// P (and the packed weights) of p_dtype = RTSDS_BF16 or RTSDS_F16 (eval-mode inference).
extern "C" int rtsds_stem_s2d_conv_fwd_dt(const void* P, int p_dtype, int n, int oh, int ow, const void* w_packed, int cout,
                                          const float* scale, const float* shift, int act, void* y, int out_ld, int out_dtype,
                                          rtsds_stream_t s) {
    return stem_s2d_conv_fwd_impl(P, p_dtype, n, oh, ow, w_packed, cout, scale, shift, act, nullptr, y, out_ld, out_dtype, s);
}

extern "C" int rtsds_stem_s2d_conv_wgrad(const void* P, int n, int oh, int ow, const void* dy, int dy_ld, int cout,
                                         float* dw_packed, rtsds_stream_t s) {
    RTSDS_REQUIRE(P && dy && dw_packed && n > 0 && oh > 0 && ow > 0 && cout > 0 && dy_ld >= cout && dy_ld % 8 == 0, "stem_s2d_conv_wgrad: bad argument");
    int rc = rtsds_check_device();
    if (rc != RTSDS_OK) return rc;
    TapProblem t;
    stem_s2d_problem(P, n, oh, ow, &t);
    return wgrad_run(t, 64, cout, dy, dy_ld, dw_packed, as_stream(s));
}
