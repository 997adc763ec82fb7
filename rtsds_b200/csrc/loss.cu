// Loss / prediction kernels.
//   nn.CrossEntropyLoss(ignore_index) as built at main.py:124-130 and applied at
//   train.py:86-92,202-204: mean over non-ignored pixels of -log softmax[target];
//   main_output.max(1) / predicted.eq(targets).sum() (train.py:102-106);
//   torch.argmax(outputs, 1) (validation.py:51).
// The fused variants evaluate the bilinear resize of the low-resolution logits
// z (build_bisenet.py:158-159,166) on the fly, so the three full-resolution
// [N,19,H,W] fp32 tensors (39.8 MB/image each at 512x1024) are never written
// or re-read: HBM traffic is the int64 target map (8 B/pixel) plus z.
#include "common.cuh"

namespace rtsds {

constexpr int LS_THREADS = 256, LS_PX = 4, LS_TILE = LS_THREADS * LS_PX, LS_MAXC = 32;

struct Online {
    float m, s, best;
    int arg;
    __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; best = -INFINITY; arg = 0; }
    __device__ __forceinline__ void push(float x, int k) {
        if (x > m) { s = s * expf(m - x) + 1.f; m = x; } else { s += expf(x - m); }
        if (x > best || (x != x && best == best)) { best = x; arg = k; }
    }
    __device__ __forceinline__ float lse() const { return m + logf(s); }
};

__device__ __forceinline__ void block_acc3(double a, double b, double c, double* acc) {
    __shared__ double sh[3][LS_THREADS / 32];
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; sh[2][warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0, tb = 0, tc = 0;
        for (int i = 0; i < LS_THREADS / 32; ++i) { ta += sh[0][i]; tb += sh[1][i]; tc += sh[2][i]; }
        if (ta != 0.0) atomicAdd(&acc[0], ta);
        if (tb != 0.0) atomicAdd(&acc[1], tb);
        if (tc != 0.0) atomicAdd(&acc[2], tc);
    }
}

// grid (ceil(ow/1024), oh, n)
__global__ void __launch_bounds__(LS_THREADS)
resize_ce_fwd_kernel(const float* __restrict__ z, int h, int w, int c, int z_ld, int oh, int ow, float rh, float rw,
                     const long long* __restrict__ target, long long ignore_index, double* acc,
                     long long* __restrict__ pred_out) {
    extern __shared__ float sm[];
    const int img = blockIdx.z, oy = blockIdx.y;
    const int ox0 = blockIdx.x * LS_TILE;
    const int ox_last = min(ox0 + LS_TILE, ow) - 1;
    const Lerp ly = lerp_src(oy, rh, h);
    const int xs = lerp_src(ox0, rw, w).i0;
    const int xe = lerp_src(ox_last, rw, w).i1;
    const int ncols = xe - xs + 1;
    float* s0 = sm;
    float* s1 = sm + ncols * c;
    const float* r0 = z + (static_cast<long long>(img) * h + ly.i0) * w * z_ld;
    const float* r1 = z + (static_cast<long long>(img) * h + ly.i1) * w * z_ld;
    for (int i = threadIdx.x; i < ncols * c; i += LS_THREADS) {
        const int col = i / c, ch = i - col * c;
        s0[i] = __ldg(r0 + static_cast<long long>(xs + col) * z_ld + ch);
        s1[i] = __ldg(r1 + static_cast<long long>(xs + col) * z_ld + ch);
    }
    __syncthreads();
    double loss = 0.0, nvalid = 0.0, ncorrect = 0.0;
    const int ox = ox0 + threadIdx.x * LS_PX;
    if (ox < ow) {
        const long long pix0 = (static_cast<long long>(img) * oh + oy) * ow + ox;
#pragma unroll
        for (int j = 0; j < LS_PX; ++j) {
            if (ox + j >= ow) break;
            Lerp lx = lerp_src(ox + j, rw, w);
            lx.i0 -= xs; lx.i1 -= xs;
            const long long t = target ? __ldg(target + pix0 + j) : ignore_index;
            Online o; o.init();
            float vt = 0.f;
            for (int ch = 0; ch < c; ++ch) {
                const float v = ly.l0 * (lx.l0 * s0[lx.i0 * c + ch] + lx.l1 * s0[lx.i1 * c + ch]) +
                                ly.l1 * (lx.l0 * s1[lx.i0 * c + ch] + lx.l1 * s1[lx.i1 * c + ch]);
                o.push(v, ch);
                if (ch == t) vt = v;
            }
            if (pred_out) pred_out[pix0 + j] = o.arg;
            if (t != ignore_index && t >= 0 && t < c) {
                loss += static_cast<double>(o.lse() - vt);
                nvalid += 1.0;
            }
            if (o.arg == t) ncorrect += 1.0;
        }
    }
    if (acc) block_acc3(loss, nvalid, ncorrect, acc);
}

// Backward of mean-CE through the bilinear resize, scattered to z resolution.
// grid (ceil(ow/1024), ceil(oh/rows_per_block), n)
__global__ void __launch_bounds__(LS_THREADS)
resize_ce_bwd_kernel(const float* __restrict__ z, int h, int w, int c, int z_ld, int oh, int ow, float rh, float rw,
                     int rows_per_block, const long long* __restrict__ target, long long ignore_index,
                     const float* __restrict__ grad_scale, float* __restrict__ dz) {
    extern __shared__ float sm[];
    const int img = blockIdx.z;
    const int oy0 = blockIdx.y * rows_per_block;
    const int oy1 = min(oy0 + rows_per_block, oh) - 1;
    const int ox0 = blockIdx.x * LS_TILE;
    const int ox_last = min(ox0 + LS_TILE, ow) - 1;
    const int ys = lerp_src(oy0, rh, h).i0, ye = lerp_src(oy1, rh, h).i1;
    const int xs = lerp_src(ox0, rw, w).i0, xe = lerp_src(ox_last, rw, w).i1;
    const int nrows = ye - ys + 1, ncols = xe - xs + 1;
    float* s_z = sm;                               // [nrows][ncols][c]
    float* s_d = sm + nrows * ncols * c;
    for (int i = threadIdx.x; i < nrows * ncols * c; i += LS_THREADS) {
        const int ch = i % c;
        const int rc = i / c;
        const int col = rc % ncols, row = rc / ncols;
        s_z[i] = __ldg(z + ((static_cast<long long>(img) * h + ys + row) * w + xs + col) * z_ld + ch);
        s_d[i] = 0.f;
    }
    __syncthreads();
    const float gs = __ldg(grad_scale);
    const int ox = ox0 + threadIdx.x * LS_PX;
    if (ox < ow) {
        for (int oy = oy0; oy <= oy1; ++oy) {
            Lerp ly = lerp_src(oy, rh, h);
            ly.i0 -= ys; ly.i1 -= ys;
            const long long pix0 = (static_cast<long long>(img) * oh + oy) * ow + ox;
#pragma unroll
            for (int j = 0; j < LS_PX; ++j) {
                if (ox + j >= ow) break;
                const long long t = __ldg(target + pix0 + j);
                if (t == ignore_index || t < 0 || t >= c) continue;
                Lerp lx = lerp_src(ox + j, rw, w);
                lx.i0 -= xs; lx.i1 -= xs;
                const float w00 = ly.l0 * lx.l0, w01 = ly.l0 * lx.l1, w10 = ly.l1 * lx.l0, w11 = ly.l1 * lx.l1;
                const int o00 = (ly.i0 * ncols + lx.i0) * c, o01 = (ly.i0 * ncols + lx.i1) * c;
                const int o10 = (ly.i1 * ncols + lx.i0) * c, o11 = (ly.i1 * ncols + lx.i1) * c;
                float v[LS_MAXC];
                float m = -INFINITY;
#pragma unroll
                for (int ch = 0; ch < LS_MAXC; ++ch) {
                    if (ch < c) {
                        v[ch] = ly.l0 * (lx.l0 * s_z[o00 + ch] + lx.l1 * s_z[o01 + ch]) +
                                ly.l1 * (lx.l0 * s_z[o10 + ch] + lx.l1 * s_z[o11 + ch]);
                        m = fmaxf(m, v[ch]);
                    }
                }
                float s = 0.f;
#pragma unroll
                for (int ch = 0; ch < LS_MAXC; ++ch)
                    if (ch < c) { v[ch] = expf(v[ch] - m); s += v[ch]; }
                const float inv = gs / s;
#pragma unroll
                for (int ch = 0; ch < LS_MAXC; ++ch) {
                    if (ch < c) {
                        const float g = v[ch] * inv - (ch == t ? gs : 0.f);
                        atomicAdd(&s_d[o00 + ch], w00 * g);
                        atomicAdd(&s_d[o01 + ch], w01 * g);
                        atomicAdd(&s_d[o10 + ch], w10 * g);
                        atomicAdd(&s_d[o11 + ch], w11 * g);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nrows * ncols * c; i += LS_THREADS) {
        const float g = s_d[i];
        if (g != 0.f) {
            const int ch = i % c;
            const int rc = i / c;
            const int col = rc % ncols, row = rc / ncols;
            atomicAdd(dz + ((static_cast<long long>(img) * h + ys + row) * w + xs + col) * z_ld + ch, g);
        }
    }
}

// CE + argmax over materialised fp32 NCHW logits; 4 pixels per thread.
__global__ void __launch_bounds__(LS_THREADS)
ce_nchw_fwd_kernel(const float* __restrict__ logits, int n, int c, long long hw, const long long* __restrict__ target,
                   long long ignore_index, double* acc, long long* __restrict__ pred_out, int vec) {
    const long long gpi = (hw + 3) / 4;
    const long long total = gpi * n;
    double loss = 0.0, nvalid = 0.0, ncorrect = 0.0;
    for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long img = g / gpi;
        const long long p0 = (g - img * gpi) * 4;
        const float* base = logits + img * c * hw + p0;
        Online o[4];
        float vt[4] = {0, 0, 0, 0};
        long long t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j].init();
            t[j] = (target && p0 + j < hw) ? __ldg(target + img * hw + p0 + j) : ignore_index;
        }
        const bool full = vec && (p0 + 4 <= hw);
        for (int ch = 0; ch < c; ++ch) {
            float x[4];
            if (full) {
                float4 q = __ldcs(reinterpret_cast<const float4*>(base + ch * hw));
                x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = (p0 + j < hw) ? base[ch * hw + j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j].push(x[j], ch);
                if (ch == t[j]) vt[j] = x[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (p0 + j >= hw) break;
            if (pred_out) pred_out[img * hw + p0 + j] = o[j].arg;
            if (t[j] != ignore_index && t[j] >= 0 && t[j] < c) {
                loss += static_cast<double>(o[j].lse() - vt[j]);
                nvalid += 1.0;
            }
            if (target && o[j].arg == t[j]) ncorrect += 1.0;
        }
    }
    if (acc) block_acc3(loss, nvalid, ncorrect, acc);
}

}  // namespace rtsds

using namespace rtsds;

static size_t ls_cols(int w, int ow) { return static_cast<size_t>(static_cast<double>(LS_TILE) * w / ow) + 4; }

extern "C" int rtsds_resize_ce_argmax_fwd(const float* z, int n, int h, int w, int c, int z_ld, int oh, int ow,
                                          const int64_t* target, int64_t ignore_index, double* acc,
                                          int64_t* pred_out, rtsds_stream_t s) {
    RTSDS_REQUIRE(z && n > 0 && h > 0 && w > 0 && c > 0 && z_ld >= c && oh > 0 && ow > 0, "resize_ce_argmax_fwd: bad argument");
    RTSDS_REQUIRE(target || pred_out, "resize_ce_argmax_fwd: nothing to compute");
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    size_t cols = ls_cols(w, ow);
    if (cols > static_cast<size_t>(w)) cols = w;
    size_t smem = sizeof(float) * 2 * cols * c;
    RTSDS_REQUIRE(smem <= 200 * 1024, "resize_ce_argmax_fwd: tile needs %zu bytes of shared memory", smem);
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(resize_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); done = true; }
    dim3 grid(static_cast<unsigned>(cdiv(ow, LS_TILE)), oh, n);
    resize_ce_fwd_kernel<<<grid, LS_THREADS, smem, as_stream(s)>>>(z, h, w, c, z_ld, oh, ow, rh, rw,
                                                                   reinterpret_cast<const long long*>(target), ignore_index, acc,
                                                                   reinterpret_cast<long long*>(pred_out));
    count_launch();
    return check_launch("resize_ce_fwd_kernel");
}

extern "C" int rtsds_resize_ce_bwd(const float* z, int n, int h, int w, int c, int z_ld, int oh, int ow,
                                   const int64_t* target, int64_t ignore_index, const float* grad_scale, float* dz,
                                   rtsds_stream_t s) {
    RTSDS_REQUIRE(z && target && grad_scale && dz, "resize_ce_bwd: NULL argument");
    RTSDS_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c <= LS_MAXC && z_ld >= c && oh > 0 && ow > 0, "resize_ce_bwd: bad shape");
    const float rh = resize_scale(h, oh), rw = resize_scale(w, ow);
    size_t cols = ls_cols(w, ow);
    if (cols > static_cast<size_t>(w)) cols = w;
    int rpb = 8;
    size_t smem = 0;
    for (; rpb >= 1; rpb >>= 1) {
        size_t rows = static_cast<size_t>(static_cast<double>(rpb) * h / oh) + 3;
        if (rows > static_cast<size_t>(h)) rows = h;
        smem = sizeof(float) * 2 * rows * cols * c;
        if (smem <= 160 * 1024) break;
    }
    RTSDS_REQUIRE(rpb >= 1, "resize_ce_bwd: tile does not fit shared memory (%zu bytes)", smem);
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(resize_ce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); done = true; }
    dim3 grid(static_cast<unsigned>(cdiv(ow, LS_TILE)), static_cast<unsigned>(cdiv(oh, rpb)), n);
    resize_ce_bwd_kernel<<<grid, LS_THREADS, smem, as_stream(s)>>>(z, h, w, c, z_ld, oh, ow, rh, rw, rpb,
                                                                   reinterpret_cast<const long long*>(target), ignore_index,
                                                                   grad_scale, dz);
    count_launch();
    return check_launch("resize_ce_bwd_kernel");
}

extern "C" int rtsds_ce_argmax_nchw_fwd(const float* logits, int n, int c, int64_t hw, const int64_t* target,
                                        int64_t ignore_index, double* acc, int64_t* pred_out, rtsds_stream_t s) {
    RTSDS_REQUIRE(logits && n > 0 && c > 0 && hw > 0, "ce_argmax_nchw_fwd: bad argument");
    RTSDS_REQUIRE(target || pred_out, "ce_argmax_nchw_fwd: nothing to compute");
    const int vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
    const long long total = cdiv(hw, 4) * n;
    long long want = cdiv(total, LS_THREADS);
    int grid = static_cast<int>(want > 8LL * num_sms() ? 8LL * num_sms() : want);
    ce_nchw_fwd_kernel<<<grid, LS_THREADS, 0, as_stream(s)>>>(logits, n, c, hw, reinterpret_cast<const long long*>(target),
                                                              ignore_index, acc, reinterpret_cast<long long*>(pred_out), vec);
    count_launch();
    return check_launch("ce_nchw_fwd_kernel");
}
