// Fused multi-tensor optimizer step + packed-operand refresh.
//
// Replaces, per training step (train.py:96 `optimizer.step()`, :269-270 both optimizers):
//   * torch.optim.Adam / SGD as built at main.py:110-120 (one fused `multi_tensor_apply` launch plus ~30 fill /
//     bookkeeping kernels), and
//   * the re-packing of every conv weight into the tensor-core operand layouts that follows it
//     (rtsds_pack_conv_weights_batch: forward [cout_pad][taps][cin] and dgrad [cin_pad][taps][ck] copies in bf16)
// by ONE launch over a table of tensors: a block owns either a flat chunk of a plain parameter (BatchNorm affine, biases,
// the fp32 ARM / FFM 1x1 convs) or a 32-output-channel x ci_t-input-channel x taps tile of an OIHW conv weight.  The tile
// is read once (p, g and the moment buffers, coalesced along OIHW rows), updated, written back in place, staged in shared
// memory and written out transposed into both packed layouts — the fp32 master never makes a second trip through HBM.
// Per-parameter-group learning rate / weight decay arrive by value every step, so utils.poly_lr_scheduler (utils.py:33-48),
// which rewrites param_groups[0]['lr'] only, keeps working on the stock optimizer object.
//
// Arithmetic follows torch.optim exactly (non-amsgrad Adam with L2 weight decay folded into the gradient; SGD with
// momentum, dampening 0, no Nesterov); tests/test_gpu_optim.py holds 10 steps against torch.optim.
#include "common.cuh"
#include <cstring>

namespace rtsds {

constexpr int OPT_CO = 32, OPT_ROW = 288;            // same tile as pack_batch_kernel (conv_simt.cu)
constexpr int OPT_FLAT = 2048;                       // elements of a plain parameter per block
__host__ __device__ inline int opt_ci_tile(int taps) {
    const int t = OPT_ROW / taps;
    return t < 1 ? 1 : (t > 256 ? 256 : t);
}

__host__ __device__ inline int opt_cin_max(const RtsdsOptJob& j) {
    int c = j.cin;
    if (j.out_fwd && j.cin_pad_fwd > c) c = j.cin_pad_fwd;
    if (j.out_dgrad && j.cin_pad_dgrad > c) c = j.cin_pad_dgrad;
    return c;
}

struct Upd {
    float lr, wd, beta1, beta2, eps, step_size_scale, inv_bc2_sqrt, momentum;
    int kind, first;
};

__device__ __forceinline__ float opt_update(const Upd& u, float p, float g, float* m, float* v) {
    if (u.wd != 0.f) g = fmaf(u.wd, p, g);
    if (u.kind == RTSDS_OPT_ADAM) {
        const float mm = *m + (g - *m) * (1.f - u.beta1);               // exp_avg.lerp_(grad, 1 - beta1)
        const float vv = fmaf(g * (1.f - u.beta2), g, *v * u.beta2);    // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
        *m = mm; *v = vv;
        const float denom = sqrtf(vv) * u.inv_bc2_sqrt + u.eps;
        return p - (u.lr * u.step_size_scale) * (mm / denom);            // step_size = lr / bias_correction1
    }
    const float b = u.momentum != 0.f ? (u.first ? g : fmaf(u.momentum, *m, g)) : g;      // SGD: buf = mu*buf + g (first step: g)
    *m = b;
    return p - u.lr * b;
}

template <typename T>
__global__ void __launch_bounds__(256)
optim_step_kernel(const RtsdsOptJob* __restrict__ jobs, const int* __restrict__ first_block, int n_jobs,
                  const __grid_constant__ RtsdsOptHyper h) {
    __shared__ float s_w[OPT_CO][OPT_ROW + 1];
    // block -> job (binary search over the prefix sums)
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (static_cast<int>(blockIdx.x) >= first_block[mid]) lo = mid; else hi = mid - 1;
    }
    const RtsdsOptJob j = jobs[lo];
    const int local = blockIdx.x - first_block[lo];
    Upd u;
    u.lr = h.lr[j.group]; u.wd = h.weight_decay[j.group]; u.beta1 = h.beta1; u.beta2 = h.beta2; u.eps = h.eps;
    u.step_size_scale = h.inv_bias_correction1; u.inv_bc2_sqrt = h.inv_bias_correction2_sqrt; u.momentum = h.momentum;
    u.kind = h.kind; u.first = h.first_step;
    const bool adam = h.kind == RTSDS_OPT_ADAM;

    if (j.taps == 0) {                                  // ---- plain parameter: flat chunk ----
        const long long base = static_cast<long long>(local) * OPT_FLAT;
#pragma unroll
        for (int k = 0; k < OPT_FLAT / 256; ++k) {
            const long long i = base + k * 256 + threadIdx.x;
            if (i < j.numel) {
                float m = j.m ? j.m[i] : 0.f, v = adam ? j.v[i] : 0.f;
                const float p = opt_update(u, j.p[i], j.g[i], &m, &v);
                j.p[i] = p;
                if (j.m) j.m[i] = m;
                if (adam) j.v[i] = v;
            }
        }
        return;
    }
    // ---- conv weight OIHW [cout][cin][taps]: tile of OPT_CO output channels x ci_t input channels x all taps ----
    const int taps = j.taps, ci_t = opt_ci_tile(taps);
    const int n_ci_tiles = (opt_cin_max(j) + ci_t - 1) / ci_t;
    const int co0 = (local / n_ci_tiles) * OPT_CO, ci0 = (local % n_ci_tiles) * ci_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nci = min(ci_t, j.cin - ci0);            // valid input channels of this tile (may be <= 0)
    const int row = nci > 0 ? nci * taps : 0;
#pragma unroll
    for (int r = warp; r < OPT_CO; r += 8) {
        const int co = co0 + r;
        if (co < j.cout) {
            const long long off = (static_cast<long long>(co) * j.cin + ci0) * taps;
            for (int e = lane; e < row; e += 32) {
                float p = j.p[off + e];
                if (j.g) {                              // frozen / gradient-less tensors are only re-packed
                    float m = j.m ? j.m[off + e] : 0.f, v = adam ? j.v[off + e] : 0.f;
                    p = opt_update(u, p, j.g[off + e], &m, &v);
                    j.p[off + e] = p;
                    if (j.m) j.m[off + e] = m;
                    if (adam) j.v[off + e] = v;
                }
                s_w[r][e] = p;
            }
        }
    }
    __syncthreads();
    if (j.out_fwd) {                // [cout_pad][taps][cin_pad_fwd]
        T* out = reinterpret_cast<T*>(j.out_fwd);
        const int ci_n = min(ci_t, j.cin_pad_fwd - ci0);
        const int co_n = min(OPT_CO, j.cout_pad - co0);
        for (int rt = warp; rt < co_n * taps; rt += 8) {
            const int r = rt / taps, t = rt - r * taps;
            const bool co_ok = co0 + r < j.cout;
            T* dst = out + (static_cast<long long>(co0 + r) * taps + t) * j.cin_pad_fwd + ci0;
            for (int c = lane; c < ci_n; c += 32) dst[c] = from_f32<T>((co_ok && c < nci) ? s_w[r][c * taps + t] : 0.f);
        }
    }
    if (j.out_dgrad) {              // dgrad operand [cin_pad_dgrad][taps][ck]
        T* out = reinterpret_cast<T*>(j.out_dgrad);
        const int ci_n = min(ci_t, j.cin_pad_dgrad - ci0);
        for (int ct = warp; ct < ci_n * taps; ct += 8) {
            const int c = ct / taps, t = ct - c * taps;
            const int co = co0 + lane;
            if (co < j.ck)
                out[(static_cast<long long>(ci0 + c) * taps + t) * j.ck + co] =
                    from_f32<T>((co < j.cout && c < nci) ? s_w[lane][c * taps + t] : 0.f);
        }
    }
}

}  // namespace rtsds

using namespace rtsds;

extern "C" int rtsds_optim_job_blocks(const RtsdsOptJob* j) {
    if (!j) return 0;
    if (j->taps == 0) return static_cast<int>(cdiv(j->numel, OPT_FLAT));
    const int cin_max = opt_cin_max(*j);
    int co_lim = j->cout;
    if (j->out_fwd && j->cout_pad > co_lim) co_lim = j->cout_pad;
    if (j->out_dgrad && j->ck > co_lim) co_lim = j->ck;
    return static_cast<int>(cdiv(co_lim, OPT_CO) * cdiv(cin_max, opt_ci_tile(j->taps)));
}

extern "C" int rtsds_optim_step(const RtsdsOptJob* jobs_dev, const int* first_block_dev, int n_jobs, int total_blocks,
                                const RtsdsOptHyper* hyper, int pack_dtype, rtsds_stream_t s) {
    RTSDS_REQUIRE(jobs_dev && first_block_dev && hyper && n_jobs > 0 && total_blocks > 0, "optim_step: bad argument");
    RTSDS_REQUIRE(hyper->kind == RTSDS_OPT_ADAM || hyper->kind == RTSDS_OPT_SGD, "optim_step: unknown optimizer kind %d", hyper->kind);
    RTSDS_REQUIRE(pack_dtype == RTSDS_BF16 || pack_dtype == RTSDS_F32 || pack_dtype == RTSDS_F16, "optim_step: bad pack dtype");
    if (pack_dtype == RTSDS_BF16) optim_step_kernel<__nv_bfloat16><<<total_blocks, 256, 0, as_stream(s)>>>(jobs_dev, first_block_dev, n_jobs, *hyper);
    else if (pack_dtype == RTSDS_F16) optim_step_kernel<__half><<<total_blocks, 256, 0, as_stream(s)>>>(jobs_dev, first_block_dev, n_jobs, *hyper);
    else optim_step_kernel<float><<<total_blocks, 256, 0, as_stream(s)>>>(jobs_dev, first_block_dev, n_jobs, *hyper);
    count_launch();
    return check_launch("optim_step_kernel");
}
