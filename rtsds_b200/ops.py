"""Thin tensor-level wrappers over the C ABI (include/rtsds_b200.h).

PyTorch is used for device memory and streams only: every function passes
`tensor.data_ptr()`, sizes and the current CUDA stream to librtsds_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F16, F32, ConvDesc, check, lib

__all__ = [
    "F32", "BF16", "F16", "ACT_NONE", "ACT_RELU", "ACT_LRELU", "ConvDesc", "dtype_code", "torch_dtype",
    "confusion_hist", "argmax_hist", "conv_out_size", "cout_pad", "pack_conv_weight", "conv2d_tc",
    "conv2d_simt", "stem_conv", "maxpool3x3s2", "maxpool3x3s2_bwd_idx", "bn_fold", "bn_finalize", "bn_finalize_apply_ptr", "scale_shift_act",
    "global_avgpool", "arm_gate", "gate_resize_nhwc", "ffm_head", "resize_to_nchw",
    "resize_ce_argmax_fwd", "resize_ce_bwd", "ce_argmax_nchw_fwd", "launch_count",
]


def _p(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _s():
    """Raw cudaStream_t of torch's current stream on the current device (the C accessors: torch.cuda.current_stream()
    costs ~12 us of Python per call, which is a third of a batch-1 frame's host time)."""
    if _lib.dry_run():
        return None
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _cuda(*ts):
    for t in ts:
        if t is not None and not isinstance(t, int) and not t.is_cuda and not _lib.dry_run():
            raise _lib.RtsdsError("rtsds_b200 kernels need CUDA tensors (there is no CPU fallback)")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    raise TypeError(f"unsupported activation dtype {dt}")


def torch_dtype(code: int) -> torch.dtype:
    return torch.bfloat16 if code == BF16 else torch.float16 if code == F16 else torch.float32


def launch_count() -> int:
    return int(lib().rtsds_launch_count())


def set_deterministic(on: bool = True) -> None:
    """Order-independent (exact fixed-point) accumulation for the train-mode BatchNorm sums, weight-gradient partials and
    BatchNorm-backward sums instead of fp32 atomics (include/rtsds_b200.h: rtsds_set_deterministic).  For parity runs:
    the reference's CPU path sums in a fixed order, so its results are reproducible run to run."""
    lib().rtsds_set_deterministic(1 if on else 0)


def is_deterministic() -> bool:
    return bool(lib().rtsds_get_deterministic())


# ----------------------------------------------------------------------------- metric
def confusion_hist(label: torch.Tensor, pred: torch.Tensor, n_cls: int, hist: torch.Tensor,
                   n_bad: torch.Tensor | None = None) -> None:
    """hist[int64 n_cls*n_cls] += fast_hist(label, pred, n_cls) (utils.py:52-58)."""
    _cuda(label, pred, hist)
    assert label.dtype == torch.int64 and pred.dtype == torch.int64 and hist.dtype == torch.int64
    assert label.is_contiguous() and pred.is_contiguous() and label.numel() == pred.numel()
    check(lib().rtsds_confusion_hist(_p(label), _p(pred), label.numel(), n_cls, _p(hist), _p(n_bad), _s()),
          "confusion_hist")


def argmax_hist(logits: torch.Tensor, label: torch.Tensor | None, hist: torch.Tensor | None,
                pred_out: torch.Tensor | None = None) -> None:
    """pred_out: int64 [n,h,w] (torch.argmax's type) or uint8 [n,h,w] (one byte per pixel for the trip back to the host)."""
    _cuda(logits, label, hist, pred_out)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 4
    n, c, h, w = logits.shape
    if pred_out is not None and pred_out.dtype == torch.uint8:
        check(lib().rtsds_argmax_hist_u8(_p(logits), _p(label), n, c, h * w, _p(pred_out), _p(hist), _s()), "argmax_hist_u8")
        return
    check(lib().rtsds_argmax_hist(_p(logits), _p(label), n, c, h * w, _p(pred_out), _p(hist), _s()), "argmax_hist")


# ----------------------------------------------------------------------------- conv
def conv_out_size(i: int, k: int, stride: int, pad: int, dil: int = 1) -> int:
    return (i + 2 * pad - dil * (k - 1) - 1) // stride + 1


def cout_pad(cout: int) -> int:
    return int(lib().rtsds_conv_cout_pad(cout))


def pack_conv_weight(w: torch.Tensor, dtype: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """OIHW fp32 -> [cout_pad][kh*kw][cin] in `dtype`."""
    _cuda(w)
    w = w.detach()
    assert w.dtype == torch.float32 and w.is_contiguous()
    co, ci, kh, kw = w.shape
    cp = cout_pad(co)
    if out is None:
        out = torch.empty((cp, kh * kw, ci), dtype=torch_dtype(dtype), device=w.device)
    check(lib().rtsds_pack_conv_weight(_p(w), co, ci, kh, kw, cp, dtype, _p(out), _s()), "pack_conv_weight")
    return out


def make_conv_desc(n, h, w, cin, in_ld, cout, out_ld, k, stride, pad, dil=1, act=ACT_NONE, slope=0.0,
                   in_dtype=BF16, out_dtype=BF16, res_ld=0, split_k=0, kw=None) -> ConvDesc:
    kw = k if kw is None else kw
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.in_ld = n, h, w, cin, in_ld
    d.cout, d.out_ld, d.res_ld = cout, out_ld, res_ld
    d.kh, d.kw, d.stride, d.pad, d.dil = k, kw, stride, pad, dil
    d.oh, d.ow = conv_out_size(h, k, stride, pad, dil), conv_out_size(w, kw, stride, pad, dil)
    d.act, d.slope = act, slope
    d.in_dtype, d.out_dtype, d.split_k = in_dtype, out_dtype, split_k
    return d


def conv2d_tc(d: ConvDesc, x, w, y, scale=None, shift=None, residual=None, stats=None, workspace=None) -> None:
    ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    check(lib().rtsds_conv2d_tc_fwd(C.byref(d), _p(x), _p(w), _p(scale), _p(shift), _p(residual), _p(stats), _p(y),
                                    _p(workspace), ws_bytes, _s()), "conv2d_tc_fwd")


def conv2d_simt(d: ConvDesc, x, w, y, scale=None, shift=None, residual=None, stats=None) -> None:
    check(lib().rtsds_conv2d_simt_fwd(C.byref(d), _p(x), _p(w), _p(scale), _p(shift), _p(residual), _p(stats), _p(y),
                                      _s()), "conv2d_simt_fwd")


def dgrad_ck(cout: int, tc: bool) -> int:
    """Channel extent of the dy operand / dgrad weight K dimension."""
    return (cout + 63) // 64 * 64 if tc else cout


def pack_conv_weight_dgrad(w: torch.Tensor, dtype: int, tc: bool, out: torch.Tensor | None = None) -> torch.Tensor:
    """OIHW fp32 -> dgrad operand [cout_pad(cin)][kh*kw][ck] in `dtype`."""
    w = w.detach()
    co, ci, kh, kw = w.shape
    cp, ck = cout_pad(ci), dgrad_ck(co, tc)
    if out is None:
        out = torch.empty((cp, kh * kw, ck), dtype=torch_dtype(dtype), device=w.device)
    check(lib().rtsds_pack_conv_weight_dgrad(_p(w), co, ci, kh, kw, cp, ck, dtype, _p(out), _s()), "pack_conv_weight_dgrad")
    return out


def conv2d_dgrad(d: ConvDesc, dy, w_dgrad, dx, dx_dtype, tc: bool, residual=None, workspace=None) -> None:
    if tc:
        ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
        check(lib().rtsds_conv2d_tc_dgrad(C.byref(d), _p(dy), _p(w_dgrad), _p(residual), _p(dx), dx_dtype, _p(workspace),
                                          ws_bytes, _s()), "conv2d_tc_dgrad")
    else:
        check(lib().rtsds_conv2d_simt_dgrad(C.byref(d), _p(dy), _p(w_dgrad), _p(residual), _p(dx), dx_dtype, _s()),
              "conv2d_simt_dgrad")


def conv2d_wgrad(d: ConvDesc, x, dy, dw_packed, tc: bool) -> None:
    fn = lib().rtsds_conv2d_tc_wgrad if tc else lib().rtsds_conv2d_simt_wgrad
    check(fn(C.byref(d), _p(x), _p(dy), _p(dw_packed), _s()), "conv2d_wgrad")


def unpack_conv_wgrad(dw_packed, grad_oihw: torch.Tensor, accumulate: bool) -> None:
    co, ci, kh, kw = grad_oihw.shape
    check(lib().rtsds_unpack_conv_wgrad(_p(dw_packed), co, ci, kh, kw, int(accumulate), _p(grad_oihw), _s()),
          "unpack_conv_wgrad")


def pack_conv_weights_batch(jobs, dtype: int, tc: bool) -> None:
    """jobs: list of (weight OIHW fp32 parameter, out tensor, kind) with kind 0 = forward layout
    [cout_pad][taps][cin], 1 = dgrad operand [cout_pad(cin)][taps][ck]; one launch per 40 jobs."""
    if not jobs:
        return
    arr = (_lib.PackJob * len(jobs))()
    for a, (w, out, kind) in zip(arr, jobs):
        w = w.detach()
        co, ci, kh, kw = w.shape
        a.w, a.out, a.cout, a.cin, a.taps, a.kind = _p(w), _p(out), co, ci, kh * kw, kind
        if kind == 0:
            a.cin_pad, a.cout_pad, a.ck = ci, cout_pad(co), 0
        else:
            a.cin_pad, a.cout_pad, a.ck = cout_pad(ci), co, dgrad_ck(co, tc)
    check(lib().rtsds_pack_conv_weights_batch(C.cast(arr, C.c_void_p), len(jobs), dtype, _s()), "pack_conv_weights_batch")


def unpack_conv_wgrads_batch(jobs) -> None:
    """jobs: list of (dw_packed fp32 tensor, grad OIHW view, accumulate); one launch per 48 layers."""
    if not jobs:
        return
    arr = (_lib.UnpackJob * len(jobs))()
    for a, (dw, g, acc) in zip(arr, jobs):
        co, ci, kh, kw = g.shape
        a.dw_packed, a.grad, a.cout, a.cin, a.cin_src, a.taps, a.accumulate = _p(dw), _p(g), co, ci, ci, kh * kw, int(acc)
    check(lib().rtsds_unpack_conv_wgrads_batch(C.cast(arr, C.c_void_p), len(jobs), _s()), "unpack_conv_wgrads_batch")


def stem_conv(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, k: int, stride: int, pad: int, scale=None,
              shift=None, act=ACT_NONE, slope=0.0, softmax_in=False, stats=None) -> None:
    _cuda(x, w, y)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    w = w.detach()
    assert w.dtype == torch.float32 and w.is_contiguous()
    n, cin, h, ww = x.shape
    check(lib().rtsds_stem_conv_fwd(_p(x), _p(w), n, cin, h, ww, w.shape[0], k, stride, pad, _p(scale), _p(shift), act,
                                    slope, int(softmax_in), _p(stats), dtype_code(y.dtype), _p(y), _s()), "stem_conv_fwd")


def stem_pack_weights(w7: torch.Tensor, w3: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((128, 192), dtype=torch.bfloat16, device=w7.device)
    check(lib().rtsds_stem_pack_weights(_p(w7.detach()), _p(w3.detach()), dtype_code(out.dtype), _p(out), _s()), "stem_pack_weights")
    return out


def stem_pair_tc_fwd(x, wpk, y_cp, y_sp, scale=None, shift=None, relu=True, stats_cp=None, stats_sp=None) -> None:
    n, _, h, w = x.shape
    assert wpk.dtype == y_cp.dtype == y_sp.dtype
    check(lib().rtsds_stem_pair_tc_fwd(_p(x), n, h, w, _p(wpk), _p(scale), _p(shift), int(relu), _p(stats_cp), _p(stats_sp),
                                       dtype_code(wpk.dtype), _p(y_cp), _p(y_sp), _s()), "stem_pair_tc_fwd")


def stem_pair_tc_fwd_pool(x, wpk, pool, y_sp, scale, shift) -> None:
    """Both stems + folded BN + ReLU, with MaxPool2d(3,2,1) of the context-path map fused in; `pool` must be zero on entry."""
    n, _, h, w = x.shape
    assert wpk.dtype == pool.dtype == y_sp.dtype
    check(lib().rtsds_stem_pair_tc_fwd_pool(_p(x), n, h, w, _p(wpk), _p(scale), _p(shift), dtype_code(wpk.dtype), _p(pool), _p(y_sp),
                                            _s()), "stem_pair_tc_fwd_pool")


def stem_pair_tc_wgrad(x, d_raw_cp, d_raw_sp, dw_ws, g7, g3) -> None:
    n, _, h, w = x.shape
    check(lib().rtsds_stem_pair_tc_wgrad(_p(x), n, h, w, _p(d_raw_cp), _p(d_raw_sp), _p(dw_ws), _p(g7), _p(g3), _s()),
          "stem_pair_tc_wgrad")


def maxpool3x3s2(x: torch.Tensor, y: torch.Tensor, ceil_mode: bool = False, idx: torch.Tensor | None = None) -> None:
    """idx (int32 [n,oh,ow,c/8], training): receives the first-maximum window positions for maxpool3x3s2_bwd_idx."""
    n, h, w, c = x.shape
    if idx is None:
        check(lib().rtsds_maxpool3x3s2_fwd(_p(x), n, h, w, c, dtype_code(x.dtype), int(ceil_mode), _p(y), _s()), "maxpool")
    else:
        assert idx.dtype == torch.int32 and idx.numel() == y.numel() // 8
        check(lib().rtsds_maxpool3x3s2_fwd_idx(_p(x), n, h, w, c, dtype_code(x.dtype), int(ceil_mode), _p(y), _p(idx), _s()),
              "maxpool_idx")


def maxpool3x3s2_bwd_idx(idx: torch.Tensor, dy: torch.Tensor, dx: torch.Tensor, ceil_mode: bool = False) -> None:
    """dx [n,h,w,c] <- gradient of the max-pool whose forward recorded idx; dy [n,oh,ow,c], same dtype as dx."""
    n, h, w, c = dx.shape
    check(lib().rtsds_maxpool3x3s2_bwd_idx(_p(idx), _p(dy), n, h, w, c, dtype_code(dx.dtype), int(ceil_mode), _p(dx), _s()),
          "maxpool_bwd_idx")


def maxpool_out_size(i: int, ceil_mode: bool = False) -> int:
    if not ceil_mode:
        return (i - 1) // 2 + 1
    o = i // 2 + 1
    if (o - 1) * 2 >= i + 1:
        o -= 1
    return o


# ----------------------------------------------------------------------------- batch norm
def bn_fold(bn, scale: torch.Tensor, shift: torch.Tensor, conv_bias=None) -> None:
    c = scale.numel()
    check(lib().rtsds_bn_fold(_p(bn.weight.detach() if bn.weight is not None else None),
                              _p(bn.bias.detach() if bn.bias is not None else None), _p(bn.running_mean),
                              _p(bn.running_var), _p(conv_bias.detach() if conv_bias is not None else None),
                              float(bn.eps), c, _p(scale), _p(shift), _s()), "bn_fold")


def bn_finalize(stats, count, bn, scale, shift, save_mean=None, save_invstd=None, update_running=True) -> None:
    c = scale.numel()
    mom = bn.momentum if bn.momentum is not None else 0.1
    check(lib().rtsds_bn_finalize(_p(stats), float(count), _p(bn.weight.detach()), _p(bn.bias.detach()), float(bn.eps),
                                  float(mom), c, _p(bn.running_mean if update_running else None),
                                  _p(bn.running_var if update_running else None), _p(scale), _p(shift), _p(save_mean),
                                  _p(save_invstd), _s()), "bn_finalize")


def bn_finalize_apply_ptr(stats, count, bn, scale, shift, save_mean, save_invstd, xp, yp, n_pix, c, resp, act, slope, x_ld,
                          y_ld, res_ld, x_dtype, y_dtype, update_running=True) -> None:
    """bn_finalize + scale_shift_act_ptr as ONE launch (train-mode BatchNorm [+ residual] [+ activation])."""
    mom = bn.momentum if bn.momentum is not None else 0.1
    check(lib().rtsds_bn_finalize_apply(_p(stats), float(count), _p(bn.weight.detach()), _p(bn.bias.detach()), float(bn.eps),
                                        float(mom), c, _p(bn.running_mean if update_running else None),
                                        _p(bn.running_var if update_running else None), _p(scale), _p(shift), _p(save_mean),
                                        _p(save_invstd), _p(xp), _p(resp), n_pix, x_ld, res_ld, y_ld, act, slope, x_dtype,
                                        y_dtype, _p(yp), _s()), "bn_finalize_apply")


def scale_shift_act(x, y, n_pix, c, scale=None, shift=None, residual=None, act=ACT_NONE, slope=0.0, x_ld=None,
                    y_ld=None, res_ld=None) -> None:
    x_ld = c if x_ld is None else x_ld
    y_ld = c if y_ld is None else y_ld
    res_ld = c if res_ld is None else res_ld
    check(lib().rtsds_scale_shift_act(_p(x), _p(scale), _p(shift), _p(residual), n_pix, c, x_ld, res_ld, y_ld, act,
                                      slope, dtype_code(x.dtype), dtype_code(y.dtype), _p(y), _s()), "scale_shift_act")


def scale_shift_act_ptr(xp, yp, n_pix, c, scale, shift, resp, act, slope, x_ld, y_ld, res_ld, x_dtype, y_dtype) -> None:
    """Pointer-level variant (views into larger buffers)."""
    check(lib().rtsds_scale_shift_act(_p(xp), _p(scale), _p(shift), _p(resp), n_pix, c, x_ld, res_ld, y_ld, act, slope,
                                      x_dtype, y_dtype, _p(yp), _s()), "scale_shift_act")


# ----------------------------------------------------------------------------- BiSeNet glue
def global_avgpool(x, n, hw, c, ld, out, dtype=None) -> None:
    dt = dtype_code(x.dtype) if dtype is None else dtype
    check(lib().rtsds_global_avgpool(_p(x), n, hw, c, ld, dt, _p(out), _s()), "global_avgpool")


def arm_gate(pooled, conv, bn, train, n, c, gate, mul=None, lin_out=None, xhat_out=None) -> None:
    mom = bn.momentum if bn.momentum is not None else 0.1
    check(lib().rtsds_arm_gate(_p(pooled), _p(conv.weight.detach()), _p(conv.bias.detach() if conv.bias is not None else None),
                               _p(bn.weight.detach()), _p(bn.bias.detach()), _p(bn.running_mean), _p(bn.running_var),
                               float(bn.eps), float(mom), int(train), n, c, _p(mul), _p(gate), _p(lin_out),
                               _p(xhat_out), _s()), "arm_gate")


def gate_resize_nhwc(src, n, h, w, c, src_ld, gate, oh, ow, dst, dst_ld, dst_coff, dtype, gate_scale=1.0) -> None:
    check(lib().rtsds_gate_resize_nhwc(_p(src), n, h, w, c, src_ld, _p(gate), float(gate_scale), oh, ow, _p(dst), dst_ld,
                                       dst_coff, dtype, _s()), "gate_resize_nhwc")


def scale_packed_channels(wpk, rows, cin, c0, c1, factor) -> None:
    """Input channels [c0, c1) of a packed weight [rows][cin] *= factor (block exponent of a scaled activation slot)."""
    check(lib().rtsds_scale_packed_channels(_p(wpk), dtype_code(wpk.dtype), rows, cin, c0, c1, float(factor), _s()),
          "scale_packed_channels")


def ffm_head(f, f_dtype, f_ld, pooled, n, hw, c, conv1, conv2, final_conv, z, z_ld, attn_out=None) -> None:
    wc = final_conv.weight.detach() if final_conv is not None else None
    bc = final_conv.bias.detach() if final_conv is not None and final_conv.bias is not None else None
    check(lib().rtsds_ffm_head(_p(f), f_dtype, f_ld, _p(pooled), n, hw, c, _p(conv1.weight.detach()),
                               _p(conv1.bias.detach()), _p(conv2.weight.detach()), _p(conv2.bias.detach()), _p(wc), _p(bc),
                               _p(attn_out), _p(z), z_ld, _s()), "ffm_head")


def conv2d_tc_gap(d: ConvDesc, x, w, y, scale, shift, residual, gap_out, workspace=None) -> None:
    """conv2d_tc with AdaptiveAvgPool2d(1) of its output fused into the epilogue: gap_out fp32 [n, gap_parts(d), cout] receives
    one partial mean per CTA (deterministic; the consumer adds the parts in order)."""
    ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    check(lib().rtsds_conv2d_tc_fwd_gap(C.byref(d), _p(x), _p(w), _p(scale), _p(shift), _p(residual), _p(y), _p(gap_out),
                                        _p(workspace), ws_bytes, _s()), "conv2d_tc_fwd_gap")


def conv2d_tc_gap_parts(d: ConvDesc) -> int:
    return int(lib().rtsds_conv2d_tc_gap_parts(C.byref(d)))


def arm_side(src, pooled, arm, h, w, c, dst_coff, mul_pooled=False, out_scale=1.0, parts=1):
    """RtsdsArmSide of one AttentionRefinementModule in eval mode (folded BatchNorm from the running statistics)."""
    a = _lib.ArmSide()
    bn, conv = arm.bn, arm.conv
    a.src, a.pooled = _p(src), _p(pooled)
    a.w, a.b = _p(conv.weight.detach()), _p(conv.bias.detach() if conv.bias is not None else None)
    a.gamma, a.beta = _p(bn.weight.detach()), _p(bn.bias.detach())
    a.running_mean, a.running_var = _p(bn.running_mean), _p(bn.running_var)
    a.eps, a.out_scale = float(bn.eps), float(out_scale)
    a.h, a.w_in, a.c, a.dst_coff, a.mul_pooled, a.pooled_parts = h, w, c, dst_coff, int(mul_pooled), int(parts)
    return a


def arm_gate_resize(side3, side4, dtype, n, oh, ow, dst, dst_ld) -> None:
    check(lib().rtsds_arm_gate_resize(C.byref(side3), C.byref(side4), dtype, n, oh, ow, _p(dst), dst_ld, _s()), "arm_gate_resize")


def ffm_head_resize(f, f_ld, pooled, n, h, w, c, conv1, conv2, final_conv, out, attn_out=None, parts=1) -> None:
    """FFM attention + final 1x1 conv + bilinear resize to `out` (fp32 NCHW [n,c,oh,ow]) in one kernel; pooled fp32
    [n, parts, c] partial means of f."""
    oh, ow = out.shape[-2:]
    bc = final_conv.bias.detach() if final_conv.bias is not None else None
    check(lib().rtsds_ffm_head_resize(_p(f), f_ld, _p(pooled), n, h, w, c, _p(conv1.weight.detach()), _p(conv1.bias.detach()),
                                      _p(conv2.weight.detach()), _p(conv2.bias.detach()), _p(final_conv.weight.detach()), _p(bc),
                                      _p(attn_out), int(parts), oh, ow, _p(out), _s()), "ffm_head_resize")


def resize_to_nchw(z, n, h, w, c, z_ld, out) -> None:
    oh, ow = out.shape[-2:]
    check(lib().rtsds_resize_to_nchw(_p(z), n, h, w, c, z_ld, oh, ow, _p(out), _s()), "resize_to_nchw")


# ----------------------------------------------------------------------------- loss
def resize_ce_argmax_fwd(z, n, h, w, c, z_ld, oh, ow, target, ignore_index, acc, pred_out=None) -> None:
    check(lib().rtsds_resize_ce_argmax_fwd(_p(z), n, h, w, c, z_ld, oh, ow, _p(target), int(ignore_index), _p(acc),
                                           _p(pred_out), _s()), "resize_ce_argmax_fwd")


def resize_ce_bwd(z, n, h, w, c, z_ld, oh, ow, target, ignore_index, grad_scale, dz) -> None:
    check(lib().rtsds_resize_ce_bwd(_p(z), n, h, w, c, z_ld, oh, ow, _p(target), int(ignore_index), _p(grad_scale),
                                    _p(dz), _s()), "resize_ce_bwd")


def resize_ce_fused_supported(h, w, c, oh, ow) -> bool:
    return bool(lib().rtsds_resize_ce_fused_supported(h, w, c, oh, ow))


def resize_ce_fused(z, n, h, w, c, z_ld, oh, ow, target, ignore_index, acc, pred_out, dz_unnorm) -> None:
    check(lib().rtsds_resize_ce_fused(_p(z), n, h, w, c, z_ld, oh, ow, _p(target), int(ignore_index), _p(acc), _p(pred_out),
                                      _p(dz_unnorm), _s()), "resize_ce_fused")


def scale_by_device_scalar(x, scale) -> None:
    check(lib().rtsds_scale_by_device_scalar(_p(x), x.numel(), _p(scale), _s()), "scale_by_device_scalar")


def ce_argmax_nchw_fwd(logits, target, ignore_index, acc, pred_out=None) -> None:
    n, c, h, w = logits.shape
    check(lib().rtsds_ce_argmax_nchw_fwd(_p(logits), n, c, h * w, _p(target), int(ignore_index), _p(acc), _p(pred_out),
                                         _s()), "ce_argmax_nchw_fwd")


# ----------------------------------------------------------------------------- space-to-depth stems
def stem_s2d_shape(n, h, w):
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return oh, ow, (n, oh + 3, ow + 3, 16)


def stem_s2d_pack(x: torch.Tensor, P: torch.Tensor) -> None:
    n, _, h, w = x.shape
    check(lib().rtsds_stem_s2d_pack(_p(x), n, h, w, _p(P), _s()), "stem_s2d_pack")


def stem_s2d_pack_ex(x: torch.Tensor, P: torch.Tensor, scale3=None, bias3=None) -> None:
    """x fp32 or RAW uint8 NCHW -> P (bf16 / fp16) = scale3*x + bias3 inside the image (ctypes float[3], None = 1 / 0)."""
    n, _, h, w = x.shape
    check(lib().rtsds_stem_s2d_pack_ex(_p(x), int(x.dtype == torch.uint8), scale3, bias3, n, h, w, dtype_code(P.dtype), _p(P), _s()),
          "stem_s2d_pack")


def stem_s2d_conv_fwd_dt(P, n, oh, ow, wpk, cout, y, out_ld, out_dtype, scale=None, shift=None, act=ACT_NONE) -> None:
    check(lib().rtsds_stem_s2d_conv_fwd_dt(_p(P), dtype_code(P.dtype), n, oh, ow, _p(wpk), cout, _p(scale), _p(shift), act, _p(y),
                                           out_ld, out_dtype, _s()), "stem_s2d_conv_fwd")


def maxpool3x3s2_ld(x_ptr, n, h, w, c, x_ld, dtype, y: torch.Tensor, ceil_mode: bool = False) -> None:
    """MaxPool2d(3,2,1) of a channel slice (pixel pitch x_ld) of a wider NHWC buffer."""
    check(lib().rtsds_maxpool3x3s2_fwd_ld(_p(x_ptr), n, h, w, c, x_ld, dtype, int(ceil_mode), _p(y), _s()), "maxpool_ld")


def stem_s2d_weight(w: torch.Tensor, w2: torch.Tensor) -> None:
    co, _, k, _ = w.shape
    check(lib().rtsds_stem_s2d_weight(_p(w.detach()), co, k, k // 2, _p(w2), _s()), "stem_s2d_weight")


def stem_s2d_weight_grad(g2: torch.Tensor, grad: torch.Tensor) -> None:
    co, _, k, _ = grad.shape
    check(lib().rtsds_stem_s2d_weight_grad(_p(g2), co, k, k // 2, _p(grad), _s()), "stem_s2d_weight_grad")


def stem_s2d_conv_fwd(P, n, oh, ow, wpk, cout, y, out_ld, out_dtype, scale=None, shift=None, act=ACT_NONE, stats=None) -> None:
    check(lib().rtsds_stem_s2d_conv_fwd(_p(P), n, oh, ow, _p(wpk), cout, _p(scale), _p(shift), act, _p(stats), _p(y), out_ld,
                                        out_dtype, _s()), "stem_s2d_conv_fwd")


def stem_s2d_conv_wgrad(P, n, oh, ow, dy, dy_ld, cout, dw_packed) -> None:
    check(lib().rtsds_stem_s2d_conv_wgrad(_p(P), n, oh, ow, _p(dy), dy_ld, cout, _p(dw_packed), _s()), "stem_s2d_conv_wgrad")


# ----------------------------------------------------------------------------- adaptive average pool
class _AdaptiveAvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow):
        x = x.contiguous()
        n, c, h, w = x.shape
        y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        check(lib().rtsds_adaptive_avgpool_nchw_fwd(_p(x), n * c, h, w, oh, ow, _p(y), _s()), "adaptive_avgpool_fwd")
        ctx.shape = (n, c, h, w, oh, ow)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, oh, ow = ctx.shape
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dy.device)
        check(lib().rtsds_adaptive_avgpool_nchw_bwd(_p(dy.contiguous()), n * c, h, w, oh, ow, _p(dx), _s()), "adaptive_avgpool_bwd")
        return dx, None, None


def adaptive_avg_pool2d(x: torch.Tensor, output_size) -> torch.Tensor:
    """F.adaptive_avg_pool2d(x, output_size) for fp32 NCHW CUDA tensors (train.py:410,438,445), differentiable."""
    _cuda(x)
    if x.dtype != torch.float32 or x.dim() != 4:
        raise TypeError("adaptive_avg_pool2d: fp32 [N,C,H,W] expected")
    oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
    return _AdaptiveAvgPool.apply(x, int(oh), int(ow))
