""""Taps as N" form of a k x k stride-1 convolution with few output channels (csrc/tapn.cu): the FeatureFusionModule's
3x3 1024 -> num_classes conv (reference models/bisenet/build_bisenet.py:64,74).  One 1x1 tcgen05 GEMM with
N = k*k*cout reads the wide input once; a gather sums the k*k shifted planes (and produces the train-mode BatchNorm
statistics); backward builds the shifted copies of the (narrow) output gradient once and runs 1x1 dgrad / wgrad GEMMs."""
from __future__ import annotations

import torch

from . import ops
from .ops import ACT_NONE, BF16, F32, _p, check, lib


def applicable(conv) -> bool:
    k = conv.kernel_size[0]
    return (k > 1 and conv.stride[0] == 1 and conv.weight.shape[0] <= 32 and conv.weight.shape[1] >= 256
            and conv.weight.shape[1] % 64 == 0 and conv.kernel_size[0] == conv.kernel_size[1])


class TapNConv:
    def __init__(self, plan, conv, x_ptr, xshape, in_ld, train: bool, in_scale=None):
        self.plan, self.conv, self.x_ptr = plan, conv, x_ptr
        self.in_scale = in_scale        # (c0, c1, factor): input channels [c0, c1) are stored divided by factor
        assert in_scale is None or not train
        n, h, w, cin = xshape
        self.n, self.h, self.w, self.cin, self.in_ld = n, h, w, cin, in_ld
        self.c = conv.weight.shape[0]
        self.k, self.pad, self.dil = conv.kernel_size[0], conv.padding[0], conv.dilation[0]
        assert 2 * self.pad == self.dil * (self.k - 1), "taps-as-N needs a 'same' convolution"
        self.nt = self.k * self.k * self.c
        self.t_ld = (self.nt + 15) // 16 * 16
        self.kpad = (self.nt + 63) // 64 * 64
        dt, f32 = plan.dt, torch.float32
        self.dt, self.tc = dt, plan.use_tc
        self.w_fwd = plan.buf(self.nt, cin, 1, 1, dtype=f32)
        self.wpk_fwd = plan.buf(ops.cout_pad(self.nt), 1, cin)
        self.T = plan.buf(n, h, w, self.t_ld, dtype=f32)
        self.d_fwd = ops.make_conv_desc(n, h, w, cin, in_ld, self.nt, self.t_ld, 1, 1, 0, 1, in_dtype=dt, out_dtype=F32)
        self.w_bwd = plan.buf(cin, self.kpad, 1, 1, dtype=f32) if train else None
        if train:
            self.wpk_bwd = plan.buf(ops.cout_pad(cin), 1, self.kpad)
            self.G = plan.zeros(n, h, w, self.kpad)
            self.dw2 = plan.zeros(self.nt, cin, dtype=f32)
            self.d_wg = ops.make_conv_desc(n, h, w, cin, in_ld, self.nt, self.kpad, 1, 1, 0, 1, in_dtype=dt, out_dtype=dt)
        plan.pack_steps.append(self._pack)
        if self.tc:
            plan.note_ws(int(lib().rtsds_conv2d_tc_workspace_bytes(self.d_fwd)))
            if train:
                d = ops.make_conv_desc(n, h, w, self.kpad, self.kpad, cin, cin, 1, 1, 0, 1, in_dtype=dt, out_dtype=dt)
                plan.note_ws(int(lib().rtsds_conv2d_tc_workspace_bytes(d)))

    def _pack(self):
        check(lib().rtsds_tapn_weights(_p(self.conv.weight.detach()), self.c, self.cin, self.k, self.kpad, _p(self.w_fwd),
                                       _p(self.w_bwd), ops._s()), "tapn_weights")
        ops.pack_conv_weight(self.w_fwd, self.dt, self.wpk_fwd)
        if self.in_scale is not None:
            c0, c1, fac = self.in_scale
            ops.scale_packed_channels(self.wpk_fwd, self.wpk_fwd.shape[0], self.cin, c0, c1, fac)
        if self.w_bwd is not None:
            ops.pack_conv_weight(self.w_bwd, self.dt, self.wpk_bwd)

    def forward(self, scale, shift, act, stats, y_ptr, y_ld, gap_out=None):
        """y (fp32 NHWC, pitch y_ld) = act(scale * conv(x) + shift); stats: train-mode sum / sum of squares of conv(x);
        gap_out: fp32 [n,c] += mean over h*w of y (AdaptiveAvgPool2d(1) fused into the gather; caller zeroes it)."""
        if self.tc:
            ops.conv2d_tc(self.d_fwd, self.x_ptr, self.wpk_fwd, self.T, None, None, None, None, self.plan.ws)
        else:
            ops.conv2d_simt(self.d_fwd, self.x_ptr, self.wpk_fwd, self.T, None, None, None, None)
        check(lib().rtsds_tapn_gather(_p(self.T), self.t_ld, self.n, self.h, self.w, self.c, self.k, self.pad, self.dil, _p(scale),
                                      _p(shift), act, _p(stats), _p(y_ptr), y_ld, _p(gap_out), ops._s()), "tapn_gather")

    def scatter(self, dy_ptr, dy_ld, dy_dtype):
        check(lib().rtsds_tapn_scatter(_p(dy_ptr), dy_ld, dy_dtype, self.n, self.h, self.w, self.c, self.k, self.pad, self.dil,
                                       _p(self.G), self.kpad, self.dt, ops._s()), "tapn_scatter")

    def weight_grad(self, gwt):
        """gwt [c, cin, k, k] += dW (call after scatter)."""
        ops.conv2d_wgrad(self.d_wg, self.x_ptr, self.G, self.dw2, self.tc)
        check(lib().rtsds_tapn_weight_grad(_p(self.dw2), self.c, self.cin, self.k, _p(gwt), ops._s()), "tapn_weight_grad")
        self.dw2.zero_()

    def input_grad(self, dx_ptr, dx_ld, dx_dtype, accumulate):
        """dx [n,h,w,cin] (=|+=) dL/dx (call after scatter)."""
        d = ops.make_conv_desc(self.n, self.h, self.w, self.kpad, self.kpad, self.cin, dx_ld, 1, 1, 0, 1, in_dtype=self.dt,
                               out_dtype=dx_dtype, res_ld=dx_ld if accumulate else 0)
        res = dx_ptr if accumulate else None
        if self.tc:
            ops.conv2d_tc(d, self.G, self.wpk_bwd, dx_ptr, None, None, res, None, self.plan.ws)
        else:
            ops.conv2d_simt(d, self.G, self.wpk_bwd, dx_ptr, None, None, res, None)
