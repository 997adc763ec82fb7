"""Stand-alone forwards of the reference's SUB-modules (the whole-model forwards in bisenet_engine / deeplab_engine never
go through here): models/bisenet/build_bisenet.py ConvBlock (:16-18), Spatial_path (:28-32), AttentionRefinementModule
(:44-53), FeatureFusionModule (:71-81); the context path and its ResNet blocks (build_contextpath.py:18-29, torchvision
BasicBlock / Bottleneck); models/deeplabv2/deeplabv2.py Bottleneck (:30-47), ClassifierModule (:62-66);
models/domain_shift/adversarial/model.py UpSampler (:25-28).

Each call converts the NCHW fp32 argument(s) to the path's NHWC layout, runs the same sm_100a kernels the fused plans
use (tcgen05 convs when cin % 64 == 0, the CUDA-core conv otherwise), and converts the result back.  BatchNorm follows
`module.training` (batch statistics + running-buffer update, or folded running statistics).  These are forward-only
entry points: gradients flow through the whole-model forwards (BiSeNet / ResNetMulti / the discriminators), so a call
that would need autograd raises instead of silently returning a detached tensor.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU, BF16, F32, _p, check, lib


def _precision(mod):
    return getattr(mod, "rtsds_precision", "bf16")


class _Ctx:
    """Per-call scratch: dtype, device, and the tensors that must outlive the asynchronous launches."""

    def __init__(self, mod, *inputs):
        for x in inputs:
            if not x.is_cuda and not ops._lib.dry_run():
                raise ops._lib.RtsdsError("rtsds_b200 sub-module forward needs CUDA tensors (there is no CPU fallback)")
        if torch.is_grad_enabled() and (any(x.requires_grad for x in inputs) or any(p.requires_grad for p in mod.parameters())):
            raise ops._lib.RtsdsError(
                f"{type(mod).__name__}: stand-alone sub-module forwards are inference-only; gradients flow through the "
                "whole-model forward (BiSeNet / ResNetMulti / discriminator).  Wrap the call in torch.no_grad().")
        check(lib().rtsds_check_device(), "device check")
        self.dt = F32 if _precision(mod) == "fp32" else BF16
        self.tdt = ops.torch_dtype(self.dt)
        self.dev = inputs[0].device
        self.ws = None

    def to_nhwc(self, x, out=None, c_off=0):
        x = x.float().contiguous()
        n, c, h, w = x.shape
        if out is None:
            ld = (c + 7) // 8 * 8
            out = torch.zeros((n, h, w, ld), dtype=self.tdt, device=self.dev)
        check(lib().rtsds_nchw_to_nhwc(_p(x), n, c, h * w, ops.dtype_code(out.dtype), _p(out), out.shape[-1], c_off, ops._s()),
              "nchw_to_nhwc")
        return out

    def to_nchw(self, t, c, c_off=0):
        n, h, w, ld = t.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=self.dev)
        check(lib().rtsds_nhwc_to_nchw(_p(t), ops.dtype_code(t.dtype), ld, c_off, n, c, h * w, _p(out), ops._s()), "nhwc_to_nchw")
        return out

    def conv(self, conv, bn, x, cin, act, residual=None, train=False, out_dtype=None, out_ld=None):
        """x NHWC [n,h,w,ld] (first cin channels valid) -> act(BN(conv(x)) + residual) NHWC [n,oh,ow,out_ld]."""
        n, h, w, ld = x.shape
        cout, k = conv.weight.shape[0], conv.kernel_size[0]
        out_dtype = self.dt if out_dtype is None else out_dtype
        if out_ld is None:
            out_ld = (cout + 7) // 8 * 8
        tc = self.dt == BF16 and cin % 64 == 0 and ld % 8 == 0
        d = ops.make_conv_desc(n, h, w, cin, ld, cout, out_ld, k, conv.stride[0], conv.padding[0], conv.dilation[0], act=ACT_NONE,
                               in_dtype=self.dt, out_dtype=out_dtype, res_ld=residual.shape[-1] if residual is not None else 0)
        y = torch.zeros((n, d.oh, d.ow, out_ld), dtype=ops.torch_dtype(out_dtype), device=self.dev)
        wpk = ops.pack_conv_weight(conv.weight, self.dt)
        f32 = torch.float32
        scale = shift = stats = None
        if bn is not None and not train:
            scale, shift = torch.empty(cout, dtype=f32, device=self.dev), torch.empty(cout, dtype=f32, device=self.dev)
            ops.bn_fold(bn, scale, shift, conv.bias)
        elif bn is None and conv.bias is not None:
            shift = conv.bias.detach()
        fused_epilogue = bn is None or not train
        if fused_epilogue:
            d.act = act
        else:
            stats = torch.zeros(2 * cout, dtype=f32, device=self.dev)
        res = residual if fused_epilogue else None
        if tc:
            ws = torch.empty(max(int(lib().rtsds_conv2d_tc_workspace_bytes(d)), 16), dtype=torch.uint8, device=self.dev)
            ops.conv2d_tc(d, x, wpk, y, scale, shift, res, stats, ws)
        else:
            ops.conv2d_simt(d, x, wpk, y, scale, shift, res, stats)
        if not fused_epilogue:                       # train-mode BatchNorm: batch statistics, running buffers updated
            n_pix = n * d.oh * d.ow
            scale, shift = torch.empty(cout, dtype=f32, device=self.dev), torch.empty(cout, dtype=f32, device=self.dev)
            ops.bn_finalize(stats, n_pix, bn, scale, shift)
            ops.scale_shift_act_ptr(y, y, n_pix, cout, scale, shift, residual, act, 0.0, out_ld, out_ld,
                                    residual.shape[-1] if residual is not None else cout, out_dtype, out_dtype)
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        return y


    def stem(self, conv, bn, x, act, train):
        """Few-channel conv (cin <= 32, e.g. the 3-channel image) straight from the NCHW fp32 tensor: direct kernel."""
        x = x.float().contiguous()
        n, cin, h, w = x.shape
        cout, k, st, pad = conv.weight.shape[0], conv.kernel_size[0], conv.stride[0], conv.padding[0]
        y = torch.empty((n, ops.conv_out_size(h, k, st, pad), ops.conv_out_size(w, k, st, pad), cout), dtype=self.tdt, device=self.dev)
        f32 = torch.float32
        scale, shift = torch.empty(cout, dtype=f32, device=self.dev), torch.empty(cout, dtype=f32, device=self.dev)
        if not train:
            ops.bn_fold(bn, scale, shift, conv.bias)
            ops.stem_conv(x, conv.weight, y, k, st, pad, scale, shift, act)
        else:
            stats = torch.zeros(2 * cout, dtype=f32, device=self.dev)
            ops.stem_conv(x, conv.weight, y, k, st, pad, stats=stats)
            n_pix = y.shape[0] * y.shape[1] * y.shape[2]
            ops.bn_finalize(stats, n_pix, bn, scale, shift)
            ops.scale_shift_act(y, y, n_pix, cout, scale, shift, None, act)
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        return y

    def conv_any(self, conv, bn, x_nchw, t, cin, act, train):
        """First conv of a chain: from the NCHW tensor when it has few channels, else from its NHWC copy."""
        if cin <= 32 and bn is not None and conv.dilation[0] == 1:
            return self.stem(conv, bn, x_nchw, act, train)
        t = self.to_nhwc(x_nchw) if t is None else t
        return self.conv(conv, bn, t, cin, act, train=train)


# ----------------------------------------------------------------------------- BiSeNet sub-modules
def convblock_forward(mod, x):
    c = _Ctx(mod, x)
    y = c.conv_any(mod.conv1, mod.bn, x, None, x.shape[1], ACT_RELU, mod.training)
    return c.to_nchw(y, mod.conv1.weight.shape[0])


def spatial_path_forward(mod, x):
    c = _Ctx(mod, x)
    t = c.conv_any(mod.convblock1.conv1, mod.convblock1.bn, x, None, x.shape[1], ACT_RELU, mod.training)
    cin = mod.convblock1.conv1.weight.shape[0]
    for blk in (mod.convblock2, mod.convblock3):
        t = c.conv(blk.conv1, blk.bn, t, cin, ACT_RELU, train=mod.training)
        cin = blk.conv1.weight.shape[0]
    return c.to_nchw(t, cin)


def arm_forward(mod, x):
    c = _Ctx(mod, x)
    n, ch, h, w = x.shape
    t = c.to_nhwc(x)
    f32 = torch.float32
    pooled = torch.empty(n, ch, dtype=f32, device=c.dev)
    gate = torch.empty(n, ch, dtype=f32, device=c.dev)
    ops.global_avgpool(t, n, h * w, ch, t.shape[-1], pooled)
    ops.arm_gate(pooled, mod.conv, mod.bn, mod.training, n, ch, gate)
    if mod.training and mod.bn.num_batches_tracked is not None:
        mod.bn.num_batches_tracked += 1
    y = torch.empty_like(t)
    ops.gate_resize_nhwc(t, n, h, w, ch, t.shape[-1], gate, h, w, y, y.shape[-1], 0, c.dt)     # same size: x * gate
    return c.to_nchw(y, ch)


def ffm_forward(mod, input_1, input_2):
    c = _Ctx(mod, input_1, input_2)
    n, c1, h, w = input_1.shape
    c2 = input_2.shape[1]
    cat = torch.zeros((n, h, w, (c1 + c2 + 7) // 8 * 8), dtype=c.tdt, device=c.dev)
    c.to_nhwc(input_1, cat, 0)
    c.to_nhwc(input_2, cat, c1)                                                              # torch.cat((sx, cx), 1)
    nc = mod.convblock.conv1.weight.shape[0]
    if nc > 32:
        raise ops._lib.RtsdsError("num_classes > 32 is not supported by the fused head kernels")
    feat = c.conv(mod.convblock.conv1, mod.convblock.bn, cat, c1 + c2, ACT_RELU, train=mod.training, out_dtype=F32, out_ld=32)
    f32 = torch.float32
    pooled = torch.empty(n, nc, dtype=f32, device=c.dev)
    ops.global_avgpool(feat, n, h * w, nc, 32, pooled)
    z = torch.zeros((n, h, w, 32), dtype=f32, device=c.dev)
    ops.ffm_head(feat, F32, 32, pooled, n, h * w, nc, mod.conv1, mod.conv2, None, z, 32)       # f*a + f
    return c.to_nchw(z, nc)


def basic_block_forward(mod, x):
    c = _Ctx(mod, x)
    t = c.to_nhwc(x)
    y, cout = _basic_block(c, mod, t, x.shape[1], mod.training)
    return c.to_nchw(y, cout)


def _basic_block(c, mod, t, cin, train):
    cout = mod.conv1.weight.shape[0]
    a = c.conv(mod.conv1, mod.bn1, t, cin, ACT_RELU, train=train)
    res = c.conv(mod.downsample[0], mod.downsample[1], t, cin, ACT_NONE, train=train) if mod.downsample is not None else t
    return c.conv(mod.conv2, mod.bn2, a, cout, ACT_RELU, residual=res, train=train), cout


def _bottleneck(c, mod, t, cin, train):
    planes = mod.conv1.weight.shape[0]
    a = c.conv(mod.conv1, mod.bn1, t, cin, ACT_RELU, train=train)
    b = c.conv(mod.conv2, mod.bn2, a, planes, ACT_RELU, train=train)
    res = c.conv(mod.downsample[0], mod.downsample[1], t, cin, ACT_NONE, train=train) if mod.downsample is not None else t
    return c.conv(mod.conv3, mod.bn3, b, planes, ACT_RELU, residual=res, train=train), planes * 4


def bottleneck_forward(mod, x):
    """torchvision-style Bottleneck (stride on conv2) and the DeepLabV2 one (stride on conv1, dilated conv2): both are
    conv1-bn-relu, conv2-bn-relu, conv3-bn, (+downsample), add, relu; the conv modules carry their own strides."""
    c = _Ctx(mod, x)
    y, cout = _bottleneck(c, mod, c.to_nhwc(x), x.shape[1], mod.training)
    return c.to_nchw(y, cout)


def context_path_forward(mod, x):
    """-> (feature3 @1/16, feature4 @1/32, tail = mean over H, W of feature4) — build_contextpath.py:18-29."""
    c = _Ctx(mod, x)
    train = mod.training
    t = c.conv_any(mod.conv1, mod.bn1, x, None, x.shape[1], ACT_RELU, train)
    n, h, w, ch = t.shape
    p = torch.empty((n, ops.maxpool_out_size(h), ops.maxpool_out_size(w), ch), dtype=t.dtype, device=c.dev)
    ops.maxpool3x3s2(t, p)
    t, cin = p, ch
    feats = []
    for layer in (mod.layer1, mod.layer2, mod.layer3, mod.layer4):
        for blk in layer:
            t, cin = (_basic_block if hasattr(blk, "conv3") is False else _bottleneck)(c, blk, t, cin, train)
        feats.append((t, cin))
    (f3, c3), (f4, c4) = feats[2], feats[3]
    tail = torch.empty(f4.shape[0], c4, dtype=torch.float32, device=c.dev)
    ops.global_avgpool(f4, f4.shape[0], f4.shape[1] * f4.shape[2], c4, f4.shape[-1], tail)
    return c.to_nchw(f3, c3), c.to_nchw(f4, c4), tail.view(f4.shape[0], c4, 1, 1)


# ----------------------------------------------------------------------------- DeepLabV2 / discriminator helpers
def classifier_forward(mod, x):
    """ClassifierModule (deeplabv2.py:62-66): sum of the dilated 3x3 branches, chained through the residual epilogue."""
    c = _Ctx(mod, x)
    t, cin = c.to_nhwc(x), x.shape[1]
    nc = mod.conv2d_list[0].weight.shape[0]
    z = None
    for conv in mod.conv2d_list:
        z = c.conv(conv, None, t, cin, ACT_NONE, residual=z, out_dtype=F32, out_ld=32)
    return c.to_nchw(z, nc)


def upsampler_forward(mod, x):
    """UpSampler (model.py:25-28): x8 bilinear then 1x1 conv with bias == 1x1 conv then x8 bilinear (they commute)."""
    c = _Ctx(mod, x)
    n, ch, h, w = x.shape
    t = c.to_nhwc(x)
    z = c.conv(mod.conv, None, t, ch, ACT_NONE, out_dtype=F32, out_ld=32)
    out = torch.empty((n, mod.conv.weight.shape[0], h * 8, w * 8), dtype=torch.float32, device=c.dev)
    ops.resize_to_nchw(z, n, h, w, mod.conv.weight.shape[0], 32, out)
    return out
