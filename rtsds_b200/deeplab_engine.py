"""Execution plans of DeepLabV2-ResNet101 (reference models/deeplabv2/deeplabv2.py:69-131,
driven by train.py:77-96 and validation.py:45).

Same building blocks as the BiSeNet path: NHWC activations (bf16, or fp32 in the check mode), every
convolution a tcgen05 implicit GEMM (1x1 = one tap; dilated 3x3 = nine taps whose TMA boxes are offset by
r*dilation with out-of-bounds zero fill as the padding; the stride-2 1x1 of layer2 reads the (even, even)
parity view), BatchNorm folded into the conv epilogue in eval mode and computed from conv-epilogue
statistics in train mode (batch statistics, running buffers updated; the affine parameters are frozen by
the reference, deeplabv2.py:15-27, so they get no gradient), hand-written backward.

Data layout in HBM:
  image      NCHW fp32 [N,3,H,W]                   API boundary (direct 7x7 s2 stem kernel reads it)
  stem       NHWC [N,H/2,W/2,64] -> ceil-mode max-pool -> [N,~H/4,~W/4,64]
  layer1..4  NHWC, 33 bottlenecks x (t1[planes], t2[planes], y[4*planes]) (+ raw pre-BN copies in train mode)
  z          NHWC fp32 [N,hf,wf,32]                ASPP logits (19 valid channels): the 4 dilated branches are
                                                   chained through the conv epilogue's residual input
  logits     NCHW fp32 [N,19,H,W]                  API boundary (generic-scale bilinear, 65x129 -> 512x1024)
"""
from __future__ import annotations

import torch

from . import ops, weights_epoch
from .bisenet_train import _Buf, _ConvBN, _s
from .ops import ACT_NONE, ACT_RELU, BF16, F32, _p, check, lib

DEEPLAB_GROUPS = (("conv1", "layer1"), ("bn1", "layer1"), ("layer1", "layer1"), ("layer2", "layer2"), ("layer3", "layer3"),
                  ("layer4", "layer4"), ("layer6", "head"))


class _PlanBase:
    def _init_base(self, model, n, h, w, precision, train):
        p0 = model.conv1.weight
        if not p0.is_cuda and not ops._lib.dry_run():
            raise ops._lib.RtsdsError("DeepLabV2 parameters must live on a CUDA device (no CPU fallback)")
        check(lib().rtsds_check_device(), "device check")
        self.model, self.device = model, p0.device
        self.n, self.h, self.w = n, h, w
        self.train = train
        self.dt = F32 if precision == "fp32" else BF16
        self.use_tc = precision == "bf16"
        self.tdt = ops.torch_dtype(self.dt)
        self.nc = model.layer6.conv2d_list[0].weight.shape[0]
        if self.nc > 32:
            raise ops._lib.RtsdsError("num_classes > 32 is not supported by the head kernels")
        self._keep, self.pack_steps = [], []
        self.stem_P = None
        if train:
            self.pack_jobs, self.pending_unpack = [], []
            weights_epoch.register_plan(self)    # the fused optimizers refresh this plan's packed operands (optim.py)
        self._stats_total = 0
        self._scratch_act, self._scratch_w, self._ws_bytes = 0, 0, 0
        self._param_version = None
        self.generation = 0
        self.ws = None

    def buf(self, *shape, dtype=None):
        t = torch.empty(shape, dtype=self.tdt if dtype is None else dtype, device=self.device)
        self._keep.append(t)
        return t

    def zeros(self, *shape, dtype=None):
        t = torch.zeros(shape, dtype=self.tdt if dtype is None else dtype, device=self.device)
        self._keep.append(t)
        return t

    def alloc_stats(self, c):
        off = self._stats_total
        self._stats_total += 2 * c
        return off

    def stats_view(self, off, c):
        return self.stats_all[off:off + 2 * c]

    def note_scratch(self, act_elems, w_elems):
        self._scratch_act = max(self._scratch_act, act_elems)
        self._scratch_w = max(self._scratch_w, w_elems)

    def note_ws(self, b):
        self._ws_bytes = max(self._ws_bytes, b)

    def d_raw_view(self, ld):
        return _Buf(self.d_raw_scratch, ld=ld, dtype=self.dt)

    def dw_view(self, numel):
        return self.dw_scratch[:numel]

    def _params_version(self):
        v = 0
        for p in self.model.parameters():
            v += p._version
        if not self.train:
            for b in self.model.buffers():
                v += b._version
        return (v, self.model.conv1.weight.data_ptr(), weights_epoch.value())

    def refresh_weights(self):
        ver = self._params_version()
        if ver != self._param_version:
            for s in self.pack_steps:
                s()
            if self.train:
                ops.pack_conv_weights_batch([(c.weight, out, kind) for c, out, kind in self.pack_jobs], self.dt, self.use_tc)
            self._param_version = self._params_version()

    def flush_unpack(self):
        if self.pending_unpack:
            ops.unpack_conv_wgrads_batch(self.pending_unpack)
            self.pending_unpack = []

    def out_hw(self):
        h2, w2 = ops.conv_out_size(self.h, 7, 2, 3), ops.conv_out_size(self.w, 7, 2, 3)
        return h2, w2, ops.maxpool_out_size(h2, True), ops.maxpool_out_size(w2, True)


# ======================================================================================= eval
class DeepLabPlan(_PlanBase):
    """Eval-mode forward: BatchNorm folded into the conv epilogues, everything after the stem in one CUDA graph."""

    def __init__(self, model, n, h, w, precision="bf16"):
        self._init_base(model, n, h, w, precision, False)
        self.pre_steps, self.steps = [], []
        self.graph = None
        self._build()

    def _conv(self, conv, bn, x, xshape, y, act, residual=None, out_dtype=None, bias=None, out_ld=None, res_ld=0):
        n, h, w, cin = xshape
        cout = conv.weight.shape[0]
        k = conv.kernel_size[0]
        out_dtype = self.dt if out_dtype is None else out_dtype
        out_ld = cout if out_ld is None else out_ld
        d = ops.make_conv_desc(n, h, w, cin, cin, cout, out_ld, k, conv.stride[0], conv.padding[0], conv.dilation[0], act=act,
                               in_dtype=self.dt, out_dtype=out_dtype, res_ld=res_ld if residual is not None else 0)
        wpk = self.buf(ops.cout_pad(cout), k * k, cin)
        self.pack_steps.append(lambda: ops.pack_conv_weight(conv.weight, self.dt, wpk))
        if self.use_tc:
            self.note_ws(int(lib().rtsds_conv2d_tc_workspace_bytes(d)))
        if bn is not None:
            scale = self.buf(cout, dtype=torch.float32)
            shift = self.buf(cout, dtype=torch.float32)
            self.pack_steps.append(lambda: ops.bn_fold(bn, scale, shift))
        else:
            scale, shift = None, None
        use_tc = self.use_tc

        def run():
            sh = shift if bn is not None else (bias.detach() if bias is not None else None)
            if use_tc:
                ops.conv2d_tc(d, x, wpk, y, scale, sh, residual, None, self.ws)
            else:
                ops.conv2d_simt(d, x, wpk, y, scale, sh, residual, None)

        self.steps.append(run)
        return d.oh, d.ow

    def _build(self):
        m, n = self.model, self.n
        h2, w2, ph, pw = self.out_hw()
        stem = self.buf(n, h2, w2, 64)
        sscale, sshift = self.buf(64, dtype=torch.float32), self.buf(64, dtype=torch.float32)
        self.pack_steps.append(lambda: ops.bn_fold(m.bn1, sscale, sshift))
        if self.use_tc:      # 7x7 s2 stem as a 4-tap implicit GEMM over the padded space-to-depth image
            oh, ow, pshape = ops.stem_s2d_shape(n, self.h, self.w)
            P = self.buf(*pshape, dtype=torch.bfloat16)
            w2 = self.buf(64, 64, 4, 1, dtype=torch.float32)
            wpk = self.buf(64, 4, 64, dtype=torch.bfloat16)
            self.pack_steps.append(lambda: (ops.stem_s2d_weight(m.conv1.weight, w2), ops.pack_conv_weight(w2, BF16, wpk)))
            self.pre_steps.append(lambda x: ops.stem_s2d_pack(x, P))
            self.steps.append(lambda: ops.stem_s2d_conv_fwd(P, n, oh, ow, wpk, 64, stem, 64, BF16, sscale, sshift, ACT_RELU))
        else:
            self.pre_steps.append(lambda x: ops.stem_conv(x, m.conv1.weight, stem, 7, 2, 3, sscale, sshift, ACT_RELU))
        x = self.buf(n, ph, pw, 64)
        self.steps.append(lambda x=x: ops.maxpool3x3s2(stem, x, True))
        shape = (n, ph, pw, 64)
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                x, shape = self._bottleneck(blk, x, shape)
        if min(shape[1], shape[2]) < 1:
            raise ops._lib.RtsdsError("input too small for DeepLabV2")
        self.hf, self.wf = shape[1], shape[2]
        self.z = self.zeros(n, self.hf, self.wf, 32, dtype=torch.float32)
        for i, conv in enumerate(m.layer6.conv2d_list):
            self._conv(conv, None, x, shape, self.z, ACT_NONE, residual=self.z if i else None, out_dtype=F32, bias=conv.bias,
                       out_ld=32, res_ld=32)
        if self._ws_bytes:
            self.ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)

    def _bottleneck(self, blk, x, shape):
        n, h, w, cin = shape
        planes = blk.conv1.weight.shape[0]
        st = blk.conv1.stride[0]
        oh, ow = ops.conv_out_size(h, 1, st, 0), ops.conv_out_size(w, 1, st, 0)
        t1 = self.buf(n, oh, ow, planes)
        t2 = self.buf(n, oh, ow, planes)
        y = self.buf(n, oh, ow, planes * 4)
        self._conv(blk.conv1, blk.bn1, x, shape, t1, ACT_RELU)
        self._conv(blk.conv2, blk.bn2, t1, (n, oh, ow, planes), t2, ACT_RELU)
        if blk.downsample is not None:
            res = self.buf(n, oh, ow, planes * 4)
            self._conv(blk.downsample[0], blk.downsample[1], x, shape, res, ACT_NONE)
        else:
            res = x
        self._conv(blk.conv3, blk.bn3, t2, (n, oh, ow, planes), y, ACT_RELU, residual=res, res_ld=planes * 4)
        return y, (n, oh, ow, planes * 4)

    def forward_lowres(self, x, use_graph):
        self.refresh_weights()
        for s in self.pre_steps:
            s(x)
        if use_graph and not ops._lib.dry_run():
            if self.graph is None:
                for s in self.steps:
                    s()
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for s in self.steps:
                        s()
                self.graph = g
            self.graph.replay()
        else:
            for s in self.steps:
                s()

    def logits(self):
        out = torch.empty((self.n, self.nc, self.h, self.w), dtype=torch.float32, device=self.device)
        ops.resize_to_nchw(self.z, self.n, self.hf, self.wf, self.nc, 32, out)
        return out


# ======================================================================================= train
class _StemS2D:
    """k x k stride-2 conv on the NCHW fp32 image + BatchNorm + ReLU (no input gradient).  Tensor-core mode: a 4-tap
    implicit GEMM over the plan's padded space-to-depth image `plan.stem_P` (csrc/conv_tc.cu: rtsds_stem_s2d_*), shared by
    every stem of the plan; check mode (fp32): the direct CUDA-core kernels."""

    def __init__(self, plan, conv, bn, k, pad, n, h, w):
        self.plan, self.conv, self.bn, self.k, self.pad, self.n = plan, conv, bn, k, pad, n
        self.oh, self.ow = ops.conv_out_size(h, k, 2, pad), ops.conv_out_size(w, k, 2, pad)
        self.n_pix = n * self.oh * self.ow
        self.raw = _Buf(plan.buf(n, self.oh, self.ow, 64))
        self.y = _Buf(plan.buf(n, self.oh, self.ow, 64))
        self.scale, self.shift = plan.buf(64, dtype=torch.float32), plan.buf(64, dtype=torch.float32)
        self.save_mean, self.save_invstd = plan.buf(64, dtype=torch.float32), plan.buf(64, dtype=torch.float32)
        self.stats = plan.alloc_stats(64)
        self.sums = plan.buf(128, dtype=torch.float32)
        plan.note_scratch(self.n_pix * 64, 0)
        self.s2d = plan.use_tc
        if self.s2d:
            oh, ow, pshape = ops.stem_s2d_shape(n, h, w)
            assert (oh, ow) == (self.oh, self.ow)
            if getattr(plan, "stem_P", None) is None:
                plan.stem_P = plan.buf(*pshape, dtype=torch.bfloat16)
            f32 = torch.float32
            self.w2 = plan.buf(64, 64, 4, 1, dtype=f32)
            self.g2 = plan.buf(64, 64, 4, 1, dtype=f32)
            self.wpk = plan.buf(64, 4, 64, dtype=torch.bfloat16)
            self.dw = plan.zeros(64 * 4 * 64, dtype=f32)          # kept zero: unpack clears what it reads
            plan.pack_steps.append(self._pack)

    def _pack(self):
        ops.stem_s2d_weight(self.conv.weight, self.w2)
        ops.pack_conv_weight(self.w2, BF16, self.wpk)

    def forward(self, x):
        """plan.stem_P must hold the space-to-depth form of x in tensor-core mode (BiSeNetTrainPlan.forward packs it once)."""
        p = self.plan
        st = p.stats_view(self.stats, 64)
        if self.s2d:
            ops.stem_s2d_conv_fwd(p.stem_P, self.n, self.oh, self.ow, self.wpk, 64, self.raw.t, 64, BF16, stats=st)
        else:
            ops.stem_conv(x, self.conv.weight, self.raw.t, self.k, 2, self.pad, stats=st)
        ops.bn_finalize_apply_ptr(st, self.n_pix, self.bn, self.scale, self.shift, self.save_mean, self.save_invstd, self.raw.t,
                                  self.y.t, self.n_pix, 64, None, ACT_RELU, 0.0, 64, 64, 64, ops.dtype_code(self.raw.t.dtype),
                                  ops.dtype_code(self.y.t.dtype))

    def backward(self, x, dy: _Buf, gw):
        p = self.plan
        s = _s()
        d_raw = p.d_raw_view(64)
        check(lib().rtsds_bn_bwd_reduce_rawmask(dy.ptr, dy.ld, self.raw.ptr, 64, _p(self.save_mean), _p(self.save_invstd),
                                                _p(self.scale), _p(self.shift), self.n_pix, 64, dy.dtype, _p(self.sums), s),
              "bn_bwd_reduce_rawmask")
        check(lib().rtsds_bn_bwd_apply_rawmask(dy.ptr, dy.ld, self.raw.ptr, 64, _p(self.save_mean), _p(self.save_invstd),
                                               _p(self.bn.weight.detach()), _p(self.sums), _p(self.scale), _p(self.shift),
                                               self.n_pix, 64, dy.dtype, d_raw.ptr, 64, p.dt, None, 0,
                                               _p(gw.get(self.bn.weight)), _p(gw.get(self.bn.bias)), s), "bn_bwd_apply_rawmask")
        gwt = gw.get(self.conv.weight)
        if gwt is None:
            return
        if self.s2d:
            ops.stem_s2d_conv_wgrad(p.stem_P, self.n, self.oh, self.ow, d_raw.ptr, 64, 64, self.dw)
            ops.unpack_conv_wgrad(self.dw, self.g2, False)
            ops.stem_s2d_weight_grad(self.g2, gwt)
        else:
            n, cin, h, w = x.shape
            check(lib().rtsds_stem_conv_wgrad(_p(x), d_raw.ptr, p.dt, n, cin, h, w, 64, self.k, 2, self.pad, _p(gwt), s),
                  "stem_conv_wgrad")


class DeepLabTrainPlan(_PlanBase):
    def __init__(self, model, n, h, w, precision="bf16"):
        self._init_base(model, n, h, w, precision, True)
        self._build()

    def _build(self):
        m, n, H, W = self.model, self.n, self.h, self.w
        f32 = torch.float32
        self.stem = _StemS2D(self, m.conv1, m.bn1, 7, 3, n, H, W)
        h2, w2, ph, pw = self.out_hw()
        self.pool = _Buf(self.buf(n, ph, pw, 64))
        self.pool_shape = (n, ph, pw, 64)
        self.pool_idx = self.buf(n, ph, pw, 8, dtype=torch.int32)
        self.blocks = []
        self.layer_first = {}
        x, shape = self.pool, self.pool_shape
        max_act = self.stem.n_pix * 64
        for li, layer in enumerate((m.layer1, m.layer2, m.layer3, m.layer4), start=1):
            self.layer_first[len(self.blocks)] = f"layer{li}"
            for blk in layer:
                x, shape = self._bottleneck(blk, x, shape)
                max_act = max(max_act, shape[0] * shape[1] * shape[2] * shape[3])
        self.feat, self.fshape = x, shape
        self.hf, self.wf = shape[1], shape[2]
        if min(self.hf, self.wf) < 1:
            raise ops._lib.RtsdsError("input too small for DeepLabV2")
        self.z = self.zeros(n, self.hf, self.wf, 32, dtype=f32)
        zb = _Buf(self.z, dtype=F32)
        self.aspp = [_ConvBN(self, conv, None, x, shape, zb, False) for conv in m.layer6.conv2d_list]
        for i, u in enumerate(self.aspp):
            u.d.res_ld = 32 if i else 0
        self.dz = self.zeros(n, self.hf, self.wf, 32, dtype=f32)
        self.dzb = self.zeros(n, self.hf, self.wf, 64 if self.dt == BF16 else 32, dtype=self.tdt)
        self.gA, self.gG, self.gT1, self.gT2 = (self.buf(max_act) for _ in range(4))
        self.stats_all = torch.zeros(max(self._stats_total, 1), dtype=f32, device=self.device)
        self.d_raw_scratch = self.zeros(max(self._scratch_act, 1))
        self.dw_scratch = self.zeros(max(self._scratch_w, 1), dtype=f32)
        if self._ws_bytes:
            self.ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
        self.acc = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.gscale = torch.zeros(1, dtype=f32, device=self.device)
        self.bias_sum = self.buf(32, dtype=f32)
        self.params = list(m.parameters())
        self._grad_numel = sum(p.numel() for p in self.params)

    def _bottleneck(self, blk, x: _Buf, shape):
        n, h, w, cin = shape
        planes = blk.conv1.weight.shape[0]
        st = blk.conv1.stride[0]
        oh, ow = ops.conv_out_size(h, 1, st, 0), ops.conv_out_size(w, 1, st, 0)
        t1 = _Buf(self.buf(n, oh, ow, planes))
        t2 = _Buf(self.buf(n, oh, ow, planes))
        y = _Buf(self.buf(n, oh, ow, planes * 4))
        first = len(self.blocks) == 0
        c1 = _ConvBN(self, blk.conv1, blk.bn1, x, shape, t1, True)
        c2 = _ConvBN(self, blk.conv2, blk.bn2, t1, (n, oh, ow, planes), t2, True)
        ds = None
        if blk.downsample is not None:
            dsy = _Buf(self.buf(n, oh, ow, planes * 4))
            ds = _ConvBN(self, blk.downsample[0], blk.downsample[1], x, shape, dsy, False)
            res = dsy
        else:
            res = x
        c3 = _ConvBN(self, blk.conv3, blk.bn3, t2, (n, oh, ow, planes), y, True, residual=res)
        self.blocks.append(dict(c1=c1, c2=c2, c3=c3, ds=ds, x=x, xshape=shape, y=y, shape=(n, oh, ow, planes * 4), first=first))
        return y, (n, oh, ow, planes * 4)

    # ---------------- forward ----------------
    def forward(self, x):
        self.refresh_weights()
        self.stats_all.zero_()
        self.generation += 1
        self.x = x
        if self.use_tc:
            ops.stem_s2d_pack(x, self.stem_P)
        self.stem.forward(x)
        ops.maxpool3x3s2(self.stem.y.t, self.pool.t, True, self.pool_idx)
        for b in self.blocks:
            b["c1"].forward()
            b["c2"].forward()
            if b["ds"] is not None:
                b["ds"].forward()
            b["c3"].forward()
        for i, u in enumerate(self.aspp):
            bias = u.conv.bias.detach() if u.conv.bias is not None else None
            u._launch(u.d, None, bias, self.z.data_ptr() if i else None, None, self.z.data_ptr())

    def logits(self):
        out = torch.empty((self.n, self.nc, self.h, self.w), dtype=torch.float32, device=self.device)
        ops.resize_to_nchw(self.z, self.n, self.hf, self.wf, self.nc, 32, out)
        return out

    # ---------------- backward ----------------
    def new_grads(self):
        flat = torch.zeros(self._grad_numel, dtype=torch.float32, device=self.device)
        gw, off = {}, 0
        for p in self.params:
            if p.requires_grad:
                gw[p] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        return flat, gw

    def backward_from_dz(self, gw, ready=lambda group: None):
        """self.dz holds the gradient w.r.t. the low-resolution logits z (fp32 NHWC, pitch 32)."""
        weights_epoch.note_backward()
        n, dt, nc = self.n, self.dt, self.nc
        s = ops._s()
        user_ready = ready

        def ready(group):
            self.flush_unpack()
            user_ready(group)

        self.pending_unpack = []
        npix = n * self.hf * self.wf
        # ---- ASPP: four dilated branches share dz ----
        biases = [u.conv.bias for u in self.aspp if u.conv.bias is not None and u.conv.bias in gw]
        if biases:
            self.bias_sum.zero_()
            check(lib().rtsds_channel_sum(_p(self.dz), 32, npix, nc, F32, _p(self.bias_sum), s), "channel_sum")
            for b in biases:
                gw[b].copy_(self.bias_sum[:nc])
        ops.scale_shift_act_ptr(self.dz, self.dzb, npix, nc, None, None, None, ACT_NONE, 0.0, 32, self.dzb.shape[-1], nc, F32, dt)
        cf = self.fshape[3]
        dfeat = _Buf(self.gA, ld=cf, dtype=dt)
        for i, u in enumerate(self.aspp):
            u.backward(_Buf(self.dzb), gw, dx=dfeat, dx_accumulate=i > 0)
        ready("head")
        # ---- bottlenecks in reverse: dy lives in gA and is consumed before dx (aliasing it) is written ----
        dy = dfeat
        for bi in range(len(self.blocks) - 1, -1, -1):
            b = self.blocks[bi]
            cout, cin = b["shape"][3], b["xshape"][3]
            planes = cout // 4
            g = _Buf(self.gG, ld=cout, dtype=dt)
            dT2 = _Buf(self.gT2, ld=planes, dtype=dt)
            dT1 = _Buf(self.gT1, ld=planes, dtype=dt)
            dx = _Buf(self.gA, ld=cin, dtype=dt)
            b["c3"].backward(dy, gw, dx=dT2, dx_accumulate=False, g_out=g)
            if b["ds"] is not None:
                b["ds"].backward(g, gw, dx=dx, dx_accumulate=False)
            else:
                nn_, hh, ww, cc = b["xshape"]
                ops.scale_shift_act_ptr(g.ptr, dx.ptr, nn_ * hh * ww, cc, None, None, None, ACT_NONE, 0.0, g.ld, dx.ld, cc, g.dtype, dx.dtype)
            b["c2"].backward(dT2, gw, dx=dT1, dx_accumulate=False)
            b["c1"].backward(dT1, gw, dx=dx, dx_accumulate=True)
            dy = dx
            if bi in self.layer_first and bi > 0:
                ready(self.layer_first[bi])
        # ---- ceil-mode max-pool and the 7x7 stem ----
        dstem = _Buf(self.gT1, ld=64, dtype=dt)
        check(lib().rtsds_maxpool3x3s2_bwd_idx(self.pool_idx.data_ptr(), dy.ptr, n, self.stem.oh, self.stem.ow, 64, dt, 1, dstem.ptr, s),
              "maxpool_bwd_idx")
        self.stem.backward(self.x, dstem, gw)
        self.flush_unpack()


# ======================================================================================= autograd boundary
_MAX_TRAIN_SLOTS = 2
_stamp = 0


def _get_plan(model, x, train):
    """One plan per (shape, mode).  Train plans: a second slot only when the first still holds the activations of a
    forward whose backward has not run (same rule as rtsds_b200/bisenet_autograd.py)."""
    global _stamp
    plans = model.__dict__.setdefault("_rtsds_plans", {})
    n, _, h, w = x.shape
    base = (n, h, w, bool(train), model.rtsds_precision, x.device.index)
    if not train:
        plan = plans.get(base)
        if plan is None:
            plan = plans[base] = DeepLabPlan(model, n, h, w, model.rtsds_precision)
        return plan
    oldest = None
    for slot in range(_MAX_TRAIN_SLOTS):
        plan = plans.get(base + (slot,))
        if plan is None:
            plan = DeepLabTrainPlan(model, n, h, w, model.rtsds_precision)
            plan.awaiting_backward = False
            plans[base + (slot,)] = plan
            break
        if not plan.awaiting_backward:
            break
        if oldest is None or plan.slot_stamp < oldest.slot_stamp:
            oldest = plan
    else:
        plan = oldest
    _stamp += 1
    plan.slot_stamp = _stamp
    return plan


def _bump_bn_counters(model):
    counters = model.__dict__.get("_rtsds_bn_counters")
    if counters is None or (counters and counters[0].device != model.conv1.weight.device):
        counters = [m.num_batches_tracked for m in model.modules()
                    if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None]
        model.__dict__["_rtsds_bn_counters"] = counters
    if counters:
        torch._foreach_add_(counters, 1)


def _run_backward(plan, gw, flat):
    from . import ddp

    if getattr(plan.model, "rtsds_ddp", False) and ddp.is_distributed():
        if getattr(plan, "_buckets", None) is None:
            plan._buckets = ddp.param_buckets(list(plan.model.named_parameters()), DEEPLAB_GROUPS)
        red = ddp.BucketedAllReduce(flat, plan._buckets)
        plan.backward_from_dz(gw, red.ready)
        red.finish()
    else:
        plan.backward_from_dz(gw)


class _DeepLabTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, *params):
        plan.forward(x)
        ctx.plan, ctx.gen, ctx.params = plan, plan.generation, params
        return plan.logits()

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        if plan.generation != ctx.gen:
            raise ops._lib.RtsdsError("DeepLabV2 backward called after another forward of the same shape reused the plan's "
                                      "saved activations; call backward() before the next forward()")
        plan.awaiting_backward = False
        dout = dout.contiguous()
        check(lib().rtsds_resize_to_nchw_bwd(_p(dout), plan.n, plan.nc, plan.h, plan.w, plan.hf, plan.wf, _p(plan.dz), 32, ops._s()),
              "resize_to_nchw_bwd")
        flat, gw = plan.new_grads()
        _run_backward(plan, gw, flat)
        return (None, None) + tuple(gw.get(p) for p in ctx.params)


class _DeepLabFusedCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, target, ignore_index, *params):
        plan.forward(x)
        plan.acc.zero_()
        pred = torch.empty((plan.n, plan.h, plan.w), dtype=torch.int64, device=plan.device)
        ctx.one_pass = ops.resize_ce_fused_supported(plan.hf, plan.wf, plan.nc, plan.h, plan.w)
        if ctx.one_pass:
            plan.dz.zero_()
            ops.resize_ce_fused(plan.z, plan.n, plan.hf, plan.wf, plan.nc, 32, plan.h, plan.w, target, ignore_index, plan.acc, pred, plan.dz)
        else:
            ops.resize_ce_argmax_fwd(plan.z, plan.n, plan.hf, plan.wf, plan.nc, 32, plan.h, plan.w, target, ignore_index, plan.acc, pred)
        loss = (plan.acc[0] / plan.acc[1]).float()
        stats = plan.acc.clone()
        ctx.plan, ctx.gen, ctx.params, ctx.stats = plan, plan.generation, params, stats
        ctx.target, ctx.ignore_index = target, ignore_index
        ctx.mark_non_differentiable(pred, stats)
        return loss, pred, stats

    @staticmethod
    def backward(ctx, dloss, _dpred, _dstats):
        plan = ctx.plan
        if plan.generation != ctx.gen:
            raise ops._lib.RtsdsError("DeepLabV2 backward called after another forward reused the plan's saved activations")
        plan.awaiting_backward = False
        plan.gscale.copy_((dloss.double() / ctx.stats[1]).float().view(1))
        if ctx.one_pass:
            ops.scale_by_device_scalar(plan.dz, plan.gscale)
        else:
            plan.dz.zero_()
            ops.resize_ce_bwd(plan.z, plan.n, plan.hf, plan.wf, plan.nc, 32, plan.h, plan.w, ctx.target, ctx.ignore_index, plan.gscale, plan.dz)
        flat, gw = plan.new_grads()
        _run_backward(plan, gw, flat)
        return (None, None, None, None) + tuple(gw.get(p) for p in ctx.params)


def deeplab_fused_ce(model, x, target, ignore_index=255):
    """CrossEntropyLoss(ignore_index) of the (single) DeepLabV2 head (train.py:86) evaluated from the low-resolution
    logits without materialising [N,19,H,W].  Returns (loss, argmax [N,H,W] int64, stats float64 [4] =
    {sum of -log p, valid pixels, pixels with argmax == target, 0})."""
    if not model.training:
        raise ops._lib.RtsdsError("deeplab_fused_ce is the train-mode fast path; call model.train() first")
    if not x.is_cuda and not ops._lib.dry_run():
        raise ops._lib.RtsdsError("deeplab_fused_ce needs CUDA tensors: rtsds_b200 has no CPU fallback")
    x = x.float().contiguous()
    target = target.contiguous()
    assert target.dtype == torch.int64
    plan = _get_plan(model, x, True)
    out = _DeepLabFusedCEFn.apply(plan, x, target, int(ignore_index), *plan.params)
    plan.awaiting_backward = bool(out[0].requires_grad)
    _bump_bn_counters(model)
    return out


def deeplab_forward(model, x):
    """ResNetMulti.forward (reference :113-131): train -> (logits, None, None); eval -> logits."""
    if not x.is_cuda and not ops._lib.dry_run():
        raise ops._lib.RtsdsError("DeepLabV2 forward needs a CUDA tensor: rtsds_b200 has no CPU fallback")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected input [N,3,H,W], got {tuple(x.shape)}")
    x = x.float().contiguous()
    if model.training:
        plan = _get_plan(model, x, True)
        if torch.is_grad_enabled():
            out = _DeepLabTrainFn.apply(plan, x, *plan.params)
            plan.awaiting_backward = bool(out.requires_grad)
        else:
            with torch.no_grad():
                plan.forward(x)
                out = plan.logits()
        _bump_bn_counters(model)
        return out, None, None
    if torch.is_grad_enabled() and x.requires_grad:
        # as BiSeNet's eval forward: inference only — never a silently detached tensor for a caller who wants input gradients
        raise ops._lib.RtsdsError("DeepLabV2 eval-mode forward is inference only (no backward through the folded-BatchNorm plan): "
                                  "call it under torch.no_grad(), or use model.train() for gradients")
    plan = _get_plan(model, x, False)
    with torch.no_grad():
        plan.forward_lowres(x, model.rtsds_cuda_graph)
        return plan.logits()
