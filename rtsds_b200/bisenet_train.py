"""Train-mode execution plan for BiSeNet-ResNet18: forward with batch-statistics
BatchNorm and the hand-written backward (reference models/bisenet/build_bisenet.py
:141-172 driven by train.py:77-96 / :199-233).

Forward keeps, per conv layer, the raw convolution output and the post-activation
output (both NHWC); backward walks the layers in reverse:
   BN(+ReLU) backward (reduce + apply)  ->  wgrad (tcgen05, MN-major operands)
                                        ->  dgrad (tcgen05 tap-GEMM, residual-add epilogue)
plus the adjoint glue kernels (bilinear resize, ARM / FFM attention, max-pool).
Parameter gradients are accumulated into one flat fp32 buffer whose views are
returned to autograd (so grad accumulation across passes, requires_grad toggles
and optimizers behave exactly as with the reference modules).
"""
from __future__ import annotations

import torch

from . import ops, weights_epoch
from .ops import ACT_NONE, ACT_RELU, BF16, F32, _p, check, lib


def _s():
    return ops._s()


class _Buf:
    """A view (pointer + pixel pitch) into an NHWC buffer."""

    __slots__ = ("t", "ptr", "ld", "dtype")

    def __init__(self, t, ld=None, off=0, dtype=None):
        self.t = t
        self.ptr = t.data_ptr() + off * t.element_size()
        self.ld = t.shape[-1] if ld is None else ld
        self.dtype = ops.dtype_code(t.dtype) if dtype is None else dtype


class _ConvBN:
    """conv (+BatchNorm | +bias) (+residual) (+ReLU) with saved tensors for backward."""

    def __init__(self, plan, conv, bn, x: _Buf, xshape, y: _Buf, relu, residual: _Buf | None = None, raw_dtype=None,
                 need_dx=True):
        self.plan, self.conv, self.bn = plan, conv, bn
        self.x, self.y, self.res, self.relu, self.need_dx = x, y, residual, relu, need_dx
        n, h, w, cin = xshape
        self.cin, self.cout = cin, conv.weight.shape[0]
        k = conv.kernel_size[0]
        self.k = k
        dt = plan.dt
        self.tc = plan.use_tc
        raw_dtype = dt if raw_dtype is None else raw_dtype
        self.raw_dtype = raw_dtype
        self.d = ops.make_conv_desc(n, h, w, cin, x.ld, self.cout, 0, k, conv.stride[0], conv.padding[0], conv.dilation[0],
                                    in_dtype=dt, out_dtype=raw_dtype)
        self.oh, self.ow = self.d.oh, self.d.ow
        self.n_pix = n * self.oh * self.ow
        cpad8 = (self.cout + 7) // 8 * 8
        if bn is not None:
            raw_c = max(cpad8, 32) if raw_dtype == F32 and self.cout < 32 else cpad8
            self.raw = _Buf(plan.zeros(n, self.oh, self.ow, raw_c, dtype=ops.torch_dtype(raw_dtype)))
            self.d.out_ld = self.raw.ld
            self.scale = plan.buf(self.cout, dtype=torch.float32)
            self.shift = plan.buf(self.cout, dtype=torch.float32)
            self.save_mean = plan.buf(self.cout, dtype=torch.float32)
            self.save_invstd = plan.buf(self.cout, dtype=torch.float32)
            self.stats = plan.alloc_stats(self.cout)
            self.sums = plan.buf(2 * self.cout, dtype=torch.float32)
        else:
            self.raw = y                       # bias-only conv writes straight to y
            self.d.out_ld = y.ld
            self.d.out_dtype = y.dtype
        if getattr(self, "_no_std_weights", False):      # taps-as-N subclass: its own weight forms (rtsds_b200/tapn.py)
            self.wpk = self.wdg = self.dw = None
            self.ck = ops.dgrad_ck(self.cout, self.tc)
            self.dyld = max(self.ck, cpad8)
            plan.note_scratch(n * self.oh * self.ow * self.dyld, 0)
            return
        self.wpk = plan.buf(ops.cout_pad(self.cout), k * k, cin)
        batched = hasattr(plan, "pack_jobs")        # plans that batch weight packing / gradient unpacking into few launches
        if batched:
            plan.pack_jobs.append((conv, self.wpk, 0))
            self.dw = plan.zeros(self.cout * k * k * cin, dtype=torch.float32)     # kept zero: unpack clears what it reads
        else:
            plan.pack_steps.append(lambda: ops.pack_conv_weight(conv.weight, dt, self.wpk))
            self.dw = None
        self.ck = ops.dgrad_ck(self.cout, self.tc)
        if need_dx:
            self.wdg = plan.buf(ops.cout_pad(cin), k * k, self.ck)
            if batched:
                plan.pack_jobs.append((conv, self.wdg, 1))
            else:
                plan.pack_steps.append(lambda: ops.pack_conv_weight_dgrad(conv.weight, dt, self.tc, self.wdg))
        # backward geometry: dy operand (d_raw) pitch = ck (tensor cores) / cout rounded to 8
        self.dyld = max(self.ck, cpad8)
        plan.note_scratch(n * self.oh * self.ow * self.dyld, self.cout * k * k * cin)
        if self.tc:
            plan.note_ws(int(lib().rtsds_conv2d_tc_workspace_bytes(self.d)))
            dd = self.bwd_desc()
            plan.note_ws(int(lib().rtsds_conv2d_tc_dgrad_workspace_bytes(dd)))

    def bwd_desc(self):
        d = self.d
        dd = ops.make_conv_desc(d.n, d.h, d.w, d.cin, d.in_ld, d.cout, self.dyld, d.kh, d.stride, d.pad, d.dil,
                                in_dtype=self.plan.dt, out_dtype=self.plan.dt)
        return dd

    # ---- forward ----
    def forward(self):
        p = self.plan
        if self.bn is None:
            b = self.conv.bias
            self._launch(self.d, None, b.detach() if b is not None else None, None, None, self.y.ptr)
            return
        st = p.stats_view(self.stats, self.cout)
        self._launch(self.d, None, None, None, st, self.raw.ptr)
        # batch statistics -> scale/shift (+ running buffers) and the normalise [+ residual] [+ ReLU] pass in ONE launch
        ops.bn_finalize_apply_ptr(st, self.n_pix, self.bn, self.scale, self.shift, self.save_mean, self.save_invstd,
                                  self.raw.ptr, self.y.ptr, self.n_pix, self.cout,
                                  self.res.ptr if self.res is not None else None, ACT_RELU if self.relu else ACT_NONE, 0.0,
                                  self.raw.ld, self.y.ld, self.res.ld if self.res is not None else self.cout, self.raw.dtype,
                                  self.y.dtype)

    def _launch(self, d, scale, shift, res, stats, yptr):
        if self.tc:
            ops.conv2d_tc(d, self.x.ptr, self.wpk, yptr, scale, shift, res, stats, self.plan.ws)
        else:
            ops.conv2d_simt(d, self.x.ptr, self.wpk, yptr, scale, shift, res, stats)

    # ---- backward ----
    def backward(self, dy: _Buf, gw, dx: _Buf | None = None, dx_accumulate=False, g_out: _Buf | None = None):
        """dy: gradient w.r.t. this layer's output y.  gw: dict param -> fp32 grad view (missing = frozen).
        dx: where to write the input gradient (dx_accumulate: add to what is there).
        g_out: receives the ReLU-masked gradient (the residual / identity branch)."""
        p = self.plan
        s = _s()
        dt = p.dt
        if self.bn is not None:
            d_raw = p.d_raw_view(self.dyld)
            gamma = self.bn.weight
            dgam, dbet = gw.get(self.bn.weight), gw.get(self.bn.bias)
            if self.relu and self.res is None and self.raw.dtype == dy.dtype:
                # no residual: the ReLU mask is a function of raw alone -> the activation is not re-read
                check(lib().rtsds_bn_bwd_reduce_rawmask(dy.ptr, dy.ld, self.raw.ptr, self.raw.ld, _p(self.save_mean),
                                                        _p(self.save_invstd), _p(self.scale), _p(self.shift), self.n_pix,
                                                        self.cout, dy.dtype, _p(self.sums), s), "bn_bwd_reduce_rawmask")
                check(lib().rtsds_bn_bwd_apply_rawmask(dy.ptr, dy.ld, self.raw.ptr, self.raw.ld, _p(self.save_mean),
                                                       _p(self.save_invstd), _p(gamma.detach()), _p(self.sums), _p(self.scale),
                                                       _p(self.shift), self.n_pix, self.cout, dy.dtype, d_raw.ptr, self.dyld, dt,
                                                       g_out.ptr if g_out is not None else None,
                                                       g_out.ld if g_out is not None else 0, _p(dgam), _p(dbet), s),
                      "bn_bwd_apply_rawmask")
            else:
                check(lib().rtsds_bn_bwd_reduce(dy.ptr, dy.ld, self.y.ptr, self.y.ld, self.raw.ptr, self.raw.ld,
                                                _p(self.save_mean), _p(self.save_invstd), self.n_pix, self.cout, int(self.relu),
                                                dy.dtype, _p(self.sums), s), "bn_bwd_reduce")
                check(lib().rtsds_bn_bwd_apply(dy.ptr, dy.ld, self.y.ptr, self.y.ld, self.raw.ptr, self.raw.ld, _p(self.save_mean),
                                               _p(self.save_invstd), _p(gamma.detach()), _p(self.sums), self.n_pix, self.cout,
                                               int(self.relu), dy.dtype, d_raw.ptr, self.dyld, dt,
                                               g_out.ptr if g_out is not None else None, g_out.ld if g_out is not None else 0,
                                               _p(dgam), _p(dbet), s), "bn_bwd_apply")
        else:
            d_raw = dy                                   # bias-only conv: dy already is the conv-output gradient
            assert dy.ld >= self.dyld and dy.dtype == dt
        gwt = gw.get(self.conv.weight)
        if gwt is not None:
            self._weight_grad(d_raw, gwt)
        if dx is not None:
            self._input_grad(d_raw, dx, dx_accumulate)

    def _weight_grad(self, d_raw: _Buf, gwt):
        p = self.plan
        dd = self.bwd_desc()
        dd.out_ld = d_raw.ld
        if self.dw is not None:
            ops.conv2d_wgrad(dd, self.x.ptr, d_raw.ptr, self.dw, self.tc)
            # converted in one launch per bucket; ASSIGNED: the flat gradient buffer of this backward pass is fresh and
            # every conv weight receives exactly one contribution per pass (autograd accumulates across passes)
            p.pending_unpack.append((self.dw, gwt, False))
        else:
            dwp = p.dw_view(self.cout * self.k * self.k * self.cin)     # kept zero: unpack clears what it reads
            ops.conv2d_wgrad(dd, self.x.ptr, d_raw.ptr, dwp, self.tc)
            ops.unpack_conv_wgrad(dwp, gwt, True)

    def _input_grad(self, d_raw: _Buf, dx: _Buf, dx_accumulate):
        dd = self.bwd_desc()
        dd.out_ld = d_raw.ld
        dd.in_ld = dx.ld
        dd.res_ld = dx.ld
        ops.conv2d_dgrad(dd, d_raw.ptr, self.wdg, dx.ptr, dx.dtype, self.tc, dx.ptr if dx_accumulate else None, self.plan.ws)


class _TapNConvBN(_ConvBN):
    """_ConvBN whose convolution runs in the taps-as-N form (rtsds_b200/tapn.py): skinny-output k x k convs."""

    def __init__(self, plan, conv, bn, x: _Buf, xshape, y: _Buf, relu, residual=None, raw_dtype=None, need_dx=True):
        from .tapn import TapNConv

        assert bn is not None and raw_dtype == F32 and residual is None
        self._no_std_weights = True
        super().__init__(plan, conv, bn, x, xshape, y, relu, residual, raw_dtype, need_dx)
        self.tapn = TapNConv(plan, conv, x.ptr, xshape, x.ld, train=True)

    def _launch(self, d, scale, shift, res, stats, yptr):
        assert scale is None and shift is None and res is None
        self.tapn.forward(None, None, ACT_NONE, stats, yptr, self.raw.ld)

    def backward(self, dy, gw, dx=None, dx_accumulate=False, g_out=None):
        self._scattered = False
        super().backward(dy, gw, dx, dx_accumulate, g_out)

    def _scatter(self, d_raw):
        if not self._scattered:
            self.tapn.scatter(d_raw.ptr, d_raw.ld, d_raw.dtype)
            self._scattered = True

    def _weight_grad(self, d_raw, gwt):
        self._scatter(d_raw)
        self.tapn.weight_grad(gwt)

    def _input_grad(self, d_raw, dx, dx_accumulate):
        self._scatter(d_raw)
        self.tapn.input_grad(dx.ptr, dx.ld, dx.dtype, dx_accumulate)


class _Stem:
    """Direct conv on the NCHW fp32 image + BatchNorm + ReLU (no input gradient)."""

    def __init__(self, plan, conv, bn, k, pad, n, h, w):
        self.plan, self.conv, self.bn, self.k, self.pad = plan, conv, bn, k, pad
        self.oh, self.ow = ops.conv_out_size(h, k, 2, pad), ops.conv_out_size(w, k, 2, pad)
        self.n_pix = n * self.oh * self.ow
        self.raw = _Buf(plan.buf(n, self.oh, self.ow, 64))
        self.y = _Buf(plan.buf(n, self.oh, self.ow, 64))
        self.scale, self.shift = plan.buf(64, dtype=torch.float32), plan.buf(64, dtype=torch.float32)
        self.save_mean, self.save_invstd = plan.buf(64, dtype=torch.float32), plan.buf(64, dtype=torch.float32)
        self.stats = plan.alloc_stats(64)
        self.sums = plan.buf(128, dtype=torch.float32)
        plan.note_scratch(self.n_pix * 64, 0)
        # the fused tensor-core weight gradient needs both stems' conv-output gradients at once
        self.d_raw = _Buf(plan.buf(n, self.oh, self.ow, 64)) if plan.use_tc else None

    def forward(self, x, conv_done=False):
        p = self.plan
        st = p.stats_view(self.stats, 64)
        if not conv_done:
            ops.stem_conv(x, self.conv.weight, self.raw.t, self.k, 2, self.pad, stats=st)
        ops.bn_finalize_apply_ptr(st, self.n_pix, self.bn, self.scale, self.shift, self.save_mean, self.save_invstd, self.raw.t,
                                  self.y.t, self.n_pix, 64, None, ACT_RELU, 0.0, 64, 64, 64, ops.dtype_code(self.raw.t.dtype),
                                  ops.dtype_code(self.y.t.dtype))

    def backward(self, x, dy: _Buf, gw, wgrad=True):
        p = self.plan
        s = _s()
        d_raw = self.d_raw if self.d_raw is not None else p.d_raw_view(64)
        check(lib().rtsds_bn_bwd_reduce_rawmask(dy.ptr, dy.ld, self.raw.ptr, 64, _p(self.save_mean), _p(self.save_invstd),
                                                _p(self.scale), _p(self.shift), self.n_pix, 64, dy.dtype, _p(self.sums), s),
              "bn_bwd_reduce_rawmask")
        check(lib().rtsds_bn_bwd_apply_rawmask(dy.ptr, dy.ld, self.raw.ptr, 64, _p(self.save_mean), _p(self.save_invstd),
                                               _p(self.bn.weight.detach()), _p(self.sums), _p(self.scale), _p(self.shift),
                                               self.n_pix, 64, dy.dtype, d_raw.ptr, 64, p.dt, None, 0,
                                               _p(gw.get(self.bn.weight)), _p(gw.get(self.bn.bias)), s), "bn_bwd_apply_rawmask")
        gwt = gw.get(self.conv.weight)
        if wgrad and gwt is not None:
            n, cin, h, w = x.shape
            check(lib().rtsds_stem_conv_wgrad(_p(x), d_raw.ptr, p.dt, n, cin, h, w, 64, self.k, 2, self.pad, _p(gwt), s),
                  "stem_conv_wgrad")


class BiSeNetTrainPlan:
    def __init__(self, model, n, h, w, precision="bf16"):
        p0 = model.conv.weight
        if not p0.is_cuda and not ops._lib.dry_run():
            raise ops._lib.RtsdsError("BiSeNet parameters must live on a CUDA device (no CPU fallback)")
        check(lib().rtsds_check_device(), "device check")
        if model._context_name not in ("resnet18", "resnet101"):
            raise ops._lib.RtsdsError(f"context paths: resnet18 and resnet101 (build_bisenet.py:95-113); got {model._context_name!r}")
        self.model, self.device = model, p0.device
        self.n, self.h, self.w = n, h, w
        # "bf16": tensor-core kernels; "fp32": CUDA-core check mode; "bf16_simt": bf16 storage with the CUDA-core
        # conv kernels (cross-checks the tcgen05 kernels on identical operands)
        self.dt = F32 if precision == "fp32" else BF16
        self.use_tc = precision == "bf16"
        self.tdt = ops.torch_dtype(self.dt)
        self.nc = model.conv.weight.shape[0]
        self._keep, self.pack_steps = [], []
        self.pack_jobs, self.pending_unpack = [], []
        self._stats_total = 0
        self._scratch_act, self._scratch_w, self._ws_bytes = 0, 0, 0
        self._param_version = None
        self.generation = 0
        self.ws = None
        self._build()
        weights_epoch.register_plan(self)        # the fused optimizers refresh this plan's packed operands (optim.py)

    def flush_unpack(self):
        if self.pending_unpack:
            ops.unpack_conv_wgrads_batch(self.pending_unpack)
            self.pending_unpack = []

    # ---------------- allocation helpers ----------------
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

    # ---------------- construction ----------------
    def _build(self):
        m, n, H, W = self.model, self.n, self.h, self.w
        cs = ops.conv_out_size
        f32 = torch.float32
        nc = self.nc
        sp, cp = m.saptial_path, m.context_path
        # spatial path
        self.sp1 = _Stem(self, sp.convblock1.conv1, sp.convblock1.bn, 3, 1, n, H, W)
        h2, w2 = self.sp1.oh, self.sp1.ow
        h4, w4 = cs(h2, 3, 2, 1), cs(w2, 3, 2, 1)
        h8, w8 = cs(h4, 3, 2, 1), cs(w4, 3, 2, 1)
        self.h8, self.w8 = h8, w8
        # concat buffer of build_bisenet.py:153,72: 256 spatial-path channels | cx1 | cx2 (1024 wide for resnet18, 3328 for
        # resnet101 — Bottleneck context path)
        self.r101 = hasattr(cp.layer1[0], "conv3")
        self.catw = catw = 256 + (1024 + 2048 if self.r101 else 256 + 512)
        self.cat = self.buf(n, h8, w8, catw)
        sp2y = _Buf(self.buf(n, h4, w4, 128))
        self.sp2 = _ConvBN(self, sp.convblock2.conv1, sp.convblock2.bn, self.sp1.y, (n, h2, w2, 64), sp2y, True)
        self.sp3 = _ConvBN(self, sp.convblock3.conv1, sp.convblock3.bn, sp2y, (n, h4, w4, 128), _Buf(self.cat, ld=catw), True)
        # context path
        self.cp0 = _Stem(self, cp.conv1, cp.bn1, 7, 3, n, H, W)
        ph, pw = ops.maxpool_out_size(self.cp0.oh), ops.maxpool_out_size(self.cp0.ow)
        self.pool = _Buf(self.buf(n, ph, pw, 64))
        self.pool_shape = (n, ph, pw, 64)
        self.pool_idx = self.buf(n, ph, pw, 8, dtype=torch.int32)
        self.blocks = []
        self.layer_first = {}                    # index of the first block of layer2..4 -> its gradient bucket
        x, shape = self.pool, self.pool_shape
        for li, layer in enumerate((cp.layer1, cp.layer2, cp.layer3, cp.layer4), 1):
            if li > 1:
                self.layer_first[len(self.blocks)] = f"layer{li}"
            for blk in layer:
                x, shape = (self._bottleneck if self.r101 else self._block)(blk, x, shape)
        self.i_l4 = len(self.blocks) - len(cp.layer4)       # first block of layer4: its input is feature3 (also read by ARM1)
        self.f3, self.s3 = self.blocks[self.i_l4 - 1]["y"], self.blocks[self.i_l4 - 1]["shape"]
        self.f4, self.s4 = self.blocks[-1]["y"], self.blocks[-1]["shape"]
        c3, c4 = self.s3[3], self.s4[3]
        self.c3, self.c4 = c3, c4
        self.arm = {}
        for tag, c in (("3", c3), ("4", c4)):
            for nm in ("pooled", "gate", "lin", "xhat", "dgate", "dlin", "dmul", "dpooled"):
                self.arm[nm + tag] = self.buf(n, c, dtype=f32)
        # heads
        self.z, self.z1, self.z2 = (self.zeros(n, h8, w8, 32, dtype=f32) for _ in range(3))
        zdt = self.dt
        z1b = _Buf(self.z1, dtype=F32)
        z2b = _Buf(self.z2, dtype=F32)
        self.sup1 = _ConvBN(self, m.supervision1, None, _Buf(self.cat, ld=catw, off=256), (n, h8, w8, c3), z1b, False)
        self.sup2 = _ConvBN(self, m.supervision2, None, _Buf(self.cat, ld=catw, off=256 + c3), (n, h8, w8, c4), z2b, False)
        ffm = m.feature_fusion_module
        self.feat = _Buf(self.zeros(n, h8, w8, 32, dtype=f32))
        from . import tapn

        ffm_cls = _TapNConvBN if tapn.applicable(ffm.convblock.conv1) else _ConvBN
        self.ffm = ffm_cls(self, ffm.convblock.conv1, ffm.convblock.bn, _Buf(self.cat, ld=catw), (n, h8, w8, catw), self.feat, True,
                           raw_dtype=F32)
        self.pooled_f = self.buf(n, nc, dtype=f32)
        self.attn = self.buf(n, nc, dtype=f32)
        self.da_ws, self.dpf_ws = self.buf(n, nc, dtype=f32), self.buf(n, nc, dtype=f32)
        # gradient buffers
        self.dz = [self.zeros(n, h8, w8, 32, dtype=f32) for _ in range(3)]
        self.dzb = [self.zeros(n, h8, w8, 64 if self.dt == BF16 else 32, dtype=self.tdt) for _ in range(2)]   # aux dy operands
        self.dfeat = self.zeros(n, h8, w8, 32, dtype=f32)
        self.dcat = self.zeros(n, h8, w8, catw)
        self.dg3 = self.buf(n, self.s3[1], self.s3[2], c3, dtype=f32)
        self.dg4 = self.buf(n, self.s4[1], self.s4[2], c4, dtype=f32)
        max_act = max([self.sp1.n_pix * 64, self.cp0.n_pix * 64] + [b["shape"][0] * b["shape"][1] * b["shape"][2] * b["shape"][3] for b in self.blocks]
                      + [b["xshape"][0] * b["xshape"][1] * b["xshape"][2] * b["xshape"][3] for b in self.blocks])
        self.gA, self.gB, self.gT, self.gG = (self.buf(max_act) for _ in range(4))
        if self.r101:
            self.gT1 = self.buf(max_act)
        self.stats_all = torch.zeros(max(self._stats_total, 1), dtype=f32, device=self.device)
        self.d_raw_scratch = self.zeros(max(self._scratch_act, 1))
        self.dw_scratch = self.zeros(max(self._scratch_w, 1), dtype=f32)
        if self._ws_bytes:
            self.ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
        if self.use_tc:
            self.stem_wpk = self.buf(128, 192, dtype=torch.bfloat16)
            self.stem_dw_ws = self.zeros(128 * 192, dtype=f32)
            self.pack_steps.append(lambda: ops.stem_pack_weights(cp.conv1.weight, sp.convblock1.conv1.weight, self.stem_wpk))
        self.acc = torch.zeros(3, 4, dtype=torch.float64, device=self.device)      # fused-loss accumulators
        self.gscale = torch.zeros(3, dtype=f32, device=self.device)
        # flat parameter-gradient buffer
        self.params = [p for p in m.parameters()]
        self._grad_numel = sum(p.numel() for p in self.params)
        fc = m.context_path.features.fc            # the reference never uses it: its grads stay None
        self.unused = {fc.weight, fc.bias}

    def _block(self, blk, x: _Buf, shape):
        n, h, w, cin = shape
        cout = blk.conv1.weight.shape[0]
        st = blk.conv1.stride[0]
        oh, ow = ops.conv_out_size(h, 3, st, 1), ops.conv_out_size(w, 3, st, 1)
        t = _Buf(self.buf(n, oh, ow, cout))
        y = _Buf(self.buf(n, oh, ow, cout))
        c1 = _ConvBN(self, blk.conv1, blk.bn1, x, shape, t, True)
        ds = None
        if blk.downsample is not None:
            dsy = _Buf(self.buf(n, oh, ow, cout))
            ds = _ConvBN(self, blk.downsample[0], blk.downsample[1], x, shape, dsy, False)
            res = dsy
        else:
            res = x
        c2 = _ConvBN(self, blk.conv2, blk.bn2, t, (n, oh, ow, cout), y, True, residual=res)
        self.blocks.append(dict(c1=c1, c2=c2, ds=ds, x=x, xshape=shape, y=y, shape=(n, oh, ow, cout)))
        return y, (n, oh, ow, cout)

    def _bottleneck(self, blk, x: _Buf, shape):
        """torchvision Bottleneck (build_contextpath.py:32-56): 1x1 -> 3x3 (carries the stride) -> 1x1 (+ shortcut) -> ReLU."""
        n, h, w, cin = shape
        planes = blk.conv1.weight.shape[0]
        st = blk.conv2.stride[0]
        oh, ow = ops.conv_out_size(h, 3, st, 1), ops.conv_out_size(w, 3, st, 1)
        t1 = _Buf(self.buf(n, h, w, planes))
        t2 = _Buf(self.buf(n, oh, ow, planes))
        y = _Buf(self.buf(n, oh, ow, planes * 4))
        c1 = _ConvBN(self, blk.conv1, blk.bn1, x, shape, t1, True)
        c2 = _ConvBN(self, blk.conv2, blk.bn2, t1, (n, h, w, planes), t2, True)
        ds = None
        if blk.downsample is not None:
            dsy = _Buf(self.buf(n, oh, ow, planes * 4))
            ds = _ConvBN(self, blk.downsample[0], blk.downsample[1], x, shape, dsy, False)
            res = dsy
        else:
            res = x
        c3 = _ConvBN(self, blk.conv3, blk.bn3, t2, (n, oh, ow, planes), y, True, residual=res)
        self.blocks.append(dict(c1=c1, c2=c2, c3=c3, ds=ds, x=x, xshape=shape, y=y, shape=(n, oh, ow, planes * 4), tshape=(n, h, w, planes)))
        return y, (n, oh, ow, planes * 4)

    # ---------------- weights ----------------
    def _params_version(self):
        v = 0
        for p in self.model.parameters():
            v += p._version
        return (v, self.model.conv.weight.data_ptr(), weights_epoch.value())

    def refresh_weights(self):
        ver = self._params_version()
        if ver != self._param_version:
            for s in self.pack_steps:
                s()
            ops.pack_conv_weights_batch([(c.weight, out, kind) for c, out, kind in self.pack_jobs], self.dt, self.use_tc)
            self._param_version = self._params_version()

    # ---------------- forward ----------------
    def forward(self, x):
        m, n = self.model, self.n
        self.refresh_weights()
        self.stats_all.zero_()
        self.generation += 1
        self.x = x
        if self.use_tc:      # both stems in one tensor-core kernel (raw outputs + BatchNorm statistics)
            ops.stem_pair_tc_fwd(x, self.stem_wpk, self.cp0.raw.t, self.sp1.raw.t, None, None, False,
                                 self.stats_view(self.cp0.stats, 64), self.stats_view(self.sp1.stats, 64))
        self.sp1.forward(x, conv_done=self.use_tc)
        self.sp2.forward()
        self.sp3.forward()
        self.cp0.forward(x, conv_done=self.use_tc)
        ops.maxpool3x3s2(self.cp0.y.t, self.pool.t, False, self.pool_idx)
        for b in self.blocks:
            b["c1"].forward()
            if b["ds"] is not None:
                b["ds"].forward()
            b["c2"].forward()
            if "c3" in b:
                b["c3"].forward()
        a, dt = self.arm, self.dt
        s3, s4, c3, c4, h8, w8 = self.s3, self.s4, self.c3, self.c4, self.h8, self.w8
        arm1, arm2 = m.attention_refinement_module1, m.attention_refinement_module2
        ops.global_avgpool(self.f3.t, n, s3[1] * s3[2], c3, c3, a["pooled3"])
        ops.global_avgpool(self.f4.t, n, s4[1] * s4[2], c4, c4, a["pooled4"])
        ops.arm_gate(a["pooled3"], arm1.conv, arm1.bn, True, n, c3, a["gate3"], None, a["lin3"], a["xhat3"])
        ops.arm_gate(a["pooled4"], arm2.conv, arm2.bn, True, n, c4, a["gate4"], a["pooled4"], a["lin4"], a["xhat4"])
        ops.gate_resize_nhwc(self.f3.t, n, s3[1], s3[2], c3, c3, a["gate3"], h8, w8, self.cat, self.catw, 256, dt)
        ops.gate_resize_nhwc(self.f4.t, n, s4[1], s4[2], c4, c4, a["gate4"], h8, w8, self.cat, self.catw, 256 + c3, dt)
        self.sup1.forward()
        self.sup2.forward()
        self.ffm.forward()
        ffm = m.feature_fusion_module
        ops.global_avgpool(self.feat.t, n, h8 * w8, self.nc, 32, self.pooled_f)
        final = m.conv if m.with_interpolation else None
        ops.ffm_head(self.feat.t, F32, 32, self.pooled_f, n, h8 * w8, self.nc, ffm.conv1, ffm.conv2, final, self.z, 32, self.attn)

    def out_sizes(self):
        main = (self.h8 * 8, self.w8 * 8) if self.model.with_interpolation else (self.h8, self.w8)
        return main, (self.h, self.w)

    def logits(self):
        main, aux = self.out_sizes()
        outs = []
        for z, (oh, ow) in ((self.z, main), (self.z1, aux), (self.z2, aux)):
            o = torch.empty((self.n, self.nc, oh, ow), dtype=torch.float32, device=self.device)
            ops.resize_to_nchw(z, self.n, self.h8, self.w8, self.nc, 32, o)
            outs.append(o)
        return outs

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
        """self.dz[0..2] hold the gradients w.r.t. z, z1, z2 (fp32 NHWC pitch 32).  `ready(group)` is called
        as soon as every parameter gradient of a bucket (rtsds_b200/ddp.py) is final."""
        weights_epoch.note_backward()          # an optimizer step may follow: every plan re-packs at its next forward
        m, n, dt = self.model, self.n, self.dt
        s = _s()
        a = self.arm
        user_ready = ready

        def ready(group):                       # a bucket is final only once its wgrad results are in OIHW form
            self.flush_unpack()
            user_ready(group)

        self.pending_unpack = []
        h8, w8, nc, c3, c4, s3, s4 = self.h8, self.w8, self.nc, self.c3, self.c4, self.s3, self.s4
        npix8 = n * h8 * w8
        ffm = m.feature_fusion_module
        final = m.conv if m.with_interpolation else None
        # ---- FFM head: dz -> dfeat, grads of ffm.conv1/conv2 and the final conv ----
        check(lib().rtsds_ffm_head_bwd(_p(self.dz[0]), 32, self.feat.ptr, 32, _p(self.pooled_f), _p(self.attn), n, h8 * w8, nc,
                                       _p(ffm.conv1.weight.detach()), _p(ffm.conv1.bias.detach()), _p(ffm.conv2.weight.detach()),
                                       _p(final.weight.detach()) if final is not None else None, _p(self.da_ws), _p(self.dpf_ws),
                                       _p(self.dfeat), 32, _p(self._g(gw, ffm.conv1.weight)), _p(self._g(gw, ffm.conv1.bias)),
                                       _p(self._g(gw, ffm.conv2.weight)), _p(self._g(gw, ffm.conv2.bias)),
                                       _p(gw.get(final.weight)) if final is not None else None,
                                       _p(gw.get(final.bias)) if final is not None else None, s), "ffm_head_bwd")
        # ---- FFM ConvBlock: -> dcat (assign over all 1024 channels) ----
        catw = self.catw
        dcat = _Buf(self.dcat, ld=catw)
        self.ffm.backward(_Buf(self.dfeat, dtype=F32), gw, dx=dcat, dx_accumulate=False)
        # ---- auxiliary heads: 1x1 convs on the cx1 / cx2 slots of the concat buffer ----
        for i, (layer, off) in enumerate(((self.sup1, 256), (self.sup2, 256 + c3))):
            dzi = self.dz[1 + i]
            gb = gw.get(layer.conv.bias)
            if gb is not None:
                check(lib().rtsds_channel_sum(_p(dzi), 32, npix8, nc, F32, _p(gb), s), "channel_sum")
            dyb = self.dzb[i]
            ops.scale_shift_act_ptr(dzi, dyb, npix8, nc, None, None, None, ACT_NONE, 0.0, 32, dyb.shape[-1], nc, F32, dt)
            layer.backward(_Buf(dyb), gw, dx=_Buf(self.dcat, ld=catw, off=off), dx_accumulate=True)
        # ---- gated resize + ARM backward -> gradients of f3 / f4 ----
        dF = {}
        for tag, f, shp, c, off, arm in (("3", self.f3, s3, c3, 256, m.attention_refinement_module1),
                                         ("4", self.f4, s4, c4, 256 + c3, m.attention_refinement_module2)):
            hw = shp[1] * shp[2]
            dgt = self.dg3 if tag == "3" else self.dg4
            check(lib().rtsds_resize_bwd_nhwc(_p(self.dcat), catw, off, n, shp[1], shp[2], c, h8, w8, f.ptr, dt, _p(dgt),
                                              _p(a["dgate" + tag]), s), "resize_bwd_nhwc")
            mul = a["pooled4"] if tag == "4" else None
            check(lib().rtsds_arm_gate_bwd(_p(a["dgate" + tag]), _p(a["pooled" + tag]), _p(a["lin" + tag]), _p(a["xhat" + tag]),
                                           _p(arm.conv.weight.detach()), _p(arm.bn.weight.detach()), _p(arm.bn.bias.detach()),
                                           _p(mul), float(arm.bn.eps), n, c, _p(a["dlin" + tag]), _p(a["dmul" + tag]),
                                           _p(a["dpooled" + tag]), _p(gw.get(arm.conv.weight)), _p(gw.get(arm.conv.bias)),
                                           _p(gw.get(arm.bn.weight)), _p(gw.get(arm.bn.bias)), s), "arm_gate_bwd")
            dfb = _Buf(self.gA if tag == "4" else self.gB, ld=c, dtype=dt)
            check(lib().rtsds_gate_bwd_finish(_p(dgt), _p(a["gate" + tag]), _p(a["dpooled" + tag]), 1.0 / hw, n, hw, c, dt,
                                              dfb.ptr, s), "gate_bwd_finish")
            dF[tag] = dfb
        ready("head")
        # ---- ResNet-18 stages in reverse ----
        # gA holds the gradient of f4 (block 7 output); gB the ARM1 part of the gradient of f3 (block 5
        # output = block 6 input).  A block consumes dy completely (BN backward of conv2) before its input
        # gradient is written, so dx may alias dy; only block 6 writes elsewhere: it ACCUMULATES into gB.
        dy = dF["4"]
        for bi in range(len(self.blocks) - 1, -1, -1):
            b = self.blocks[bi]
            cout, cin = b["shape"][3], b["xshape"][3]
            g = _Buf(self.gG, ld=cout, dtype=dt)
            acc_in = bi == self.i_l4
            dx = _Buf(self.gB if acc_in else dy.t, ld=cin, dtype=dt)
            if "c3" in b:                                   # Bottleneck: conv3 (+shortcut) -> conv2 -> conv1
                planes = cout // 4
                dT2 = _Buf(self.gT, ld=planes, dtype=dt)
                dT1 = _Buf(self.gT1, ld=planes, dtype=dt)
                b["c3"].backward(dy, gw, dx=dT2, dx_accumulate=False, g_out=g)
                if b["ds"] is not None:
                    b["ds"].backward(g, gw, dx=dx, dx_accumulate=acc_in)
                else:
                    self._copy(g, dx, b["xshape"])
                b["c2"].backward(dT2, gw, dx=dT1, dx_accumulate=False)
                b["c1"].backward(dT1, gw, dx=dx, dx_accumulate=True)
            else:
                dT = _Buf(self.gT, ld=cout, dtype=dt)
                b["c2"].backward(dy, gw, dx=dT, dx_accumulate=False, g_out=g)
                if b["ds"] is not None:
                    b["ds"].backward(g, gw, dx=dx, dx_accumulate=acc_in)
                else:
                    self._copy(g, dx, b["xshape"])            # identity shortcut (never the accumulating block)
                b["c1"].backward(dT, gw, dx=dx, dx_accumulate=True)
            dy = dx
            if bi in self.layer_first:
                ready(self.layer_first[bi])
        # ---- max-pool and the 7x7 stem ----
        n_, ph, pw, _ = self.pool_shape
        dcp0 = _Buf(self.gT, ld=64, dtype=dt)
        check(lib().rtsds_maxpool3x3s2_bwd_idx(self.pool_idx.data_ptr(), dy.ptr, n, self.cp0.oh, self.cp0.ow, 64, dt, 0, dcp0.ptr, s),
              "maxpool_bwd_idx")
        self.cp0.backward(self.x, dcp0, gw, wgrad=not self.use_tc)
        # ---- spatial path ----
        d2 = _Buf(self.gA, ld=128, dtype=dt)
        self.sp3.backward(_Buf(self.dcat, ld=catw), gw, dx=d2, dx_accumulate=False)
        d1 = _Buf(self.gB, ld=64, dtype=dt)
        self.sp2.backward(d2, gw, dx=d1, dx_accumulate=False)
        self.sp1.backward(self.x, d1, gw, wgrad=not self.use_tc)
        self.flush_unpack()
        if self.use_tc:
            g7, g3 = gw.get(self.cp0.conv.weight), gw.get(self.sp1.conv.weight)
            if g7 is not None or g3 is not None:
                ops.stem_pair_tc_wgrad(self.x, self.cp0.d_raw.t, self.sp1.d_raw.t, self.stem_dw_ws, g7, g3)

    def _g(self, gw, p):
        g = gw.get(p)
        if g is None:           # frozen parameter: kernels still need somewhere to write — ONE scratch target per parameter,
            cache = self.__dict__.setdefault("_frozen_grad_scratch", {})     # not a new plan-owned buffer every backward
            g = cache.get(p)
            if g is None:
                g = cache[p] = self.buf(*p.shape, dtype=torch.float32)
        return g

    def _copy(self, src: _Buf, dst: _Buf, shape):
        n, h, w, c = shape
        ops.scale_shift_act_ptr(src.ptr, dst.ptr, n * h * w, c, None, None, None, ACT_NONE, 0.0, src.ld, dst.ld, c, src.dtype,
                                dst.dtype)

    def _add_into(self, dst: _Buf, src: _Buf, shape):
        n, h, w, c = shape
        ops.scale_shift_act_ptr(src.ptr, dst.ptr, n * h * w, c, None, None, dst.ptr, ACT_NONE, 0.0, src.ld, dst.ld, dst.ld,
                                src.dtype, dst.dtype)
