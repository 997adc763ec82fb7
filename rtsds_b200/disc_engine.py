"""Execution plan of the domain discriminators (reference
models/domain_shift/adversarial/model.py:30-83) as driven by the adversarial step
(train.py:218-263): forward on the [N,19,H,W] class-probability map, BCE-with-logits
gradient back to the discriminator's parameters and/or to the generator's logits.

Data layout in HBM (bf16 production mode; fp32 in check mode):
  probs / logits   NCHW fp32   [N,C,H,W]              API boundary (C = num_classes <= 32)
  xs               NHWC        [N,H/2+1,W/2+1,128]    space-to-depth of the padded input (4 parities x 32 ch);
                                                      softmax fused when the caller hands over logits
  y1               NHWC        [N,H/2,W/2,64]         conv1 = 2x2 s1 tcgen05 conv over xs, + bias + LeakyReLU
  y2..y4           NHWC        64->128->256->512, each 4x4 s2 p1 (+bias, LeakyReLU)       (full D only)
  tapsum           fp32        [N,16,C_last]          classifier + global-average-pool collapse (csrc/disc.cu)
  out              fp32        [N,1,1,1]              API boundary
"""
from __future__ import annotations

import ctypes as C

import torch

from . import ops, weights_epoch
from .ops import ACT_LRELU, BF16, F32, _p, check, lib


class _Layer:
    """conv (+bias) + LeakyReLU with what backward needs."""

    def __init__(self, plan, conv, w_src, x, xshape, k, stride, pad):
        self.plan, self.conv = plan, conv
        n, h, w, cin = xshape
        self.cin, self.cout, self.k = cin, conv.weight.shape[0], k
        dt = plan.dt
        self.x = x
        self.d = ops.make_conv_desc(n, h, w, cin, cin, self.cout, self.cout, k, stride, pad, 1, act=ACT_LRELU, slope=plan.slope,
                                    in_dtype=dt, out_dtype=dt)
        self.oh, self.ow = self.d.oh, self.d.ow
        self.n_pix = n * self.oh * self.ow
        self.y = plan.buf(n, self.oh, self.ow, self.cout)
        self.wpk = plan.buf(ops.cout_pad(self.cout), k * k, cin)
        self.ck = ops.dgrad_ck(self.cout, plan.use_tc)
        self.wdg = plan.buf(ops.cout_pad(cin), k * k, self.ck)
        self.w_src = w_src                        # callable -> OIHW fp32 weight of THIS conv geometry
        plan.pack_steps.append(lambda: ops.pack_conv_weight(self.w_src(), dt, self.wpk))
        plan.pack_steps.append(lambda: ops.pack_conv_weight_dgrad(self.w_src(), dt, plan.use_tc, self.wdg))
        plan.note_scratch(self.n_pix * self.cout, self.cout * k * k * cin)
        if plan.use_tc:
            plan.note_ws(int(lib().rtsds_conv2d_tc_workspace_bytes(self.d)))
            plan.note_ws(int(lib().rtsds_conv2d_tc_dgrad_workspace_bytes(self.d)))

    def forward(self):
        p = self.plan
        b = self.conv.bias.detach() if self.conv.bias is not None else None
        if p.use_tc:
            ops.conv2d_tc(self.d, self.x, self.wpk, self.y, None, b, None, None, p.ws)
        else:
            ops.conv2d_simt(self.d, self.x, self.wpk, self.y, None, b, None, None)


class DiscPlan:
    def __init__(self, model, n, c, h, w, precision="bf16"):
        p0 = model.conv1.weight
        if not p0.is_cuda and not ops._lib.dry_run():
            raise ops._lib.RtsdsError("discriminator parameters must live on a CUDA device (no CPU fallback)")
        check(lib().rtsds_check_device(), "device check")
        if c != model.conv1.weight.shape[1]:
            raise ValueError(f"expected {model.conv1.weight.shape[1]} input channels, got {c}")
        if c > 32:
            raise ops._lib.RtsdsError("num_classes > 32 is not supported by the space-to-depth conv1")
        if h < 4 or w < 4:
            raise ops._lib.RtsdsError("input too small for the discriminator")
        self.model, self.device = model, p0.device
        self.n, self.c, self.h, self.w = n, c, h, w
        self.dt = F32 if precision == "fp32" else BF16
        self.use_tc = precision == "bf16"
        self.tdt = ops.torch_dtype(self.dt)
        self.slope = float(model.leaky_relu.negative_slope)
        self._keep, self.pack_steps = [], []
        self._scratch_act, self._scratch_w, self._ws_bytes = 0, 0, 0
        self._param_version = None
        self.generation = 0
        self.ws = None
        self._build()

    def buf(self, *shape, dtype=None):
        t = torch.empty(shape, dtype=self.tdt if dtype is None else dtype, device=self.device)
        self._keep.append(t)
        return t

    def zeros(self, *shape, dtype=None):
        t = torch.zeros(shape, dtype=self.tdt if dtype is None else dtype, device=self.device)
        self._keep.append(t)
        return t

    def note_scratch(self, act_elems, w_elems):
        self._scratch_act = max(self._scratch_act, act_elems)
        self._scratch_w = max(self._scratch_w, w_elems)

    def note_ws(self, b):
        self._ws_bytes = max(self._ws_bytes, b)

    def _build(self):
        m, n, c, h, w = self.model, self.n, self.c, self.h, self.w
        f32 = torch.float32
        self.hs, self.ws_ = h // 2 + 1, w // 2 + 1
        self.xs = self.buf(n, self.hs, self.ws_, 128)
        convs = [m.conv1] + [getattr(m, nm) for nm in ("conv2", "conv3", "conv4") if hasattr(m, nm)]
        self.convs = convs
        # conv1 as a 2x2 stride-1 conv over the space-to-depth input
        c1 = convs[0]
        self.w2 = self.buf(c1.weight.shape[0], 128, 2, 2, dtype=f32)
        self.gw2 = self.buf(c1.weight.shape[0], 128, 2, 2, dtype=f32)
        self.pack_steps.append(lambda: check(lib().rtsds_s2d_weight(_p(c1.weight.detach()), c1.weight.shape[0], c, _p(self.w2),
                                                                    ops._s()), "s2d_weight"))
        self.layers = [_Layer(self, c1, lambda: self.w2, self.xs, (n, self.hs, self.ws_, 128), 2, 1, 0)]
        for conv in convs[1:]:
            prev = self.layers[-1]
            self.layers.append(_Layer(self, conv, (lambda cv=conv: cv.weight), prev.y, (n, prev.oh, prev.ow, prev.cout), 4, 2, 1))
        last = self.layers[-1]
        if last.oh < 2 or last.ow < 2:
            raise ops._lib.RtsdsError("input too small for the discriminator")
        self.tapsum = self.buf(n, 16, last.cout, dtype=f32)
        self.out = self.buf(n, dtype=f32)
        self.gxs = None                                   # gradient w.r.t. xs, allocated on first use
        self.d_a = self.zeros(max(self._scratch_act, 1))
        self.d_b = self.zeros(max(self._scratch_act, 1))
        self.dw_scratch = self.zeros(max(self._scratch_w, 1), dtype=f32)
        if self._ws_bytes:
            self.ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
        self.params = list(m.parameters())
        self._grad_numel = sum(p.numel() for p in self.params)

    # ---------------- weights ----------------
    def _params_version(self):
        return (sum(p._version for p in self.params), self.params[0].data_ptr(), weights_epoch.value())

    def refresh_weights(self):
        ver = self._params_version()
        if ver != self._param_version:
            for s in self.pack_steps:
                s()
            self._param_version = ver

    # ---------------- forward ----------------
    def forward(self, x, softmax_in):
        self.refresh_weights()
        self.generation += 1
        self.softmax_in = bool(softmax_in)
        check(lib().rtsds_s2d_fwd(_p(x), self.n, self.c, self.h, self.w, int(self.softmax_in), self.dt, _p(self.xs), ops._s()),
              "s2d_fwd")
        for layer in self.layers:
            layer.forward()
        last = self.layers[-1]
        cls = self.model.classifier
        check(lib().rtsds_disc_cls_fwd(_p(last.y), last.cout, self.dt, self.n, last.oh, last.ow, last.cout,
                                       _p(cls.weight.detach()), _p(cls.bias.detach() if cls.bias is not None else None),
                                       _p(self.tapsum), _p(self.out), ops._s()), "disc_cls_fwd")
        return self.out.clone().view(self.n, 1, 1, 1)

    # ---------------- backward ----------------
    def new_grads(self):
        flat = torch.zeros(self._grad_numel, dtype=torch.float32, device=self.device)
        gw, off = {}, 0
        for p in self.params:
            if p.requires_grad:
                gw[p] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        return flat, gw

    def backward(self, g, gw, need_dx, g_scale=1.0):
        """g: fp32 [n] = dL/dout.  gw: param -> fp32 grad view (missing = frozen).  Returns dL/dx (fp32 NCHW) or None."""
        weights_epoch.note_backward()
        s = ops._s()
        dt = self.dt
        last = self.layers[-1]
        cls = self.model.classifier
        any_conv_grad = any(l.conv.weight in gw or (l.conv.bias is not None and l.conv.bias in gw) for l in self.layers)
        need_chain = need_dx or any_conv_grad
        d_cur, d_nxt = self.d_a, self.d_b
        check(lib().rtsds_disc_cls_bwd(_p(g), float(g_scale), _p(self.tapsum), _p(cls.weight.detach()), _p(last.y), last.cout, dt,
                                       self.n, last.oh, last.ow, last.cout, 1, self.slope,
                                       _p(d_cur) if need_chain else None, last.cout,
                                       _p(gw.get(last.conv.bias)) if last.conv.bias is not None else None,
                                       _p(gw.get(cls.weight)), _p(gw.get(cls.bias)) if cls.bias is not None else None, s),
              "disc_cls_bwd")
        if not need_chain:
            return None
        dx = None
        for li in range(len(self.layers) - 1, -1, -1):
            L = self.layers[li]
            dd = ops.make_conv_desc(L.d.n, L.d.h, L.d.w, L.cin, L.cin, L.cout, L.cout, L.k, L.d.stride, L.d.pad, 1,
                                    in_dtype=dt, out_dtype=dt)
            gwt = gw.get(L.conv.weight)
            if gwt is not None:
                dwp = self.dw_scratch[:L.cout * L.k * L.k * L.cin]
                ops.conv2d_wgrad(dd, L.x, d_cur, dwp, self.use_tc)
                if li == 0:
                    ops.unpack_conv_wgrad(dwp, self.gw2, False)
                    check(lib().rtsds_s2d_weight_grad(_p(self.gw2), L.cout, self.c, _p(gwt), s), "s2d_weight_grad")
                else:
                    ops.unpack_conv_wgrad(dwp, gwt, True)
            lower_needs = need_dx or any(l.conv.weight in gw or (l.conv.bias is not None and l.conv.bias in gw)
                                         for l in self.layers[:li])
            if li == 0:
                if need_dx:
                    if self.gxs is None:
                        self.gxs = self.buf(self.n, self.hs, self.ws_, 128, dtype=torch.float32)
                    dd.res_ld = 128
                    ops.conv2d_dgrad(dd, d_cur, L.wdg, self.gxs, F32, self.use_tc, None, self.ws)
                    dx = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.device)
                    check(lib().rtsds_s2d_bwd(_p(self.gxs), _p(self.xs) if self.softmax_in else None, self.n, self.c, self.h,
                                              self.w, dt, _p(dx), s), "s2d_bwd")
            elif lower_needs:
                prev = self.layers[li - 1]
                dd.res_ld = L.cin
                ops.conv2d_dgrad(dd, d_cur, L.wdg, d_nxt, dt, self.use_tc, None, self.ws)        # dL/dy_prev
                pb = prev.conv.bias
                check(lib().rtsds_act_bwd(_p(d_nxt), prev.cout, _p(prev.y), prev.cout, prev.n_pix, prev.cout, ACT_LRELU,
                                          self.slope, dt, _p(d_cur), prev.cout, _p(gw.get(pb)) if pb is not None else None, s),
                      "act_bwd")
            else:
                break
        return dx


_MAX_SLOTS = 4
_stamp = 0


def _get_plan(model, x):
    """One plan per input shape -- and a further one (up to _MAX_SLOTS) whenever every existing plan of that shape
    still holds the activations of a forward whose backward has not run yet: train.py:447-458 (adversarial_train_2)
    calls the discriminator twice and then backpropagates through both calls at once."""
    global _stamp
    plans = model.__dict__.setdefault("_rtsds_plans", {})
    n, c, h, w = x.shape
    base = (n, c, h, w, model.rtsds_precision, x.device.index)
    oldest = None
    for slot in range(_MAX_SLOTS):
        plan = plans.get(base + (slot,))
        if plan is None:
            plan = DiscPlan(model, n, c, h, w, model.rtsds_precision)
            plan.awaiting_backward = False
            plans[base + (slot,)] = plan
            break
        if not plan.awaiting_backward:
            break
        if oldest is None or plan.slot_stamp < oldest.slot_stamp:
            oldest = plan
    else:
        plan = oldest          # every slot is pending (graphs dropped without backward): recycle the oldest one
    _stamp += 1
    plan.slot_stamp = _stamp
    return plan


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, softmax_in, *params):
        out = plan.forward(x, softmax_in)
        ctx.plan, ctx.gen, ctx.params = plan, plan.generation, params
        plan.awaiting_backward = True
        return out

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        if plan.generation != ctx.gen:
            raise ops._lib.RtsdsError("discriminator backward called after another forward of the same shape reused the "
                                      "plan's saved activations; call backward() before the next forward()")
        plan.awaiting_backward = False
        g = dout.contiguous().view(-1).float()
        flat, gw = plan.new_grads()
        dx = plan.backward(g, gw, ctx.needs_input_grad[1])
        if gw and getattr(plan.model, "rtsds_ddp", False):
            # data parallel: average this pass's parameter gradients over the ranks (one small bucket: 20 K / 2.8 M floats);
            # the input gradient stays local — it belongs to this rank's generator batch
            from . import ddp

            if ddp.is_distributed():
                red = ddp.BucketedAllReduce(flat, {"all": (0, flat.numel())})
                red.ready("all")
                red.finish()
        return (None, dx, None) + tuple(gw.get(p) for p in ctx.params)


def disc_forward(model, x, softmax_in=False):
    """DomainDiscriminator / TinyDomainDiscriminator forward (model.py:53-64 / :77-83) on NCHW fp32 probabilities
    (softmax_in=False, the stock call site) or logits (softmax_in=True: F.softmax(dim=1) of train.py:225 fused)."""
    if not x.is_cuda and not ops._lib.dry_run():
        raise ops._lib.RtsdsError("discriminator forward needs a CUDA tensor: rtsds_b200 has no CPU fallback")
    if x.dim() != 4:
        raise ValueError(f"expected input [N,C,H,W], got {tuple(x.shape)}")
    xin = x if x.dtype == torch.float32 else x.float()
    xin = xin.contiguous()
    plan = _get_plan(model, xin)
    params = tuple(plan.params)
    if torch.is_grad_enabled() and (xin.requires_grad or any(p.requires_grad for p in params)):
        return _DiscFn.apply(plan, xin, bool(softmax_in), *params)
    with torch.no_grad():
        return plan.forward(xin, softmax_in)


def bce_with_logits_const(logit: torch.Tensor, target: float, scale: float = 1.0):
    """nn.BCEWithLogitsLoss()(logit, full_like(logit, target)) * scale as one kernel (train.py:228-229)."""
    return _BceFn.apply(logit, float(target), float(scale))


class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit, target, scale):
        x = logit.contiguous().view(-1).float()
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        dl = torch.empty_like(x)
        check(lib().rtsds_bce_logits(_p(x), x.numel(), target, scale, _p(loss), _p(dl), ops._s()), "bce_logits")
        ctx.save_for_backward(dl)
        ctx.shape = logit.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        (dl,) = ctx.saved_tensors
        return (dl * dloss).view(ctx.shape), None, None
