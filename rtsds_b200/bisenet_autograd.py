"""Autograd boundary of the train-mode BiSeNet path.

`bisenet_train_forward` is what `BiSeNet.forward` calls in train mode: it returns the
reference's `(result, cx1_sup, cx2_sup)` tuple (build_bisenet.py:169-170) as ordinary fp32 NCHW
autograd tensors, so the stock call sites work unchanged (`criterion(out, target)`,
`loss.backward()`, `F.softmax(out)` -> discriminator, `.detach()`, `.max(1)`; train.py:86-106,
:199-233).  Their backward is one hand-written pass (rtsds_b200/bisenet_train.py).

`bisenet_fused_ce` is the fast path for the supervised step: bilinear resize + CrossEntropyLoss
(ignore_index) + argmax + pixel-accuracy of all three heads are evaluated from the 1/8-resolution
logits in one kernel each, and the backward starts from there — the three [N,19,H,W] fp32 tensors
(119 MB/image at 512x1024) and their gradients are never materialised.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import _p, check, lib


_MAX_SLOTS = 2
_stamp = 0


def _get_train_plan(model, x):
    """One train plan per input shape; a second one only when the first still holds the activations of a forward whose
    backward has not run (train.py:199-213 with equal source / target sizes: two generator forwards, one backward)."""
    from .bisenet_train import BiSeNetTrainPlan

    global _stamp
    plans = model.__dict__.setdefault("_rtsds_train_plans", {})
    n, _, h, w = x.shape
    base = (n, h, w, model.rtsds_precision, x.device.index)
    oldest = None
    for slot in range(_MAX_SLOTS):
        plan = plans.get(base + (slot,))
        if plan is None:
            plan = BiSeNetTrainPlan(model, n, h, w, model.rtsds_precision)
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
    if counters is None or (counters and counters[0].device != model.conv.weight.device):
        counters = [m.num_batches_tracked for m in model.modules()
                    if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None]
        model.__dict__["_rtsds_bn_counters"] = counters
    if counters:
        torch._foreach_add_(counters, 1)


def _grad_tuple(plan, params, gw):
    return tuple(None if (p in plan.unused or p not in gw) else gw[p] for p in params)


def _aliases(plan, prev) -> bool:
    """True when every live parameter's .grad still IS the view of `prev` this module handed to autograd."""
    base, off = prev.data_ptr(), 0
    for p in plan.params:
        if p.requires_grad and p not in plan.unused:
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * off or g.dtype != torch.float32:
                return False
        off += p.numel()
    return True


def _finish_backward(plan, params):
    """Run the hand-written backward pass and hand the parameter gradients to autograd.

    * Data parallel (model.rtsds_ddp, process group active): the flat fp32 gradient is all-reduced bucket by bucket while
      backward still runs (rtsds_b200/ddp.py).
    * Accumulation — a second backward before optimizer.step(), train.py:213 then :233 — when every p.grad still is the
      view of the flat buffer of the previous backward, the new gradients are added there with ONE kernel and autograd
      receives None (instead of one `add` kernel per tensor: 86 launches per adversarial iteration).
    * model.rtsds_ddp_hold = True on the first of the two backward calls defers its all-reduce: the SUM is reduced once,
      after the second ("allreduce G once after train.py:233"); rtsds_b200.train_steps.adversarial_step sets it.
    """
    from . import ddp

    model = plan.model
    use_ddp = getattr(model, "rtsds_ddp", False) and ddp.is_distributed()
    hold = use_ddp and getattr(model, "rtsds_ddp_hold", False)
    pending = use_ddp and model.__dict__.get("_rtsds_grad_pending_reduce", False)
    prev = model.__dict__.get("_rtsds_grad_flat")
    accumulate = prev is not None and prev.numel() == plan._grad_numel and prev.device == plan.device and _aliases(plan, prev)
    if not accumulate:
        prev = model.__dict__["_rtsds_grad_flat"] = None     # let the allocator hand the same block out again (stable pointers)
    flat, gw = plan.new_grads()
    if use_ddp and getattr(plan, "_buckets", None) is None:
        plan._buckets = ddp.param_buckets(list(model.named_parameters()), ddp.BISENET_GROUPS)
    if use_ddp and not hold and not pending:
        red = ddp.BucketedAllReduce(flat, plan._buckets)
        plan.backward_from_dz(gw, red.ready)
        red.finish()
    else:
        plan.backward_from_dz(gw)
    if accumulate:
        prev.add_(flat)
        delivered = tuple(None for _ in params)
        total = prev
    else:
        if pending:
            raise ops._lib.RtsdsError("rtsds_ddp_hold: the gradients of the held backward were replaced before the next backward; "
                                      "keep p.grad in place between the two backward calls of one iteration")
        model.__dict__["_rtsds_grad_flat"] = flat
        delivered = _grad_tuple(plan, params, gw)
        total = flat
    if hold:
        model.__dict__["_rtsds_grad_pending_reduce"] = True
    elif pending:
        red = ddp.BucketedAllReduce(total, plan._buckets)        # the accumulated sum, once
        red.finish()
        model.__dict__["_rtsds_grad_pending_reduce"] = False
    return delivered


class _BiSeNetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, *params):
        plan.forward(x)
        outs = plan.logits()
        ctx.plan, ctx.gen, ctx.params = plan, plan.generation, params
        # heads nobody differentiates (the adversarial loss uses the main head only, train.py:218-229) arrive as None in
        # backward instead of as materialised full-resolution zero tensors
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        plan = ctx.plan
        if plan.generation != ctx.gen:
            raise ops._lib.RtsdsError("BiSeNet backward called after another forward of the same shape reused the plan's "
                                      "saved activations; call backward() before the next forward()")
        plan.awaiting_backward = False
        s = ops._s()
        main, aux = plan.out_sizes()
        for i, d in enumerate(douts):
            oh, ow = main if i == 0 else aux
            if d is None:
                plan.dz[i].zero_()
            else:
                d = d.contiguous()
                check(lib().rtsds_resize_to_nchw_bwd(_p(d), plan.n, plan.nc, oh, ow, plan.h8, plan.w8, _p(plan.dz[i]), 32, s),
                      "resize_to_nchw_bwd")
        return (None, None) + _finish_backward(plan, ctx.params)


def bisenet_train_forward(model, x):
    plan = _get_train_plan(model, x)
    params = tuple(plan.params)
    if not torch.is_grad_enabled():
        with torch.no_grad():
            plan.forward(x)
            outs = tuple(plan.logits())
    else:
        outs = _BiSeNetTrainFn.apply(plan, x, *params)
        plan.awaiting_backward = bool(outs[0].requires_grad)
    _bump_bn_counters(model)
    return outs


class _BiSeNetFusedCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, target, ignore_index, *params):
        plan.forward(x)
        main, aux = plan.out_sizes()
        plan.acc.zero_()
        pred = torch.empty((plan.n,) + main, dtype=torch.int64, device=plan.device)
        # ~x8 heads: loss, argmax and the unnormalised gradient in ONE pass over the labels (csrc/loss.cu)
        one_pass = torch.is_grad_enabled() or any(ctx.needs_input_grad)
        ctx.one_pass = []
        for i, z in enumerate((plan.z, plan.z1, plan.z2)):
            oh, ow = main if i == 0 else aux
            fused = one_pass and ops.resize_ce_fused_supported(plan.h8, plan.w8, plan.nc, oh, ow)
            ctx.one_pass.append(fused)
            if fused:
                plan.dz[i].zero_()
                ops.resize_ce_fused(z, plan.n, plan.h8, plan.w8, plan.nc, 32, oh, ow, target, ignore_index, plan.acc[i],
                                    pred if i == 0 else None, plan.dz[i])
            else:
                ops.resize_ce_argmax_fwd(z, plan.n, plan.h8, plan.w8, plan.nc, 32, oh, ow, target, ignore_index, plan.acc[i],
                                         pred if i == 0 else None)
        stats = plan.acc.clone()
        ctx.den = stats[:, 1]
        per_head = (plan.acc[:, 0] / plan.acc[:, 1]).float()          # mean over valid pixels, per head
        if getattr(plan.model, "rtsds_loss_norm", "local") == "global" and getattr(plan.model, "rtsds_ddp", False):
            # nn.DataParallel semantics (utils.py:104-105): the criterion sees the GATHERED batch, i.e. every head's loss is
            # sum(-log p) over all ranks / valid pixels of all ranks.  One tiny all-reduce (3 x 2 doubles) gives every rank
            # that loss; its gradient is this rank's unnormalised gradient * world / global count, which the gradient
            # all-reduce (AVG) turns into sum over ranks / global count (SURVEY 8e `global_valid_count`).
            from . import ddp

            if ddp.is_distributed():
                import torch.distributed as dist

                tot = plan.acc[:, :2].clone()
                dist.all_reduce(tot)
                per_head = (tot[:, 0] / tot[:, 1]).float()
                ctx.den = tot[:, 1] / dist.get_world_size()
        loss = per_head.sum()
        ctx.plan, ctx.gen, ctx.params = plan, plan.generation, params
        ctx.target, ctx.ignore_index = target, ignore_index
        ctx.stats = stats
        ctx.mark_non_differentiable(pred, stats)
        return loss, pred, stats

    @staticmethod
    def backward(ctx, dloss, _dpred, _dstats):
        plan = ctx.plan
        if plan.generation != ctx.gen:
            raise ops._lib.RtsdsError("BiSeNet backward called after another forward reused the plan's saved activations")
        plan.awaiting_backward = False
        main, aux = plan.out_sizes()
        plan.gscale.copy_((dloss.double() / ctx.den).float())
        for i, z in enumerate((plan.z, plan.z1, plan.z2)):
            oh, ow = main if i == 0 else aux
            if ctx.one_pass[i]:
                ops.scale_by_device_scalar(plan.dz[i], plan.gscale[i:i + 1])
            else:
                plan.dz[i].zero_()
                ops.resize_ce_bwd(z, plan.n, plan.h8, plan.w8, plan.nc, 32, oh, ow, ctx.target, ctx.ignore_index,
                                  plan.gscale[i:i + 1], plan.dz[i])
        return (None, None, None, None) + _finish_backward(plan, ctx.params)


def bisenet_fused_ce(model, x, target, ignore_index=255, return_logits=False):
    """Sum of the three heads' CrossEntropyLoss(ignore_index) (train.py:86-92) without materialising
    the full-resolution logits.  Returns (loss, argmax of the main head [N,H,W] int64,
    stats [3,4] float64 = per head {sum of -log p, valid pixels, pixels with argmax == target, 0}).
    return_logits=True appends the main head's full-resolution logits, DETACHED (what the adversarial
    loop feeds to the discriminator after `.detach()`, train.py:242,245)."""
    if not model.training:
        raise ops._lib.RtsdsError("bisenet_fused_ce is the train-mode fast path; call model.train() first")
    if not x.is_cuda and not ops._lib.dry_run():
        raise ops._lib.RtsdsError("bisenet_fused_ce needs CUDA tensors: rtsds_b200 has no CPU fallback")
    x = x.float().contiguous()
    target = target.contiguous()
    assert target.dtype == torch.int64
    plan = _get_train_plan(model, x)
    out = _BiSeNetFusedCEFn.apply(plan, x, target, int(ignore_index), *plan.params)
    plan.awaiting_backward = bool(out[0].requires_grad)
    _bump_bn_counters(model)
    if return_logits:
        (oh, ow), _ = plan.out_sizes()
        logits = torch.empty((plan.n, plan.nc, oh, ow), dtype=torch.float32, device=plan.device)
        ops.resize_to_nchw(plan.z, plan.n, plan.h8, plan.w8, plan.nc, 32, logits)
        return out + (logits,)
    return out
