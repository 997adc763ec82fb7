"""Fused multi-tensor optimizers: drop-in replacements for the `torch.optim.Adam` / `torch.optim.SGD`
objects the reference builds at main.py:110-120 and steps at train.py:96 / :269-270.

One kernel launch per step (csrc/optim.cu) updates EVERY parameter tensor and, for the conv weights of an attached
training plan, writes the new values straight into the packed bf16 tensor-core operands (forward and dgrad layouts) the
next forward / backward read — the separate re-pack launches of the plan and torch's ~30 optimizer / fill kernels
disappear from the step.

They are real `torch.optim.Optimizer`s: `param_groups` (so `utils.poly_lr_scheduler`, utils.py:33-48, which rewrites
`param_groups[0]['lr']` only, works unchanged), `zero_grad`, `state_dict` / `load_state_dict` with torch's state keys
(`step`, `exp_avg`, `exp_avg_sq` / `momentum_buffer`).  `fuse_(optimizer)` converts an existing stock optimizer object in
place, which is how the reference's `train(model, optimizer, ...)` call sites get the fused step without being edited.
"""
from __future__ import annotations

import ctypes as C
import math
import types

import torch

from . import _lib, ops
from ._lib import OptHyper, OptJob, check, lib

ADAM, SGD = 0, 1
MAX_GROUPS = 8


class _Fused(torch.optim.Optimizer):
    _kind = ADAM

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        if len(self.param_groups) > MAX_GROUPS:
            raise ValueError(f"at most {MAX_GROUPS} parameter groups")
        self._plans = []
        self._table_key = None
        self._tables = {}

    # ------------------------------------------------------------------ plans whose packed operands this step refreshes
    def attach(self, plan) -> None:
        """`plan`: a training plan with batched weight packing (`pack_jobs` = [(conv, packed tensor, kind)]).  After every
        step its packed conv operands are current, so its own re-pack is skipped; one plan can be attached per parameter
        (further plans of the same model simply re-pack themselves as before)."""
        if plan not in self._plans and hasattr(plan, "pack_jobs"):
            self._plans.append(plan)
            self._table_key = None

    def _auto_attach(self):
        """Training plans register themselves when they are built (weights_epoch.register_plan); one whose conv weights
        this optimizer owns, and that no attached plan already covers, is attached without the caller doing anything."""
        from . import weights_epoch

        live = weights_epoch.live_plans()
        if len(live) == getattr(self, "_n_live_seen", -1):
            return
        self._n_live_seen = len(live)
        mine = {p for g in self.param_groups for p in g["params"]}
        covered = {conv.weight for pl in self._plans for conv, _, _ in pl.pack_jobs}
        for plan in live:
            if plan in self._plans or not getattr(plan, "pack_jobs", None):
                continue
            ws = {conv.weight for conv, _, _ in plan.pack_jobs}
            if ws <= mine and not (ws & covered):
                self.attach(plan)
                covered |= ws

    def _pack_targets(self):
        tgt = {}
        for plan in self._plans:
            for conv, out, kind in plan.pack_jobs:
                slot = tgt.setdefault(conv.weight, {"plan": plan})
                if slot["plan"] is plan and kind not in slot:
                    slot[kind] = out
        return tgt

    # ------------------------------------------------------------------ state
    def _moments(self, p, group):
        st = self.state[p]
        if not st:
            st["step"] = 0
            if self._kind == ADAM:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif group["momentum"] != 0:
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _build_table(self, entries, targets, dev):
        """entries: [(param, grad, state, group index)] -> device job table + block prefix sums."""
        n = len(entries)
        jobs = (OptJob * n)()
        first = (C.c_int * n)()
        blocks = 0
        pack_dtype = None
        for i, (p, g, st, gi) in enumerate(entries):
            j = jobs[i]
            j.p, j.g, j.numel, j.group = p.data_ptr(), g.data_ptr(), p.numel(), gi
            if self._kind == ADAM:
                j.m, j.v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            else:
                mb = st.get("momentum_buffer")
                j.m, j.v = (mb.data_ptr() if mb is not None else None), None
            t = targets.get(p)
            if t is not None and p.dim() == 4:
                co, ci, kh, kw = p.shape
                j.taps, j.cout, j.cin = kh * kw, co, ci
                fwd, dg = t.get(0), t.get(1)
                if fwd is not None:
                    j.out_fwd, j.cout_pad, j.cin_pad_fwd = fwd.data_ptr(), fwd.shape[0], fwd.shape[2]
                    pack_dtype = fwd.dtype
                if dg is not None:
                    j.out_dgrad, j.cin_pad_dgrad, j.ck = dg.data_ptr(), dg.shape[0], dg.shape[2]
                    pack_dtype = dg.dtype
            first[i] = blocks
            blocks += int(lib().rtsds_optim_job_blocks(C.byref(j)))
        jt = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(dev)
        ft = torch.frombuffer(bytearray(bytes(first)), dtype=torch.uint8).to(dev)
        return dict(jobs=jt, first=ft, n=n, blocks=blocks, dtype=ops.dtype_code(pack_dtype) if pack_dtype is not None else ops.F32)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if _lib.dry_run():
            return loss
        by_step = {}
        dev = None
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda:
                    raise _lib.RtsdsError("rtsds_b200 optimizers need CUDA parameters (there is no CPU fallback)")
                if g.is_sparse or g.dtype != torch.float32 or p.dtype != torch.float32:
                    raise _lib.RtsdsError("fused optimizer: dense fp32 parameters and gradients only")
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                st = self._moments(p, group)
                by_step.setdefault(int(st["step"]), []).append((p, g, st, gi))
                dev = p.device
        if not by_step:
            return loss
        self._auto_attach()
        for t_prev, entries in by_step.items():
            key = (t_prev == 0 and self._kind == SGD, tuple((p.data_ptr(), g.data_ptr()) for p, g, _, _ in entries),
                   tuple(id(pl) for pl in self._plans))
            tab = self._tables.get(key)
            if tab is None:
                if len(self._tables) >= 4:                    # gradient buffers normally cycle through one or two addresses
                    self._tables.pop(next(iter(self._tables)))
                tab = self._tables[key] = self._build_table(entries, self._pack_targets(), dev)
            h = OptHyper()
            h.kind, h.first_step = self._kind, int(t_prev == 0)
            t = t_prev + 1
            for gi, group in enumerate(self.param_groups):
                h.lr[gi] = float(group["lr"])
                h.weight_decay[gi] = float(group.get("weight_decay", 0.0))
            g0 = self.param_groups[0]
            if self._kind == ADAM:
                b1, b2 = g0["betas"]
                h.beta1, h.beta2, h.eps = b1, b2, g0["eps"]
                h.inv_bias_correction1 = 1.0 / (1.0 - b1 ** t)
                h.inv_bias_correction2_sqrt = 1.0 / math.sqrt(1.0 - b2 ** t)
            else:
                h.momentum = g0["momentum"]
            with torch.cuda.device(dev):
                check(lib().rtsds_optim_step(tab["jobs"].data_ptr(), tab["first"].data_ptr(), tab["n"], tab["blocks"], C.byref(h),
                                             tab["dtype"], ops._s()), "optim_step")
            for _, _, st, _ in entries:
                st["step"] = t
        # the attached plans' batched conv operands are current; whatever else they pack (fused stems, taps-as-N forms)
        # is refreshed here, and their next forward skips its own re-pack
        for plan in self._plans:
            for s in plan.pack_steps:
                s()
            plan._param_version = plan._params_version()
        return loss


class FusedAdam(_Fused):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) — non-amsgrad, L2 weight decay (main.py:116-118)."""

    _kind = ADAM

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        for g in self.param_groups:
            if g["betas"] != self.param_groups[0]["betas"] or g["eps"] != self.param_groups[0]["eps"]:
                raise ValueError("betas / eps must be the same for every parameter group")


class FusedSGD(_Fused):
    """torch.optim.SGD(params, lr, momentum, weight_decay) — dampening 0, no Nesterov (main.py:119-120)."""

    _kind = SGD

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        for g in self.param_groups:
            if g["momentum"] != self.param_groups[0]["momentum"]:
                raise ValueError("momentum must be the same for every parameter group")


def fuse_(optimizer, *plans):
    """Convert a stock `torch.optim.Adam` / `torch.optim.SGD` OBJECT (as built at main.py:116-120) to the fused step in
    place: the caller's reference, its `param_groups` and `state` stay the same objects; only `step` is re-bound.
    Unsupported configurations (amsgrad, Nesterov, dampening, maximize, ...) are left untouched and returned as they are."""
    if isinstance(optimizer, _Fused):
        fused = optimizer
    else:
        g0 = optimizer.param_groups[0]
        if type(optimizer) is torch.optim.Adam and not any(g.get("amsgrad") or g.get("maximize") or g.get("capturable") or
                                                           g.get("decoupled_weight_decay") for g in optimizer.param_groups):
            fused = FusedAdam.__new__(FusedAdam)
        elif type(optimizer) is torch.optim.SGD and not any(g.get("nesterov") or g.get("dampening") or g.get("maximize")
                                                            for g in optimizer.param_groups):
            fused = FusedSGD.__new__(FusedSGD)
        else:
            return optimizer
        if len(optimizer.param_groups) > MAX_GROUPS or any(isinstance(g["lr"], torch.Tensor) for g in optimizer.param_groups):
            return optimizer
        fused.__dict__.update(optimizer.__dict__)            # same param_groups list, same state dict, same hooks
        fused._plans, fused._table_key, fused._tables = [], None, {}
        for st in fused.state.values():                       # torch keeps `step` as a tensor; the fused step counts in Python
            if isinstance(st.get("step"), torch.Tensor):
                st["step"] = int(st["step"].item())
        optimizer.step = types.MethodType(lambda self, closure=None, _f=fused: _f.step(closure), optimizer)
        optimizer._rtsds_fused = fused
        del g0
    for plan in plans:
        fused.attach(plan)
    return optimizer
