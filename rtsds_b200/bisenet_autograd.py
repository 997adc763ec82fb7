"""Train-mode BiSeNet forward/backward as one autograd node
(reference models/bisenet/build_bisenet.py:141-172, train.py:77-95).

forward: batch-statistics BatchNorm (running buffers updated), auxiliary heads,
returns `(result, cx1_sup, cx2_sup)` as fp32 NCHW tensors.
"""
from __future__ import annotations

import torch

from . import ops
from .bisenet_engine import _get_plan


def bisenet_train_forward(model, x):
    plan = _get_plan(model, x, True)
    with torch.no_grad():
        plan.forward_lowres(x, use_graph=False)
        outs = (plan.logits(plan.z), plan.logits_aux(plan.z1), plan.logits_aux(plan.z2))
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None:
                m.num_batches_tracked += 1
    return outs
