"""ORACLE (test infrastructure, not product): CPU restatement of the reference's domain
discriminators and of the adversarial losses around them, in plain functional PyTorch fp32.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this; the product path (rtsds_b200, models/) never does.

Restates (paths relative to /root/reference):
  models/domain_shift/adversarial/model.py:53-64   DomainDiscriminator.forward
  models/domain_shift/adversarial/model.py:77-83   TinyDomainDiscriminator.forward
  models/domain_shift/adversarial/model.py:9-17    GradientReversalFunction
  train.py:225-229, 245-247, 256-258               softmax -> D -> BCEWithLogits vs ones / zeros
The reference ships no tests; parity is pinned against outputs of the reference itself on seeded
weights (tests/golden/discriminators.npz, written by oracle/gen_golden.py in the build container).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def discriminator_forward(x, sd, slope=0.2):
    """x: [N,C,H,W] class probabilities -> [N,1,1,1] logits.  Works for both topologies: every
    `conv{i}` present in `sd` is applied (4x4 s2 p1 + bias + LeakyReLU), then `classifier`, then the
    global average pool (model.py:54-60 / :78-81)."""
    i = 1
    while f"conv{i}.weight" in sd:
        x = F.leaky_relu(F.conv2d(x, sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], stride=2, padding=1), slope)
        i += 1
    x = F.conv2d(x, sd["classifier.weight"], sd["classifier.bias"], stride=2, padding=1)
    return F.adaptive_avg_pool2d(x, (1, 1))


def adversarial_bce(logits, sd, target, scale=1.0):
    """scale * BCEWithLogitsLoss(D(softmax(logits, 1)), full(target)) — train.py:225-229 (target 1, scale
    lambda/iterations), :245-247 (1), :256-258 (0)."""
    p = discriminator_forward(F.softmax(logits, dim=1), sd)
    return scale * F.binary_cross_entropy_with_logits(p, torch.full_like(p, float(target))), p


# ---- ideal-bf16 emulation --------------------------------------------------------------------------
# What an IDEAL bf16 pipeline (exact fp32 accumulation, tensors rounded to bf16 only where the CUDA path
# stores them in bf16) gives on the CPU.  Its distance from the fp32 oracle is the round-off floor of
# bf16 itself — e.g. LeakyReLU masks of activations within round-off of zero flip — against which the
# bf16 CUDA path is judged (tests/test_gpu_disc.py).
class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def discriminator_forward_bf16(x, sd, slope=0.2):
    """discriminator_forward with bf16 storage of: the probability map, conv weights, every activation and every
    activation / pre-activation gradient (the classifier and its pooled output stay fp32)."""
    x = _RoundFwd.apply(x)
    i = 1
    while f"conv{i}.weight" in sd:
        raw = F.conv2d(x, _RoundFwd.apply(sd[f"conv{i}.weight"]), sd[f"conv{i}.bias"], stride=2, padding=1)
        x = _RoundBoth.apply(F.leaky_relu(_RoundBwd.apply(raw), slope))
        i += 1
    x = F.conv2d(x, sd["classifier.weight"], sd["classifier.bias"], stride=2, padding=1)
    return F.adaptive_avg_pool2d(x, (1, 1))
