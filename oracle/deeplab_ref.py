"""ORACLE (test infrastructure, not product): CPU restatement of the reference's DeepLabV2-ResNet101
(models/deeplabv2/deeplabv2.py) in plain functional PyTorch fp32, optionally with ideal-bf16 rounding of
the tensors the CUDA path stores in bf16.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The reference ships no tests; parity is pinned against outputs of the reference itself on seeded weights
(tests/golden/deeplab_*.npz, written by oracle/gen_golden.py in the build container).

Restates (paths relative to /root/reference/models/deeplabv2/deeplabv2.py):
  :30-47   Bottleneck.forward      conv1x1(stride) -> BN -> ReLU -> conv3x3(dilation) -> BN -> ReLU -> conv1x1 -> BN
                                   (+ downsample(x) | x) -> ReLU
  :62-66   ClassifierModule.forward  sum of 4 dilated 3x3 convs (6, 12, 18, 24) with bias
  :77-84   layers [3,4,23,3]; layer2 stride 2; layer3 / layer4 stride 1, dilation 2 / 4; every first block has a
           1x1 downsample (+BN); MaxPool2d(3, 2, 1, ceil_mode=True)
  :113-131 ResNetMulti.forward     -> bilinear resize to the input size; train: (x, None, None)
BatchNorm runs in batch-statistics mode when `train` (its affine parameters are frozen, :15-27, but the layer
itself is in train mode) and updates the running buffers in `sd` in place.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LAYERS = (("layer1", 64, 3, 1, 1), ("layer2", 128, 4, 2, 1), ("layer3", 256, 23, 1, 2), ("layer4", 512, 3, 1, 4))
ASPP_DILATIONS = (6, 12, 18, 24)


def _r(t, bf16):
    return t.to(torch.bfloat16).float() if bf16 else t


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding in forward, identity in backward (ideal-bf16 emulation with an exact fp32 backward)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def _rs(t, bf16):
    return _RoundSTE.apply(t) if bf16 else t


def _bn(x, sd, prefix, train, eps=1e-5, momentum=0.1):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=train, momentum=momentum, eps=eps)


def bottleneck(x, sd, prefix, stride, dilation, train, bf16=False):
    """Bottleneck.forward — deeplabv2.py:30-47.  bf16: round conv weights, the raw conv outputs (train) and every
    activation to bf16, as the CUDA path stores them."""
    raw = (lambda t: _rs(t, bf16 and train))
    out = raw(F.conv2d(x, _rs(sd[prefix + ".conv1.weight"], bf16), None, stride=stride))
    out = _rs(F.relu(_bn(out, sd, prefix + ".bn1", train)), bf16)
    out = raw(F.conv2d(out, _rs(sd[prefix + ".conv2.weight"], bf16), None, stride=1, padding=dilation, dilation=dilation))
    out = _rs(F.relu(_bn(out, sd, prefix + ".bn2", train)), bf16)
    out = raw(F.conv2d(out, _rs(sd[prefix + ".conv3.weight"], bf16), None))
    out = _bn(out, sd, prefix + ".bn3", train)
    if prefix + ".downsample.0.weight" in sd:
        res = raw(F.conv2d(x, _rs(sd[prefix + ".downsample.0.weight"], bf16), None, stride=stride))
        res = _rs(_bn(res, sd, prefix + ".downsample.1", train), bf16)
    else:
        res = x
    return _rs(F.relu(out + res), bf16)


def classifier(x, sd, prefix="layer6", bf16=False):
    """ClassifierModule.forward — deeplabv2.py:62-66."""
    out = None
    for i, d in enumerate(ASPP_DILATIONS):
        y = F.conv2d(x, _rs(sd[f"{prefix}.conv2d_list.{i}.weight"], bf16), sd[f"{prefix}.conv2d_list.{i}.bias"], stride=1, padding=d, dilation=d)
        out = y if out is None else out + y
    return out


def deeplab_forward(x, sd, train, bf16=False, return_lowres=False):
    """ResNetMulti.forward — deeplabv2.py:113-131.  Returns the full-resolution logits (the reference returns
    `(x, None, None)` in train mode)."""
    H, W = x.shape[-2:]
    y = F.conv2d(x, sd["conv1.weight"], None, stride=2, padding=3)          # stem stays fp32 on the CUDA path too
    y = _rs(F.relu(_bn(_rs(y, bf16 and train), sd, "bn1", train)), bf16)
    y = F.max_pool2d(y, 3, 2, 1, ceil_mode=True)
    for name, planes, blocks, stride, dil in LAYERS:
        for b in range(blocks):
            y = bottleneck(y, sd, f"{name}.{b}", stride if b == 0 else 1, dil, train, bf16)
    z = classifier(y, sd, bf16=bf16)
    if return_lowres:
        return z
    return F.interpolate(z, size=(H, W), mode="bilinear")
