"""ORACLE (test infrastructure, not product): CPU restatement of the reference's
BiSeNet-ResNet18 forward in plain functional PyTorch fp32.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.  The product path (rtsds_b200, models/) never
does and has no CPU fallback.

The reference algorithm lives partly in third-party code that is not under
/root/reference: torchvision==0.18.0 (pinned, requirements.txt:85) ResNet-18
(`BasicBlock`, resnet.py:59-105; `_forward_impl` :266-283) and torch==2.3.0 ATen
operators (requirements.txt:81).  Their published algorithms are restated here
with torch.nn.functional on CPU.  The reference ships no tests or golden
vectors; parity is pinned instead against outputs of the reference
itself, run in the build container on seeded weights and committed as
tests/golden/bisenet_*.npz by oracle/gen_golden.py.

Every function cites the reference lines it follows (paths relative to
/root/reference).  `sd` is a state_dict with the reference's key names; running
BatchNorm buffers are updated in place in train mode, exactly like nn.BatchNorm2d.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn(x, sd, prefix, train, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d forward (batch statistics + running-buffer update in train mode)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=train, momentum=momentum, eps=eps)


def conv_block(x, sd, prefix, stride, train):
    """ConvBlock.forward — models/bisenet/build_bisenet.py:16-18 (conv no-bias k3 p1 -> BN -> ReLU)."""
    x = F.conv2d(x, sd[prefix + ".conv1.weight"], None, stride=stride, padding=1)
    return F.relu(_bn(x, sd, prefix + ".bn", train))


def spatial_path(x, sd, train, prefix="saptial_path"):
    """Spatial_path.forward — build_bisenet.py:28-32."""
    x = conv_block(x, sd, prefix + ".convblock1", 2, train)
    x = conv_block(x, sd, prefix + ".convblock2", 2, train)
    return conv_block(x, sd, prefix + ".convblock3", 2, train)


def basic_block(x, sd, prefix, stride, train):
    """torchvision BasicBlock.forward (resnet.py:86-105)."""
    out = F.conv2d(x, sd[prefix + ".conv1.weight"], None, stride=stride, padding=1)
    out = F.relu(_bn(out, sd, prefix + ".bn1", train))
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], None, stride=1, padding=1)
    out = _bn(out, sd, prefix + ".bn2", train)
    if prefix + ".downsample.0.weight" in sd:
        idn = F.conv2d(x, sd[prefix + ".downsample.0.weight"], None, stride=stride)
        idn = _bn(idn, sd, prefix + ".downsample.1", train)
    else:
        idn = x
    return F.relu(out + idn)


def context_path_r18(x, sd, train, prefix="context_path.features"):
    """resnet18.forward — models/bisenet/build_contextpath.py:18-29."""
    x = F.conv2d(x, sd[prefix + ".conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, prefix + ".bn1", train))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = basic_block(x, sd, f"{prefix}.layer{li}.0", stride, train)
        x = basic_block(x, sd, f"{prefix}.layer{li}.1", 1, train)
        feats.append(x)
    f3, f4 = feats[2], feats[3]
    tail = torch.mean(f4, 3, keepdim=True)
    tail = torch.mean(tail, 2, keepdim=True)
    return f3, f4, tail


def bottleneck(x, sd, prefix, stride, train):
    """torchvision Bottleneck.forward (resnet.py:143-163; v1.5: the stride sits on the 3x3 conv2)."""
    out = F.relu(_bn(F.conv2d(x, sd[prefix + ".conv1.weight"]), sd, prefix + ".bn1", train))
    out = F.relu(_bn(F.conv2d(out, sd[prefix + ".conv2.weight"], None, stride=stride, padding=1), sd, prefix + ".bn2", train))
    out = _bn(F.conv2d(out, sd[prefix + ".conv3.weight"]), sd, prefix + ".bn3", train)
    if prefix + ".downsample.0.weight" in sd:
        idn = _bn(F.conv2d(x, sd[prefix + ".downsample.0.weight"], None, stride=stride), sd, prefix + ".downsample.1", train)
    else:
        idn = x
    return F.relu(out + idn)


def context_path_r101(x, sd, train, prefix="context_path.features"):
    """resnet101.forward — models/bisenet/build_contextpath.py:45-56 (Bottleneck blocks [3, 4, 23, 3])."""
    x = F.conv2d(x, sd[prefix + ".conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, prefix + ".bn1", train))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, stride, blocks in ((1, 1, 3), (2, 2, 4), (3, 2, 23), (4, 2, 3)):
        for b in range(blocks):
            x = bottleneck(x, sd, f"{prefix}.layer{li}.{b}", stride if b == 0 else 1, train)
        feats.append(x)
    f3, f4 = feats[2], feats[3]
    tail = torch.mean(f4, 3, keepdim=True)
    tail = torch.mean(tail, 2, keepdim=True)
    return f3, f4, tail


def arm(x, sd, prefix, train):
    """AttentionRefinementModule.forward — build_bisenet.py:44-53."""
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.conv2d(g, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"])
    g = torch.sigmoid(_bn(g, sd, prefix + ".bn", train))
    return torch.mul(x, g)


def ffm(sx, cx, sd, train, prefix="feature_fusion_module"):
    """FeatureFusionModule.forward — build_bisenet.py:71-81."""
    x = torch.cat((sx, cx), dim=1)
    feature = conv_block(x, sd, prefix + ".convblock", 1, train)
    a = F.adaptive_avg_pool2d(feature, 1)
    a = F.relu(F.conv2d(a, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"]))
    a = torch.sigmoid(F.conv2d(a, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"]))
    return torch.add(torch.mul(feature, a), feature)


def bisenet_forward(x, sd, train, with_interpolation=True, return_intermediates=False):
    """BiSeNet.forward — build_bisenet.py:141-172.  train -> (result, cx1_sup, cx2_sup); eval -> result."""
    sx = spatial_path(x, sd, train)
    # BiSeNet(num_classes, 'resnet101') (build_bisenet.py:95-102) is recognised by its Bottleneck conv3 weights
    r101 = "context_path.features.layer1.0.conv3.weight" in sd
    cx1, cx2, tail = (context_path_r101 if r101 else context_path_r18)(x, sd, train)
    f3, f4 = cx1, cx2
    cx1 = arm(cx1, sd, "attention_refinement_module1", train)
    cx2 = arm(cx2, sd, "attention_refinement_module2", train)
    cx2 = torch.mul(cx2, tail)
    cx1 = F.interpolate(cx1, size=sx.shape[-2:], mode="bilinear")
    cx2 = F.interpolate(cx2, size=sx.shape[-2:], mode="bilinear")
    cx = torch.cat((cx1, cx2), dim=1)
    if train:
        cx1_sup = F.conv2d(cx1, sd["supervision1.weight"], sd["supervision1.bias"])
        cx2_sup = F.conv2d(cx2, sd["supervision2.weight"], sd["supervision2.bias"])
        cx1_sup = F.interpolate(cx1_sup, size=x.shape[-2:], mode="bilinear")
        cx2_sup = F.interpolate(cx2_sup, size=x.shape[-2:], mode="bilinear")
    result = ffm(sx, cx, sd, train)
    if with_interpolation:
        result = F.interpolate(result, scale_factor=8, mode="bilinear")
        result = F.conv2d(result, sd["conv.weight"], sd["conv.bias"])
    if return_intermediates:
        return dict(sx=sx, f3=f3, f4=f4, tail=tail, cx=cx, result=result)
    if train:
        return result, cx1_sup, cx2_sup
    return result


def ce_loss(logits, target, ignore_index):
    """nn.CrossEntropyLoss(ignore_index) — main.py:124-130, train.py:86-92."""
    return F.cross_entropy(logits, target, ignore_index=ignore_index)
