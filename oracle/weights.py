"""ORACLE (test infrastructure): deterministic, order-independent synthetic
state_dicts with the reference's key names.

Each tensor is drawn from its own CPU generator seeded by crc32(name) ^ seed, so
the same values are produced here, on the GPU box, and inside oracle/gen_golden.py
(where they are loaded into the REAL reference modules) regardless of module
construction order.  Shapes restate the reference's constructors:
  BiSeNet-R18: models/bisenet/build_bisenet.py:85-119 + torchvision resnet18
  discriminators: models/domain_shift/adversarial/model.py:39-52,69-76
"""
from __future__ import annotations

import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _conv_w(name, seed, co, ci, k, gain=2.0):
    fan_in = ci * k * k
    return torch.randn(co, ci, k, k, generator=_gen(name, seed)) * (gain / fan_in) ** 0.5


def _vec(name, seed, c, mean=0.0, std=0.05):
    return mean + std * torch.randn(c, generator=_gen(name, seed))


def _bn(sd, prefix, seed, c):
    sd[prefix + ".weight"] = 0.6 + 0.8 * torch.rand(c, generator=_gen(prefix + ".weight", seed))
    sd[prefix + ".bias"] = _vec(prefix + ".bias", seed, c, 0.0, 0.1)
    sd[prefix + ".running_mean"] = _vec(prefix + ".running_mean", seed, c, 0.0, 0.1)
    sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=_gen(prefix + ".running_var", seed))
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def resnet18_state(seed: int, prefix: str) -> dict:
    """torchvision resnet18 parameter tree (conv1, bn1, layer1-4 BasicBlocks, fc)."""
    sd = {}
    sd[prefix + ".conv1.weight"] = _conv_w(prefix + ".conv1.weight", seed, 64, 3, 7)
    _bn(sd, prefix + ".bn1", seed, 64)
    inpl = 64
    for li, planes in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in range(2):
            p = f"{prefix}.layer{li}.{bi}"
            cin = inpl if bi == 0 else planes
            sd[p + ".conv1.weight"] = _conv_w(p + ".conv1.weight", seed, planes, cin, 3)
            _bn(sd, p + ".bn1", seed, planes)
            sd[p + ".conv2.weight"] = _conv_w(p + ".conv2.weight", seed, planes, planes, 3)
            _bn(sd, p + ".bn2", seed, planes)
            if bi == 0 and li > 1:
                sd[p + ".downsample.0.weight"] = _conv_w(p + ".downsample.0.weight", seed, planes, cin, 1)
                _bn(sd, p + ".downsample.1", seed, planes)
        inpl = planes
    sd[prefix + ".fc.weight"] = torch.randn(1000, 512, generator=_gen(prefix + ".fc.weight", seed)) * 0.02
    sd[prefix + ".fc.bias"] = _vec(prefix + ".fc.bias", seed, 1000)
    return sd


def bisenet_r18_state(seed: int = 0, num_classes: int = 19) -> dict:
    """All 290 state_dict keys of the reference's BiSeNet(num_classes, 'resnet18'), incl.
    the aliased context_path.{conv1,bn1,layer1..4}.* duplicates."""
    sd = {}
    nc = num_classes
    for i, (ci, co) in enumerate(((3, 64), (64, 128), (128, 256)), start=1):
        p = f"saptial_path.convblock{i}"
        sd[p + ".conv1.weight"] = _conv_w(p + ".conv1.weight", seed, co, ci, 3)
        _bn(sd, p + ".bn", seed, co)
    feats = resnet18_state(seed, "context_path.features")
    sd.update(feats)
    for k, v in feats.items():   # aliases: same tensors under the shortcut attribute names
        rest = k[len("context_path.features."):]
        if rest.split(".")[0] in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4"):
            sd["context_path." + rest] = v
    for i, c in ((1, 256), (2, 512)):
        p = f"attention_refinement_module{i}"
        sd[p + ".conv.weight"] = _conv_w(p + ".conv.weight", seed, c, c, 1)
        sd[p + ".conv.bias"] = _vec(p + ".conv.bias", seed, c)
        _bn(sd, p + ".bn", seed, c)
        sd[f"supervision{i}.weight"] = _conv_w(f"supervision{i}.weight", seed, nc, c, 1)
        sd[f"supervision{i}.bias"] = _vec(f"supervision{i}.bias", seed, nc)
    p = "feature_fusion_module"
    sd[p + ".convblock.conv1.weight"] = _conv_w(p + ".convblock.conv1.weight", seed, nc, 1024, 3)
    _bn(sd, p + ".convblock.bn", seed, nc)
    for j in (1, 2):
        sd[f"{p}.conv{j}.weight"] = _conv_w(f"{p}.conv{j}.weight", seed, nc, nc, 1)
        sd[f"{p}.conv{j}.bias"] = _vec(f"{p}.conv{j}.bias", seed, nc)
    sd["conv.weight"] = _conv_w("conv.weight", seed, nc, nc, 1)
    sd["conv.bias"] = _vec("conv.bias", seed, nc)
    return sd


def resnet101_state(seed: int, prefix: str) -> dict:
    """torchvision resnet101 parameter tree (conv1, bn1, layer1-4 Bottlenecks [3,4,23,3], fc); stride on conv2 (v1.5)."""
    sd = {}
    sd[prefix + ".conv1.weight"] = _conv_w(prefix + ".conv1.weight", seed, 64, 3, 7)
    _bn(sd, prefix + ".bn1", seed, 64)
    inpl = 64
    for li, planes, blocks in ((1, 64, 3), (2, 128, 4), (3, 256, 23), (4, 512, 3)):
        for b in range(blocks):
            p = f"{prefix}.layer{li}.{b}"
            cin = inpl if b == 0 else planes * 4
            sd[p + ".conv1.weight"] = _conv_w(p + ".conv1.weight", seed, planes, cin, 1)
            _bn(sd, p + ".bn1", seed, planes)
            sd[p + ".conv2.weight"] = _conv_w(p + ".conv2.weight", seed, planes, planes, 3)
            _bn(sd, p + ".bn2", seed, planes)
            sd[p + ".conv3.weight"] = _conv_w(p + ".conv3.weight", seed, planes * 4, planes, 1, gain=1.0)
            _bn(sd, p + ".bn3", seed, planes * 4)
            sd[p + ".bn3.weight"] = sd[p + ".bn3.weight"] * 0.25      # keeps the 33 residual sums bounded
            if b == 0:
                sd[p + ".downsample.0.weight"] = _conv_w(p + ".downsample.0.weight", seed, planes * 4, cin, 1, gain=1.0)
                _bn(sd, p + ".downsample.1", seed, planes * 4)
        inpl = planes * 4
    sd[prefix + ".fc.weight"] = torch.randn(1000, 2048, generator=_gen(prefix + ".fc.weight", seed)) * 0.02
    sd[prefix + ".fc.bias"] = _vec(prefix + ".fc.bias", seed, 1000)
    return sd


def bisenet_r101_state(seed: int = 0, num_classes: int = 19) -> dict:
    """state_dict of the reference's BiSeNet(num_classes, 'resnet101') (build_bisenet.py:95-102: ARMs 1024 / 2048,
    FFM over 3328 channels), incl. the aliased context_path.{conv1,bn1,layer1..4}.* duplicates."""
    sd = {}
    nc = num_classes
    for i, (ci, co) in enumerate(((3, 64), (64, 128), (128, 256)), start=1):
        p = f"saptial_path.convblock{i}"
        sd[p + ".conv1.weight"] = _conv_w(p + ".conv1.weight", seed, co, ci, 3)
        _bn(sd, p + ".bn", seed, co)
    feats = resnet101_state(seed, "context_path.features")
    sd.update(feats)
    for k, v in feats.items():
        rest = k[len("context_path.features."):]
        if rest.split(".")[0] in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4"):
            sd["context_path." + rest] = v
    for i, c in ((1, 1024), (2, 2048)):
        p = f"attention_refinement_module{i}"
        sd[p + ".conv.weight"] = _conv_w(p + ".conv.weight", seed, c, c, 1)
        sd[p + ".conv.bias"] = _vec(p + ".conv.bias", seed, c)
        _bn(sd, p + ".bn", seed, c)
        sd[f"supervision{i}.weight"] = _conv_w(f"supervision{i}.weight", seed, nc, c, 1)
        sd[f"supervision{i}.bias"] = _vec(f"supervision{i}.bias", seed, nc)
    p = "feature_fusion_module"
    sd[p + ".convblock.conv1.weight"] = _conv_w(p + ".convblock.conv1.weight", seed, nc, 3328, 3)
    _bn(sd, p + ".convblock.bn", seed, nc)
    for j in (1, 2):
        sd[f"{p}.conv{j}.weight"] = _conv_w(f"{p}.conv{j}.weight", seed, nc, nc, 1)
        sd[f"{p}.conv{j}.bias"] = _vec(f"{p}.conv{j}.bias", seed, nc)
    sd["conv.weight"] = _conv_w("conv.weight", seed, nc, nc, 1)
    sd["conv.bias"] = _vec("conv.bias", seed, nc)
    return sd


def discriminator_state(seed: int = 0, tiny: bool = False, num_classes: int = 19) -> dict:
    """DomainDiscriminator (model.py:45-49) / TinyDomainDiscriminator (:72-73)."""
    sd = {}
    chans = [(num_classes, 64)] if tiny else [(num_classes, 64), (64, 128), (128, 256), (256, 512)]
    for i, (ci, co) in enumerate(chans, start=1):
        sd[f"conv{i}.weight"] = _conv_w(f"conv{i}.weight", seed, co, ci, 4)
        sd[f"conv{i}.bias"] = _vec(f"conv{i}.bias", seed, co)
    ci = chans[-1][1]
    sd["classifier.weight"] = _conv_w("classifier.weight", seed, 1, ci, 4)
    sd["classifier.bias"] = _vec("classifier.bias", seed, 1)
    return sd


def deeplab_state(seed: int = 0, num_classes: int = 19) -> dict:
    """State dict of the reference's get_deeplab_v2 (ResNetMulti(Bottleneck, [3,4,23,3]), deeplabv2.py:69-111): same
    keys, shapes and order as model.state_dict().  Conv weights get a He-style scale (the reference's N(0, 0.01)
    init makes the activations vanish through 101 layers, which would not exercise the arithmetic); BatchNorm
    parameters and buffers are non-trivial so that eval-mode folding is tested."""
    sd = {}
    sd["conv1.weight"] = _conv_w("conv1.weight", seed, 64, 3, 7)
    _bn(sd, "bn1", seed, 64)
    inpl = 64
    for name, planes, blocks in (("layer1", 64, 3), ("layer2", 128, 4), ("layer3", 256, 23), ("layer4", 512, 3)):
        for b in range(blocks):
            p = f"{name}.{b}"
            cin = inpl if b == 0 else planes * 4
            sd[p + ".conv1.weight"] = _conv_w(p + ".conv1.weight", seed, planes, cin, 1)
            _bn(sd, p + ".bn1", seed, planes)
            sd[p + ".conv2.weight"] = _conv_w(p + ".conv2.weight", seed, planes, planes, 3)
            _bn(sd, p + ".bn2", seed, planes)
            sd[p + ".conv3.weight"] = _conv_w(p + ".conv3.weight", seed, planes * 4, planes, 1, gain=1.0)
            _bn(sd, p + ".bn3", seed, planes * 4)
            sd[p + ".bn3.weight"] = sd[p + ".bn3.weight"] * 0.25      # keeps the 33 residual sums bounded in eval mode
            if b == 0:
                sd[p + ".downsample.0.weight"] = _conv_w(p + ".downsample.0.weight", seed, planes * 4, cin, 1, gain=1.0)
                _bn(sd, p + ".downsample.1", seed, planes * 4)
        inpl = planes * 4
    for i in range(4):
        sd[f"layer6.conv2d_list.{i}.weight"] = _conv_w(f"layer6.conv2d_list.{i}.weight", seed, num_classes, 2048, 3, gain=0.5)
        sd[f"layer6.conv2d_list.{i}.bias"] = _vec(f"layer6.conv2d_list.{i}.bias", seed, num_classes)
    return sd


def clone_state(sd: dict) -> dict:
    """Deep copy preserving aliasing between duplicated keys."""
    memo = {}
    out = {}
    for k, v in sd.items():
        if id(v) not in memo:
            memo[id(v)] = v.clone()
        out[k] = memo[id(v)]
    return out
