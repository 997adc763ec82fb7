"""Context path of BiSeNet — drop-in for the reference's
models/bisenet/build_contextpath.py (resnet18 :5-29, resnet101 :32-56,
build_contextpath :59-64).

Same class names, constructor signatures, attribute aliases and state_dict keys
(`features.*` plus the duplicated `conv1/bn1/layer1..4` aliases).  The modules
only OWN parameters; arithmetic runs in hand-written sm_100a kernels reached
through rtsds_b200 (the ResNet-18 stages as tcgen05 implicit GEMMs).  The
parameter containers below restate torchvision's ResNet layout (BasicBlock /
Bottleneck, `_make_layer`, Kaiming fan_out init) without importing torchvision,
so the model builds offline.
"""
from __future__ import annotations

import warnings

import torch
from torch import nn


class BasicBlock(nn.Module):
    """Parameter layout of torchvision.models.resnet.BasicBlock (resnet.py:59-105)."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        from rtsds_b200.module_ops import basic_block_forward

        return basic_block_forward(self, x)


class Bottleneck(nn.Module):
    """Parameter layout of torchvision.models.resnet.Bottleneck (stride on conv2)."""

    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        from rtsds_b200.module_ops import bottleneck_forward

        return bottleneck_forward(self, x)


class _ResNet(nn.Module):
    """torchvision.models.ResNet parameter tree (names, shapes, init), incl. the
    unused avgpool/fc the reference keeps."""

    def __init__(self, block, layers, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):  # ImageNet classifier head: not part of the RTSDS hot path
        raise NotImplementedError("the torchvision classifier forward is not part of the RTSDS hot path")


def _load_pretrained(net: nn.Module, arch: str) -> None:
    """`pretrained=True` in the reference downloads ImageNet weights
    (build_contextpath.py:8,35).  Offline we load them only if torchvision can
    serve them from its local hub cache; otherwise keep the seeded random init."""
    import os

    files = {"resnet18": "resnet18-f37072fd.pth", "resnet101": "resnet101-63fe2227.pth"}
    path = os.path.join(torch.hub.get_dir(), "checkpoints", files[arch])
    if not os.path.exists(path):  # never touch the network from the hot path
        # The reference fails here without a network (URLError from torch.hub).  Silently training from a random backbone
        # would change mIoU completely, so random init needs an explicit opt-in; build_bisenet.py disables `warnings`
        # globally, hence a plain stderr line rather than warnings.warn.
        if os.environ.get("RTSDS_ALLOW_RANDOM_INIT") != "1":
            raise RuntimeError(f"{arch}: pretrained=True but the ImageNet weights are not in the local hub cache ({path}). "
                               "Place the file there, or set RTSDS_ALLOW_RANDOM_INIT=1 to keep the seeded random init "
                               "(synthetic benchmarks / parity tests).")
        import sys
        print(f"[rtsds_b200] {arch}: ImageNet weights not found ({path}); RTSDS_ALLOW_RANDOM_INIT=1 -> random init",
              file=sys.stderr)
        return
    net.load_state_dict(torch.load(path, map_location="cpu"))


class _ContextPath(nn.Module):
    _arch = ""
    _block = BasicBlock
    _layers = (2, 2, 2, 2)

    def __init__(self, pretrained=True):
        super().__init__()
        self.features = _ResNet(self._block, self._layers)
        if pretrained:
            _load_pretrained(self.features, self._arch)
        self.conv1 = self.features.conv1
        self.bn1 = self.features.bn1
        self.relu = self.features.relu
        self.maxpool1 = self.features.maxpool
        self.layer1 = self.features.layer1
        self.layer2 = self.features.layer2
        self.layer3 = self.features.layer3
        self.layer4 = self.features.layer4

    def forward(self, input):
        """-> (feature3 @1/16, feature4 @1/32, tail = GAP(feature4)) as NCHW fp32."""
        from rtsds_b200.module_ops import context_path_forward

        return context_path_forward(self, input)


class resnet18(_ContextPath):
    _arch = "resnet18"
    _block = BasicBlock
    _layers = (2, 2, 2, 2)


class resnet101(_ContextPath):
    _arch = "resnet101"
    _block = Bottleneck
    _layers = (3, 4, 23, 3)


def build_contextpath(name):
    # The reference eagerly builds BOTH backbones, resnet18 first (build_contextpath.py:59-64), and returns one.  The
    # discarded one still consumes the global RNG stream, so it is built here too: under torch.manual_seed(s) the
    # drop-in then draws exactly the reference's random-init weights (tests/test_config1_cpu.py pins this against
    # tests/golden/config1.npz, produced by the real reference).
    if name not in ("resnet18", "resnet101"):
        raise KeyError(name)
    model = {"resnet18": resnet18(pretrained=True), "resnet101": resnet101(pretrained=True)}
    return model[name]
