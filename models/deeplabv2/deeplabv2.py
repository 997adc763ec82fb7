"""DeepLabV2-ResNet101 — drop-in for the reference's models/deeplabv2/deeplabv2.py.

Class names, constructor signatures, module tree (hence the 632 state_dict keys), initialisation and
the frozen BatchNorm affine parameters are the reference's (:7-111).  `ResNetMulti.forward(x)` keeps
the reference contract (:113-131): NCHW fp32 in; train -> `(logits, None, None)`, eval -> `logits`,
full-resolution fp32 NCHW, autograd-tracked.  The modules own parameters only; every FLOP runs in
hand-written sm_100a kernels through the C ABI (rtsds_b200/deeplab_engine.py): 1x1 and dilated 3x3
convolutions as tcgen05 implicit GEMMs (dilation = TMA tap offsets, padding = TMA out-of-bounds
fill), the four ASPP branches chained through the residual epilogue, ceil-mode max-pool, train-mode
BatchNorm from conv-epilogue statistics.  No CPU fallback: a non-CUDA input raises.
"""
from __future__ import annotations

import torch
from rtsds_b200.weights_epoch import PlanOwner
import torch.nn as nn

affine_par = True

# (planes, stride, dilation) of layer1..layer4 -- output stride 8: the last two stages trade their stride for dilation (:78-81)
_STAGES = ((64, 1, 1), (128, 2, 1), (256, 1, 2), (512, 1, 4))
_ASPP_RATES = (6, 12, 18, 24)


def _lazy(path):
    def call(*args):
        import importlib

        mod, fn = path.rsplit(".", 1)
        return getattr(importlib.import_module(mod), fn)(*args)
    return call


_bottleneck_fwd = _lazy("rtsds_b200.module_ops.bottleneck_forward")
_classifier_fwd = _lazy("rtsds_b200.module_ops.classifier_forward")
_deeplab_fwd = _lazy("rtsds_b200.deeplab_engine.deeplab_forward")


def _frozen_bn(c):
    """BatchNorm whose affine parameters never train (the reference switches requires_grad off one by one, :16-44)."""
    bn = nn.BatchNorm2d(c, affine=affine_par)
    bn.requires_grad_(False)
    return bn


def _conv(cin, cout, k, **kw):
    return nn.Conv2d(cin, cout, kernel_size=k, bias=False, **kw)


class Bottleneck(nn.Module):
    """1x1 (stride) -> 3x3 (dilation) -> 1x1 (x4), BatchNorm affine frozen (reference :7-47)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        wide = planes * self.expansion
        pieces = (("conv1", _conv(inplanes, planes, 1, stride=stride)), ("bn1", _frozen_bn(planes)),
                  ("conv2", _conv(planes, planes, 3, stride=1, padding=dilation, dilation=dilation)), ("bn2", _frozen_bn(planes)),
                  ("conv3", _conv(planes, wide, 1)), ("bn3", _frozen_bn(wide)), ("relu", nn.ReLU(inplace=True)))
        for name, mod in pieces:
            self.add_module(name, mod)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        return _bottleneck_fwd(self, x)


class ClassifierModule(nn.Module):
    """ASPP head: sum of dilated 3x3 convs with bias, N(0, 0.01) init (reference :50-66)."""

    def __init__(self, inplanes, dilation_series, padding_series, num_classes):
        super().__init__()
        self.conv2d_list = nn.ModuleList(
            nn.Conv2d(inplanes, num_classes, kernel_size=3, stride=1, padding=pad, dilation=rate, bias=True)
            for rate, pad in zip(dilation_series, padding_series))
        for branch in self.conv2d_list:
            nn.init.normal_(branch.weight, 0, 0.01)

    def forward(self, x):
        return _classifier_fwd(self, x)


class ResNetMulti(PlanOwner, nn.Module):
    def __init__(self, block, layers, num_classes):
        super().__init__()
        self.inplanes = 64
        stem = (("conv1", _conv(3, 64, 7, stride=2, padding=3)), ("bn1", _frozen_bn(64)), ("relu", nn.ReLU(inplace=True)),
                ("maxpool", nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=True)))
        for name, mod in stem:
            self.add_module(name, mod)
        for i, ((planes, stride, dilation), depth) in enumerate(zip(_STAGES, layers), start=1):
            self.add_module(f"layer{i}", self._make_layer(block, planes, depth, stride=stride, dilation=dilation))
        self.layer6 = ClassifierModule(512 * block.expansion, list(_ASPP_RATES), list(_ASPP_RATES), num_classes)
        # reference :84-90: every conv N(0, 0.01) (the classifier's a second time), every BatchNorm gamma=1 beta=0
        for mod in self.modules():
            if isinstance(mod, nn.Conv2d):
                nn.init.normal_(mod.weight, 0, 0.01)
            elif isinstance(mod, nn.BatchNorm2d):
                nn.init.ones_(mod.weight)
                nn.init.zeros_(mod.bias)
        # rtsds_b200 execution options (not part of the reference API)
        self.rtsds_precision = "bf16"       # "fp32": CUDA-core check mode (BASELINE.json 1e-4 tolerance)
        self.rtsds_cuda_graph = True

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        # every stage of the reference ends up with a 1x1 projection shortcut on its first block (:88-97)
        wide = planes * block.expansion
        shortcut = nn.Sequential(_conv(self.inplanes, wide, 1, stride=stride), _frozen_bn(wide))
        chain = [block(self.inplanes, planes, stride, dilation=dilation, downsample=shortcut)]
        self.inplanes = wide
        chain += [block(wide, planes, dilation=dilation) for _ in range(blocks - 1)]
        return nn.Sequential(*chain)

    def forward(self, x):
        return _deeplab_fwd(self, x)

    def get_1x_lr_params_no_scale(self):
        """Trainable parameters of everything but the classifier (reference :133-156)."""
        for name in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4"):
            yield from (p for p in getattr(self, name).parameters() if p.requires_grad)

    def get_10x_lr_params(self):
        """Parameters of the classifier (reference :158-170; its `self.multi_level` branch is dead code there)."""
        yield from self.layer6.parameters()

    def optim_parameters(self, lr):
        return [dict(params=self.get_1x_lr_params_no_scale(), lr=lr), dict(params=self.get_10x_lr_params(), lr=10 * lr)]


def get_deeplab_v2(num_classes=19, pretrain=True, pretrain_model_path='DeepLab_resnet_pretrained_imagenet.pth'):
    model = ResNetMulti(Bottleneck, [3, 4, 23, 3], num_classes)
    if pretrain:
        print('Deeplab pretraining loading...')
        checkpoint = torch.load(pretrain_model_path)
        merged = dict(model.state_dict())
        # checkpoint keys carry one leading scope component (reference :185-188)
        merged.update({key.split('.', 1)[1]: value for key, value in checkpoint.items()})
        model.load_state_dict(merged, strict=False)
    return model
