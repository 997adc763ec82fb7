"""BiSeNet — drop-in for the reference's models/bisenet/build_bisenet.py.

Class names, constructor signatures, attribute names (incl. the reference's
`saptial_path` spelling, :90), parameter/buffer names and shapes, `init_weight`
(:130-139) and `mul_lr` (:121-128) are the reference's.  `forward(input)` keeps
the reference contract (:141-172): NCHW fp32 in; train -> `(result, cx1_sup,
cx2_sup)`, eval -> `result`, NCHW fp32 autograd-tracked tensors.  The modules own
parameters only; every FLOP runs in hand-written sm_100a CUDA kernels through
the C-ABI library (rtsds_b200).  There is no CPU or eager-PyTorch fallback:
a non-CUDA input raises.
"""
from __future__ import annotations

import warnings

import torch
from rtsds_b200.weights_epoch import PlanOwner
from torch import nn

from .build_contextpath import build_contextpath

warnings.filterwarnings(action="ignore")


def _bind(name):
    """Forward bodies live in rtsds_b200 (hand-written CUDA through the C ABI); resolved lazily so that importing the
    model file does not load the library."""
    def call(*args):
        import importlib

        mod, fn = name.rsplit(".", 1)
        return getattr(importlib.import_module(mod), fn)(*args)
    return call


_convblock = _bind("rtsds_b200.module_ops.convblock_forward")
_spatial = _bind("rtsds_b200.module_ops.spatial_path_forward")
_arm = _bind("rtsds_b200.module_ops.arm_forward")
_ffm = _bind("rtsds_b200.module_ops.ffm_forward")
_bisenet = _bind("rtsds_b200.bisenet_engine.bisenet_forward")

# channels of (feature3, feature4) handed over by each context path (reference :95-113)
_CONTEXT_CHANNELS = {"resnet18": (256, 512), "resnet101": (1024, 2048)}
_SPATIAL_CHANNELS = 256


class ConvBlock(torch.nn.Module):
    """conv(k, stride, pad=1, no bias) -> BN -> ReLU (reference :8-18)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=2, padding=1):
        super().__init__()
        geometry = dict(kernel_size=kernel_size, stride=stride, padding=padding)
        self.add_module("conv1", nn.Conv2d(in_channels, out_channels, bias=False, **geometry))
        self.add_module("bn", nn.BatchNorm2d(out_channels))
        self.add_module("relu", nn.ReLU())

    def forward(self, input):
        return _convblock(self, input)


class Spatial_path(torch.nn.Module):
    """3 ConvBlocks 3->64->128->256, each stride 2 (reference :21-32)."""

    def __init__(self):
        super().__init__()
        widths = (3, 64, 128, _SPATIAL_CHANNELS)
        for i, (cin, cout) in enumerate(zip(widths[:-1], widths[1:]), start=1):
            self.add_module(f"convblock{i}", ConvBlock(cin, cout))

    def forward(self, input):
        return _spatial(self, input)


class AttentionRefinementModule(torch.nn.Module):
    """x * sigmoid(BN(conv1x1(GAP(x)))) (reference :35-53)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels = in_channels
        for name, mod in (("conv", nn.Conv2d(in_channels, out_channels, 1)), ("bn", nn.BatchNorm2d(out_channels)),
                          ("sigmoid", nn.Sigmoid()), ("avgpool", nn.AdaptiveAvgPool2d((1, 1)))):
            self.add_module(name, mod)

    def forward(self, input):
        if input.size(1) != self.in_channels:          # the reference asserts with this message (:46)
            raise AssertionError('in_channels and out_channels should all be {}'.format(input.size(1)))
        return _arm(self, input)


class FeatureFusionModule(torch.nn.Module):
    """ConvBlock(cat(sx, cx)) re-weighted by its own channel attention (reference :56-81)."""

    def __init__(self, num_classes, in_channels):
        super().__init__()
        self.in_channels = in_channels
        squeeze = lambda: nn.Conv2d(num_classes, num_classes, 1)
        for name, mod in (("convblock", ConvBlock(in_channels, num_classes, stride=1)), ("conv1", squeeze()), ("relu", nn.ReLU()),
                          ("conv2", squeeze()), ("sigmoid", nn.Sigmoid()), ("avgpool", nn.AdaptiveAvgPool2d((1, 1)))):
            self.add_module(name, mod)

    def forward(self, input_1, input_2):
        total = input_1.size(1) + input_2.size(1)
        if total != self.in_channels:                  # reference :73
            raise AssertionError('in_channels of ConvBlock should be {}'.format(total))
        return _ffm(self, input_1, input_2)


class BiSeNet(PlanOwner, torch.nn.Module):
    def __init__(self, num_classes, context_path, with_interpolation=True):
        super().__init__()
        self.with_interpolation = with_interpolation
        self.saptial_path = Spatial_path()                      # (sic) the reference's attribute name, :90
        self.context_path = build_contextpath(name=context_path)
        heads = _CONTEXT_CHANNELS.get(context_path)
        if heads is None:
            print('Error: unspport context_path network \n')   # the reference prints and carries on (:114-115)
        else:
            # same creation order as the reference (ARM 1, ARM 2, the two auxiliary heads, the fusion module): the
            # state_dict key order and the sequence of default-init RNG draws stay identical
            for i, c in enumerate(heads, start=1):
                self.add_module(f"attention_refinement_module{i}", AttentionRefinementModule(c, c))
            for i, c in enumerate(heads, start=1):
                self.add_module(f"supervision{i}", nn.Conv2d(c, num_classes, 1))
            self.feature_fusion_module = FeatureFusionModule(num_classes, _SPATIAL_CHANNELS + sum(heads))
        self.conv = nn.Conv2d(num_classes, num_classes, 1)
        self.init_weight()
        # modules whose learning rate the reference's optimiser setup may scale (:121-128)
        self.mul_lr = [getattr(self, n) for n in ("saptial_path", "attention_refinement_module1", "attention_refinement_module2",
                                                  "supervision1", "supervision2", "feature_fusion_module", "conv")]
        self._context_name, self._num_classes = context_path, num_classes
        # rtsds_b200 execution options (not part of the reference API):
        #   precision "bf16" (tcgen05 path: bf16 training, eval-mode inference in rtsds_eval_precision) or "fp32" (check
        #   mode, BASELINE.json 1e-4 tolerance); eval precision "fp16" (default) or "bf16"
        self.rtsds_precision = "bf16"
        self.rtsds_eval_precision = "fp16"
        #   uint8 input frames: (mean, std) of the transforms.Normalize applied on the device (None: plain .float())
        self.rtsds_input_norm = None
        self.rtsds_cuda_graph = True

    def init_weight(self):
        """Kaiming-normal(fan_in, relu) on every conv outside the backbone, BatchNorm gamma=1 beta=0 eps=1e-5 momentum=0.1
        (reference :130-139)."""
        for name, mod in self.named_modules():
            if 'context_path' in name:
                continue
            if isinstance(mod, nn.Conv2d):
                nn.init.kaiming_normal_(mod.weight, mode='fan_in', nonlinearity='relu')
            elif isinstance(mod, nn.BatchNorm2d):
                mod.eps, mod.momentum = 1e-5, 0.1
                nn.init.ones_(mod.weight)
                nn.init.zeros_(mod.bias)

    def forward(self, input):
        return _bisenet(self, input)
