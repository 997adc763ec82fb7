"""ORACLE (test infrastructure): bf16-emulating restatement of the eval-mode BiSeNet-R18
forward.  Same algorithm as oracle/bisenet_ref.py (reference models/bisenet/build_bisenet.py
:141-172) but every tensor the CUDA path stores in bf16 is rounded to bf16 here too (weights
of the tensor-core convs, every NHWC activation buffer, the concat slots), with fp32
accumulation.  It separates *expected* bf16 rounding error (this emulation vs the fp32
oracle) from *implementation* error (CUDA path vs this emulation, which must be tiny).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _r(t):
    return t.to(torch.bfloat16).float()


def _fold(sd, p, eps=1e-5):
    sc = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + eps)
    sh = sd[p + ".bias"] - sd[p + ".running_mean"] * sc
    return sc.view(1, -1, 1, 1), sh.view(1, -1, 1, 1)


def bisenet_eval_bf16(x, sd, return_intermediates=False):
    # both stems run on the tensor cores too: bf16 image and weights, fp32 accumulation, bf16 output buffer
    x = _r(x)
    s = F.conv2d(x, _r(sd["saptial_path.convblock1.conv1.weight"]), None, 2, 1)
    sc, sh = _fold(sd, "saptial_path.convblock1.bn")
    s = _r(F.relu(s * sc + sh))
    for i in (2, 3):
        p = f"saptial_path.convblock{i}"
        s = F.conv2d(s, _r(sd[p + ".conv1.weight"]), None, 2, 1)
        sc, sh = _fold(sd, p + ".bn")
        s = _r(F.relu(s * sc + sh))
    P = "context_path.features"
    c = F.conv2d(x, _r(sd[P + ".conv1.weight"]), None, 2, 3)
    sc, sh = _fold(sd, P + ".bn1")
    c = F.max_pool2d(_r(F.relu(c * sc + sh)), 3, 2, 1)
    feats = []
    for li, st in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in range(2):
            p = f"{P}.layer{li}.{bi}"
            stride = st if bi == 0 else 1
            o = F.conv2d(c, _r(sd[p + ".conv1.weight"]), None, stride, 1)
            sc, sh = _fold(sd, p + ".bn1")
            o = _r(F.relu(o * sc + sh))
            o = F.conv2d(o, _r(sd[p + ".conv2.weight"]), None, 1, 1)
            sc, sh = _fold(sd, p + ".bn2")
            o = o * sc + sh
            if p + ".downsample.0.weight" in sd:
                idn = F.conv2d(c, _r(sd[p + ".downsample.0.weight"]), None, stride)
                sc, sh = _fold(sd, p + ".downsample.1")
                idn = _r(idn * sc + sh)
            else:
                idn = c
            c = _r(F.relu(o + idn))
        feats.append(c)
    f3, f4 = feats[2], feats[3]

    def gate(xx, p):
        g = F.adaptive_avg_pool2d(xx, 1)
        g = F.conv2d(g, sd[p + ".conv.weight"], sd[p + ".conv.bias"])
        sc, sh = _fold(sd, p + ".bn")
        return torch.sigmoid(g * sc + sh)

    g3 = gate(f3, "attention_refinement_module1")
    g4 = gate(f4, "attention_refinement_module2") * F.adaptive_avg_pool2d(f4, 1)
    cx1 = _r(F.interpolate(f3, size=s.shape[-2:], mode="bilinear") * g3)
    cx2 = _r(F.interpolate(f4, size=s.shape[-2:], mode="bilinear") * g4)
    cat = torch.cat((s, cx1, cx2), 1)
    p = "feature_fusion_module"
    f = F.conv2d(cat, _r(sd[p + ".convblock.conv1.weight"]), None, 1, 1)
    sc, sh = _fold(sd, p + ".convblock.bn")
    f = F.relu(f * sc + sh)
    a = F.adaptive_avg_pool2d(f, 1)
    a = F.relu(F.conv2d(a, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"]))
    a = torch.sigmoid(F.conv2d(a, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"]))
    g = f * a + f
    z = F.conv2d(g, sd["conv.weight"], sd["conv.bias"])
    out = F.interpolate(z, scale_factor=8, mode="bilinear")
    if return_intermediates:
        return dict(sx=s, f3=f3, f4=f4, cat=cat, feat=f, z=z, out=out, f1=feats[0], f2=feats[1])
    return out


# ----------------------------------------------------------------------------- train mode
class _Bf16Functional:
    """torch.nn.functional with straight-through bf16 rounding where the CUDA train path stores bf16:
    tensor-core conv weights, raw conv outputs and post-activation buffers.  Backward stays exact fp32,
    so gradients of this emulation show how far bf16 FORWARD round-off alone moves the gradients."""

    def __getattr__(self, k):
        return getattr(F, k)

    @staticmethod
    def _ste(t):
        return t + (t.to(torch.bfloat16).float() - t).detach()

    def conv2d(self, x, w, b=None, *a, **k):
        if x.shape[-1] == 1 and x.shape[-2] == 1:          # ARM / FFM attention 1x1 on pooled vectors: fp32 kernels
            return F.conv2d(x, w, b, *a, **k)
        if w.shape[1] == 3:                                 # fused tensor-core stems: bf16 image
            x = self._ste(x)
        return self._ste(F.conv2d(x, self._ste(w), b, *a, **k))   # bf16 weights, bf16 raw output

    def relu(self, x):
        return self._ste(F.relu(x)) if x.shape[-1] > 1 else F.relu(x)


def bisenet_train_bf16(x, sd):
    """Train-mode forward of oracle/bisenet_ref.py under bf16 storage emulation (autograd-enabled)."""
    from oracle import bisenet_ref

    old = bisenet_ref.F
    bisenet_ref.F = _Bf16Functional()
    try:
        return bisenet_ref.bisenet_forward(x, sd, train=True)
    finally:
        bisenet_ref.F = old
