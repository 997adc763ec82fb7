"""Import the REAL reference (read-only at /root/reference) in the build container.

Used only by oracle/gen_golden.py and by tests that are skipped when the
reference is absent (it does not exist on the GPU box).  The shim follows
SURVEY §8c: torchvision.models.resnet18/101 are wrapped to build with
weights=None *before* models.bisenet.build_bisenet is imported (the reference
would otherwise try to download ImageNet weights, build_contextpath.py:59-64),
and fast_hist / per_class_iou are ast-extracted from utils.py because
`import utils` needs fvcore/matplotlib.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys

REF = os.environ.get("RTSDS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models", "bisenet"))


def _import_ref(modname: str):
    """Import `modname` from the reference tree even though this repo has same-named packages."""
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module(modname)
        loaded = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k == "models" or k.startswith("models."):
                del sys.modules[k]
        sys.modules.update(saved)
    return mod, loaded


def load_models():
    """-> dict(BiSeNet=..., DomainDiscriminator=..., TinyDomainDiscriminator=..., get_deeplab_v2=...)."""
    import torchvision

    orig18, orig101 = torchvision.models.resnet18, torchvision.models.resnet101
    torchvision.models.resnet18 = lambda *a, **k: orig18(weights=None)
    torchvision.models.resnet101 = lambda *a, **k: orig101(weights=None)
    try:
        bis, _ = _import_ref("models.bisenet.build_bisenet")
        adv, _ = _import_ref("models.domain_shift.adversarial.model")
        dl, _ = _import_ref("models.deeplabv2.deeplabv2")
    finally:
        pass  # keep the wrappers: BiSeNet() calls build_contextpath at construction time
    return dict(BiSeNet=bis.BiSeNet, DomainDiscriminator=adv.DomainDiscriminator,
                TinyDomainDiscriminator=adv.TinyDomainDiscriminator, get_deeplab_v2=dl.get_deeplab_v2,
                GradientReversalFunction=adv.GradientReversalFunction)


def load_utils_functions():
    """ast-extract fast_hist, per_class_iou, poly_lr_scheduler from the reference's utils.py."""
    import numpy as np

    src = open(os.path.join(REF, "utils.py")).read()
    tree = ast.parse(src)
    want = {"fast_hist", "per_class_iou", "poly_lr_scheduler"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), os.path.join(REF, "utils.py"), "exec"), ns)
    return {k: ns[k] for k in want}
