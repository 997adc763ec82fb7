"""Device-side input pipeline: the per-sample CPU work of the reference's Dataset + torchvision transforms
(main.py:60-108; datasets/cityscapes.py:66-72; datasets/gta5.py:68-82; utils.py:67-75), evaluated on the GPU from the
raw uint8 planes (csrc/input.cu), so a frame crosses PCIe as 3 bytes per pixel instead of 12 and a label as 1 instead of 8.

    pipe = DeviceInputPipeline(size=(512, 1024))              # Cityscapes: Resize(antialias=True) + Normalize + clamp [0, 19]
    x = pipe.images(u8_batch)                                  # uint8 [N,3,h,w] cuda -> fp32 [N,3,512,1024]
    y = pipe.labels(u8_labels, clamp=(0, 19))                  # uint8/int64 [N,1,h,w] or [N,h,w] -> int64 [N,512,1024]

and, for inference, `model(u8_frame)`: the eval-mode BiSeNet forward accepts a uint8 NCHW tensor and applies
`model.rtsds_input_norm = (mean, std)` in a 3 us convert + normalise pass in front of its fused stem kernel (reading the
bytes inside the stem kernel was measured slower: the conversion lands on that kernel's critical gather warps).

The reference normalises the 0..255 float image with the ImageNet mean / std of 0..1 images (main.py:70: `.float()` is
never divided by 255); that quirk is kept: the defaults below are the reference's numbers on the reference's scale."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from ._lib import check, lib

MEAN = (0.485, 0.456, 0.406)       # main.py:70,82
STD = (0.229, 0.224, 0.225)


def _affine(mean, std):
    sc = (C.c_float * 3)(*[1.0 / s for s in std])
    bi = (C.c_float * 3)(*[-m / s for m, s in zip(mean, std)])
    return sc, bi


class DeviceInputPipeline:
    def __init__(self, size=None, mean=MEAN, std=STD):
        """size (H, W) of transforms.Resize (None: keep the input size); mean / std of transforms.Normalize (None: skip)."""
        self.size = tuple(size) if size is not None else None
        self.mean = tuple(mean) if mean is not None else (0.0, 0.0, 0.0)
        self.std = tuple(std) if std is not None else (1.0, 1.0, 1.0)
        self._sc, self._bi = _affine(self.mean, self.std)

    def images(self, u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """read_image(...).float() -> Resize(size, antialias=True) -> Normalize(mean, std)  (datasets/cityscapes.py:66, main.py:68-71)."""
        if not u8.is_cuda and not _lib.dry_run():
            raise _lib.RtsdsError("DeviceInputPipeline needs CUDA tensors (there is no CPU fallback)")
        if u8.dtype != torch.uint8 or u8.dim() != 4 or u8.shape[1] > 3:
            raise ValueError(f"expected uint8 [N,C<=3,h,w], got {u8.dtype} {tuple(u8.shape)}")
        u8 = u8.contiguous()
        n, c, h, w = u8.shape
        oh, ow = self.size if self.size is not None else (h, w)
        if out is None:
            out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=u8.device)
        check(lib().rtsds_image_u8_to_f32(ops._p(u8), n, c, h, w, oh, ow, self._sc, self._bi, ops._p(out), ops._s()), "image_u8_to_f32")
        return out

    def labels(self, lab: torch.Tensor, clamp=None, out: torch.Tensor | None = None) -> torch.Tensor:
        """read_image(...).long() -> Resize(size, antialias=True) [-> IntRangeTransformer(lo, hi)] (main.py:73-76, utils.py:67-75).
        Returns int64 [N,H,W] — the `.squeeze(1)` of train.py:72 included."""
        if not lab.is_cuda and not _lib.dry_run():
            raise _lib.RtsdsError("DeviceInputPipeline needs CUDA tensors (there is no CPU fallback)")
        if lab.dim() == 4:
            if lab.shape[1] != 1:
                raise ValueError("label tensor must be [N,1,h,w] or [N,h,w]")
            lab = lab[:, 0]
        if lab.dtype not in (torch.uint8, torch.int64):
            raise ValueError("labels must be uint8 or int64")
        lab = lab.contiguous()
        n, h, w = lab.shape
        oh, ow = self.size if self.size is not None else (h, w)
        if out is None:
            out = torch.empty((n, oh, ow), dtype=torch.int64, device=lab.device)
        lo, hi = clamp if clamp is not None else (0, 0)
        check(lib().rtsds_label_resize_clamp(ops._p(lab), int(lab.dtype == torch.uint8), n, h, w, oh, ow, int(clamp is not None),
                                             int(lo), int(hi), ops._p(out), ops._s()), "label_resize_clamp")
        return out


def stem_affine(model):
    """(scale3, bias3) ctypes arrays of model.rtsds_input_norm = (mean, std) (None: plain .float())."""
    norm = getattr(model, "rtsds_input_norm", None)
    key = None if norm is None else (tuple(norm[0]), tuple(norm[1]))
    cached = model.__dict__.get("_rtsds_input_affine")
    if cached is None or cached[0] != key:
        cached = (key, _affine(*key) if key is not None else _affine((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)))
        model.__dict__["_rtsds_input_affine"] = cached
    return cached[1]
