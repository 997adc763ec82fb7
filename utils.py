"""Drop-in for the hot-path functions of the reference's utils.py: `fast_hist` (:52-58),
`per_class_iou` (:61-63), `poly_lr_scheduler` (:33-48) and `forModel` (:97-107).

`fast_hist(a, b, n)` keeps the reference's signature and result (int64 [n,n], rows = label `a`,
columns = prediction `b`, labels outside [0,n) dropped) but counts on the GPU with the
shared-memory-atomic histogram kernel (rtsds_b200/csrc/hist.cu) — bit-exact integer arithmetic.
It accepts what validation.py:54,124 passes (numpy arrays, uploaded once) as well as CUDA tensors
(no copy; the result then stays on the device so a validation loop can accumulate without
synchronising).  There is no CPU implementation here: without a CUDA device it raises.
The colour palette, plotting and fvcore helpers of the reference's utils.py are host-side tooling
outside the hot path and are not reproduced.
"""
from __future__ import annotations

import numpy as np
import torch


def poly_lr_scheduler(optimizer, init_lr, iter, lr_decay_iter=1, max_iter=300, power=0.9):
    """Polynomial decay lr = init_lr * (1 - iter/max_iter)^power, written to param_groups[0] ONLY,
    as the reference does (utils.py:33-48)."""
    lr = init_lr * (1 - iter / max_iter) ** power
    optimizer.param_groups[0]['lr'] = lr
    return lr


def fast_hist(a, b, n):
    """Confusion matrix of labels `a` against predictions `b` over n classes (utils.py:52-58)."""
    from rtsds_b200 import ops

    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.is_cuda:
        hist = torch.zeros(n * n, dtype=torch.int64, device=a.device)
        ops.confusion_hist(a.to(torch.int64).contiguous().view(-1), b.to(torch.int64).contiguous().view(-1), n, hist)
        return hist.view(n, n)
    if not torch.cuda.is_available():
        raise ops._lib.RtsdsError("fast_hist runs on the GPU (rtsds_b200 has no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    ta = torch.as_tensor(np.ascontiguousarray(a)).to(torch.int64).to(dev).view(-1)
    tb = torch.as_tensor(np.ascontiguousarray(b)).to(torch.int64).to(dev).view(-1)
    hist = torch.zeros(n * n, dtype=torch.int64, device=dev)
    if ta.numel():
        ops.confusion_hist(ta, tb, n, hist)
    return hist.view(n, n).cpu().numpy()


def per_class_iou(hist):
    """diag / (row sum + column sum - diag + 1e-5) in float64 (utils.py:61-63); 361 numbers, host arithmetic."""
    if isinstance(hist, torch.Tensor):
        hist = hist.detach().cpu().numpy()
    hist = np.asarray(hist)
    epsilon = 1e-5
    return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + epsilon)


class IntRangeTransformer:
    """Label transform of main.py:74-77 (utils.py:67-75): clamp into [min_val, max_val], cast to int64.
    Host-side dataset tooling, kept so `from utils import IntRangeTransformer, forModel` (main.py:20) resolves."""

    def __init__(self, min_val=0, max_val=255):
        self.min_val, self.max_val = min_val, max_val

    def __call__(self, sample):
        return torch.clamp(sample, self.min_val, self.max_val).long()


def tabular_print(log_dict):
    """Epoch summary printer used by train.py:288,467 (utils.py:77-94): one `key: value` row per entry."""
    width = max((len(str(k)) for k in log_dict), default=0)
    for k, v in log_dict.items():
        print(f"{str(k):<{width}} : {v}")


def forModel(model, device):
    """Device placement (utils.py:97-107).  The reference wraps the model in nn.DataParallel when several
    GPUs are visible; here multi-GPU is one process per GPU with NCCL gradient all-reduce
    (rtsds_b200/ddp.py), so inside an initialised process group the model is flagged for that
    instead; a single process just moves the model to its device."""
    from rtsds_b200 import ddp

    if device == 'cuda' or (isinstance(device, torch.device) and device.type == 'cuda'):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available")
        model = model.cuda()
        if ddp.is_distributed():
            ddp.broadcast_module(model, 0)
            model.rtsds_ddp = True
    return model
