"""ORACLE (test infrastructure): numpy restatement of the reference's mIoU metric.

Follows /root/reference/utils.py:52-63 (fast_hist, per_class_iou) and the
accumulation in validation.py:39,54-55,69-70.  Pinned against the functions
ast-extracted from the reference itself (oracle/gen_golden.py ->
tests/golden/fast_hist_*.npz): the reference has no tests of its own.
"""
from __future__ import annotations

import numpy as np


def fast_hist(a: np.ndarray, b: np.ndarray, n: int) -> np.ndarray:
    """utils.py:52-58 — a = label, b = prediction; int64 [n, n], rows = label."""
    k = (a >= 0) & (a < n)
    return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)


def per_class_iou(hist: np.ndarray) -> np.ndarray:
    """utils.py:61-63 — float64 IoU per class with the reference's +1e-5 epsilon."""
    epsilon = 1e-5
    return (np.diag(hist)) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + epsilon)


def mean_iou(hist: np.ndarray) -> float:
    """validation.py:69-70."""
    return float(np.nanmean(per_class_iou(hist)))


def fast_hist_loops(a, b, n):
    """Independent pure-Python restatement for tiny cases (cross-checks the numpy one)."""
    h = [[0] * n for _ in range(n)]
    for x, y in zip(np.asarray(a).ravel().tolist(), np.asarray(b).ravel().tolist()):
        if 0 <= x < n:
            idx = n * x + y
            h[idx // n][idx % n] += 1
    return np.array(h, dtype=np.int64)
