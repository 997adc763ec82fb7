"""Data-parallel plumbing: one process per GPU, gradients averaged with NCCL all-reduce over
NVLink / NVSwitch, issued per bucket as soon as backward has finished that part of the flat
gradient buffer so the transfer overlaps the rest of backward (replaces the reference's
nn.DataParallel, utils.py:97-107).  The compute path has no collective of its own:
images are independent, BatchNorm statistics stay per rank exactly as under DataParallel.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

# RTSDS_DDP_TIMELINE=1: CUDA events at backward start, at every bucket's "gradients final" point, at backward end and — on the
# compute stream, after the wait on each bucket's collective — at "bucket averaged"; last_timeline() turns the most recent
# step's events into milliseconds (tools/ddp_timeline.py prints them).  Off by default: no events are recorded.
TIMELINE = os.environ.get("RTSDS_DDP_TIMELINE", "0") == "1"
_last = None


def last_timeline():
    """[(bucket, MB, ready_ms, averaged_ms)], backward_end_ms of the most recent step (times since the reducer was created =
    backward start), or None."""
    if _last is None:
        return None
    torch.cuda.synchronize()
    t0, rows, end = _last
    return [(g, mb, t0.elapsed_time(r), t0.elapsed_time(d)) for g, mb, r, d in rows], t0.elapsed_time(end)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def param_buckets(named_params, groups):
    """Split the flat gradient buffer (parameters laid out in `named_params` order) into contiguous
    [start, end) ranges, one per group name; `groups` maps a parameter-name prefix to its group.
    Returns {group: (start, end)}; every parameter must fall into exactly one group and each group
    must be contiguous."""
    ranges, off = {}, 0
    for name, p in named_params:
        grp = None
        for prefix, g in groups:
            if name.startswith(prefix):
                grp = g
                break
        if grp is None:
            raise ValueError(f"parameter {name} belongs to no gradient bucket")
        if grp in ranges:
            s, e = ranges[grp]
            if e != off:
                raise ValueError(f"bucket {grp} is not contiguous at {name}")
            ranges[grp] = (s, off + p.numel())
        else:
            ranges[grp] = (off, off + p.numel())
        off += p.numel()
    return ranges


class BucketedAllReduce:
    """Averages slices of one flat tensor across ranks, asynchronously per bucket."""

    def __init__(self, flat: torch.Tensor, ranges: dict):
        self.flat, self.ranges = flat, ranges
        self.handles = []
        self.done = set()
        self.world = dist.get_world_size() if is_distributed() else 1
        self.avg = self.world > 1 and dist.get_backend() == "nccl"
        self.tl = None
        if TIMELINE and self.world > 1 and flat.is_cuda:
            self.tl = (torch.cuda.Event(enable_timing=True), [])
            self.tl[0].record()

    def ready(self, group: str):
        """Call when backward has finished writing the gradients of `group`."""
        if self.world == 1 or group in self.done:
            return
        self.done.add(group)
        s, e = self.ranges[group]
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        if self.tl is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.tl[1].append([group, (e - s) * 4 / 1e6, ev, None])
        self.handles.append(dist.all_reduce(self.flat[s:e], op=op, async_op=True))

    def finish(self):
        """Issue whatever was not signalled, wait for everything, return the averaged buffer."""
        if self.world == 1:
            return self.flat
        for g in self.ranges:
            self.ready(g)
        if self.tl is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()                                  # backward end (every bucket has been signalled)
        for i, h in enumerate(self.handles):
            h.wait()
            if self.tl is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self.tl[1][i][3] = ev
        if self.tl is not None:
            global _last
            _last = (self.tl[0], [tuple(r) for r in self.tl[1]], end)
        if not self.avg:
            self.flat.mul_(1.0 / self.world)
        return self.flat


BISENET_GROUPS = (
    ("saptial_path.", "spatial"),
    ("context_path.features.conv1", "layer1"), ("context_path.features.bn1", "layer1"),
    ("context_path.features.layer1", "layer1"), ("context_path.features.layer2", "layer2"),
    ("context_path.features.layer3", "layer3"), ("context_path.features.layer4", "layer4"),
    ("context_path.features.fc", "layer4"),
    ("attention_refinement_module", "head"), ("supervision", "head"), ("feature_fusion_module", "head"), ("conv.", "head"),
)


def broadcast_module(module: torch.nn.Module, src: int = 0):
    """Identical initial weights and buffers on every rank (what DataParallel's replicate does each step)."""
    if not is_distributed():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def allreduce_confusion(hist: torch.Tensor) -> torch.Tensor:
    """Sum the int64 confusion matrix over ranks once at the end of validation (validation.py:39,55)."""
    if is_distributed():
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist
