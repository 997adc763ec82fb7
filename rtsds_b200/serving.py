"""Host-buffer inference pipeline (the validation.py:41-54 call pattern — `inputs.to(device)`, `model(inputs)`,
`torch.argmax(outputs, 1)`, `.cpu()` — as a steady-state loop).

`PipelinedSegmenter` keeps `depth` frames in flight: the pinned-host -> device copy of frame i+1 and the device ->
pinned-host copy of frame i-1's prediction map run on their own streams while frame i is in the forward pass (one
CUDA-graph replay + the argmax kernel), so the PCIe transfers (6.3 MB in, 4.2 MB out per 512x1024 frame) hide behind
the compute instead of adding to it.  Results are returned in submission order; nothing is dropped or cached."""
from __future__ import annotations

import torch

from . import ops


class PipelinedSegmenter:
    def __init__(self, model, batch: int, height: int, width: int, depth: int = 3, lanes: int = 1, uint8_io: bool = False):
        """depth frames in flight; lanes > 1: consecutive frames run their forward on `lanes` compute streams with one
        execution plan (buffers, CUDA graph) each, so the latency-bound batch-1 kernels of neighbouring frames overlap.
        uint8_io: frames arrive as RAW uint8 [N,3,H,W] (normalised on the device by the stem kernel, model.rtsds_input_norm)
        and class maps leave as uint8 [N,H,W]: 1.5 MB in / 0.5 MB out per 512x1024 frame instead of 6.3 / 4.2."""
        self.uint8_io = uint8_io
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise ops._lib.RtsdsError("PipelinedSegmenter needs a CUDA model: rtsds_b200 has no CPU fallback")
        self.model, self.dev, self.depth = model.eval(), p0.device, depth
        self.lanes = max(1, min(lanes, depth))
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.s_c = [torch.cuda.Stream(self.dev) for _ in range(self.lanes)] if self.lanes > 1 else [None]
        # one slot more than frames in flight: the buffer handed back by submit() is not reused before the NEXT submit()
        self.slots = slots = depth + 1
        xdt, pdt = (torch.uint8, torch.uint8) if uint8_io else (torch.float32, torch.int64)
        self.x = [torch.empty(batch, 3, height, width, dtype=xdt, device=self.dev) for _ in range(slots)]
        self.pred = [torch.empty(batch, height, width, dtype=pdt, device=self.dev) for _ in range(slots)]
        self.host = [torch.empty(batch, height, width, dtype=pdt).pin_memory() for _ in range(slots)]
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_done = [torch.cuda.Event() for _ in range(slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]
        self.i = 0
        self.pending = []          # slots whose result has not been handed out yet, in order

    def submit(self, host_image: torch.Tensor):
        """Enqueue one pinned host batch [N,3,H,W] fp32.  Returns the oldest finished prediction (a pinned int64 host
        tensor, valid until the next submit()) once `depth` frames are in flight, else None."""
        out = None
        if len(self.pending) == self.depth:
            out = self._pop()
        k = self.i % self.slots
        lane = self.i % self.lanes
        self.i += 1
        # torch.cuda.stream() contexts cost ~15 us of Python each: switch streams with set_stream and restore at the end
        cur = torch.cuda.current_stream(self.dev)
        try:
            torch.cuda.set_stream(self.s_in)
            self.s_in.wait_event(self.ev_done[k])              # the forward that read this input slot has finished
            self.x[k].copy_(host_image, non_blocking=True)
            self.ev_in[k].record(self.s_in)
            cs = self.s_c[lane] if self.lanes > 1 else cur
            if cs is not cur:
                cs.wait_stream(cur)                             # whatever the caller queued before this frame (weight updates)
            cs.wait_event(self.ev_in[k])
            cs.wait_event(self.ev_out[k])                       # the previous prediction in this slot has left the device
            torch.cuda.set_stream(cs)
            self.model.rtsds_lane = lane
            with torch.no_grad():
                logits = self.model(self.x[k])
                ops.argmax_hist(logits, None, None, self.pred[k])
            self.ev_done[k].record(cs)
            torch.cuda.set_stream(self.s_out)
            self.s_out.wait_event(self.ev_done[k])
            self.host[k].copy_(self.pred[k], non_blocking=True)
            self.ev_out[k].record(self.s_out)
        finally:
            self.model.rtsds_lane = 0
            torch.cuda.set_stream(cur)
        self.pending.append(k)
        return out

    def _pop(self):
        k = self.pending.pop(0)
        self.ev_out[k].synchronize()
        return self.host[k]

    def drain(self):
        """Finish everything in flight; returns the remaining predictions in order."""
        return [self._pop() for _ in range(len(self.pending))]


class DevicePrefetcher:
    """Training-side counterpart: wraps an iterable of pinned host (image, label) batches and yields device tensors,
    copying batch i+1 on a side stream while step i computes (what train.py:71-72's `.to(device)` does serially)."""

    # Staging buffers are owned by ONE live prefetcher at a time.  A finished prefetcher returns its buffers to this
    # free pool (allocating 100+ MB per epoch is slow); a second prefetcher alive at the same time with equal batch
    # shapes (the source and target loaders of train.py:adversarial_train) checks out its OWN set, so it can never
    # overwrite tensors the first one has yielded and the step is still reading.
    _free = {}             # (device, shapes) -> [buffer sets]

    def __init__(self, batches, device, depth: int = 2):
        self.it, self.dev, self.depth = iter(batches), device, depth
        self.stream = torch.cuda.Stream(device)
        self.queue = []
        self.bufs = {}         # slot -> (pool key, device buffers) checked out by this instance

    def _buffers(self, slot, tensors):
        key = (str(self.dev), tuple((tuple(t.shape), t.dtype) for t in tensors))
        have = self.bufs.get(slot)
        if have is not None and have[0] == key:
            return have[1]
        if have is not None:                                  # ragged last batch: this slot's shapes changed
            DevicePrefetcher._free.setdefault(have[0], []).append(have[1])
        pool = DevicePrefetcher._free.get(key)
        bufs = pool.pop() if pool else [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in tensors]
        self.bufs[slot] = (key, bufs)
        return bufs

    def release(self):
        """Return the staging buffers to the free pool (called when the iterator is exhausted; later work on the
        current stream is ordered after every step that read them)."""
        for key, bufs in self.bufs.values():
            DevicePrefetcher._free.setdefault(key, []).append(bufs)
        self.bufs = {}

    def _enqueue(self, slot):
        try:
            host = next(self.it)
        except StopIteration:
            return False
        dst = self._buffers(slot, host)
        ev = torch.cuda.Event()
        # the step that last used these buffers was enqueued on the current stream before this point
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            for d, h in zip(dst, host):
                d.copy_(h, non_blocking=True)
            ev.record(self.stream)
        self.queue.append((dst, ev))
        return True

    def __iter__(self):
        slot = 0
        for _ in range(self.depth):
            if not self._enqueue(slot):
                break
            slot = (slot + 1) % (self.depth + 1)
        while self.queue:
            dst, ev = self.queue.pop(0)
            torch.cuda.current_stream(self.dev).wait_event(ev)
            self._enqueue(slot)
            slot = (slot + 1) % (self.depth + 1)
            yield tuple(dst)
        self.release()


class AsyncScalarReader:
    """Reads a device scalar (the step's loss, train.py:99 `loss.item()`) back to the host every step WITHOUT stalling
    the launch queue: the value is copied to pinned memory asynchronously and handed out one step later."""

    def __init__(self, device, depth: int = 2):
        self.host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev = [torch.cuda.Event() for _ in range(depth)]
        self.n, self.depth = 0, depth

    def push(self, t: torch.Tensor):
        """Enqueue the copy of this step's scalar; returns the value of the step `depth - 1` steps ago (or None)."""
        out = None
        k = self.n % self.depth
        if self.n >= self.depth:
            self.ev[k].synchronize()
            out = float(self.host[k][0])
        self.host[k].copy_(t.detach().reshape(1).float(), non_blocking=True)
        self.ev[k].record()
        self.n += 1
        return out

    def drain(self):
        vals = []
        for j in range(max(0, self.n - self.depth), self.n):
            k = j % self.depth
            self.ev[k].synchronize()
            vals.append(float(self.host[k][0]))
        return vals
