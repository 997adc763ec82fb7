"""Sync-free forms of the reference's loops: same signatures, same callback contract, same return
values as train.py:train (:24-128), validation.py:val (:12-77) / val_GTA5 (:79-149) and train.py:adversarial_train
(:130-318) — `from rtsds_b200.loops import train, val, val_GTA5, adversarial_train` instead of
`from train import ...` / `from validation import ...` (main.py:13,19) is the whole switch.

What changes is WHERE the numbers live.  The reference reads two to eight scalars back from the GPU in every iteration
(`loss.item()`, `predicted.eq(targets).sum().item()`, `.cpu().numpy()` of both label maps; train.py:99,106,214,234,253,
264,280-283, validation.py:54-66,120-134), each a full host<->device synchronisation that drains the launch queue.  Here

  * running loss / correct pixels / the confusion matrix are ACCUMULATED ON THE DEVICE and read once per epoch;
  * the per-batch values the callbacks receive (`on_batch_end(i, {...})`, `on_validation_batch_end(i, loss)`) are copied to
    pinned host memory asynchronously and delivered a fixed number of batches late (default 2), in order, all of them
    before the epoch-end callbacks — plain Python floats as in the reference, never a stall;
  * the loss / argmax / pixel accuracy of the three heads come from the fused resize+CE kernels (no full-resolution logits),
    argmax + fast_hist of validation from the fused argmax_hist kernel, the optimizer step from the fused multi-tensor
    kernel (rtsds_b200/optim.py) — whenever the model / criterion / optimizer are the ones the reference builds
    (main.py:110-136); anything else takes the stock call sequence, still without per-batch synchronisation.

Numerically the loops are the reference's: tests/test_gpu_loops.py runs them beside a restatement of the reference loops
(per-batch `.item()` and all) on the same data and compares every logged number.
"""
from __future__ import annotations

import numpy as np
import torch

import utils as _utils                     # this repository's drop-in utils.py (poly_lr_scheduler, per_class_iou, tabular_print)

from . import ops, optim
from .bisenet_autograd import bisenet_fused_ce

DEFER = 2                                   # batches between an iteration and the delivery of its host-side numbers


class _Deferred:
    """Per-batch device scalars -> pinned host memory, asynchronously; `push` hands back (tag, values) of the batch
    DEFER pushes ago, `drain` the rest, in order."""

    def __init__(self, width, dtype=torch.float64, depth=DEFER):
        self.depth = max(1, depth)
        self.host = [torch.zeros(width, dtype=dtype).pin_memory() for _ in range(self.depth)]
        self.ev = [torch.cuda.Event() for _ in range(self.depth)]
        self.tags = [None] * self.depth
        self.n = 0

    def push(self, tag, values: torch.Tensor):
        k = self.n % self.depth
        out = None
        if self.n >= self.depth:
            self.ev[k].synchronize()              # long finished: this copy was enqueued `depth` iterations ago
            out = (self.tags[k], self.host[k].clone())
        self.host[k].copy_(values.detach().reshape(-1).to(self.host[k].dtype), non_blocking=True)
        self.ev[k].record()
        self.tags[k] = tag
        self.n += 1
        return out

    def drain(self):
        out = []
        for j in range(max(0, self.n - self.depth), self.n):
            k = j % self.depth
            self.ev[k].synchronize()
            out.append((self.tags[k], self.host[k].clone()))
        self.n = 0
        return out


def _fusable(model, criterion):
    return (hasattr(model, "saptial_path") and hasattr(model, "rtsds_precision") and type(criterion) is torch.nn.CrossEntropyLoss
            and criterion.weight is None and criterion.reduction == "mean" and criterion.label_smoothing == 0.0)


def train(epoch, model, train_loader, criterion, optimizer, init_lr, max_iter, power=0.9, lr_decay_iter=1.0, device='cpu',
          callbacks=[]):
    """train.py:24-128.  Returns the model."""
    for callback in callbacks:
        callback.on_train_begin()
    model.train()
    optim.fuse_(optimizer)                                    # one kernel per step; a no-op for unsupported optimizers
    fused = _fusable(model, criterion)
    dev = torch.device(device)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)     # running loss, correct pixels
    total = 0
    late = _Deferred(2)

    def deliver(item):
        if item is not None:
            (idx, tot), v = item
            for callback in callbacks:
                callback.on_batch_end(idx, {'train_loss': float(v[0]), 'train_accuracy': 100. * float(v[1]) / tot})

    for batch_idx, (inputs, targets) in enumerate(train_loader):
        current_iter = epoch * len(train_loader) + batch_idx
        if current_iter % lr_decay_iter == 0 and current_iter <= max_iter:
            _utils.poly_lr_scheduler(optimizer, init_lr, current_iter, lr_decay_iter, max_iter, power)
        inputs = inputs.to(dev, non_blocking=True)
        targets = targets.to(dev, non_blocking=True).squeeze(1)
        optimizer.zero_grad()
        if fused:
            loss, _, stats = bisenet_fused_ce(model, inputs, targets, criterion.ignore_index)
            correct = stats[0, 2]
        else:
            outputs = model(inputs)
            main_output, aux1, aux2 = outputs if isinstance(outputs, tuple) else (outputs, None, None)
            loss = criterion(main_output, targets)
            if aux1 is not None:
                loss = loss + criterion(aux1, targets)
            if aux2 is not None:
                loss = loss + criterion(aux2, targets)
            correct = main_output.max(1)[1].eq(targets).sum()
        loss.backward()
        optimizer.step()
        total += targets.size(0) * targets.size(1) * targets.size(2)
        step = torch.stack((loss.detach().double(), correct.double()))
        acc += step
        # the reference reports the RUNNING accuracy (correct so far / pixels so far) with this batch's loss
        deliver(late.push((batch_idx, total), torch.stack((step[0], acc[1]))))
    for item in late.drain():
        deliver(item)
    host = acc.cpu()                                          # the one synchronising read of the epoch
    train_loss = float(host[0]) / len(train_loader)
    train_accuracy = 100. * float(host[1]) / total
    print(f'Train Epoch: {epoch + 1} Loss: {train_loss:.6f} Acc: {train_accuracy:.2f}%')
    for callback in callbacks:
        callback.on_epoch_end(epoch, {'train_loss': train_loss, 'train_accuracy': train_accuracy})
    return model


def _validate(model, val_loader, num_classes, device, callbacks):
    """Shared body of val / val_GTA5: eval forward, fused argmax + fast_hist accumulated on the device (int64, bit-exact),
    per-batch pixel-accuracy "loss" delivered late.  Returns the [n,n] int64 confusion matrix as numpy."""
    model.eval()
    dev = torch.device(device)
    hist = torch.zeros(num_classes * num_classes, dtype=torch.int64, device=dev)
    late = _Deferred(num_classes * num_classes, dtype=torch.int64)

    def deliver(item):
        if item is not None:
            idx, h = item
            h = h.numpy().reshape(num_classes, num_classes).astype(np.float64)
            loss = 1. - np.sum(np.diag(h)) / np.sum(h)
            for callback in callbacks:
                callback.on_validation_batch_end(idx, loss)

    with torch.no_grad():
        for batch_idx, (inputs, targets) in enumerate(val_loader):
            inputs = inputs.to(dev, non_blocking=True)
            targets = targets.to(dev, non_blocking=True).squeeze(1)
            outputs = model(inputs)
            if isinstance(outputs, tuple):
                outputs = outputs[0]
            ops.argmax_hist(outputs.contiguous(), targets.to(torch.int64).contiguous(), hist, None)
            deliver(late.push(batch_idx, hist))
    for item in late.drain():
        deliver(item)
    return hist.cpu().numpy().reshape(num_classes, num_classes)


def val(epoch, model, val_loader, num_classes, device='cpu', callbacks=[]):
    """validation.py:12-77.  Returns the mean IoU."""
    for callback in callbacks:
        callback.on_validation_begin()
    total_hist = _validate(model, val_loader, num_classes, device, callbacks)
    ious = _utils.per_class_iou(total_hist)
    mean_iou = np.nanmean(ious)
    print(f'Validation Mean IoU for Epoch {epoch + 1}: {mean_iou:.4f}')
    for callback in callbacks:
        callback.on_validation_end(mean_iou)
    return mean_iou


def val_GTA5(epoch, model, val_loader, num_classes, class_names, callbacks=[], device='cpu'):
    """validation.py:79-149.  Returns (mean IoU, per-class DataFrame)."""
    import pandas as pd

    for callback in callbacks:
        callback.on_validation_begin()
    confusion_matrix = _validate(model, val_loader, num_classes, device, callbacks)
    IoUs = _utils.per_class_iou(confusion_matrix)
    total_miou = np.nanmean(IoUs)
    print(f'Validation mIoU for Epoch {epoch + 1}: {total_miou:.4f}')
    class_result_df = pd.DataFrame({'Class': class_names, 'IoU': [f'{iou:.4f}' for iou in IoUs]})
    print(class_result_df)
    for callback in callbacks:
        callback.on_validation_end({'validation_mIoU': total_miou}, data=class_result_df)
    return total_miou, class_result_df


class _Cycle:
    """The reference draws `next(iter(loader))` every iteration (train.py:184-185): a fresh iterator — and with worker
    processes a fresh worker pool — per batch.  One persistent iterator, restarted when exhausted, yields the same stream of
    shuffled batches without that cost."""

    def __init__(self, loader):
        self.loader, self.it = loader, iter(loader)

    def next(self):
        try:
            return next(self.it)
        except StopIteration:
            self.it = iter(self.loader)
            return next(self.it)


def adversarial_train(iterations, epochs, generator, discriminator, generator_optimizer, discriminator_optimizer,
                      source_dataloader, target_dataloader, generator_loss, discriminator_loss, lambda_, gen_init_lr, gen_power,
                      dis_power, dis_init_lr, lr_decay_iter, num_classes, class_names, val_loader, do_validation=1, device='cpu',
                      when_print=10, callbacks=[]):
    """train.py:130-318: per iteration G(source) 3xCE, G(target) fooling the frozen D, D on both detached predictions, both
    optimizers; per epoch the discriminator's poly LR, the tabular report, val_GTA5 and the best-model checkpoints."""
    from .train_steps import adversarial_step

    optim.fuse_(generator_optimizer)
    optim.fuse_(discriminator_optimizer)
    dev = torch.device(device)
    src, tgt = _Cycle(source_dataloader), _Cycle(target_dataloader)
    keys = ('loss_gen_source', 'loss_adversarial', 'loss_disc_source', 'loss_disc_target')
    for epoch in range(epochs):
        for callback in callbacks:
            callback.on_train_begin()
        acc = torch.zeros(5, dtype=torch.float64, device=dev)        # the four running losses, correct pixels
        generator_total = 0
        best_mIoU = 0
        generator.train()
        discriminator.train()
        dis_lr = _utils.poly_lr_scheduler(discriminator_optimizer, dis_init_lr, epoch, lr_decay_iter, epochs, dis_power)
        gen_lr = None
        max_iter = epochs * iterations
        late = _Deferred(4)

        def deliver(item):
            if item is not None:
                idx, v = item
                for callback in callbacks:
                    callback.on_batch_end(idx, {k: float(v[j]) for j, k in enumerate(keys)})

        for i in range(iterations):
            current_iter = epoch * iterations + i
            if current_iter % lr_decay_iter == 0 and current_iter <= max_iter:
                gen_lr = _utils.poly_lr_scheduler(generator_optimizer, gen_init_lr, current_iter, lr_decay_iter, max_iter, gen_power)
            source_image, source_label = src.next()
            target_image, _ = tgt.next()
            source_image = source_image.to(dev, non_blocking=True)
            source_label = source_label.to(dev, non_blocking=True).squeeze(1)
            target_image = target_image.to(dev, non_blocking=True)
            out = adversarial_step(generator, discriminator, generator_optimizer, discriminator_optimizer, source_image,
                                   source_label, target_image, generator_loss, discriminator_loss, lambda_, iterations, fused=True)
            step = torch.stack([out[k].double() for k in keys] + [out['generator_correct'].double()])
            acc += step
            generator_total += source_label.size(0) * source_label.size(1) * source_label.size(2)
            deliver(late.push(i, step[:4]))
        for item in late.drain():
            deliver(item)
        host = acc.cpu()                                              # the one synchronising read of the epoch
        accuracy = 100. * float(host[4]) / generator_total
        print(f'Epoch Results {epoch}')
        _utils.tabular_print({**{k: float(host[j]) / iterations for j, k in enumerate(keys)}, 'Genrator Accuracy': accuracy,
                              'dis_lr': dis_lr if dis_lr else -1, 'gen_lr': gen_lr if gen_lr else -1})
        for callback in callbacks:
            callback.on_epoch_end(epoch, {'dis_lr': dis_lr if dis_lr else -1, 'gen_lr': gen_lr if gen_lr else -1,
                                          'Genrator Accuracy': accuracy})
        if do_validation != 0 and epoch % do_validation == 0:
            print('-' * 50, 'Validation', '-' * 50)
            validation_mIou, _ = val_GTA5(epoch, generator, val_loader, num_classes, class_names, callbacks, device=device)
            print('-' * 100)
            if validation_mIou > best_mIoU:
                best_mIoU = validation_mIou
                torch.save(generator.state_dict(), 'best_generator.pth')
                torch.save(discriminator.state_dict(), 'best_discriminator.pth')
                print(f'Best Model Saved at Epoch {epoch}')
    for callback in callbacks:
        callback.on_train_end()
