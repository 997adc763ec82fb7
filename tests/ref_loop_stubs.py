"""TEST INFRASTRUCTURE: restatement of the reference's loop bodies — train.py:train (:24-128) and validation.py:val
(:12-77) — with every per-batch host read the reference performs (`loss.item()` :99, `.sum().item()` :106,
`.cpu().numpy()` + numpy fast_hist :54), so that the GPU box, which has no /root/reference, still has the reference's
call sequence to run the drop-in modules through and to compare the sync-free loops (rtsds_b200/loops.py) against.
tests/test_reference_call_sites.py (build container, reference present) pins these restatements to the real loops: same
kernel-launch sequence, same callback calls, same arguments."""
import numpy as np
import torch

import utils


def train(epoch, model, train_loader, criterion, optimizer, init_lr, max_iter, power=0.9, lr_decay_iter=1.0, device='cpu', callbacks=[]):
    for cb in callbacks:
        cb.on_train_begin()
    model.train()
    running_loss, correct, total = 0.0, 0, 0
    for batch_idx, (inputs, targets) in enumerate(train_loader):
        it = epoch * len(train_loader) + batch_idx                                    # :66
        if it % lr_decay_iter == 0 and it <= max_iter:                                # :68-69
            utils.poly_lr_scheduler(optimizer, init_lr, it, lr_decay_iter, max_iter, power)
        inputs = inputs.to(device)                                                    # :71-72
        targets = targets.to(device).squeeze(1)
        optimizer.zero_grad()
        outputs = model(inputs)                                                       # :77
        main_output, aux1, aux2 = outputs if isinstance(outputs, tuple) else (outputs, None, None)
        loss = criterion(main_output, targets)                                        # :86-92
        if aux1 is not None:
            loss += criterion(aux1, targets)
        if aux2 is not None:
            loss += criterion(aux2, targets)
        loss.backward()                                                               # :95-96
        optimizer.step()
        running_loss += loss.item()                                                   # :99  (sync)
        _, predicted = main_output.max(1)                                             # :102
        total += targets.size(0) * targets.size(1) * targets.size(2)
        correct += predicted.eq(targets).sum().item()                                 # :106 (sync)
        for cb in callbacks:                                                          # :109-113
            cb.on_batch_end(batch_idx, {'train_loss': loss.item(), 'train_accuracy': 100. * correct / total})
    train_loss = running_loss / len(train_loader)
    train_accuracy = 100. * correct / total
    for cb in callbacks:                                                              # :122-126
        cb.on_epoch_end(epoch, {'train_loss': train_loss, 'train_accuracy': train_accuracy})
    return model


def val(epoch, model, val_loader, num_classes, device='cpu', callbacks=[]):
    for cb in callbacks:
        cb.on_validation_begin()
    model.eval()
    total_hist = np.zeros((num_classes, num_classes))
    with torch.no_grad():
        for batch_idx, (inputs, targets) in enumerate(val_loader):
            inputs = inputs.to(device)
            targets = targets.to(device).squeeze(1)
            outputs = model(inputs)                                                   # :45
            if isinstance(outputs, tuple):
                outputs = outputs[0]
            predicted = torch.argmax(outputs, dim=1)                                  # :51
            total_hist += utils.fast_hist(targets.cpu().numpy(), predicted.cpu().numpy(), num_classes)    # :54-55 (sync)
            loss = 1. - np.sum(np.diag(total_hist)) / np.sum(total_hist)              # :58-62
            for cb in callbacks:
                cb.on_validation_batch_end(batch_idx, loss)
    mean_iou = np.nanmean(utils.per_class_iou(total_hist))                            # :69-70
    for cb in callbacks:
        cb.on_validation_end(mean_iou)
    return mean_iou
