"""Loop bodies of the reference's training loops expressed on the drop-in modules.

`adversarial_step` is one iteration of train.py:adversarial_train (:177-275): generator on the source
batch with 3 x CE, generator on the target batch fooling the frozen discriminator, discriminator on the
detached source (label 1) and target (label 0) predictions, both optimizers stepped.  With
`fused=False` it issues exactly the reference's sequence of public calls (module __call__,
`F.softmax`, the stock criteria, `.backward()`, `.detach()`, `requires_grad` toggling), which is how
the parity tests check that the call sites stay intact; with `fused=True` it uses the fast paths
behind the same modules: bilinear-resize + CE + argmax fused on the 1/8-resolution logits for the
source batch, softmax fused into the discriminator's first kernel, BCE against a constant target as
one kernel.  Losses are returned as 0-dim device tensors — nothing here synchronises with the host
(the reference's `.item()` calls, train.py:214,234,253,264, are the caller's business).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _set_requires_grad(module, flag):
    for p in module.parameters():
        p.requires_grad = flag


def adversarial_step(generator, discriminator, generator_optimizer, discriminator_optimizer, source_image, source_label,
                     target_image, generator_loss, discriminator_loss, lambda_, iterations, fused=False):
    """One iteration of train.py:177-275.  Returns dict(loss_gen_source, loss_adversarial, loss_disc_source,
    loss_disc_target, generator_correct) of device tensors."""
    from .bisenet_autograd import bisenet_fused_ce
    from .disc_engine import bce_with_logits_const

    generator_optimizer.zero_grad()
    discriminator_optimizer.zero_grad()
    # the discriminator is frozen while the generator trains (train.py:192-193)
    _set_requires_grad(discriminator, False)

    # ---- generator, source batch: 3 x CE / iterations (train.py:199-213)
    correct = None
    if fused and generator.training and hasattr(generator, "rtsds_precision") and hasattr(generator, "saptial_path"):
        ignore = getattr(generator_loss, "ignore_index", -100)
        loss_sum, pred, stats, source_features = bisenet_fused_ce(generator, source_image, source_label, ignore, return_logits=True)
        loss_gen_source = loss_sum / iterations
        correct = stats[0, 2]
    else:
        out = generator(source_image)
        if isinstance(out, tuple):
            loss_gen_source = generator_loss(out[0], source_label)
            loss_gen_source = loss_gen_source + generator_loss(out[1], source_label) if out[1] is not None else loss_gen_source
            loss_gen_source = loss_gen_source + generator_loss(out[2], source_label) if out[2] is not None else loss_gen_source
            source_features = out[0]
        else:
            loss_gen_source = generator_loss(out, source_label)
            source_features = out
        loss_gen_source = loss_gen_source / iterations
    # data parallel: the generator's gradient is the sum of this backward and the adversarial one below (train.py:213,233);
    # it is all-reduced ONCE, after the second (rtsds_b200/bisenet_autograd.py:_finish_backward)
    generator.rtsds_ddp_hold = bool(getattr(generator, "rtsds_ddp", False))
    try:
        loss_gen_source.backward()
    finally:
        generator.rtsds_ddp_hold = False

    # ---- generator, target batch: lambda * BCE(D(softmax(G(target))), 1) / iterations (train.py:218-233)
    out = generator(target_image)
    target_feature = out[0] if isinstance(out, tuple) else out
    if fused and hasattr(discriminator, "forward_logits"):
        predicted = discriminator.forward_logits(target_feature)
        loss_adversarial = bce_with_logits_const(predicted, 1.0, lambda_ / iterations)
    else:
        predicted = discriminator(F.softmax(target_feature, dim=1))
        loss_adversarial = lambda_ * discriminator_loss(predicted, torch.ones(predicted.size(), device=predicted.device))
        loss_adversarial = loss_adversarial / iterations
    loss_adversarial.backward()

    # ---- discriminator on the detached predictions (train.py:238-262)
    _set_requires_grad(discriminator, True)
    source_features = source_features.detach()
    target_feature = target_feature.detach()
    losses = []
    for feat, label in ((source_features, 1.0), (target_feature, 0.0)):
        if fused and hasattr(discriminator, "forward_logits"):
            predicted = discriminator.forward_logits(feat)
            loss = bce_with_logits_const(predicted, label, 1.0 / iterations)
        else:
            predicted = discriminator(F.softmax(feat, dim=1))
            loss = discriminator_loss(predicted, torch.full(predicted.size(), label, device=predicted.device)) / iterations
        loss.backward()
        losses.append(loss)

    generator_optimizer.step()
    discriminator_optimizer.step()
    if correct is None:
        correct = source_features.argmax(dim=1).eq(source_label).sum()       # train.py:272-273
    return dict(loss_gen_source=loss_gen_source.detach(), loss_adversarial=loss_adversarial.detach(),
                loss_disc_source=losses[0].detach(), loss_disc_target=losses[1].detach(), generator_correct=correct)


def supervised_step(model, optimizer, inputs, targets, criterion, fused=False):
    """Loop body of train.py:train (:74-106): zero_grad, forward, CE over the 1 or 3 heads, backward, step,
    pixel accuracy.  Returns (loss, correct_pixels) as device tensors."""
    from .bisenet_autograd import bisenet_fused_ce

    optimizer.zero_grad()
    if fused and model.training and hasattr(model, "saptial_path"):
        loss, pred, stats = bisenet_fused_ce(model, inputs, targets, getattr(criterion, "ignore_index", -100))
        correct = stats[0, 2]
    else:
        outputs = model(inputs)
        if isinstance(outputs, tuple):
            main_output = outputs[0]
            loss = criterion(main_output, targets)
            for aux in outputs[1:]:
                if aux is not None:
                    loss = loss + criterion(aux, targets)
        else:
            main_output = outputs
            loss = criterion(main_output, targets)
        correct = main_output.max(1)[1].eq(targets).sum()
    loss.backward()
    optimizer.step()
    return loss.detach(), correct
