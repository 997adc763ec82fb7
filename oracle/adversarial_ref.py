"""ORACLE (test infrastructure, not product): CPU restatement of one iteration of the reference's
adversarial training loop, train.py:adversarial_train (:177-270), on the functional oracles
(oracle/bisenet_ref.py, oracle/disc_ref.py) with autograd.  Only tests/ and bench.py's CPU legs may
import this.  Parity is pinned through its parts: bisenet_ref and disc_ref are each checked against
outputs of the real reference (tests/golden/*.npz); this file only restates the ORDER of operations:

  :192-193  freeze D            :199-213  G(source) -> 3 x CE / iterations -> backward
  :218-233  G(target) -> softmax -> D -> lambda * BCE(., 1) / iterations -> backward (into G only)
  :238-243  unfreeze D, detach  :245-251  D(softmax(source)) -> BCE(., 1) / iterations -> backward
  :256-262  D(softmax(target)) -> BCE(., 0) / iterations -> backward
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import bisenet_ref, disc_ref


def _leaves(sd, prefixes=None):
    out = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running" not in k and (prefixes is None or k.startswith(prefixes)):
            v.requires_grad_(True)
            out[k] = v
    return out


G_PREFIXES = ("context_path.features", "saptial", "attention", "supervision", "feature_fusion", "conv.")


def adversarial_iteration(gen_sd, disc_sd, source_image, source_label, target_image, ignore_index, lambda_, iterations):
    """Returns (losses dict of floats, generator grads {name: tensor}, discriminator grads {name: tensor},
    source main-head logits).  gen_sd's running BatchNorm buffers are updated in place (two train-mode forwards)."""
    g_leaves = _leaves(gen_sd, G_PREFIXES)
    d_leaves = _leaves(disc_sd)
    # generator, source
    outs = bisenet_ref.bisenet_forward(source_image, gen_sd, train=True)
    loss_gen_source = sum(bisenet_ref.ce_loss(t, source_label, ignore_index) for t in outs) / iterations
    loss_gen_source.backward()
    source_features = outs[0]
    # generator, target, D frozen: gradients flow through D into G only
    frozen = {k: v.detach() for k, v in disc_sd.items()}
    target_feature = bisenet_ref.bisenet_forward(target_image, gen_sd, train=True)[0]
    loss_adv, _ = disc_ref.adversarial_bce(target_feature, frozen, 1.0, lambda_ / iterations)
    loss_adv.backward()
    # discriminator on detached predictions
    loss_ds, _ = disc_ref.adversarial_bce(source_features.detach(), disc_sd, 1.0, 1.0 / iterations)
    loss_ds.backward()
    loss_dt, _ = disc_ref.adversarial_bce(target_feature.detach(), disc_sd, 0.0, 1.0 / iterations)
    loss_dt.backward()
    losses = dict(loss_gen_source=loss_gen_source.item(), loss_adversarial=loss_adv.item(), loss_disc_source=loss_ds.item(),
                  loss_disc_target=loss_dt.item())
    return (losses, {k: v.grad for k, v in g_leaves.items() if v.grad is not None},
            {k: v.grad for k, v in d_leaves.items() if v.grad is not None}, source_features.detach())
