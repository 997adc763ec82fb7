"""CPU half of BASELINE config 1: the drop-in constructor draws the reference's weights, and the oracle restatement
reproduces the REAL reference's config-1 outputs (tests/golden/config1.npz) — so the live oracle the GPU tests compare
against at full resolution is pinned."""
import numpy as np
import torch

import config1
from oracle import bisenet_ref


def test_dropin_constructor_draws_the_reference_weights():
    gold = config1.golden()
    sd = config1.seeded_model().state_dict()
    keys = [str(k) for k in gold["state_keys"]]
    assert [k for k, v in sd.items() if v.dtype.is_floating_point] == keys          # same keys, same order
    for k, s1, s2 in zip(keys, gold["state_sum"], gold["state_sumsq"]):
        v = sd[k].double()
        assert abs(v.sum().item() - s1) <= 1e-9 * max(1.0, abs(s1)), k
        assert abs((v * v).sum().item() - s2) <= 1e-9 * max(1.0, abs(s2)), k


def test_oracle_reproduces_the_reference_on_config1():
    gold = config1.golden()
    x, y = config1.inputs()
    assert abs(x.double().sum().item() - gold["x_sum"][0]) < 1e-6 and y.double().sum().item() == gold["y_sum"][0]
    m = config1.seeded_model()
    S, SA = config1.S_LOGIT, config1.S_ARGMAX
    with torch.no_grad():
        ev = bisenet_ref.bisenet_forward(x, config1.state_clone(m), train=False)
        tr = bisenet_ref.bisenet_forward(x, config1.state_clone(m), train=True)
    ref = torch.from_numpy(gold["eval_result"])
    assert (ev[..., ::S, ::S] - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    assert (ev.argmax(1)[..., ::SA, ::SA].numpy() == gold["eval_argmax"]).mean() >= 0.99999
    for t, k in zip(tr, ("train_result", "train_sup1", "train_sup2")):
        ref = torch.from_numpy(gold[k])
        assert (t[..., ::S, ::S] - ref).abs().max().item() <= 2e-5 * ref.abs().max().item(), k
    for ign in (19, 255):
        yy = y.clone()
        if ign == 255:
            yy[yy == 19] = 255
        loss = sum(bisenet_ref.ce_loss(t, yy, ign) for t in tr).item()
        assert abs(loss - float(gold[f"train_loss_ign{ign}"][0])) < 1e-5, (ign, loss)
    assert abs(float(gold["train_loss_ign19"][0]) - 10.139338) < 1e-5        # SURVEY 8(c) anchor
