"""Pin the oracle (CPU restatement) against vectors produced by the REAL reference
(tests/golden/*.npz, written by oracle/gen_golden.py in the build container)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bisenet_ref, metrics_ref, weights

SUB = 3


def _input(seed, n, h, w):
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(n, 3, h, w, generator=g)
    y = torch.randint(0, 20, (n, h, w), generator=g)
    return x, y


@pytest.mark.parametrize("name", ["bisenet_64x96", "bisenet_72x104"])
def test_bisenet_oracle_matches_reference_outputs(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    sd = weights.bisenet_r18_state(seed)
    chk = sum(v.double().sum().item() for k, v in sorted(sd.items()) if v.dtype.is_floating_point)
    assert abs(chk - float(gold["weights_checksum"][0])) < 1e-6 * max(1.0, abs(chk)), "seeded weights drifted"
    x, y = _input(seed, n, h, w)
    with torch.no_grad():
        r = bisenet_ref.bisenet_forward(x, weights.clone_state(sd), train=False)
    np.testing.assert_allclose(r[..., ::SUB, ::SUB].numpy(), gold["eval_result"], rtol=1e-4, atol=1e-5)
    assert (r.argmax(1)[..., ::SUB, ::SUB].numpy() == gold["eval_argmax"]).mean() > 0.999
    s = gold["eval_result_sum"]
    assert abs(r.double().sum().item() - s[0]) <= 1e-5 * max(1.0, abs(s[0]) + s[1] ** 0.5)

    sdt = weights.clone_state(sd)
    with torch.no_grad():
        res, s1, s2 = bisenet_ref.bisenet_forward(x, sdt, train=True)
    for k, t in (("train_result", res), ("train_sup1", s1), ("train_sup2", s2)):
        np.testing.assert_allclose(t[..., ::SUB, ::SUB].numpy(), gold[k], rtol=1e-4, atol=2e-5)
    for ign in (19, 255):
        yy = y.clone()
        if ign == 255:
            yy[yy == 19] = 255
        loss = sum(bisenet_ref.ce_loss(t, yy, ign) for t in (res, s1, s2)).item()
        assert abs(loss - float(gold[f"train_loss_ign{ign}"][0])) < 1e-4
    # running buffers were updated like nn.BatchNorm2d does
    for k in gold.files:
        if k.startswith("buf:"):
            np.testing.assert_allclose(sdt[k[4:]].numpy(), gold[k], rtol=1e-4, atol=1e-6)


def test_fast_hist_oracle_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "fast_hist.npz"))
    cases = sorted({k.rsplit("_", 1)[0] for k in gold.files})
    assert len(cases) >= 6
    for c in cases:
        a, b = gold[c + "_label"], gold[c + "_pred"]
        h = metrics_ref.fast_hist(a, b, 19)
        assert h.dtype == np.int64 and h.shape == (19, 19)
        assert (h == gold[c + "_hist"]).all(), c
        assert (metrics_ref.per_class_iou(h) == gold[c + "_iou"]).all(), c
        if a.size <= 4096:
            assert (metrics_ref.fast_hist_loops(a, b, 19) == h).all()


def test_fast_hist_properties():
    rng = np.random.default_rng(0)
    a = rng.integers(-2, 25, size=5000)
    b = rng.integers(0, 19, size=5000)
    h = metrics_ref.fast_hist(a, b, 19)
    assert h.sum() == ((a >= 0) & (a < 19)).sum()
    # additivity over a split of the pixels (what validation.py:55 relies on)
    h2 = metrics_ref.fast_hist(a[:1234], b[:1234], 19) + metrics_ref.fast_hist(a[1234:], b[1234:], 19)
    assert (h == h2).all()
    iou = metrics_ref.per_class_iou(np.zeros((19, 19), dtype=np.int64))
    assert (iou == 0).all()  # empty class -> 0, not NaN (epsilon in the denominator)


@pytest.mark.parametrize("tiny", [True, False])
def test_discriminator_oracle_matches_reference_outputs(golden_dir, tiny):
    """oracle/disc_ref.py against the REAL reference's outputs, losses and gradients (gen_golden.gen_discriminators)."""
    from oracle import disc_ref

    gold = np.load(os.path.join(golden_dir, "discriminators.npz"))
    tag = "tiny" if tiny else "full"
    g = torch.Generator().manual_seed(int(gold["seed"][0]))
    logits = torch.randn(2, 19, 64, 96, generator=g) * 3
    sd = weights.discriminator_state(3, tiny=tiny)
    for target in (0.0, 1.0):
        x = logits.clone().requires_grad_(True)
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        loss, p = disc_ref.adversarial_bce(x, sdg, target)
        loss.backward()
        np.testing.assert_allclose(p.detach().numpy(), gold[tag + "_out"], rtol=1e-5, atol=1e-6)
        assert abs(loss.item() - float(gold[f"{tag}_bce{int(target)}"][0])) < 1e-6
        np.testing.assert_allclose(x.grad[..., ::SUB, ::SUB].numpy(), gold[f"{tag}_bce{int(target)}_dx"], rtol=1e-4, atol=1e-9)
        gn = np.array([sdg[k].grad.double().norm().item() for k in sd])     # state_dict order == parameters() order
        np.testing.assert_allclose(gn, gold[f"{tag}_bce{int(target)}_gnorm"], rtol=1e-5)


def test_deeplab_oracle_matches_reference_outputs(golden_dir):
    """oracle/deeplab_ref.py against the REAL reference's DeepLabV2-R101 (gen_golden.gen_deeplab): eval and train
    logits, loss, every gradient norm, selected gradient tensors, running buffers."""
    from oracle import deeplab_ref

    torch.set_num_threads(os.cpu_count() or 1)
    gold = np.load(os.path.join(golden_dir, "deeplab_72x104.npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(n, 3, h, w, generator=g)
    y = torch.randint(0, 20, (n, h, w), generator=g)
    sd = weights.deeplab_state(seed)
    assert len(sd) == 632
    with torch.no_grad():
        r = deeplab_ref.deeplab_forward(x, weights.clone_state(sd), False)
    np.testing.assert_allclose(r[..., ::SUB, ::SUB].numpy(), gold["eval_result"], rtol=1e-4, atol=1e-5)
    assert (r.argmax(1)[..., ::SUB, ::SUB].numpy() == gold["eval_argmax"]).mean() > 0.999
    sdt = weights.clone_state(sd)
    leaves = {k: v.requires_grad_(True) for k, v in sdt.items() if v.dtype.is_floating_point and "running" not in k}
    out = deeplab_ref.deeplab_forward(x, sdt, True)
    np.testing.assert_allclose(out.detach()[..., ::SUB, ::SUB].numpy(), gold["train_result"], rtol=1e-4, atol=2e-5)
    loss = F.cross_entropy(out, y, ignore_index=19)
    assert abs(loss.item() - float(gold["train_loss_ign19"][0])) < 1e-5
    loss.backward()
    for k, rn in zip((str(s) for s in gold["grad_names"]), gold["grad_norms"]):
        assert abs(leaves[k].grad.double().norm().item() - rn) <= 1e-4 * max(rn, 1e-8), k
    for k in gold.files:
        if k.startswith("grad:"):
            g_ = leaves[k[5:]].grad
            g_ = g_ if g_.numel() < 200000 else g_.flatten()[::37]
            np.testing.assert_allclose(g_.numpy(), gold[k], rtol=2e-3, atol=1e-7)
        if k.startswith("buf:"):
            np.testing.assert_allclose(sdt[k[4:]].detach().numpy(), gold[k], rtol=1e-4, atol=1e-6)


def test_bisenet_resnet101_oracle_matches_reference_outputs(golden_dir):
    """BiSeNet(19, 'resnet101') (build_bisenet.py:95-102, build_contextpath.py:32-56), eval forward."""
    gold = np.load(os.path.join(golden_dir, "bisenet_r101_64x96.npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(n, 3, h, w, generator=g)
    sd = weights.bisenet_r101_state(seed)
    assert len(sd) == int(gold["n_state_keys"][0])            # 1298 state_dict keys incl. the aliased context_path.* duplicates
    with torch.no_grad():
        r = bisenet_ref.bisenet_forward(x, weights.clone_state(sd), train=False)
    ref = torch.from_numpy(gold["eval_result"])
    assert (r[..., ::3, ::3] - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    assert (r.argmax(1)[..., ::3, ::3].numpy() == gold["eval_argmax"]).all()
