"""Generate tests/golden/*.npz by running the REAL reference (imported read-only
from /root/reference through oracle/refshim.py) on seeded synthetic inputs and
the deterministic weights of oracle/weights.py.

Run in the build container only:   python -m oracle.gen_golden
The GPU box has no /root/reference; tests there compare the oracle restatement
and the CUDA path against these committed vectors.

Outputs are stored sub-sampled (every 3rd pixel in H and W) together with
float64 checksums of the full tensors, to keep the fixtures small.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim, weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SUB = 3


def make_input(seed, n, h, w):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g)
    y = torch.randint(0, 20, (n, h, w), generator=g)
    return x, y


def summarize(t: torch.Tensor):
    t = t.detach().double()
    return np.array([t.sum().item(), (t * t).sum().item(), t.abs().max().item()], dtype=np.float64)


def sub(t: torch.Tensor):
    return t.detach()[..., ::SUB, ::SUB].contiguous().numpy().astype(np.float32)


def gen_bisenet(R, name, seed, n, h, w):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x, y = make_input(1000 + seed, n, h, w)
    out = {"shape": np.array([n, h, w]), "seed": np.array([seed])}
    sd = weights.bisenet_r18_state(seed)
    out["weights_checksum"] = np.array([sum(v.double().sum().item() for k, v in sorted(sd.items()) if v.dtype.is_floating_point)])
    # eval
    m = R["BiSeNet"](19, "resnet18")
    m.load_state_dict(weights.clone_state(sd))
    m.eval()
    with torch.no_grad():
        r = m(x)
    out["eval_result"] = sub(r)
    out["eval_result_sum"] = summarize(r)
    out["eval_argmax"] = r.argmax(1)[..., ::SUB, ::SUB].numpy().astype(np.int16)
    # train (batch statistics, running buffers updated)
    m = R["BiSeNet"](19, "resnet18")
    m.load_state_dict(weights.clone_state(sd))
    m.train()
    res, s1, s2 = m(x)
    for k, t in (("train_result", res), ("train_sup1", s1), ("train_sup2", s2)):
        out[k] = sub(t)
        out[k + "_sum"] = summarize(t)
    for ign in (19, 255):
        yy = y.clone()
        if ign == 255:
            yy[yy == 19] = 255
        loss = sum(F.cross_entropy(t, yy, ignore_index=ign) for t in (res, s1, s2))
        out[f"train_loss_ign{ign}"] = np.array([loss.item()], dtype=np.float64)
    loss = sum(F.cross_entropy(t, y, ignore_index=19) for t in (res, s1, s2))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    out["grad_names"] = np.array(sorted(grads.keys()))
    out["grad_norms"] = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    out["grad_none"] = np.array(sorted(k for k, p in m.named_parameters() if p.grad is None))
    # a few full gradient tensors (small ones) for elementwise comparison
    for k in ("conv.weight", "conv.bias", "supervision1.bias", "feature_fusion_module.conv2.weight",
              "attention_refinement_module2.bn.weight", "saptial_path.convblock1.conv1.weight",
              "context_path.features.bn1.weight", "feature_fusion_module.convblock.bn.bias"):
        out["grad:" + k] = grads[k].numpy().astype(np.float32)
    bufs = dict(m.named_buffers())
    for k in ("saptial_path.convblock1.bn.running_mean", "saptial_path.convblock1.bn.running_var",
              "context_path.features.layer4.1.bn2.running_var", "attention_refinement_module1.bn.running_mean",
              "attention_refinement_module2.bn.running_var", "feature_fusion_module.convblock.bn.running_var"):
        out["buf:" + k] = bufs[k].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.size > 8})


def gen_bisenet_r101(R, name, seed, n, h, w):
    """BiSeNet(19, 'resnet101'), eval forward only."""
    x, _ = make_input(1000 + seed, n, h, w)
    sd = weights.bisenet_r101_state(seed)
    m = R["BiSeNet"](19, "resnet101")
    missing = m.load_state_dict(weights.clone_state(sd))
    assert not missing.missing_keys and not missing.unexpected_keys
    m.eval()
    with torch.no_grad():
        r = m(x)
    out = {"shape": np.array([n, h, w]), "seed": np.array([seed]), "eval_result": sub(r), "eval_result_sum": summarize(r),
           "eval_argmax": r.argmax(1)[..., ::SUB, ::SUB].numpy().astype(np.int16), "n_state_keys": np.array([len(m.state_dict())])}
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, r.shape, float(r.abs().max()))


def gen_fast_hist(U):
    out = {}
    rng = np.random.default_rng(7)
    cases = {
        "uniform": (rng.integers(0, 20, size=(2, 37, 53)), rng.integers(0, 19, size=(2, 37, 53))),
        "with255": (np.where(rng.random((3, 16, 31)) < 0.2, 255, rng.integers(0, 19, size=(3, 16, 31))),
                    rng.integers(0, 19, size=(3, 16, 31))),
        "negative": (rng.integers(-3, 22, size=(1, 64, 9)), rng.integers(0, 19, size=(1, 64, 9))),
        "single": (np.array([[5]]), np.array([[7]])),
        "empty": (np.zeros((0,), dtype=np.int64), np.zeros((0,), dtype=np.int64)),
        "allignored": (np.full((4, 4), 19), np.zeros((4, 4), dtype=np.int64)),
    }
    for k, (a, b) in cases.items():
        a = a.astype(np.int64)
        b = b.astype(np.int64)
        h = U["fast_hist"](a, b, 19)
        out[k + "_label"] = a
        out[k + "_pred"] = b
        out[k + "_hist"] = h.astype(np.int64)
        out[k + "_iou"] = U["per_class_iou"](h).astype(np.float64)
    np.savez_compressed(os.path.join(GOLD, "fast_hist.npz"), **out)
    print("wrote fast_hist")


def gen_discriminators(R):
    out = {}
    g = torch.Generator().manual_seed(4242)
    logits = torch.randn(2, 19, 64, 96, generator=g) * 3
    out["seed"] = np.array([4242])
    for tiny in (False, True):
        cls = R["TinyDomainDiscriminator"] if tiny else R["DomainDiscriminator"]
        d = cls(19)
        d.load_state_dict(weights.discriminator_state(3, tiny=tiny))
        d.train()
        x = logits.clone().requires_grad_(True)
        p = d(F.softmax(x, dim=1))
        tag = "tiny" if tiny else "full"
        out[tag + "_out"] = p.detach().numpy().astype(np.float32)
        for target in (0.0, 1.0):
            for prm in d.parameters():
                prm.grad = None
            x.grad = None
            loss = F.binary_cross_entropy_with_logits(p, torch.full_like(p, target))
            loss.backward(retain_graph=True)
            out[f"{tag}_bce{int(target)}"] = np.array([loss.item()])
            out[f"{tag}_bce{int(target)}_dx_sum"] = summarize(x.grad)
            out[f"{tag}_bce{int(target)}_dx"] = sub(x.grad)
            out[f"{tag}_bce{int(target)}_gnorm"] = np.array([prm.grad.double().norm().item() for prm in d.parameters()])
    np.savez_compressed(os.path.join(GOLD, "discriminators.npz"), **out)
    print("wrote discriminators")


def gen_deeplab(R, name, seed, n, h, w):
    """DeepLabV2-R101 (models/deeplabv2/deeplabv2.py) eval + train forward, CE loss, gradients."""
    from oracle import deeplab_ref  # noqa: F401  (same input recipe as the tests)

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x, y = make_input(2000 + seed, n, h, w)
    out = {"shape": np.array([n, h, w]), "seed": np.array([seed])}
    sd = weights.deeplab_state(seed)
    m = R["get_deeplab_v2"](19, pretrain=False)
    m.load_state_dict(weights.clone_state(sd))
    m.eval()
    with torch.no_grad():
        r = m(x)
    out["eval_result"] = sub(r)
    out["eval_result_sum"] = summarize(r)
    out["eval_argmax"] = r.argmax(1)[..., ::SUB, ::SUB].numpy().astype(np.int16)
    m = R["get_deeplab_v2"](19, pretrain=False)
    m.load_state_dict(weights.clone_state(sd))
    m.train()
    res, a1, a2 = m(x)
    assert a1 is None and a2 is None
    out["train_result"] = sub(res)
    out["train_result_sum"] = summarize(res)
    loss = F.cross_entropy(res, y, ignore_index=19)
    out["train_loss_ign19"] = np.array([loss.item()], dtype=np.float64)
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    out["grad_names"] = np.array(sorted(grads.keys()))
    out["grad_norms"] = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    out["grad_none"] = np.array(sorted(k for k, p in m.named_parameters() if p.grad is None))
    for k in ("conv1.weight", "layer1.0.conv1.weight", "layer2.0.downsample.0.weight", "layer3.11.conv2.weight",
              "layer6.conv2d_list.3.bias", "layer6.conv2d_list.0.weight"):
        g_ = grads[k]
        out["grad:" + k] = (g_ if g_.numel() < 200000 else g_.flatten()[::37]).numpy().astype(np.float32)
    bufs = dict(m.named_buffers())
    for k in ("bn1.running_mean", "layer2.0.downsample.1.running_var", "layer3.22.bn3.running_var", "layer4.2.bn2.running_mean"):
        out["buf:" + k] = bufs[k].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name)


def config1_inputs():
    """BASELINE.json configs[0] / SURVEY 8(c) anchor, verbatim: data generator seed 1234, x ~ N(0,1) [2,3,512,1024],
    labels uniform in [0,19] (19 = ignore)."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 512, 1024, generator=g)
    y = torch.randint(0, 20, (2, 512, 1024), generator=g)
    return x, y


def gen_config1(R, name="config1"):
    """BASELINE config 1 on the REAL reference: torch.manual_seed(42) -> BiSeNet(19, 'resnet18') (the constructor's own
    random init), eval and train forward, 3xCE with ignore_index 19 / 255, gradient norms.  Stored: per-tensor
    checksums of the seeded state_dict (pins that the drop-in constructor draws the same weights), logits sub-sampled
    every 16th pixel, argmax every 4th pixel, float64 checksums of the full tensors."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x, y = config1_inputs()
    torch.manual_seed(42)
    m = R["BiSeNet"](19, "resnet18")
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    keys = [k for k, v in sd0.items() if v.dtype.is_floating_point]
    out = {"state_keys": np.array(keys),
           "state_sum": np.array([sd0[k].double().sum().item() for k in keys]),
           "state_sumsq": np.array([(sd0[k].double() ** 2).sum().item() for k in keys]),
           "x_sum": summarize(x), "y_sum": np.array([y.double().sum().item()])}
    S, SA = 16, 4
    m.eval()
    with torch.no_grad():
        r = m(x)
    out["eval_result"] = r[..., ::S, ::S].contiguous().numpy().astype(np.float32)
    out["eval_result_sum"] = summarize(r)
    out["eval_argmax"] = r.argmax(1)[..., ::SA, ::SA].numpy().astype(np.int8)
    m.train()
    res, s1, s2 = m(x)
    for k, t in (("train_result", res), ("train_sup1", s1), ("train_sup2", s2)):
        out[k] = t.detach()[..., ::S, ::S].contiguous().numpy().astype(np.float32)
        out[k + "_sum"] = summarize(t)
        out[k + "_argmax"] = t.argmax(1)[..., ::SA, ::SA].numpy().astype(np.int8)
    for ign in (19, 255):
        yy = y.clone()
        if ign == 255:
            yy[yy == 19] = 255
        loss = sum(F.cross_entropy(t, yy, ignore_index=ign) for t in (res, s1, s2))
        out[f"train_loss_ign{ign}"] = np.array([loss.item()], dtype=np.float64)
    loss = sum(F.cross_entropy(t, y, ignore_index=19) for t in (res, s1, s2))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    out["grad_names"] = np.array(sorted(grads.keys()))
    out["grad_norms"] = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, "loss ign19", out["train_loss_ign19"], "ign255", out["train_loss_ign255"])


def main():
    assert refshim.available(), "reference tree not found"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    R = refshim.load_models()
    if "config1" in sys.argv[1:]:
        gen_config1(R)
        return
    U = refshim.load_utils_functions()
    gen_fast_hist(U)
    gen_bisenet(R, "bisenet_64x96", 0, 2, 64, 96)
    gen_bisenet(R, "bisenet_72x104", 1, 2, 72, 104)
    gen_discriminators(R)
    gen_deeplab(R, "deeplab_72x104", 2, 2, 72, 104)
    gen_bisenet_r101(R, "bisenet_r101_64x96", 3, 1, 64, 96)


if __name__ == "__main__":
    main()
