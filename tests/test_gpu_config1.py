"""BASELINE.json configs[0] VERBATIM on the GPU (SURVEY 8c/8d "Config 1"): 2x3x512x1024, 19 classes,
torch.manual_seed(42) constructor weights, data seed 1234 — eval and train mode, ignore_index 19 and 255.

north_star's numbers are asserted unconditionally:
  * eval, production dtype (fp16 operands, fp32 TMEM accumulation): logits rel <= 2e-2, argmax agreement >= 99.9 %
  * eval + train, fp32 check mode: rel <= 1e-4, argmax >= 99.9 %, loss to 1e-5
against (a) the CPU oracle run live at full resolution and (b) the REAL reference's outputs committed in
tests/golden/config1.npz (the oracle itself is pinned to those by tests/test_config1_cpu.py).

bf16 — the TRAINING dtype — is measured and recorded on the same inputs.  In eval mode an ideal bf16 pipeline
reaches 99.78 % argmax agreement on this configuration (bf16 WEIGHT rounding alone gives 99.78 %, activation rounding
alone 99.92 %: oracle/bisenet_bf16.py), which is why inference runs fp16 (99.98 %).  In train mode the reference function
itself is ill-conditioned on this input: the ARM BatchNorm normalises over the N = 2 pooled vectors of two i.i.d.
noise images, d/sqrt(d^2+eps) with |d| ~ sqrt(eps), a gain of ~300 on any upstream round-off; an ideal fp16 pipeline
reaches 98.9 % there, ideal bf16 91 %, so the train-mode bar for 16-bit operands is "no further from fp32 than the ideal
emulation", with the loss (well conditioned) held to 1e-3.  Every figure goes to gpurun_out/r02_parity.json."""
import os

import pytest
import torch
import torch.nn.functional as F

import config1
from gpu_util import rel_err
from oracle import bisenet_bf16, bisenet_ref
from parity_log import record

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1():
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = config1.inputs()
    m = config1.seeded_model()
    sd = config1.state_clone(m)
    with torch.no_grad():
        ref_eval = bisenet_ref.bisenet_forward(x, config1.state_clone(m), train=False)
        ref_train = bisenet_ref.bisenet_forward(x, config1.state_clone(m), train=True)
    return dict(x=x, y=y, sd=sd, ref_eval=ref_eval, ref_train=ref_train, gold=config1.golden())


def _fresh(c1, precision, eval_precision="fp16"):
    m = config1.seeded_model()
    m.load_state_dict(c1["sd"])
    m.rtsds_precision = precision
    m.rtsds_eval_precision = eval_precision
    return m.cuda()


def _agree(a, b):
    return (a.argmax(1) == b.argmax(1)).float().mean().item()


@pytest.mark.parametrize("mode,tol", [("fp16", 2e-2), ("fp32", 1e-4)])
def test_config1_eval_meets_north_star(cuda, c1, mode, tol):
    m = _fresh(c1, "fp32" if mode == "fp32" else "bf16").eval()
    out = m(c1["x"].cuda()).cpu()
    ref, gold = c1["ref_eval"], c1["gold"]
    e, a = rel_err(out, ref), _agree(out, ref)
    S, SA = config1.S_LOGIT, config1.S_ARGMAX
    eg = rel_err(out[..., ::S, ::S], torch.from_numpy(gold["eval_result"]))
    ag = float((out.argmax(1)[..., ::SA, ::SA].numpy() == gold["eval_argmax"]).mean())
    record(f"config1/eval/{mode}", rel=e, argmax_agree=a, rel_vs_reference_golden=eg, argmax_vs_reference_golden=ag,
           tol_rel=tol, tol_argmax=0.999)
    assert e <= tol, (mode, e)
    assert a >= 0.999, (mode, a)
    assert eg <= tol and ag >= 0.999, (mode, eg, ag)


def test_config1_eval_bf16_recorded(cuda, c1):
    """bf16 inference (model.rtsds_eval_precision = 'bf16'): within north_star's logits tolerance; its argmax agreement is
    bounded by bf16 weight rounding (see module docstring) — the CUDA path must sit on that floor, not below it."""
    m = _fresh(c1, "bf16", "bf16").eval()
    out = m(c1["x"].cuda()).cpu()
    ref = c1["ref_eval"]
    with torch.no_grad():
        emu = bisenet_bf16.bisenet_eval_bf16(c1["x"], c1["sd"])
    e, a, fe, fa = rel_err(out, ref), _agree(out, ref), rel_err(emu, ref), _agree(emu, ref)
    record("config1/eval/bf16", rel=e, argmax_agree=a, ideal_bf16_rel=fe, ideal_bf16_argmax=fa, tol_rel=2e-2)
    assert e <= 2e-2, e
    assert e <= 1.6 * fe + 2e-3 and a >= fa - 0.002, (e, a, fe, fa)


def _losses(outs, y):
    res = {}
    for ign in (19, 255):
        yy = y.clone()
        if ign == 255:
            yy[yy == 19] = 255
        res[ign] = sum(F.cross_entropy(t, yy.to(t.device), ignore_index=ign) for t in outs).item()
    return res


def test_config1_train_fp32_meets_north_star(cuda, c1):
    m = _fresh(c1, "fp32").train()
    outs = [t.detach() for t in m(c1["x"].cuda())]
    gold = c1["gold"]
    fig = {}
    for t, r, k in zip(outs, c1["ref_train"], ("result", "sup1", "sup2")):
        fig[f"rel_{k}"] = rel_err(t.cpu(), r)
        fig[f"argmax_{k}"] = _agree(t.cpu(), r)
    ls = _losses(outs, c1["y"])
    fig.update(loss_ign19=ls[19], loss_ign255=ls[255], ref_loss_ign19=float(gold["train_loss_ign19"][0]),
               ref_loss_ign255=float(gold["train_loss_ign255"][0]))
    record("config1/train_forward/fp32", tol_rel=1e-4, tol_argmax=0.999, **fig)
    for k in ("result", "sup1", "sup2"):
        assert fig[f"rel_{k}"] <= 1e-4 and fig[f"argmax_{k}"] >= 0.999, (k, fig)
    for ign in (19, 255):
        assert abs(ls[ign] - float(gold[f"train_loss_ign{ign}"][0])) <= 1e-5 * 10.14, (ign, ls)


def test_config1_train_bf16_on_the_ideal_bf16_floor(cuda, c1):
    m = _fresh(c1, "bf16").train()
    outs = [t.detach() for t in m(c1["x"].cuda())]
    with torch.no_grad():
        emu = bisenet_bf16.bisenet_train_bf16(c1["x"], {k: v.clone() for k, v in c1["sd"].items()})
    gold = c1["gold"]
    fig = {}
    for t, r, em, k in zip(outs, c1["ref_train"], emu, ("result", "sup1", "sup2")):
        fig[f"rel_{k}"], fig[f"argmax_{k}"] = rel_err(t.cpu(), r), _agree(t.cpu(), r)
        fig[f"ideal_bf16_rel_{k}"], fig[f"ideal_bf16_argmax_{k}"] = rel_err(em, r), _agree(em, r)
    ls = _losses(outs, c1["y"])
    ref_loss = float(gold["train_loss_ign19"][0])
    fig.update(loss_ign19=ls[19], loss_ign255=ls[255], ref_loss=ref_loss, loss_rel=abs(ls[19] - ref_loss) / ref_loss)
    record("config1/train_forward/bf16", **fig)
    assert fig["loss_rel"] <= 1e-3 and abs(ls[255] - ref_loss) / ref_loss <= 1e-3, fig
    for k in ("result", "sup1", "sup2"):
        assert fig[f"rel_{k}"] <= 1.6 * fig[f"ideal_bf16_rel_{k}"] + 5e-3, (k, fig)
        assert fig[f"argmax_{k}"] >= fig[f"ideal_bf16_argmax_{k}"] - 0.03, (k, fig)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_config1_backward_gradient_norms_vs_reference(cuda, c1, precision):
    """loss.backward() through the stock criterion call site (train.py:86-95); every parameter's gradient norm against the
    real reference's (golden).  ARM conv bias gradients are analytically zero (the BatchNorm removes them): skipped."""
    m = _fresh(c1, precision).train()
    outs = m(c1["x"].cuda())
    loss = sum(F.cross_entropy(t, c1["y"].cuda(), ignore_index=19) for t in outs)
    loss.backward()
    gold = c1["gold"]
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert grads["context_path.features.fc.weight"] is None
    errs = {}
    for k, rn in zip((str(s) for s in gold["grad_names"]), gold["grad_norms"]):
        if k.startswith("attention_refinement_module") and (k.endswith("conv.bias") or (precision == "bf16" and k.endswith("conv.weight"))):
            continue
        errs[k] = abs(grads[k].double().norm().item() - rn) / max(rn, 1e-12)
    worst = max(errs, key=errs.get)
    med = sorted(errs.values())[len(errs) // 2]
    fig = dict(worst_param=worst, worst_grad_norm_rel=errs[worst], median_grad_norm_rel=med, loss=loss.item())
    if precision == "fp32":
        record("config1/backward/fp32", **fig)
        # 1.8e-3 .. 2.1e-3 observed on the stem BatchNorm bias (the end of the whole chain; the order of the fp32 statistics
        # atomics varies run to run); every other parameter <= 1e-3, the median 6e-5
        assert errs[worst] <= 5e-3 and med <= 5e-4, (worst, errs[worst], med)
        return
    # bf16: gradients of this configuration inherit the train-mode ill-conditioning (module docstring); the yardstick is
    # the ideal-bf16 emulation (fp32 oracle with straight-through bf16 rounding of the stored tensors, exact fp32 backward)
    sde = {k: v.clone() for k, v in c1["sd"].items()}
    leaves = {}
    for k, v in sde.items():
        if v.dtype.is_floating_point and "running" not in k and k in grads and grads[k] is not None:
            v.requires_grad_(True)
            leaves[k] = v
    eo = bisenet_bf16.bisenet_train_bf16(c1["x"], sde)
    sum(bisenet_ref.ce_loss(t, c1["y"], 19) for t in eo).backward()
    emu = {k: abs(leaves[k].grad.double().norm().item() - rn) / max(rn, 1e-12)
           for k, rn in zip((str(s) for s in gold["grad_names"]), gold["grad_norms"]) if k in errs and leaves[k].grad is not None}
    emed = sorted(emu.values())[len(emu) // 2]
    record("config1/backward/bf16", ideal_bf16_median_grad_norm_rel=emed, ideal_bf16_worst_grad_norm_rel=max(emu.values()), **fig)
    # Both are single draws from a chaotic map (train-mode BatchNorm over N = 2, fp32 atomics order): over repeated runs the
    # CUDA path's median norm error ranged 0.14 .. 0.25 with the emulation at 0.07 (worst 0.26).  What is asserted here is
    # that the gradients are of the right size everywhere; the floor-relative bf16 gradient test at a well-conditioned batch
    # (6 x 128 x 256) is tests/test_gpu_bisenet.py::test_train_backward_bf16_is_as_good_as_ideal_bf16.
    assert med <= 0.5 and errs[worst] <= 1.0, (med, worst, errs[worst], emed)
