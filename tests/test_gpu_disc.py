"""Parity of the CUDA discriminators (drop-in models.domain_shift.adversarial.model) against
(a) golden vectors produced by the REAL reference (tests/golden/discriminators.npz) and (b) the CPU
oracle (oracle/disc_ref.py) on the same seeded inputs and weights.  fp32 check mode: rel <= 1e-4
(BASELINE.json); bf16: rel <= 2e-2 on outputs, gradients compared in norm."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import disc_ref, weights

from gpu_util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
SUB = 3


def _model(tiny, precision, seed=3):
    from models.domain_shift.adversarial.model import DomainDiscriminator, TinyDomainDiscriminator

    m = (TinyDomainDiscriminator if tiny else DomainDiscriminator)(19)
    m.load_state_dict(weights.discriminator_state(seed, tiny=tiny))
    m.rtsds_precision = precision
    return m.cuda().train()


def _golden_logits():
    g = torch.Generator().manual_seed(4242)
    return torch.randn(2, 19, 64, 96, generator=g) * 3


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tiny", [True, False])
def test_discriminator_vs_reference_golden(cuda, golden_dir, tiny, precision, fused):
    gold = np.load(os.path.join(golden_dir, "discriminators.npz"))
    tag = "tiny" if tiny else "full"
    m = _model(tiny, precision)
    tol = 1e-4 if precision == "fp32" else 2e-2
    for target in (0.0, 1.0):
        x = _golden_logits().cuda().requires_grad_(True)
        p = m.forward_logits(x) if fused else m(F.softmax(x, dim=1))
        assert p.shape == (2, 1, 1, 1) and p.dtype == torch.float32
        ref_out = torch.from_numpy(gold[tag + "_out"])
        assert (p.detach().cpu() - ref_out).abs().max().item() < tol * max(1.0, ref_out.abs().max().item()), (p, ref_out)
        for prm in m.parameters():
            prm.grad = None
        loss = F.binary_cross_entropy_with_logits(p, torch.full_like(p, target))
        loss.backward()
        assert abs(loss.item() - float(gold[f"{tag}_bce{int(target)}"][0])) < tol
        dx_ref = torch.from_numpy(gold[f"{tag}_bce{int(target)}_dx"])
        if precision == "fp32":
            e = rel_err(x.grad[..., ::SUB, ::SUB].cpu(), dx_ref)
            assert e < 2e-4, e
        else:      # bf16 flips LeakyReLU masks of near-zero activations: compare the gradient as a vector
            e = rel_l2(x.grad[..., ::SUB, ::SUB].cpu(), dx_ref)     # ideal-bf16 floor: 3e-2 (tiny) / 7e-2 (full), see below
            assert e < 0.1 and rel_err(x.grad[..., ::SUB, ::SUB].cpu(), dx_ref) < 0.3, e
        gn = np.array([prm.grad.double().norm().item() for prm in m.parameters()])
        gref = gold[f"{tag}_bce{int(target)}_gnorm"]
        assert np.all(np.abs(gn - gref) <= (2e-4 if precision == "fp32" else 6e-2) * np.maximum(gref, 1e-8)), (gn, gref)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 8e-2)])
@pytest.mark.parametrize("tiny", [True, False])
@pytest.mark.parametrize("n,h,w", [(2, 90, 160), (1, 67, 131), (3, 128, 256)])
def test_discriminator_vs_oracle_elementwise(cuda, tiny, precision, tol, n, h, w):
    """Every parameter gradient and the input gradient, incl. odd sizes (h, w not multiples of 2^k): elementwise
    (max-abs) in the fp32 check mode; as vectors (relative L2) in bf16, where the tcgen05 path must also agree
    with the CUDA-core bf16 path on identical operands (the residual is bf16 round-off, not the kernels)."""
    if not tiny and min(h, w) < 64:
        pytest.skip("too small for 5 stride-2 convs")
    sd = weights.discriminator_state(5, tiny=tiny)
    g = torch.Generator().manual_seed(77)
    logits = torch.randn(n, 19, h, w, generator=g) * 2
    xr = logits.clone().requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_r, p_r = disc_ref.adversarial_bce(xr, sdr, 1.0, 0.37)
    loss_r.backward()
    m = _model(tiny, precision, seed=5)
    x = logits.cuda().requires_grad_(True)
    p = m(F.softmax(x, dim=1))
    loss = 0.37 * F.binary_cross_entropy_with_logits(p, torch.ones_like(p))
    loss.backward()
    if precision == "fp32":
        assert rel_err(p.detach().cpu(), p_r.detach()) < tol
        assert rel_err(x.grad.cpu(), xr.grad) < tol, rel_err(x.grad.cpu(), xr.grad)
        for k, prm in m.named_parameters():
            assert rel_err(prm.grad.cpu(), sdr[k].grad) < tol, (k, rel_err(prm.grad.cpu(), sdr[k].grad))
    else:
        # bf16: no further from the fp32 oracle than an IDEAL bf16 pipeline is (oracle/disc_ref.py emulation)
        xe = logits.clone().requires_grad_(True)
        sde = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        pe = disc_ref.discriminator_forward_bf16(F.softmax(xe, 1), sde)
        (0.37 * F.binary_cross_entropy_with_logits(pe, torch.ones_like(pe))).backward()
        assert rel_err(p.detach().cpu(), p_r.detach()) < 2e-2
        floor = rel_l2(xe.grad, xr.grad)
        assert rel_l2(x.grad.cpu(), xr.grad) < 1.5 * floor + 5e-3, (rel_l2(x.grad.cpu(), xr.grad), floor)
        for k, prm in m.named_parameters():
            floor = rel_l2(sde[k].grad, sdr[k].grad)
            e = rel_l2(prm.grad.cpu(), sdr[k].grad)
            assert e < 1.5 * floor + 5e-3 and e < tol, (k, e, floor)
    if precision == "bf16":
        m2 = _model(tiny, "bf16_simt", seed=5)
        x2 = logits.cuda().requires_grad_(True)
        p2 = m2(F.softmax(x2, dim=1))
        (0.37 * F.binary_cross_entropy_with_logits(p2, torch.ones_like(p2))).backward()
        assert rel_err(p2.detach(), p.detach()) < 5e-3
        assert rel_l2(x2.grad, x.grad) < 2e-2, rel_l2(x2.grad, x.grad)
        for (k, a), b in zip(m.named_parameters(), m2.parameters()):
            assert rel_l2(a.grad, b.grad) < 2e-2, (k, rel_l2(a.grad, b.grad))


def test_frozen_discriminator_and_detached_input(cuda):
    """train.py:192-193 freezes D while the adversarial loss flows into the generator; :242-243 detaches the
    logits while D itself trains."""
    m = _model(True, "fp32")
    logits = _golden_logits().cuda()
    for prm in m.parameters():
        prm.requires_grad = False
    x = logits.clone().requires_grad_(True)
    F.binary_cross_entropy_with_logits(m(F.softmax(x, 1)), torch.ones(2, 1, 1, 1, device="cuda")).backward()
    assert x.grad is not None and all(prm.grad is None for prm in m.parameters())
    dx_frozen = x.grad.clone()
    for prm in m.parameters():
        prm.requires_grad = True
    x2 = logits.clone().requires_grad_(True)
    F.binary_cross_entropy_with_logits(m(F.softmax(x2, 1)), torch.ones(2, 1, 1, 1, device="cuda")).backward()
    # same kernels either way; bit-equal except for the order of fp32 atomics in the CUDA-core dgrad (seen once in ~10 runs)
    assert (dx_frozen - x2.grad).abs().max().item() <= 1e-6 * dx_frozen.abs().max().item()
    g1 = [prm.grad.clone() for prm in m.parameters()]
    # detached input: parameter grads only; a second backward ACCUMULATES into .grad like autograd does
    p = m(F.softmax(logits, 1).detach())
    F.binary_cross_entropy_with_logits(p, torch.ones_like(p)).backward()
    for a, prm in zip(g1, m.parameters()):
        assert torch.allclose(prm.grad, 2 * a, rtol=1e-5, atol=1e-8)
    with torch.no_grad():
        assert not m(F.softmax(logits, 1)).requires_grad


def test_gradient_reversal_and_fused_bce(cuda):
    from models.domain_shift.adversarial.model import DomainDiscriminator
    from rtsds_b200.disc_engine import bce_with_logits_const

    sd = weights.discriminator_state(3, tiny=False)
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(2, 19, 64, 128, generator=g)).cuda()
    outs = []
    for grl in (False, True):
        m = DomainDiscriminator(19, with_grl=grl, lambda_=0.25)
        m.load_state_dict(sd)
        m.rtsds_precision = "fp32"
        m = m.cuda()
        x = logits.clone().requires_grad_(True)
        p = m(F.softmax(x, 1))
        bce_with_logits_const(p, 0.0, 2.0).backward()
        outs.append((p.detach(), x.grad.clone(), m.conv1.weight.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[1][1], -0.25 * outs[0][1], rtol=1e-5, atol=1e-10)
    assert torch.allclose(outs[1][2], -0.25 * outs[0][2], rtol=1e-3, atol=1e-7)      # fp32 atomics: order differs run to run
    # fused BCE == stock criterion
    p = outs[0][0].clone().requires_grad_(True)
    q = outs[0][0].clone().requires_grad_(True)
    a = bce_with_logits_const(p, 1.0, 0.5)
    b = 0.5 * F.binary_cross_entropy_with_logits(q, torch.ones_like(q))
    a.backward(); b.backward()
    assert abs(a.item() - b.item()) < 1e-6 and torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-8)


def test_discriminator_rejects_cpu_tensor():
    from models.domain_shift.adversarial.model import TinyDomainDiscriminator
    from rtsds_b200 import RtsdsError

    with pytest.raises(RtsdsError):
        TinyDomainDiscriminator(19)(torch.zeros(1, 19, 32, 32))


@pytest.mark.parametrize("tiny", [True, False])
def test_two_forwards_then_one_backward(cuda, tiny):
    """train.py:447-458 (adversarial_train_2): D(real) and D(fake) are both evaluated before ONE backward through their
    summed losses -- each forward keeps its own saved activations (plan slots); the parameter gradients equal the sum of
    the two passes done one after the other."""
    g = torch.Generator().manual_seed(11)
    xa = torch.softmax(torch.randn(2, 19, 64, 96, generator=g), 1).cuda()
    xb = torch.softmax(torch.randn(2, 19, 64, 96, generator=g), 1).cuda()
    bce = torch.nn.BCEWithLogitsLoss()
    m = _model(tiny, "fp32")
    oa, ob = m(xa), m(xb)
    (bce(oa, torch.ones_like(oa)) + bce(ob, torch.zeros_like(ob))).backward()
    both = {k: p.grad.clone() for k, p in m.named_parameters()}
    m2 = _model(tiny, "fp32")
    oa2 = m2(xa)
    bce(oa2, torch.ones_like(oa2)).backward()
    ob2 = m2(xb)
    bce(ob2, torch.zeros_like(ob2)).backward()
    assert torch.allclose(oa, oa2, rtol=1e-5, atol=1e-6) and torch.allclose(ob, ob2, rtol=1e-5, atol=1e-6)   # GAP sums are atomics
    for k, p in m2.named_parameters():
        assert rel_l2(both[k], p.grad) < 1e-4, k
    # a fifth pending forward recycles the oldest slot, whose backward then reports the stale activations
    outs = [m(xa) for _ in range(5)]
    with pytest.raises(Exception, match="backward"):
        outs[0].sum().backward()
    outs[4].sum().backward()
