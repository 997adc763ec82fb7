"""the fused multi-tensor optimizer kernel (csrc/optim.cu, rtsds_b200/optim.py) against torch.optim — the
optimizers the reference builds at main.py:110-120 — over 10 steps on the same gradients, including poly-LR on
param_groups[0] only (utils.py:33-48), L2 weight decay (the discriminator's Adam, config.yaml), parameters without a
gradient, and the refresh of a training plan's packed bf16 conv operands inside the same launch."""
import pytest
import torch

from rtsds_b200 import ops, optim

pytestmark = pytest.mark.gpu

SHAPES = [(64, 3, 7, 7), (128, 64, 3, 3), (19, 256, 1, 1), (64,), (19,), (1000, 512), (256, 256, 1, 1), (1,), (2049,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.1).cuda()) for s in SHAPES]


def _grads(step, ps):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(*p.shape, generator=g).cuda() * (0.01 + 0.1 * (i % 3)) for i, p in enumerate(ps)]


@pytest.mark.parametrize("kind,kw", [("adam", dict(lr=1e-3)), ("adam", dict(lr=1e-4, weight_decay=1e-4)),
                                     ("sgd", dict(lr=1e-2, momentum=0.9)), ("sgd", dict(lr=1e-2, momentum=0.9, weight_decay=5e-4)),
                                     ("sgd", dict(lr=1e-2))])
def test_fused_step_matches_torch_optim(cuda, kind, kw):
    a, b = _params(0), _params(0)
    # two groups: the reference's poly scheduler rewrites group 0 only
    mk = (lambda ps, cls: cls([{"params": ps[:5]}, {"params": ps[5:], "lr": kw["lr"] * 10}], **kw))
    ref = mk(a, torch.optim.Adam if kind == "adam" else torch.optim.SGD)
    mine = mk(b, optim.FusedAdam if kind == "adam" else optim.FusedSGD)
    for step in range(10):
        lr0 = kw["lr"] * (1 - step / 20) ** 0.9                  # utils.poly_lr_scheduler
        ref.param_groups[0]["lr"] = mine.param_groups[0]["lr"] = lr0
        ga = _grads(step, a)
        for i, (pa, pb, g) in enumerate(zip(a, b, ga)):
            skip = i == 3 and step < 4                           # a parameter that gets its first gradient later
            pa.grad = None if skip else g.clone()
            pb.grad = None if skip else g.clone()
        ref.step()
        mine.step()
    torch.cuda.synchronize()
    for pa, pb in zip(a, b):
        err = (pa - pb).abs().max().item() / max(pa.abs().max().item(), 1e-12)
        assert err < 2e-6, (tuple(pa.shape), err)
    # torch's state keys, so state_dict() round-trips between the two
    sd = mine.state_dict()
    k0 = sd["state"][0]
    assert ("exp_avg" in k0 and "exp_avg_sq" in k0) if kind == "adam" else ("momentum_buffer" in k0) == ("momentum" in kw)
    assert int(k0["step"]) == 10 and int(sd["state"][3]["step"]) == 6


def test_fuse_converts_a_stock_optimizer_in_place(cuda):
    a, b = _params(1), _params(1)
    ref = torch.optim.Adam(a, lr=1e-3, weight_decay=1e-4)
    stock = torch.optim.Adam(b, lr=1e-3, weight_decay=1e-4)
    for step in range(3):                                        # a few stock steps first: the moments carry over
        for pa, pb, g in zip(a, b, _grads(step, a)):
            pa.grad, pb.grad = g.clone(), g.clone()
        ref.step(); stock.step()
    same = optim.fuse_(stock)
    assert same is stock and hasattr(stock, "_rtsds_fused")
    n0 = ops.launch_count()
    for step in range(3, 8):
        stock.param_groups[0]["lr"] = ref.param_groups[0]["lr"] = 1e-3 * (1 - step / 10)
        for pa, pb, g in zip(a, b, _grads(step, a)):
            pa.grad, pb.grad = g.clone(), g.clone()
        ref.step(); stock.step()
    assert ops.launch_count() - n0 == 5                          # ONE kernel per step
    for pa, pb in zip(a, b):
        assert (pa - pb).abs().max().item() <= 2e-6 * max(pa.abs().max().item(), 1e-12)


def test_fused_step_refreshes_the_training_plans_packed_operands(cuda):
    """A BiSeNet training step with FusedAdam: after step() the plan's packed bf16 conv operands equal a fresh pack of the
    updated fp32 masters, the plan does not re-pack at its next forward, and the step matches torch.optim.Adam."""
    from models.bisenet.build_bisenet import BiSeNet
    from oracle import weights
    from rtsds_b200.bisenet_autograd import bisenet_fused_ce

    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 128, 192, generator=g).cuda()
    y = torch.randint(0, 20, (2, 128, 192), generator=g).cuda()
    models, opts = [], []
    for fused in (False, True):
        m = BiSeNet(19, "resnet18")
        m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(7)))
        m = m.cuda().train()
        opt = (optim.FusedAdam if fused else torch.optim.Adam)(m.parameters(), lr=1e-3)
        models.append(m); opts.append(opt)
    for step in range(3):
        for m, opt in zip(models, opts):
            opt.zero_grad(set_to_none=True)
            loss, _, _ = bisenet_fused_ce(m, x, y, 19)
            loss.backward()
            opt.step()
    torch.cuda.synchronize()
    m = models[1]
    plan = next(iter(m._rtsds_train_plans.values()))
    assert plan in opts[1]._plans                                # attached automatically
    assert plan._param_version == plan._params_version()         # no re-pack at the next forward
    for conv, out, kind in plan.pack_jobs:
        fresh = (ops.pack_conv_weight(conv.weight, plan.dt) if kind == 0
                 else ops.pack_conv_weight_dgrad(conv.weight, plan.dt, plan.use_tc))
        assert torch.equal(out.view(-1), fresh.view(-1)), (tuple(conv.weight.shape), kind)
    # bf16 training is chaotic over steps (see tests/test_gpu_config1.py); one optimizer step from identical state must agree:
    # rerun both models from the SAME weights for one step and compare the updates
    ma, mb = models
    mb.load_state_dict(ma.state_dict())
    oa = torch.optim.Adam(ma.parameters(), lr=1e-3)
    ob = optim.FusedAdam(mb.parameters(), lr=1e-3)
    for m_, o_ in ((ma, oa), (mb, ob)):
        m_.rtsds_precision = "fp32"
        o_.zero_grad(set_to_none=True)
        loss, _, _ = bisenet_fused_ce(m_, x, y, 19)
        loss.backward()
    ga = {k: p.grad.clone() for k, p in ma.named_parameters() if p.grad is not None}
    for k, p in mb.named_parameters():                           # identical gradients into both optimizers
        if p.grad is not None:
            p.grad.copy_(ga[k])
    oa.step(); ob.step()
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa - pb).abs().max().item() <= 2e-6 * max(pa.abs().max().item(), 1e-12), k
