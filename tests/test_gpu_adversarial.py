"""One iteration of the reference's adversarial loop (train.py:177-270) on the CUDA drop-in modules vs the CPU
oracle (oracle/adversarial_ref.py): the four losses, the generator gradient accumulated over its two backward
passes, the discriminator gradient of its two passes (and none from the pass where it is frozen).
SGD(lr, no momentum) makes the parameter update a direct read-out of the gradients."""
import os

import pytest
import torch

from oracle import adversarial_ref, weights

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu
LR = 0.5


def _data(n, hs, ws, ht, wt):
    g = torch.Generator().manual_seed(99)
    src = torch.randn(n, 3, hs, ws, generator=g)
    lbl = torch.randint(0, 20, (n, hs, ws), generator=g)
    tgt = torch.randn(n, 3, ht, wt, generator=g)
    return src, lbl, tgt


def _models(tiny, precision):
    from models.bisenet.build_bisenet import BiSeNet
    from models.domain_shift.adversarial.model import DomainDiscriminator, TinyDomainDiscriminator

    gen = BiSeNet(19, "resnet18")
    gen.load_state_dict(weights.clone_state(weights.bisenet_r18_state(7)))
    gen.rtsds_precision = precision
    dis = (TinyDomainDiscriminator if tiny else DomainDiscriminator)(19)
    dis.load_state_dict(weights.discriminator_state(7, tiny=tiny))
    dis.rtsds_precision = precision
    return gen.cuda().train(), dis.cuda().train()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("tiny,shapes", [(True, (2, 96, 160, 64, 128)), (False, (2, 128, 192, 128, 128))])
def test_adversarial_iteration_vs_oracle_fp32(cuda, tiny, shapes, fused):
    from rtsds_b200.train_steps import adversarial_step

    torch.set_num_threads(os.cpu_count() or 1)
    src, lbl, tgt = _data(*shapes)
    lam, iters = 0.1, 4
    gsd = weights.clone_state(weights.bisenet_r18_state(7))
    dsd = {k: v.clone() for k, v in weights.discriminator_state(7, tiny=tiny).items()}
    ref_losses, g_ref, d_ref, src_feat = adversarial_ref.adversarial_iteration(gsd, dsd, src, lbl, tgt, 19, lam, iters)

    gen, dis = _models(tiny, "fp32")
    before_g = {k: p.detach().clone() for k, p in gen.named_parameters()}
    before_d = {k: p.detach().clone() for k, p in dis.named_parameters()}
    gopt = torch.optim.SGD(gen.parameters(), lr=LR)
    dopt = torch.optim.SGD(dis.parameters(), lr=LR)
    out = adversarial_step(gen, dis, gopt, dopt, src.cuda(), lbl.cuda(), tgt.cuda(), torch.nn.CrossEntropyLoss(ignore_index=19),
                           torch.nn.BCEWithLogitsLoss(), lam, iters, fused=fused)
    for k, v in ref_losses.items():
        assert abs(out[k].item() - v) < 2e-4 * max(1.0, abs(v)), (k, out[k].item(), v)
    assert out["generator_correct"].item() == (src_feat.argmax(1) == lbl).sum().item() or \
        abs(out["generator_correct"].item() - (src_feat.argmax(1) == lbl).sum().item()) <= 2     # fp32 near-ties
    assert all(p.requires_grad for p in dis.parameters())              # unfrozen again (train.py:238-239)
    for k, p in dis.named_parameters():
        g = (before_d[k] - p.detach()) / LR
        e = rel_l2(g.cpu(), d_ref[k])
        assert e < 1e-2, ("D", k, e)      # source (label 1) and target (label 0) passes partly cancel in the sum
    worst = ("", 0.0)
    for k, p in gen.named_parameters():
        if k not in g_ref:
            assert torch.equal(p.detach(), before_g[k]), k              # features.fc: no gradient, untouched
            continue
        if k.startswith("attention_refinement_module") and k.endswith("conv.bias"):
            continue                                                    # analytically zero gradient (BatchNorm removes it)
        g = (before_g[k] - p.detach()) / LR
        e = rel_l2(g.cpu(), g_ref[k])
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 2e-2, worst
    # running BatchNorm buffers saw two train-mode forwards
    bufs = dict(gen.named_buffers())
    for k in ("saptial_path.convblock1.bn.running_mean", "context_path.features.layer4.1.bn2.running_var"):
        assert rel_l2(bufs[k].cpu(), gsd[k]) < 1e-4, k
    assert int(bufs["saptial_path.convblock1.bn.num_batches_tracked"]) == 2


def test_adversarial_iteration_bf16_runs_and_tracks_fp32(cuda):
    from rtsds_b200.train_steps import adversarial_step

    src, lbl, tgt = _data(2, 256, 384, 192, 256)
    res = {}
    for precision in ("fp32", "bf16"):
        for fused in (False, True):
            gen, dis = _models(True, precision)
            gopt = torch.optim.Adam(gen.parameters(), lr=1e-4)
            dopt = torch.optim.Adam(dis.parameters(), lr=1e-4, weight_decay=1e-4)
            out = adversarial_step(gen, dis, gopt, dopt, src.cuda(), lbl.cuda(), tgt.cuda(), torch.nn.CrossEntropyLoss(ignore_index=19),
                                   torch.nn.BCEWithLogitsLoss(), 0.1, 100, fused=fused)
            res[(precision, fused)] = {k: v.item() for k, v in out.items()}
    ref = res[("fp32", False)]
    for key, r in res.items():
        for k in ("loss_gen_source", "loss_adversarial", "loss_disc_source", "loss_disc_target"):
            tol = 1e-4 if key[0] == "fp32" else 3e-2
            assert abs(r[k] - ref[k]) < tol * max(abs(ref[k]), 1e-3), (key, k, r[k], ref[k])
