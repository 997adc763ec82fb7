"""End-to-end parity of the CUDA DeepLabV2-ResNet101 path (drop-in models.deeplabv2.deeplabv2) against golden
vectors produced by the REAL reference (tests/golden/deeplab_72x104.npz) and the CPU oracle
(oracle/deeplab_ref.py) on the same seeded inputs and weights.  fp32 check mode: logits rel <= 1e-4
(BASELINE.json); bf16: judged against an ideal-bf16 emulation of the same pipeline (101 stacked bf16 layers)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import deeplab_ref, weights

from gpu_util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
SUB = 3


def _input(seed, n, h, w):
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(n, 3, h, w, generator=g)
    y = torch.randint(0, 20, (n, h, w), generator=g)
    return x, y


def _model(seed, precision):
    from models.deeplabv2.deeplabv2 import get_deeplab_v2

    m = get_deeplab_v2(19, pretrain=False)
    m.load_state_dict(weights.clone_state(weights.deeplab_state(seed)))
    m.rtsds_precision = precision
    return m.cuda()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_eval_forward_vs_reference_golden(cuda, golden_dir, precision):
    gold = np.load(os.path.join(golden_dir, "deeplab_72x104.npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    x, _ = _input(seed, n, h, w)
    m = _model(seed, precision).eval()
    out = m(x.cuda())
    assert out.shape == (n, 19, h, w) and out.dtype == torch.float32
    ref = torch.from_numpy(gold["eval_result"])
    e = rel_err(out[..., ::SUB, ::SUB].cpu(), ref)
    agree = (out.argmax(1)[..., ::SUB, ::SUB].cpu().numpy() == gold["eval_argmax"]).mean()
    if precision == "fp32":
        assert e < 1e-4 and agree >= 0.999, (e, agree)
    else:
        with torch.no_grad():
            emu = deeplab_ref.deeplab_forward(x, weights.clone_state(weights.deeplab_state(seed)), False, bf16=True)
        floor = rel_err(emu[..., ::SUB, ::SUB], ref)
        floor_agree = (emu.argmax(1)[..., ::SUB, ::SUB].numpy() == gold["eval_argmax"]).mean()
        print("deeplab eval bf16: cuda rel %.4f (ideal bf16 %.4f), argmax agree %.4f (ideal %.4f)" % (e, floor, agree, floor_agree))
        assert e < 1.6 * floor + 5e-3 and agree >= min(0.999, floor_agree - 0.01), (e, floor, agree, floor_agree)
    out2 = m(x.cuda())                                   # CUDA-graph replay, plan reuse
    assert torch.equal(out, out2)


@pytest.mark.parametrize("fused", [False, True])
def test_train_step_fp32_vs_reference_golden_and_oracle(cuda, golden_dir, fused):
    from rtsds_b200.deeplab_engine import deeplab_fused_ce

    torch.set_num_threads(os.cpu_count() or 1)
    gold = np.load(os.path.join(golden_dir, "deeplab_72x104.npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    x, y = _input(seed, n, h, w)
    # CPU oracle with autograd (elementwise gradients), itself pinned to the golden loss
    sd = weights.clone_state(weights.deeplab_state(seed))
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k and ".bn" not in k
              and not k.startswith("bn1") and "downsample.1" not in k}
    ref_out = deeplab_ref.deeplab_forward(x, sd, True)
    ref_loss = F.cross_entropy(ref_out, y, ignore_index=19)
    ref_loss.backward()
    assert abs(ref_loss.item() - float(gold["train_loss_ign19"][0])) < 1e-4
    m = _model(seed, "fp32").train()
    if fused:
        loss, pred, stats = deeplab_fused_ce(m, x.cuda(), y.cuda(), 19)
        assert stats[1].item() == (y != 19).sum().item()
    else:
        outs = m(x.cuda())
        assert isinstance(outs, tuple) and outs[1] is None and outs[2] is None
        e = rel_err(outs[0][..., ::SUB, ::SUB].cpu(), torch.from_numpy(gold["train_result"]))
        assert e < 2e-4, e
        loss = F.cross_entropy(outs[0], y.cuda(), ignore_index=19)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item())), (loss.item(), ref_loss.item())
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert sorted(k for k, g in grads.items() if g is None) == sorted(str(s) for s in gold["grad_none"])     # frozen BN affine
    names = [str(s) for s in gold["grad_names"]]
    worst = ("", 0.0)
    for k, rn in zip(names, gold["grad_norms"]):
        mine = grads[k].double().norm().item()
        assert abs(mine - rn) <= 2e-2 * max(rn, 1e-7), (k, mine, rn)
        e = rel_l2(grads[k].cpu(), leaves[k].grad)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 2e-2, worst
    bufs = dict(m.named_buffers())
    for k in gold.files:
        if k.startswith("buf:"):
            assert rel_err(bufs[k[4:]].cpu(), torch.from_numpy(gold[k])) < 1e-4, k
    assert int(bufs["layer3.22.bn3.num_batches_tracked"]) == 1


def test_train_step_bf16_tracks_ideal_bf16(cuda):
    from rtsds_b200.deeplab_engine import deeplab_fused_ce

    torch.set_num_threads(os.cpu_count() or 1)
    x, y = _input(4, 2, 136, 200)
    res = {}
    for mode in ("fp32", "emu"):
        sd = weights.clone_state(weights.deeplab_state(4))
        leaves = {k: v.requires_grad_(True) for k, v in sd.items() if k.endswith("conv1.weight") or "conv2" in k or "conv3" in k
                  or "downsample.0" in k or k.startswith("layer6")}
        out = deeplab_ref.deeplab_forward(x, sd, True, bf16=(mode == "emu"))
        loss = F.cross_entropy(out, y, ignore_index=19)
        loss.backward()
        res[mode] = (loss.item(), {k: v.grad for k, v in leaves.items()})
    m = _model(4, "bf16").train()
    loss, _, _ = deeplab_fused_ce(m, x.cuda(), y.cuda(), 19)
    loss.backward()
    ref_loss, ref_g = res["fp32"]
    emu_loss, emu_g = res["emu"]
    assert abs(loss.item() - ref_loss) < max(3.0 * abs(emu_loss - ref_loss), 5e-3 * abs(ref_loss)), (loss.item(), emu_loss, ref_loss)
    grads = {k: p.grad for k, p in m.named_parameters()}
    e_gpu = {k: rel_l2(grads[k].cpu(), ref_g[k]) for k in ref_g}
    e_emu = {k: rel_l2(emu_g[k], ref_g[k]) for k in ref_g}
    med = lambda d: sorted(d.values())[len(d) // 2]
    print("deeplab bf16 grads: median rel-L2 vs fp32: cuda %.4f, ideal-bf16 emulation %.4f" % (med(e_gpu), med(e_emu)))
    assert med(e_gpu) < 1.6 * med(e_emu) + 0.02, (med(e_gpu), med(e_emu))
    for k in e_gpu:
        assert e_gpu[k] < 3.0 * max(e_emu[k], med(e_emu)) + 0.05, (k, e_gpu[k], e_emu[k])


def test_full_size_eval_bf16_vs_oracle(cuda):
    """BASELINE config 4 shape: 1x3x512x1024 -> 65x129 features (odd sizes) -> 512x1024 logits."""
    torch.set_num_threads(os.cpu_count() or 1)
    x, _ = _input(6, 1, 512, 1024)
    sd = weights.deeplab_state(6)
    with torch.no_grad():
        ref = deeplab_ref.deeplab_forward(x, weights.clone_state(sd), False)
        emu = deeplab_ref.deeplab_forward(x, weights.clone_state(sd), False, bf16=True)
    floor = rel_err(emu, ref)
    floor_agree = (emu.argmax(1) == ref.argmax(1)).float().mean().item()
    m = _model(6, "bf16").eval()
    out = m(x.cuda()).cpu()
    e = rel_err(out, ref)
    agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
    print("deeplab 512x1024 bf16: rel %.4f (ideal %.4f) argmax agree %.4f (ideal %.4f)" % (e, floor, agree, floor_agree))
    assert e < 1.6 * floor + 5e-3 and agree >= min(0.999, floor_agree - 0.005), (e, floor, agree, floor_agree)


def test_deeplab_rejects_cpu_tensor():
    from models.deeplabv2.deeplabv2 import get_deeplab_v2
    from rtsds_b200 import RtsdsError

    with pytest.raises(RtsdsError):
        get_deeplab_v2(19, pretrain=False).eval()(torch.zeros(1, 3, 64, 64))
