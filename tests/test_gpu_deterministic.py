"""Deterministic mode (include/rtsds_b200.h: rtsds_set_deterministic).

The reference's results come from ATen's CPU path, which sums in a fixed order: its numbers are reproducible run to run.
The CUDA path's cross-CTA reductions — train-mode BatchNorm sum / sum of squares in every conv and stem epilogue, the
weight-gradient partials, the BatchNorm-backward sums, the split global average pool — normally go through fp32 atomics,
whose arrival order varies.  In deterministic mode they are accumulated exactly (64.64 fixed point, integer atomics) and
rounded once, so

  * two runs are BIT-identical (asserted with torch.equal), and
  * the value is the correctly rounded sum of the per-warp partials, hence within fp32 round-off of the atomic path.
"""
import pytest
import torch
import torch.nn.functional as F

from rtsds_b200 import ops
from rtsds_b200._lib import check, lib
from rtsds_b200.ops import BF16, F32

from gpu_util import nhwc, rel_err, run_conv
from parity_log import record

pytestmark = pytest.mark.gpu


@pytest.fixture
def det():
    ops.set_deterministic(True)
    assert ops.is_deterministic()
    yield
    ops.set_deterministic(False)
    assert not ops.is_deterministic()


def _mk(n, cin, h, w, cout, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    return x, wt, g


# (kind, dtype, n, cin, h, w, cout, k, stride): one-wave grid, persistent + halo form (>= 2 waves of tiles), split-K, SIMT fp32
STATS_CASES = [("tc", BF16, 2, 64, 45, 80, 128, 3, 2), ("tc", BF16, 4, 64, 180, 320, 64, 3, 1), ("tc", BF16, 2, 512, 12, 20, 512, 3, 1),
               ("tc", BF16, 2, 128, 90, 160, 256, 1, 1), ("simt", F32, 2, 64, 45, 80, 128, 3, 2), ("simt", F32, 1, 19, 33, 47, 24, 3, 1)]


@pytest.mark.parametrize("case", STATS_CASES, ids=[f"{c[0]}_{c[3]}to{c[6]}_k{c[7]}s{c[8]}_{c[4]}x{c[5]}" for c in STATS_CASES])
def test_bn_statistics_are_bit_reproducible_and_exact(cuda, det, case):
    kind, dtype, n, cin, h, w, cout, k, stride = case
    x, wt, _ = _mk(n, cin, h, w, cout, k, seed=31)
    pad = k // 2
    out_ld = ops.cout_pad(cout) if kind == "tc" else cout
    y1, s1, _ = run_conv(kind, x, wt, stride=stride, pad=pad, dtype=dtype, out_dtype=F32, out_ld=out_ld, want_stats=True)
    y2, s2, _ = run_conv(kind, x, wt, stride=stride, pad=pad, dtype=dtype, out_dtype=F32, out_ld=out_ld, want_stats=True)
    assert torch.equal(y1, y2)
    assert torch.equal(s1, s2), (s1 - s2).abs().max().item()
    # the raw fp32 outputs are what the epilogue summed: compare with their fp64 sum
    want1, want2 = y1.double().sum((0, 2, 3)), (y1.double() ** 2).sum((0, 2, 3))
    e1 = (s1[:cout].double() - want1).abs().max().item() / want2.sqrt().max().item()
    e2 = rel_err(s1[cout:], want2)
    assert e1 < 1e-5 and e2 < 1e-5, (e1, e2)
    ops.set_deterministic(False)
    _, s3, _ = run_conv(kind, x, wt, stride=stride, pad=pad, dtype=dtype, out_dtype=F32, out_ld=out_ld, want_stats=True)
    ops.set_deterministic(True)
    assert rel_err(s3[cout:], s1[cout:]) < 1e-5          # same quantity as the fp32-atomic path, to round-off
    record(f"deterministic:bn_stats:{kind}:{cin}->{cout}:k{k}s{stride}:{n}x{h}x{w}", bit_identical=True, sum_err=e1, sumsq_rel=e2,
           vs_atomics_rel=rel_err(s3[cout:], s1[cout:]))


def _wgrad(kind, x, wt, dy, stride, pad, dtype):
    tc = kind == "tc"
    tdt = ops.torch_dtype(dtype)
    n, cin, h, w = x.shape
    cout, _, kh, kw = wt.shape
    dy_ld = max(ops.dgrad_ck(cout, tc), (cout + 7) // 8 * 8)
    d = ops.make_conv_desc(n, h, w, cin, cin, cout, dy_ld, kh, stride, pad, 1, in_dtype=dtype, out_dtype=dtype, kw=kw)
    dyg = torch.zeros(n, d.oh, d.ow, dy_ld, dtype=tdt, device="cuda")
    dyg[..., :cout] = nhwc(dy, tdt)
    xg = nhwc(x, tdt)
    dwp = torch.zeros(cout, kh * kw, cin, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(d, xg, dyg, dwp, tc)
    torch.cuda.synchronize()
    return dwp


WGRAD_CASES = [("tc", BF16, 4, 64, 90, 160, 64, 3, 1), ("tc", BF16, 2, 128, 45, 80, 256, 3, 2), ("tc", BF16, 2, 512, 12, 20, 512, 3, 1),
               ("tc", BF16, 2, 256, 23, 40, 19, 1, 1), ("simt", F32, 2, 24, 33, 47, 40, 3, 1)]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[f"{c[0]}_{c[3]}to{c[6]}_k{c[7]}s{c[8]}_{c[4]}x{c[5]}" for c in WGRAD_CASES])
def test_weight_gradient_is_bit_reproducible(cuda, det, case):
    kind, dtype, n, cin, h, w, cout, k, stride = case
    x, wt, g = _mk(n, cin, h, w, cout, k, seed=32)
    pad = k // 2
    oh, ow = ops.conv_out_size(h, k, stride, pad, 1), ops.conv_out_size(w, k, stride, pad, 1)
    dy = torch.randn(n, cout, oh, ow, generator=g)
    a = _wgrad(kind, x, wt, dy, stride, pad, dtype)
    b = _wgrad(kind, x, wt, dy, stride, pad, dtype)
    assert torch.equal(a, b), (a - b).abs().max().item()
    ops.set_deterministic(False)
    c = _wgrad(kind, x, wt, dy, stride, pad, dtype)
    ops.set_deterministic(True)
    e = rel_err(c.cpu(), a.cpu())
    assert e < 1e-5, e
    # and it is the gradient: autograd on the rounded operands
    rnd = (lambda t: t.bfloat16().float()) if dtype == BF16 else (lambda t: t)
    wr = rnd(wt).clone().requires_grad_(True)
    F.conv2d(rnd(x), wr, None, stride, pad).backward(rnd(dy))
    got = a.view(cout, k, k, cin).permute(0, 3, 1, 2).cpu()
    assert rel_err(got, wr.grad) < 2e-4
    record(f"deterministic:wgrad:{kind}:{cin}->{cout}:k{k}s{stride}:{n}x{h}x{w}", bit_identical=True, vs_atomics_rel=e)


def test_bn_backward_sums_are_bit_reproducible(cuda, det):
    n_pix, c = 40003, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    raw = (torch.randn(n_pix, c, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(n_pix, c, device="cuda", generator=g).bfloat16()
    mean = raw.float().mean(0)
    invstd = (raw.float().var(0, unbiased=False) + 1e-5).rsqrt()
    fsc = torch.rand(c, device="cuda", generator=g) + 0.5
    fsh = -mean * fsc

    def run():
        sums = torch.empty(2 * c, device="cuda")
        check(lib().rtsds_bn_bwd_reduce_rawmask(dy.data_ptr(), c, raw.data_ptr(), c, mean.data_ptr(), invstd.data_ptr(), fsc.data_ptr(),
                                                fsh.data_ptr(), n_pix, c, BF16, sums.data_ptr(), None), "reduce")
        torch.cuda.synchronize()
        return sums

    a, b = run(), run()
    assert torch.equal(a, b)
    mask = torch.addcmul(fsh, raw.float(), fsc) > 0
    gg = dy.double() * mask
    xh = (raw.double() - mean.double()) * invstd.double()
    scale = float(n_pix) ** 0.5
    assert (a[:c].double() - gg.sum(0)).abs().max().item() < 1e-4 * scale
    assert (a[c:].double() - (gg * xh).sum(0)).abs().max().item() < 1e-4 * scale


def test_stem_pair_statistics_and_gradient_are_bit_reproducible(cuda, det):
    n, h, w = 2, 72, 104
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, 3, h, w, generator=g).cuda()
    w7 = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).cuda()
    w3 = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).cuda()
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    wpk = ops.stem_pack_weights(w7, w3)
    d7 = torch.randn(n, oh, ow, 64, generator=g).bfloat16().cuda()
    d3 = torch.randn(n, oh, ow, 64, generator=g).bfloat16().cuda()

    def run():
        ycp = torch.empty(n, oh, ow, 64, dtype=torch.bfloat16, device="cuda")
        ysp = torch.empty_like(ycp)
        st7 = torch.zeros(128, device="cuda"); st3 = torch.zeros(128, device="cuda")
        ops.stem_pair_tc_fwd(x, wpk, ycp, ysp, None, None, False, st7, st3)
        ws = torch.zeros(128 * 192, device="cuda")
        g7 = torch.zeros(64, 3, 7, 7, device="cuda"); g3 = torch.zeros(64, 3, 3, 3, device="cuda")
        ops.stem_pair_tc_wgrad(x, d7, d3, ws, g7, g3)
        torch.cuda.synchronize()
        assert (ws == 0).all()
        return st7, st3, g7, g3, ycp

    a, b = run(), run()
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    raw7 = F.conv2d(x.cpu().bfloat16().float(), w7.cpu().bfloat16().float(), None, 2, 3)
    assert rel_err(a[0][:64].cpu(), raw7.sum((0, 2, 3))) < 2e-3 and rel_err(a[0][64:].cpu(), (raw7 * raw7).sum((0, 2, 3))) < 2e-3
    ops.set_deterministic(False)
    c = run()
    ops.set_deterministic(True)
    assert rel_err(c[2].cpu(), a[2].cpu()) < 1e-5 and rel_err(c[3].cpu(), a[3].cpu()) < 1e-5


def test_global_average_pool_is_bit_reproducible(cuda, det):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 45 * 80, 256, device="cuda", generator=g).bfloat16()

    def run():
        out = torch.empty(2, 256, device="cuda")
        check(lib().rtsds_global_avgpool(x.data_ptr(), 2, 45 * 80, 256, 256, BF16, out.data_ptr(), None), "gap")
        torch.cuda.synchronize()
        return out

    a, b = run(), run()
    assert torch.equal(a, b)
    assert rel_err(a.cpu(), x.float().mean(1).cpu()) < 1e-5


@pytest.mark.parametrize("fused", [False, True], ids=["stock_criteria", "fused_ce"])
def test_bisenet_train_step_is_bit_reproducible(cuda, det, fused):
    """Whole train-mode BiSeNet-R18 (bf16, the benchmarked mode) at a GTA5-shaped odd size: two forward passes from the same
    state give bit-identical logits (or argmax maps) and BatchNorm running statistics, and two backward passes give
    bit-identical gradients for every parameter — with the stock nn.CrossEntropyLoss criteria (the reference's call sequence)
    and with the fused resize + CE + argmax path."""
    from models.bisenet.build_bisenet import BiSeNet
    from oracle import weights

    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, 360, 640, generator=g).cuda()
    y = torch.randint(0, 20, (2, 360, 640), generator=g).cuda()

    def run():
        m = BiSeNet(19, "resnet18")
        m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(5)))
        m = m.cuda().train()
        if fused:                 # train.py:77-92 + :102-106 in one call: x8 resize + 3 x CE + argmax + gradient kernels
            from rtsds_b200.bisenet_autograd import bisenet_fused_ce

            loss, pred, _ = bisenet_fused_ce(m, x, y, 19)
            outs = [pred.float()]
        else:
            outs = m(x)
            loss = sum(F.cross_entropy(t, y, ignore_index=19) for t in outs)
        loss.backward()
        torch.cuda.synchronize()
        bufs = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        return [o.detach().clone() for o in outs], bufs, grads, loss.item()

    o1, b1, g1, l1 = run()
    o2, b2, g2, l2 = run()
    for u, v in zip(o1, o2):
        assert torch.equal(u, v), (u - v).abs().max().item()
    for k in b1:
        assert torch.equal(b1[k], b2[k]), k
    assert abs(l1 - l2) < 1e-5 * abs(l1)         # (torch's own nll_loss2d forward sums with atomics)
    diffs = {}
    for k in g1:
        diffs[k] = (g1[k].double() - g2[k].double()).norm().item() / max(g1[k].double().norm().item(), 1e-30)
    differing = sorted((k for k in g1 if not torch.equal(g1[k], g2[k])), key=lambda k: -diffs[k])
    record(f"deterministic:bisenet_train_2x3x360x640:{'fused_ce' if fused else 'stock_criteria'}", forward_bit_identical=True, loss=l1, grad_tensors=len(g1),
           grad_tensors_bit_identical=len(g1) - len(differing), worst_grad_rel_l2_between_runs=max(diffs.values()),
           differing=", ".join(f"{k}={diffs[k]:.2e}" for k in differing[:8]))
    # every reduction of the step (stock criterion path) is covered by the mode: ALL 90 gradient tensors are bit-identical
    assert not differing, differing[:8]


@pytest.mark.parametrize("tiny", [True, False], ids=["tiny_d", "full_d"])
def test_adversarial_iteration_is_bit_reproducible(cuda, det, tiny):
    """One iteration of the reference's adversarial loop (train.py:177-270; bf16, fused fast paths): generator + discriminator,
    two generator backward passes, two discriminator passes.  With SGD(lr, no momentum) the updated parameters are a direct
    read-out of the gradients: two runs from the same state give bit-identical generator AND discriminator parameters."""
    from models.bisenet.build_bisenet import BiSeNet
    from models.domain_shift.adversarial.model import DomainDiscriminator, TinyDomainDiscriminator
    from oracle import weights
    from rtsds_b200.train_steps import adversarial_step

    g = torch.Generator().manual_seed(99)
    src = torch.randn(2, 3, 192, 320, generator=g).cuda()
    lbl = torch.randint(0, 20, (2, 192, 320), generator=g).cuda()
    tgt = torch.randn(2, 3, 128, 256, generator=g).cuda()

    def run():
        gen = BiSeNet(19, "resnet18")
        gen.load_state_dict(weights.clone_state(weights.bisenet_r18_state(7)))
        dis = (TinyDomainDiscriminator if tiny else DomainDiscriminator)(19)
        dis.load_state_dict(weights.discriminator_state(7, tiny=tiny))
        gen, dis = gen.cuda().train(), dis.cuda().train()
        gopt = torch.optim.SGD(gen.parameters(), lr=0.5)
        dopt = torch.optim.SGD(dis.parameters(), lr=0.5)
        out = adversarial_step(gen, dis, gopt, dopt, src, lbl, tgt, torch.nn.CrossEntropyLoss(ignore_index=19),
                               torch.nn.BCEWithLogitsLoss(), 0.1, 4, fused=True)
        torch.cuda.synchronize()
        return ({k: p.detach().clone() for k, p in gen.named_parameters()}, {k: p.detach().clone() for k, p in dis.named_parameters()},
                {k: float(v) for k, v in out.items()})

    g1, d1, o1 = run()
    g2, d2, o2 = run()
    bad = [k for k in g1 if not torch.equal(g1[k], g2[k])] + ["D." + k for k in d1 if not torch.equal(d1[k], d2[k])]
    record(f"deterministic:adversarial_iteration:{'tiny' if tiny else 'full'}_d", parameters=len(g1) + len(d1),
           parameters_differing=len(bad), losses=", ".join(f"{k}={v:.6g}" for k, v in o1.items()))
    assert not bad, bad[:8]


def test_deeplab_train_step_is_bit_reproducible(cuda, det):
    """DeepLabV2-ResNet101 (bf16, fused resize + CE), 2x3x136x200: two runs from the same state give bit-identical loss-side
    argmax maps, BatchNorm running statistics and gradients for every trainable parameter."""
    from models.deeplabv2.deeplabv2 import get_deeplab_v2
    from oracle import weights
    from rtsds_b200.deeplab_engine import deeplab_fused_ce

    g = torch.Generator().manual_seed(2004)
    x = torch.randn(2, 3, 136, 200, generator=g).cuda()
    y = torch.randint(0, 20, (2, 136, 200), generator=g).cuda()

    def run():
        m = get_deeplab_v2(19, pretrain=False)
        m.load_state_dict(weights.clone_state(weights.deeplab_state(4)))
        m.rtsds_precision = "bf16"
        m = m.cuda().train()
        loss, pred, _ = deeplab_fused_ce(m, x, y, 19)
        loss.backward()
        torch.cuda.synchronize()
        bufs = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        return pred.clone(), bufs, grads

    p1, b1, g1 = run()
    p2, b2, g2 = run()
    assert torch.equal(p1, p2)
    bad = [k for k in b1 if not torch.equal(b1[k], b2[k])] + [k for k in g1 if not torch.equal(g1[k], g2[k])]
    record("deterministic:deeplab_train_2x3x136x200", grad_tensors=len(g1), tensors_differing=len(bad))
    assert len(g1) > 100 and not bad, bad[:8]
