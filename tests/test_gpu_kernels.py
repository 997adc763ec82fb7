"""Parity of the non-GEMM kernels against the CPU oracle: confusion-matrix
histogram (bit-exact), stems, max-pool, BatchNorm helpers, ARM gate, gated
bilinear resize, FFM head, resize-to-NCHW, fused resize+CE+argmax."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import metrics_ref
from rtsds_b200 import ops
from rtsds_b200.ops import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F32

from gpu_util import nchw, nhwc, rel_err

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- histogram (bit-exact)
def _hist_gpu(a, b, n):
    la = torch.from_numpy(np.ascontiguousarray(a).astype(np.int64)).cuda().reshape(-1)
    lb = torch.from_numpy(np.ascontiguousarray(b).astype(np.int64)).cuda().reshape(-1)
    hist = torch.zeros(n * n, dtype=torch.int64, device="cuda")
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.confusion_hist(la, lb, n, hist, bad)
    return hist.cpu().numpy().reshape(n, n), int(bad.item())


def test_confusion_hist_golden(cuda, golden_dir):
    gold = np.load(os.path.join(golden_dir, "fast_hist.npz"))
    for c in sorted({k.rsplit("_", 1)[0] for k in gold.files}):
        h, bad = _hist_gpu(gold[c + "_label"], gold[c + "_pred"], 19)
        assert bad == 0
        assert (h == gold[c + "_hist"]).all(), c


@pytest.mark.parametrize("shape,n", [((2, 512, 1024), 19), ((1, 720, 1280), 19), ((3, 33, 77), 19), ((1, 1, 1), 19),
                                      ((4, 100, 101), 7), ((1, 64, 64), 64)])
def test_confusion_hist_random(cuda, shape, n):
    rng = np.random.default_rng(sum(shape))
    a = rng.integers(-1, n + 2, size=shape)
    a[rng.random(shape) < 0.05] = 255
    b = rng.integers(0, n, size=shape)
    h, bad = _hist_gpu(a, b, n)
    assert bad == 0 and (h == metrics_ref.fast_hist(a, b, n)).all()


def test_confusion_hist_piecewise_constant_and_accumulation(cuda):
    """Segmentation-like maps (long runs of one bin) and accumulation over batches (validation.py:55)."""
    rng = np.random.default_rng(5)
    a = np.repeat(rng.integers(0, 19, size=(4, 64, 16)), 64, axis=2)
    b = np.repeat(rng.integers(0, 19, size=(4, 64, 8)), 128, axis=2)
    la, lb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    hist = torch.zeros(361, dtype=torch.int64, device="cuda")
    for i in range(4):
        ops.confusion_hist(la[i].reshape(-1), lb[i].reshape(-1), 19, hist)
    assert (hist.cpu().numpy().reshape(19, 19) == metrics_ref.fast_hist(a, b, 19)).all()
    # unaligned views take the scalar path
    h2 = torch.zeros(361, dtype=torch.int64, device="cuda")
    ops.confusion_hist(la.reshape(-1)[1:], lb.reshape(-1)[1:], 19, h2)
    assert (h2.cpu().numpy().reshape(19, 19) == metrics_ref.fast_hist(a.reshape(-1)[1:], b.reshape(-1)[1:], 19)).all()


def test_confusion_hist_out_of_range_prediction_is_reported(cuda):
    a = np.array([0, 1, 18, 18]); b = np.array([0, 400, 18, -400])
    h, bad = _hist_gpu(a, b, 19)
    assert bad == 2 and h.sum() == 2


@pytest.mark.parametrize("n,c,h,w", [(2, 19, 64, 128), (1, 19, 33, 51), (1, 7, 8, 8)])
def test_argmax_hist(cuda, n, c, h, w):
    g = torch.Generator().manual_seed(n * h)
    logits = torch.randn(n, c, h, w, generator=g)
    logits[:, 3] = logits[:, 5]            # exact ties: first index must win (torch.argmax)
    label = torch.randint(-1, c + 1, (n, h, w), generator=g)
    pred = torch.empty(n, h, w, dtype=torch.int64, device="cuda")
    hist = torch.zeros(c * c, dtype=torch.int64, device="cuda")
    ops.argmax_hist(logits.cuda(), label.cuda(), hist, pred)
    ref_pred = logits.argmax(1)
    assert (pred.cpu() == ref_pred).all()
    assert (hist.cpu().numpy().reshape(c, c) == metrics_ref.fast_hist(label.numpy(), ref_pred.numpy(), c)).all()


# ----------------------------------------------------------------------------- stems / pool
@pytest.mark.parametrize("k,pad,cin,h,w", [(3, 1, 3, 64, 96), (7, 3, 3, 72, 104), (3, 1, 3, 45, 81), (7, 3, 3, 33, 47),
                                           (4, 1, 19, 64, 96)])
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_stem_conv(cuda, k, pad, cin, h, w, dtype):
    g = torch.Generator().manual_seed(k * h)
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(64, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    scale = torch.rand(64, generator=g) + 0.5
    shift = torch.randn(64, generator=g)
    oh, ow = ops.conv_out_size(h, k, 2, pad), ops.conv_out_size(w, k, 2, pad)
    y = torch.empty(2, oh, ow, 64, dtype=ops.torch_dtype(dtype), device="cuda")
    stats = torch.zeros(128, dtype=torch.float32, device="cuda")
    ops.stem_conv(x.cuda(), wt.cuda(), y, k, 2, pad, scale.cuda(), shift.cuda(), ACT_RELU, stats=stats)
    raw = F.conv2d(x, wt, None, 2, pad)
    ref = F.relu(raw * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert rel_err(nchw(y), ref) < (1e-5 if dtype == F32 else 1e-2)
    assert rel_err(stats[:64].cpu(), raw.sum((0, 2, 3))) < 1e-4
    assert rel_err(stats[64:].cpu(), (raw * raw).sum((0, 2, 3))) < 1e-4


def test_stem_conv_softmax_input_bias_leaky(cuda):
    """Discriminator conv1: softmax over the 19 logits fused into the loader (train.py:225)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 19, 40, 56, generator=g) * 4
    wt = torch.randn(64, 19, 4, 4, generator=g) * 0.1
    bias = torch.randn(64, generator=g)
    y = torch.empty(2, 20, 28, 64, dtype=torch.float32, device="cuda")
    ops.stem_conv(x.cuda(), wt.cuda(), y, 4, 2, 1, None, bias.cuda(), ACT_LRELU, 0.2, softmax_in=True)
    ref = F.leaky_relu(F.conv2d(F.softmax(x, 1), wt, bias, 2, 1), 0.2)
    assert rel_err(nchw(y), ref) < 1e-5


@pytest.mark.parametrize("h,w,ceil", [(32, 48, False), (45, 81, False), (45, 81, True), (65, 129, True), (64, 128, True)])
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_maxpool(cuda, h, w, ceil, dtype):
    x = torch.randn(2, 64, h, w, generator=torch.Generator().manual_seed(h))
    tdt = ops.torch_dtype(dtype)
    ref = F.max_pool2d(x.to(tdt).float(), 3, 2, 1, ceil_mode=ceil)
    y = torch.empty(2, ref.shape[2], ref.shape[3], 64, dtype=tdt, device="cuda")
    assert ops.maxpool_out_size(h, ceil) == ref.shape[2] and ops.maxpool_out_size(w, ceil) == ref.shape[3]
    ops.maxpool3x3s2(nhwc(x, tdt), y, ceil)
    assert torch.equal(nchw(y), ref)


@pytest.mark.parametrize("h,w,c,ceil", [(32, 48, 64, False), (45, 81, 64, False), (45, 81, 64, True), (65, 129, 64, True), (33, 35, 24, False),
                                        (18, 70, 128, True)])
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_maxpool_backward(cuda, h, w, c, ceil, dtype):
    """Gradient goes to the FIRST maximum of each window (ties: values drawn from 5 levels), ceil_mode windows that
    hang over the border, channel counts that are not a multiple of 64."""
    g = torch.Generator().manual_seed(h * 7 + w)
    tdt = ops.torch_dtype(dtype)
    x = torch.randint(0, 5, (2, c, h, w), generator=g).float().requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1, ceil_mode=ceil)
    dy = torch.randint(-8, 9, y.shape, generator=g).float()
    y.backward(dy)
    dx = torch.full((2, h, w, c), float("nan"), dtype=tdt, device="cuda")
    xg, dyg = nhwc(x.detach(), tdt), nhwc(dy, tdt)           # keep the device copies alive across the launch
    ops.check(ops.lib().rtsds_maxpool3x3s2_bwd(xg.data_ptr(), dyg.data_ptr(), 2, h, w, c, dtype, int(ceil), dx.data_ptr(), ops._s()),
              "maxpool_bwd")
    assert torch.equal(nchw(dx), x.grad)


@pytest.mark.parametrize("h,w,c,ceil", [(32, 48, 64, False), (45, 81, 64, False), (45, 81, 64, True), (65, 129, 64, True), (33, 35, 24, False),
                                        (18, 70, 128, True), (1, 1, 8, False), (2, 3, 8, True)])
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_maxpool_indexed_forward_backward(cuda, h, w, c, ceil, dtype):
    """Training pair: the forward records the first-maximum positions, the backward routes dy with them alone."""
    g = torch.Generator().manual_seed(h * 5 + w)
    tdt = ops.torch_dtype(dtype)
    x = torch.randint(0, 5, (2, c, h, w), generator=g).float().requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1, ceil_mode=ceil)
    dy = torch.randint(-8, 9, y.shape, generator=g).float()
    y.backward(dy)
    oh, ow = y.shape[2], y.shape[3]
    yg = torch.empty(2, oh, ow, c, dtype=tdt, device="cuda")
    idx = torch.full((2, oh, ow, c // 8), -1, dtype=torch.int32, device="cuda")
    ops.maxpool3x3s2(nhwc(x.detach(), tdt), yg, ceil, idx)
    assert torch.equal(nchw(yg), y.detach())
    dx = torch.full((2, h, w, c), float("nan"), dtype=tdt, device="cuda")
    ops.maxpool3x3s2_bwd_idx(idx, nhwc(dy, tdt), dx, ceil)
    assert torch.equal(nchw(dx), x.grad)


# ----------------------------------------------------------------------------- batch norm helpers
def test_bn_fold_and_finalize(cuda):
    g = torch.Generator().manual_seed(2)
    c = 96
    bn = torch.nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5); bn.bias.copy_(torch.randn(c, generator=g))
        bn.running_mean.copy_(torch.randn(c, generator=g)); bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    scale = torch.empty(c, device="cuda"); shift = torch.empty(c, device="cuda")
    ops.bn_fold(bn, scale, shift)
    rs = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    assert rel_err(scale.cpu(), rs.detach().cpu()) < 1e-6
    assert rel_err(shift.cpu(), (bn.bias - bn.running_mean * rs).detach().cpu()) < 1e-6
    # finalize == F.batch_norm(training=True) incl. running-stat update
    x = torch.randn(4, c, 9, 11, generator=g) * 2 + 1
    stats = torch.cat([x.sum((0, 2, 3)), (x * x).sum((0, 2, 3))]).cuda()
    rm, rv = bn.running_mean.clone().cpu(), bn.running_var.clone().cpu()
    ref = F.batch_norm(x, rm, rv, bn.weight.detach().cpu(), bn.bias.detach().cpu(), True, 0.1, bn.eps)
    sm = torch.empty(c, device="cuda"); si = torch.empty(c, device="cuda")
    ops.bn_finalize(stats, 4 * 9 * 11, bn, scale, shift, sm, si)
    y = torch.empty(4, 9, 11, c, device="cuda")
    ops.scale_shift_act(nhwc(x, torch.float32), y, 4 * 9 * 11, c, scale, shift)
    assert rel_err(nchw(y), ref) < 1e-5
    assert rel_err(bn.running_mean.cpu(), rm) < 1e-5 and rel_err(bn.running_var.cpu(), rv) < 1e-5


# ----------------------------------------------------------------------------- BiSeNet glue
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_global_avgpool(cuda, dtype):
    tdt = ops.torch_dtype(dtype)
    for (n, c, h, w, ld) in [(2, 256, 32, 64, 256), (1, 512, 23, 40, 512), (2, 19, 64, 128, 32), (1, 64, 3, 5, 64)]:
        x = torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(c + h)) + 0.3
        xg = torch.zeros(n, h, w, ld, dtype=tdt, device="cuda")
        xg[..., :c] = nhwc(x, tdt)
        out = torch.empty(n, c, dtype=torch.float32, device="cuda")
        ops.global_avgpool(xg, n, h * w, c, ld, out)
        ref = x.to(tdt).float().mean((2, 3))
        assert rel_err(out.cpu(), ref) < 1e-5


@pytest.mark.parametrize("train", [False, True])
def test_arm_gate(cuda, train):
    from models.bisenet.build_bisenet import AttentionRefinementModule

    g = torch.Generator().manual_seed(3)
    n, c = 3, 256
    arm = AttentionRefinementModule(c, c)
    with torch.no_grad():
        arm.bn.weight.copy_(torch.rand(c, generator=g) + 0.5); arm.bn.bias.copy_(torch.randn(c, generator=g) * 0.2)
        arm.bn.running_mean.copy_(torch.randn(c, generator=g) * 0.1); arm.bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    pooled = torch.randn(n, c, generator=g)
    mul = torch.randn(n, c, generator=g)
    rm, rv = arm.bn.running_mean.clone(), arm.bn.running_var.clone()
    lin = F.conv2d(pooled.view(n, c, 1, 1), arm.conv.weight.detach(), arm.conv.bias.detach())
    ref = torch.sigmoid(F.batch_norm(lin, rm, rv, arm.bn.weight.detach(), arm.bn.bias.detach(), train, 0.1, 1e-5)).view(n, c) * mul
    arm = arm.cuda()
    gate = torch.empty(n, c, device="cuda")
    ops.arm_gate(pooled.cuda(), arm.conv, arm.bn, train, n, c, gate, mul.cuda())
    assert rel_err(gate.cpu(), ref) < 2e-5
    assert rel_err(arm.bn.running_mean.cpu(), rm) < 1e-5 and rel_err(arm.bn.running_var.cpu(), rv) < 1e-5
    if train:
        with pytest.raises(ops._lib.RtsdsError, match="N >= 2"):
            ops.arm_gate(pooled.cuda(), arm.conv, arm.bn, True, 1, c, gate)


@pytest.mark.parametrize("h,w,oh,ow", [(32, 64, 64, 128), (16, 32, 64, 128), (23, 40, 90, 160), (45, 80, 90, 160), (5, 7, 9, 13)])
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_gate_resize_nhwc(cuda, h, w, oh, ow, dtype):
    tdt = ops.torch_dtype(dtype)
    g = torch.Generator().manual_seed(h * w)
    n, c = 2, 64
    x = torch.randn(n, c, h, w, generator=g)
    gate = torch.rand(n, c, generator=g)
    dst = torch.zeros(n, oh, ow, 160, dtype=tdt, device="cuda")
    ops.gate_resize_nhwc(nhwc(x, tdt), n, h, w, c, c, gate.cuda(), oh, ow, dst, 160, 32, dtype)
    ref = F.interpolate(x.to(tdt).float() * gate.view(n, c, 1, 1), size=(oh, ow), mode="bilinear")
    assert rel_err(nchw(dst[..., 32:96]), ref) < (1e-5 if dtype == F32 else 1e-2)
    assert (dst[..., :32] == 0).all() and (dst[..., 96:] == 0).all()


def test_ffm_head_and_resize_to_nchw(cuda):
    from models.bisenet.build_bisenet import FeatureFusionModule

    g = torch.Generator().manual_seed(9)
    n, c, h, w = 2, 19, 12, 20
    ffm = FeatureFusionModule(19, 1024)
    final = torch.nn.Conv2d(19, 19, 1)
    feat = F.relu(torch.randn(n, c, h, w, generator=g))
    a = F.adaptive_avg_pool2d(feat, 1)
    a = torch.sigmoid(ffm.conv2(F.relu(ffm.conv1(a))))
    gref = feat * a + feat
    ref_lo = final(gref).detach()
    ref = final(F.interpolate(gref, scale_factor=8, mode="bilinear")).detach()
    ffm, final = ffm.cuda(), final.cuda()
    fg = torch.zeros(n, h, w, 32, device="cuda"); fg[..., :c] = nhwc(feat, torch.float32)
    pooled = torch.empty(n, c, device="cuda")
    ops.global_avgpool(fg, n, h * w, c, 32, pooled)
    z = torch.zeros(n, h, w, 32, device="cuda")
    ops.ffm_head(fg, F32, 32, pooled, n, h * w, c, ffm.conv1, ffm.conv2, final, z, 32)
    assert rel_err(nchw(z[..., :c]), ref_lo) < 1e-5
    out = torch.empty(n, c, h * 8, w * 8, device="cuda")
    ops.resize_to_nchw(z, n, h, w, c, 32, out)
    # 1x1 conv commutes with bilinear interpolation (weights sum to 1): only fp32 rounding differs
    assert rel_err(out.cpu(), ref) < 1e-5
    z2 = torch.zeros(n, h, w, 32, device="cuda")
    ops.ffm_head(fg, F32, 32, pooled, n, h * w, c, ffm.conv1, ffm.conv2, None, z2, 32)
    assert rel_err(nchw(z2[..., :c]), gref.detach()) < 1e-5


@pytest.mark.parametrize("h,w,oh,ow", [(64, 128, 512, 1024), (65, 129, 512, 1024), (9, 13, 72, 104), (90, 160, 720, 1280),
                                       (12, 20, 12, 20)])
def test_resize_to_nchw_generic_scales(cuda, h, w, oh, ow):
    g = torch.Generator().manual_seed(h + w)
    z = torch.randn(1, 19, h, w, generator=g)
    zg = torch.zeros(1, h, w, 32, device="cuda"); zg[..., :19] = nhwc(z, torch.float32)
    out = torch.empty(1, 19, oh, ow, device="cuda")
    ops.resize_to_nchw(zg, 1, h, w, 19, 32, out)
    ref = F.interpolate(z, size=(oh, ow), mode="bilinear")
    assert rel_err(out.cpu(), ref) < 1e-5


# ----------------------------------------------------------------------------- loss
@pytest.mark.parametrize("ignore", [19, 255])
@pytest.mark.parametrize("h,w,s", [(16, 24, 8), (9, 13, 8), (23, 31, 4)])
def test_fused_resize_ce_argmax(cuda, ignore, h, w, s):
    g = torch.Generator().manual_seed(h + ignore)
    n, c = 2, 19
    z = (torch.randn(n, c, h, w, generator=g) * 2).requires_grad_(True)
    oh, ow = h * s, w * s
    target = torch.randint(0, 20, (n, oh, ow), generator=g)
    target[target == 19] = ignore
    logits = F.interpolate(z, size=(oh, ow), mode="bilinear")
    loss = F.cross_entropy(logits, target, ignore_index=ignore)
    loss.backward()
    zg = torch.zeros(n, h, w, 32, device="cuda"); zg[..., :c] = nhwc(z.detach(), torch.float32)
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    pred = torch.empty(n, oh, ow, dtype=torch.int64, device="cuda")
    ops.resize_ce_argmax_fwd(zg, n, h, w, c, 32, oh, ow, target.cuda(), ignore, acc, pred)
    a = acc.cpu()
    valid = (target != ignore).sum().item()
    assert a[1].item() == valid
    assert abs(a[0].item() / valid - loss.item()) < 1e-5 * max(1.0, loss.item())
    ref_pred = logits.argmax(1)
    assert (pred.cpu() == ref_pred).float().mean().item() > 0.9999
    assert a[2].item() == (pred.cpu() == target).sum().item()
    # backward straight to z resolution
    gs = torch.tensor([1.0 / valid], device="cuda")
    dz = torch.zeros(n, h, w, 32, device="cuda")
    ops.resize_ce_bwd(zg, n, h, w, c, 32, oh, ow, target.cuda(), ignore, gs, dz)
    assert rel_err(nchw(dz[..., :c]), z.grad) < 1e-4
    assert (dz[..., c:] == 0).all()


@pytest.mark.parametrize("ignore", [19, 255])
@pytest.mark.parametrize("h,w,oh,ow,c", [(16, 24, 128, 192, 19), (9, 13, 72, 104, 19), (23, 40, 180, 320, 19), (65, 129, 512, 1024, 19),
                                         (12, 20, 96, 160, 7), (90, 160, 720, 1280, 19)])
def test_one_pass_resize_ce_forward_and_gradient(cuda, ignore, h, w, oh, ow, c):
    """rtsds_resize_ce_fused: loss sums, argmax and the unnormalised gradient in one pass (~x8 heads, incl. the
    non-integer 7.8x of the aux heads at 720x1280 and DeepLab's 65x129 -> 512x1024)."""
    g = torch.Generator().manual_seed(h + ignore)
    n = 2 if oh < 700 else 1
    z = (torch.randn(n, c, h, w, generator=g) * 2).requires_grad_(True)
    target = torch.randint(0, c + 1, (n, oh, ow), generator=g)
    target[target == c] = ignore
    assert ops.resize_ce_fused_supported(h, w, c, oh, ow) and not ops.resize_ce_fused_supported(h, w, c, 4 * h, 4 * w)
    logits = F.interpolate(z, size=(oh, ow), mode="bilinear")
    loss = F.cross_entropy(logits, target, ignore_index=ignore)
    (3.0 * loss).backward()
    zg = torch.zeros(n, h, w, 32, device="cuda"); zg[..., :c] = nhwc(z.detach(), torch.float32)
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    pred = torch.empty(n, oh, ow, dtype=torch.int64, device="cuda")
    dz = torch.zeros(n, h, w, 32, device="cuda")
    ops.resize_ce_fused(zg, n, h, w, c, 32, oh, ow, target.cuda(), ignore, acc, pred, dz)
    a = acc.cpu()
    valid = (target != ignore).sum().item()
    assert a[1].item() == valid
    assert abs(a[0].item() / valid - loss.item()) < 1e-5 * max(1.0, loss.item())
    assert (pred.cpu() == logits.argmax(1)).float().mean().item() > 0.9999
    assert a[2].item() == (pred.cpu() == target).sum().item()
    ops.scale_by_device_scalar(dz, torch.tensor([3.0 / valid], device="cuda"))
    assert rel_err(nchw(dz[..., :c]), z.grad) < 1e-4, rel_err(nchw(dz[..., :c]), z.grad)
    assert (dz[..., c:] == 0).all()
    # agrees with the two-kernel generic path
    acc2 = torch.zeros(4, dtype=torch.float64, device="cuda")
    ops.resize_ce_argmax_fwd(zg, n, h, w, c, 32, oh, ow, target.cuda(), ignore, acc2, None)
    assert abs(acc2[0].item() - a[0].item()) < 1e-5 * abs(a[0].item()) and acc2[1].item() == a[1].item()


def test_ce_argmax_nchw(cuda):
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(2, 19, 33, 47, generator=g) * 3
    target = torch.randint(0, 20, (2, 33, 47), generator=g)
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    pred = torch.empty(2, 33, 47, dtype=torch.int64, device="cuda")
    ops.ce_argmax_nchw_fwd(logits.cuda(), target.cuda(), 19, acc, pred)
    ref = F.cross_entropy(logits, target, ignore_index=19)
    a = acc.cpu()
    assert abs(a[0].item() / a[1].item() - ref.item()) < 1e-5
    assert (pred.cpu() == logits.argmax(1)).all()


# ----------------------------------------------------------------------------- fused tensor-core stems
@pytest.mark.parametrize("n,h,w", [(1, 64, 96), (2, 72, 104), (1, 45, 81), (1, 512, 1024), (2, 40, 50), (1, 70, 24), (3, 9, 11)])
@pytest.mark.parametrize("train", [False, True])
def test_stem_pair_tc(cuda, n, h, w, train):
    g = torch.Generator().manual_seed(h + w)
    x = torch.randn(n, 3, h, w, generator=g)
    w7 = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    w3 = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    scale = torch.rand(128, generator=g) + 0.5
    shift = torch.randn(128, generator=g)
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    wpk = ops.stem_pack_weights(w7.cuda(), w3.cuda())
    ycp = torch.full((n, oh, ow, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ysp = torch.full((n, oh, ow, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    xr, w7r, w3r = x.bfloat16().float(), w7.bfloat16().float(), w3.bfloat16().float()
    raw7 = F.conv2d(xr, w7r, None, 2, 3)
    raw3 = F.conv2d(xr, w3r, None, 2, 1)
    if train:
        st7 = torch.zeros(128, device="cuda"); st3 = torch.zeros(128, device="cuda")
        ops.stem_pair_tc_fwd(x.cuda(), wpk, ycp, ysp, None, None, False, st7, st3)
        assert rel_err(nchw(ycp), raw7) < 1e-2 and rel_err(nchw(ysp), raw3) < 1e-2
        assert rel_err(st7[:64].cpu(), raw7.sum((0, 2, 3))) < 2e-3 and rel_err(st7[64:].cpu(), (raw7 * raw7).sum((0, 2, 3))) < 2e-3
        assert rel_err(st3[:64].cpu(), raw3.sum((0, 2, 3))) < 2e-3 and rel_err(st3[64:].cpu(), (raw3 * raw3).sum((0, 2, 3))) < 2e-3
    else:
        ops.stem_pair_tc_fwd(x.cuda(), wpk, ycp, ysp, scale.cuda(), shift.cuda(), True)
        r7 = F.relu(raw7 * scale[:64].view(1, -1, 1, 1) + shift[:64].view(1, -1, 1, 1))
        r3 = F.relu(raw3 * scale[64:].view(1, -1, 1, 1) + shift[64:].view(1, -1, 1, 1))
        assert rel_err(nchw(ycp), r7) < 1e-2 and rel_err(nchw(ysp), r3) < 1e-2


@pytest.mark.parametrize("n,h,w", [(1, 64, 96), (2, 72, 104), (1, 45, 81), (2, 256, 512), (2, 40, 50), (1, 70, 24), (3, 9, 11)])
def test_stem_pair_tc_wgrad(cuda, n, h, w):
    g = torch.Generator().manual_seed(h * 3 + w)
    x = torch.randn(n, 3, h, w, generator=g)
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    d7 = torch.randn(n, 64, oh, ow, generator=g)
    d3 = torch.randn(n, 64, oh, ow, generator=g)
    xr = x.bfloat16().float()
    w7 = torch.zeros(64, 3, 7, 7, requires_grad=True); w3 = torch.zeros(64, 3, 3, 3, requires_grad=True)
    (F.conv2d(xr, w7, None, 2, 3) * d7.bfloat16().float()).sum().backward()
    (F.conv2d(xr, w3, None, 2, 1) * d3.bfloat16().float()).sum().backward()
    ws = torch.zeros(128 * 192, device="cuda")
    g7 = torch.ones(64, 3, 7, 7, device="cuda"); g3 = torch.ones(64, 3, 3, 3, device="cuda")
    ops.stem_pair_tc_wgrad(x.cuda(), nhwc(d7, torch.bfloat16), nhwc(d3, torch.bfloat16), ws, g7, g3)
    assert rel_err(g7.cpu() - 1, w7.grad) < 2e-3 and rel_err(g3.cpu() - 1, w3.grad) < 2e-3
    assert (ws == 0).all()


# ----------------------------------------------------------------------------- space-to-depth stems
@pytest.mark.parametrize("n,h,w", [(1, 64, 96), (2, 72, 104), (1, 45, 81), (1, 512, 1024)])
@pytest.mark.parametrize("k", [7, 3])
def test_stem_s2d_forward_and_wgrad(cuda, n, h, w, k):
    """The 4-tap implicit GEMM over the padded space-to-depth image == conv k x k stride 2 (7x7 p3 / 3x3 p1) on the image,
    forward (+BN statistics) and weight gradient, incl. odd sizes."""
    g = torch.Generator().manual_seed(h + k)
    x = torch.randn(n, 3, h, w, generator=g)
    wt = torch.randn(64, 3, k, k, generator=g) * (2.0 / (3 * k * k)) ** 0.5
    xr = x.to(torch.bfloat16).float()
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, 2, k // 2)
    oh, ow, pshape = ops.stem_s2d_shape(n, h, w)
    assert ref.shape[-2:] == (oh, ow)
    P = torch.full(pshape, float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.stem_s2d_pack(x.cuda(), P)
    w2 = torch.empty(64, 64, 4, 1, device="cuda")
    ops.stem_s2d_weight(wt.cuda(), w2)
    wpk = ops.pack_conv_weight(w2, BF16)
    y = torch.full((n, oh, ow, 64), float("nan"), dtype=torch.float32, device="cuda")
    stats = torch.zeros(128, device="cuda")
    ops.stem_s2d_conv_fwd(P, n, oh, ow, wpk, 64, y, 64, F32, stats=stats)
    assert rel_err(nchw(y), ref.detach()) < 2e-5
    assert rel_err(stats[:64].cpu(), ref.detach().sum((0, 2, 3))) < 1e-3
    dy = torch.randn(n, 64, oh, ow, generator=g)
    ref.backward(dy.to(torch.bfloat16).float())
    dw = torch.zeros(64 * 4 * 64, device="cuda")
    ops.stem_s2d_conv_wgrad(P, n, oh, ow, nhwc(dy, torch.bfloat16), 64, 64, dw)
    g2 = torch.zeros(64, 64, 4, 1, device="cuda")
    ops.unpack_conv_wgrad(dw, g2, False)
    gw = torch.zeros(64, 3, k, k, device="cuda")
    ops.stem_s2d_weight_grad(g2, gw)
    assert rel_err(gw.cpu(), wr.grad) < 2e-3, rel_err(gw.cpu(), wr.grad)


@pytest.mark.parametrize("n,c,h,w,oh,ow", [(2, 19, 16, 24, 128, 192), (1, 19, 23, 40, 184, 320), (2, 7, 9, 70, 72, 560), (1, 19, 12, 20, 50, 77),
                                            (1, 19, 5, 6, 160, 192), (2, 19, 33, 35, 33, 35), (1, 3, 64, 128, 512, 1024)])
def test_resize_to_nchw_backward(cuda, n, c, h, w, oh, ow):
    """Adjoint of the logits writer (bilinear, align_corners=False, NHWC pitch-32 source -> NCHW) against torch.autograd of
    F.interpolate: integer and fractional factors, x32 (more candidates than the unrolled part), identity."""
    from rtsds_b200._lib import check, lib
    g = torch.Generator().manual_seed(h * w + oh)
    z = torch.randn(n, c, h, w, generator=g, requires_grad=True)
    dout = torch.randn(n, c, oh, ow, generator=g)
    F.interpolate(z, size=(oh, ow), mode="bilinear", align_corners=False).backward(dout)
    dz = torch.full((n, h, w, 32), float("nan"), device="cuda")
    dg = dout.cuda()
    check(lib().rtsds_resize_to_nchw_bwd(dg.data_ptr(), n, c, oh, ow, h, w, dz.data_ptr(), 32, None), "resize_to_nchw_bwd")
    torch.cuda.synchronize()
    got = dz[..., :c].permute(0, 3, 1, 2).cpu()
    assert rel_err(got, z.grad) < 1e-5


@pytest.mark.parametrize("shape,out", [((2, 19, 720, 1280), (512, 1024)), ((1, 19, 96, 130), (64, 96)), ((2, 5, 37, 53), (37, 53)),
                                       ((1, 3, 40, 60), (64, 96))])
def test_adaptive_avg_pool2d_matches_torch(cuda, shape, out):
    """train.py:410,438,445 (adversarial_train_2): F.adaptive_avg_pool2d of the logits, forward and backward."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(*shape, generator=g)
    dy = torch.randn(shape[0], shape[1], *out, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.adaptive_avg_pool2d(xr, out)
    yr.backward(dy)
    xg = x.cuda().requires_grad_(True)
    yg = ops.adaptive_avg_pool2d(xg, out)
    yg.backward(dy.cuda())
    assert (yg.cpu() - yr).abs().max().item() <= 1e-6 * max(1.0, yr.abs().max().item())
    assert (xg.grad.cpu() - xr.grad).abs().max().item() <= 1e-6 * max(1.0, xr.grad.abs().max().item())
