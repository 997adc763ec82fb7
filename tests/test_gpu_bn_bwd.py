"""BatchNorm(+ReLU) backward kernels called through the C ABI, both forms: the register form
(small maps, fp32, pitched views) and the bulk-copy streamed form (contiguous bf16, >= 1 Mi elements).
Reference: the closed-form BN backward in fp64 on the same bf16-rounded inputs, i.e. what
torch.autograd computes for nn.BatchNorm2d(train) -> ReLU in models/bisenet/build_bisenet.py ConvBlock."""
import pytest
import torch

from rtsds_b200._lib import check, lib
from rtsds_b200.ops import BF16, F32

pytestmark = pytest.mark.gpu


def _p(t):
    return None if t is None else t.data_ptr()


def _reference(dy, raw, y, mean, invstd, gamma, fsc, fsh, mode):
    d, r = dy.double(), raw.double()
    if mode == "rawmask":
        mask = torch.addcmul(fsh.float(), raw.float(), fsc.float()) > 0     # fp32, like the forward
    elif mode == "y":
        mask = y.double() > 0
    else:
        mask = torch.ones_like(d, dtype=torch.bool)
    g = d * mask
    xh = (r - mean.double()) * invstd.double()
    m = dy.shape[0]
    s1, s2 = g.sum(0), (g * xh).sum(0)
    d_raw = gamma.double() * invstd.double() * (g - s1 / m - xh * s2 / m)
    return s1, s2, d_raw, g


CASES = [(40003, 64, "rawmask"), (40003, 64, "y"), (40003, 64, "none"), (9001, 256, "rawmask"), (4099, 512, "y"),
         (1031, 2048, "rawmask"), (16384, 64, "rawmask"), (300, 64, "rawmask"), (777, 24, "y"), (50000, 48, "rawmask")]


@pytest.mark.parametrize("n_pix,c,mode", CASES)
@pytest.mark.parametrize("dtype", [BF16, F32])
def test_bn_bwd(cuda, n_pix, c, mode, dtype):
    g = torch.Generator(device="cuda").manual_seed(n_pix + c)
    tdt = torch.bfloat16 if dtype == BF16 else torch.float32
    raw = (torch.randn(n_pix, c, device="cuda", generator=g) * 1.5 + 0.3).to(tdt)
    dy = torch.randn(n_pix, c, device="cuda", generator=g).to(tdt)
    mean = raw.float().mean(0)
    invstd = (raw.float().var(0, unbiased=False) + 1e-5).rsqrt()
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g) * 0.2
    fsc = gamma * invstd
    fsh = beta - mean * fsc
    y = torch.relu(torch.addcmul(fsh, raw.float(), fsc) + 0.1 * torch.randn(n_pix, c, device="cuda", generator=g)).to(tdt)
    sums = torch.empty(2 * c, device="cuda")
    d_raw = torch.full((n_pix, c), 7.0, device="cuda", dtype=tdt)
    g_out = torch.full((n_pix, c), 7.0, device="cuda", dtype=tdt)
    dgam = torch.ones(c, device="cuda")
    dbet = torch.ones(c, device="cuda")
    if mode == "rawmask":
        check(lib().rtsds_bn_bwd_reduce_rawmask(_p(dy), c, _p(raw), c, _p(mean), _p(invstd), _p(fsc), _p(fsh), n_pix, c, dtype,
                                                _p(sums), None), "reduce")
        check(lib().rtsds_bn_bwd_apply_rawmask(_p(dy), c, _p(raw), c, _p(mean), _p(invstd), _p(gamma), _p(sums), _p(fsc), _p(fsh),
                                               n_pix, c, dtype, _p(d_raw), c, dtype, _p(g_out), c, _p(dgam), _p(dbet), None),
              "apply")
    else:
        relu = int(mode == "y")
        check(lib().rtsds_bn_bwd_reduce(_p(dy), c, _p(y), c, _p(raw), c, _p(mean), _p(invstd), n_pix, c, relu, dtype, _p(sums),
                                        None), "reduce")
        check(lib().rtsds_bn_bwd_apply(_p(dy), c, _p(y), c, _p(raw), c, _p(mean), _p(invstd), _p(gamma), _p(sums), n_pix, c, relu,
                                       dtype, _p(d_raw), c, dtype, _p(g_out), c, _p(dgam), _p(dbet), None), "apply")
    torch.cuda.synchronize()
    s1, s2, want, gref = _reference(dy, raw, y, mean, invstd, gamma, fsc, fsh, mode)
    scale = float(n_pix) ** 0.5
    assert (sums[:c].double() - s1).abs().max().item() < 2e-3 * scale
    assert (sums[c:].double() - s2).abs().max().item() < 2e-3 * scale
    tol = 1.5e-2 if dtype == BF16 else 2e-4
    err = (d_raw.double() - want).abs().max().item() / want.abs().max().item()
    assert err < tol, err
    # the rawmask sign can flip where the forward value is within an ulp of zero: allow a handful of such pixels
    assert int(((g_out.double() - gref).abs() > 1e-6).sum().item()) <= 4
    assert (dgam.double() - 1 - s2).abs().max().item() < 2e-3 * scale
    assert (dbet.double() - 1 - s1).abs().max().item() < 2e-3 * scale


# ----------------------------------------------------------------------------- batched weight pack / gradient unpack
PACK_SHAPES = [(64, 64, 3), (19, 1024, 3), (512, 512, 3), (128, 64, 1), (256, 19, 4), (1, 256, 4), (64, 3, 7), (2048, 512, 1),
               (40, 24, 3), (19, 2048, 1)]


@pytest.mark.parametrize("dtype", [BF16, F32])
def test_pack_batch_matches_reference_layout(cuda, dtype):
    """[cout_pad][taps][cin] and the dgrad operand [cin_pad][taps][ck] against a torch permute of the OIHW weight."""
    from rtsds_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    tdt = torch.bfloat16 if dtype == BF16 else torch.float32
    jobs, want = [], []
    for co, ci, k in PACK_SHAPES:
        w = torch.randn(co, ci, k, k, device="cuda", generator=g)
        cp = ops.cout_pad(co)
        out0 = torch.full((cp, k * k, ci), 3.0, device="cuda", dtype=tdt)
        ref0 = torch.zeros(cp, k * k, ci, device="cuda")
        ref0[:co] = w.reshape(co, ci, k * k).permute(0, 2, 1)
        jobs.append((w, out0, 0)); want.append(ref0.to(tdt))
        cip, ck = ops.cout_pad(ci), ops.dgrad_ck(co, True)
        out1 = torch.full((cip, k * k, ck), 3.0, device="cuda", dtype=tdt)
        ref1 = torch.zeros(cip, k * k, ck, device="cuda")
        ref1[:ci, :, :co] = w.reshape(co, ci, k * k).permute(1, 2, 0)
        jobs.append((w, out1, 1)); want.append(ref1.to(tdt))
    ops.pack_conv_weights_batch(jobs, dtype, True)
    torch.cuda.synchronize()
    for (w, out, kind), ref in zip(jobs, want):
        assert torch.equal(out, ref), (tuple(w.shape), kind)


@pytest.mark.parametrize("accumulate", [False, True])
def test_unpack_batch(cuda, accumulate):
    from rtsds_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(6)
    jobs, want = [], []
    for co, ci, k in PACK_SHAPES:
        dw = torch.randn(co, k * k, ci, device="cuda", generator=g)
        grad = torch.randn(co, ci, k, k, device="cuda", generator=g)
        ref = dw.permute(0, 2, 1).reshape(co, ci, k, k) + (grad if accumulate else 0)
        jobs.append((dw, grad, accumulate)); want.append(ref.clone())
    ops.unpack_conv_wgrads_batch(jobs)
    torch.cuda.synchronize()
    for (dw, grad, _), ref in zip(jobs, want):
        assert torch.equal(grad, ref), tuple(grad.shape)
        assert not dw.any(), "scratch must be left zeroed"


# ----------------------------------------------------------------------------- train-mode BatchNorm forward in one launch
@pytest.mark.parametrize("n_pix,c", [(40003, 64), (9001, 256), (1031, 2048), (300, 128), (777, 24)])
@pytest.mark.parametrize("dtype", [BF16, F32])
@pytest.mark.parametrize("residual", [False, True])
def test_bn_finalize_apply_equals_the_two_launch_form(cuda, n_pix, c, dtype, residual):
    """rtsds_bn_finalize_apply == rtsds_bn_finalize + rtsds_scale_shift_act, bit for bit (c = 24: the fallback itself)."""
    from rtsds_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(n_pix + c)
    tdt = torch.bfloat16 if dtype == BF16 else torch.float32
    raw = (torch.randn(n_pix, c, device="cuda", generator=g) * 1.5 + 0.3).to(tdt)
    res = torch.randn(n_pix, c, device="cuda", generator=g).to(tdt) if residual else None
    stats = torch.cat([raw.float().sum(0), (raw.float() ** 2).sum(0)]).contiguous()
    outs = []
    for fused in (False, True):
        bn = torch.nn.BatchNorm2d(c).cuda()
        with torch.no_grad():
            bn.weight.copy_(torch.rand(c, device="cuda", generator=g) * 0 + torch.linspace(0.5, 1.5, c, device="cuda"))
            bn.bias.copy_(torch.linspace(-0.2, 0.2, c, device="cuda"))
        scale, shift = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        sm, si = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        y = torch.full((n_pix, c), 7.0, device="cuda", dtype=tdt)
        if fused:
            ops.bn_finalize_apply_ptr(stats, n_pix, bn, scale, shift, sm, si, raw, y, n_pix, c, res, ops.ACT_RELU, 0.0, c, c, c, dtype, dtype)
        else:
            ops.bn_finalize(stats, n_pix, bn, scale, shift, sm, si)
            ops.scale_shift_act_ptr(raw, y, n_pix, c, scale, shift, res, ops.ACT_RELU, 0.0, c, c, c, dtype, dtype)
        torch.cuda.synchronize()
        outs.append((y, scale, shift, sm, si, bn.running_mean.clone(), bn.running_var.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # and it is BatchNorm: torch on the same (rounded) input
    ref = torch.nn.functional.batch_norm(raw.float(), None, None, torch.linspace(0.5, 1.5, c, device="cuda"),
                                         torch.linspace(-0.2, 0.2, c, device="cuda"), True, 0.1, 1e-5)
    if residual:
        ref = ref + res.float()
    ref = torch.relu(ref)
    tol = 2e-2 if dtype == BF16 else 1e-4
    assert (outs[1][0].float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
