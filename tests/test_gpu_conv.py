"""Parity of the convolution kernels against the CPU oracle (torch fp32 on CPU).

tcgen05 path: operands are bf16, accumulation fp32 -> compare with the oracle on
bf16-rounded operands; tolerance 2e-2 rel for bf16 outputs (BASELINE.json), much
tighter when the output is kept in fp32.  SIMT fp32 path: 1e-4.
"""
import os

import pytest
import torch
import torch.nn.functional as F

from rtsds_b200 import ops
from rtsds_b200.ops import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F32

from gpu_util import bf16_round, conv_ref, nchw, nhwc, rel_err, run_conv

pytestmark = pytest.mark.gpu


def _mk(n, cin, h, w, cout, k, seed=0, kw=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k if kw is None else kw, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    return x, wt, g


# (name, n, cin, h, w, cout, k, stride, pad, dil)
TC_CASES = [
    ("gemm_1tile", 1, 64, 8, 16, 32, 1, 1, 0, 1),
    ("gemm_2kblk", 1, 128, 8, 16, 64, 1, 1, 0, 1),
    ("3x3_s1", 1, 64, 16, 32, 64, 3, 1, 1, 1),
    ("3x3_s1_multi", 2, 64, 64, 128, 64, 3, 1, 1, 1),
    ("3x3_s2_even", 1, 64, 32, 64, 128, 3, 2, 1, 1),
    ("3x3_s2_odd", 2, 64, 45, 80, 128, 3, 2, 1, 1),
    ("3x3_s2_odd2", 1, 128, 23, 41, 256, 3, 2, 1, 1),
    ("1x1_s2", 2, 128, 24, 40, 256, 1, 2, 0, 1),
    ("1x1_s2_odd", 1, 64, 45, 81, 128, 1, 2, 0, 1),
    ("layer4", 2, 512, 4, 6, 512, 3, 1, 1, 1),
    ("dil2", 1, 64, 17, 33, 64, 3, 1, 2, 2),
    ("dil4", 1, 128, 20, 20, 128, 3, 1, 4, 4),
    ("disc_4x4_s2", 2, 64, 32, 48, 128, 4, 2, 1, 1),
    ("ragged_w", 1, 64, 9, 13, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tc_raw_fp32_out(cuda, case):
    _, n, cin, h, w, cout, k, stride, pad, dil = case
    x, wt, _ = _mk(n, cin, h, w, cout, k)
    y, _, _ = run_conv("tc", x, wt, stride=stride, pad=pad, dil=dil, out_dtype=F32, out_ld=ops.cout_pad(cout))
    ref, _ = conv_ref(x, wt, stride=stride, pad=pad, dil=dil)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 2e-4, rel_err(y, ref)


@pytest.mark.parametrize("block_n", [32, 64, 128])
def test_conv_tc_block_n_variants(cuda, block_n):
    x, wt, _ = _mk(2, 128, 20, 36, 256, 3, seed=3)
    ops.lib().rtsds_conv2d_tc_tune(block_n, 0)
    try:
        y, _, _ = run_conv("tc", x, wt, out_dtype=F32, out_ld=256)
    finally:
        ops.lib().rtsds_conv2d_tc_tune(0, 0)
    ref, _ = conv_ref(x, wt)
    assert rel_err(y, ref) < 2e-4


@pytest.mark.parametrize("stages", [2, 3, 6])
def test_conv_tc_stage_counts(cuda, stages):
    x, wt, _ = _mk(1, 256, 16, 32, 128, 3, seed=4)
    ops.lib().rtsds_conv2d_tc_tune(0, stages)
    try:
        y, _, _ = run_conv("tc", x, wt, out_dtype=F32, out_ld=128)
    finally:
        ops.lib().rtsds_conv2d_tc_tune(0, 0)
    ref, _ = conv_ref(x, wt)
    assert rel_err(y, ref) < 2e-4


@pytest.mark.parametrize("split", [2, 4, 8])
def test_conv_tc_split_k(cuda, split):
    x, wt, g = _mk(1, 256, 16, 32, 256, 3, seed=5)
    scale = torch.rand(256, generator=g) + 0.5
    shift = torch.randn(256, generator=g)
    res = torch.randn(1, 256, 16, 32, generator=g)
    y, st, _ = run_conv("tc", x, wt, scale=scale, shift=shift, residual=res, act=ACT_RELU, split_k=split, want_stats=True)
    ref, raw = conv_ref(x, wt, scale=scale, shift=shift, residual=res, act=ACT_RELU)
    assert rel_err(y, ref) < 1e-2
    s1 = raw.sum((0, 2, 3))
    s2 = (raw * raw).sum((0, 2, 3))
    assert rel_err(st[:256], s1) < 1e-3 and rel_err(st[256:], s2) < 1e-3


def test_conv_tc_epilogue_bn_residual_relu_bf16(cuda):
    x, wt, g = _mk(2, 64, 32, 64, 64, 3, seed=6)
    scale = torch.rand(64, generator=g) + 0.5
    shift = torch.randn(64, generator=g)
    res = torch.randn(2, 64, 32, 64, generator=g)
    y, _, _ = run_conv("tc", x, wt, scale=scale, shift=shift, residual=res, act=ACT_RELU)
    ref, _ = conv_ref(x, wt, scale=scale, shift=shift, residual=res, act=ACT_RELU)
    assert rel_err(y, ref) < 1e-2          # bf16 output rounding (2^-9 relative)


def test_conv_tc_bias_leaky_relu(cuda):
    x, wt, g = _mk(2, 64, 32, 48, 128, 4, seed=7)
    bias = torch.randn(128, generator=g)
    y, _, _ = run_conv("tc", x, wt, stride=2, pad=1, shift=bias, act=ACT_LRELU, slope=0.2, out_dtype=F32, out_ld=128)
    ref, _ = conv_ref(x, wt, stride=2, pad=1, shift=bias, act=ACT_LRELU, slope=0.2)
    assert rel_err(y, ref) < 2e-4


def test_conv_tc_stats_for_train_bn(cuda):
    x, wt, _ = _mk(2, 64, 45, 80, 128, 3, seed=8)
    y, st, _ = run_conv("tc", x, wt, stride=2, want_stats=True)
    ref, raw = conv_ref(x, wt, stride=2)
    assert rel_err(y, ref) < 1e-2
    assert rel_err(st[:128], raw.sum((0, 2, 3))) < 1e-3
    assert rel_err(st[128:], (raw * raw).sum((0, 2, 3))) < 1e-3


def test_conv_tc_skinny_cout19_from_concat_view(cuda):
    """FFM conv (1024->19, fp32 out with pitch 32) and a supervision head reading a channel slice."""
    x, wt, g = _mk(2, 1024, 12, 20, 19, 3, seed=9)
    y, _, ybuf = run_conv("tc", x, wt, out_dtype=F32, out_ld=32, act=ACT_RELU)
    ref, _ = conv_ref(x, wt, act=ACT_RELU)
    assert rel_err(y, ref) < 2e-4
    assert torch.isnan(ybuf[..., 19:]).all(), "padding channels beyond cout must not be written"
    x2, w2, g = _mk(2, 256, 12, 20, 19, 1, seed=10)
    bias = torch.randn(19, generator=g)
    y2, _, _ = run_conv("tc", x2, w2, pad=0, shift=bias, out_dtype=F32, out_ld=32, in_ld=1024, x_off=256)
    ref2, _ = conv_ref(x2, w2, pad=0, shift=bias)
    assert rel_err(y2, ref2) < 2e-4


def test_conv_tc_rejects_bad_arguments(cuda):
    x, wt, _ = _mk(1, 48, 8, 8, 32, 3)
    with pytest.raises(ops._lib.RtsdsError, match="multiple of 64"):
        run_conv("tc", x, wt)


SIMT_CASES = [
    ("3x3_s1", 2, 64, 16, 32, 64, 3, 1, 1, 1),
    ("3x3_s2_odd", 1, 64, 45, 80, 128, 3, 2, 1, 1),
    ("cin19", 1, 19, 12, 20, 19, 1, 1, 0, 1),
    ("1x1_s2", 1, 128, 24, 40, 256, 1, 2, 0, 1),
    ("dil2", 1, 64, 17, 33, 64, 3, 1, 2, 2),
    ("4x4_s2", 1, 64, 32, 48, 128, 4, 2, 1, 1),
    ("ffm", 1, 1024, 12, 20, 19, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", SIMT_CASES, ids=[c[0] for c in SIMT_CASES])
def test_conv_simt_fp32_check_mode(cuda, case):
    _, n, cin, h, w, cout, k, stride, pad, dil = case
    x, wt, g = _mk(n, cin, h, w, cout, k, seed=11)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g)
    y, st, _ = run_conv("simt", x, wt, stride=stride, pad=pad, dil=dil, scale=scale, shift=shift, act=ACT_RELU, dtype=F32,
                        want_stats=True)
    ref, raw = conv_ref(x, wt, stride=stride, pad=pad, dil=dil, scale=scale, shift=shift, act=ACT_RELU, round_inputs=False)
    assert rel_err(y, ref) < 1e-5
    assert rel_err(st[:cout], raw.sum((0, 2, 3))) < 1e-4


def test_conv_simt_bf16_cross_checks_tc(cuda):
    x, wt, _ = _mk(2, 128, 23, 40, 128, 3, seed=12)
    a, _, _ = run_conv("tc", x, wt, out_dtype=F32, out_ld=128)
    b, _, _ = run_conv("simt", x, wt, out_dtype=F32, out_ld=128)
    assert rel_err(a, b) < 1e-4


# ----------------------------------------------------------------------------- backward (dgrad / wgrad)
BWD_CASES = [
    ("3x3_s1", 2, 64, 16, 32, 64, 3, 1, 1, 1),
    ("3x3_s1_c128", 1, 128, 20, 36, 256, 3, 1, 1, 1),
    ("3x3_s2_even", 2, 64, 32, 64, 128, 3, 2, 1, 1),
    ("3x3_s2_odd", 1, 64, 45, 81, 128, 3, 2, 1, 1),
    ("1x1_s2", 2, 128, 24, 40, 256, 1, 2, 0, 1),
    ("1x1_s2_odd", 1, 64, 23, 41, 128, 1, 2, 0, 1),
    ("1x1_s1_cout19", 2, 256, 12, 20, 19, 1, 1, 0, 1),
    ("ffm_cout19", 1, 1024, 12, 20, 19, 3, 1, 1, 1),
    ("layer4", 2, 512, 4, 6, 512, 3, 1, 1, 1),
    ("4x4_s2", 1, 64, 32, 48, 128, 4, 2, 1, 1),
    ("dil2", 1, 64, 17, 33, 64, 3, 1, 2, 2),
]


def _bwd_ref(x, wt, dy, stride, pad, dil, rnd):
    xr, wr, dyr = (bf16_round(x), bf16_round(wt), bf16_round(dy)) if rnd else (x, wt, dy)
    xr = xr.clone().requires_grad_(True)
    wr1 = wr.clone().requires_grad_(True)
    y = torch.nn.functional.conv2d(xr, wr1, None, stride, pad, dil)
    y.backward(dyr)
    return xr.grad, wr1.grad


def _run_bwd(kind, x, wt, dy, stride, pad, dil, dtype, residual=None):
    tc = kind == "tc"
    tdt = ops.torch_dtype(dtype)
    n, cin, h, w = x.shape
    cout, _, kh, kw = wt.shape
    ck = ops.dgrad_ck(cout, tc)
    dy_ld = max(ck, (cout + 7) // 8 * 8)
    d = ops.make_conv_desc(n, h, w, cin, cin, cout, dy_ld, kh, stride, pad, dil, in_dtype=dtype, out_dtype=dtype, kw=kw)
    dyg = torch.zeros(n, d.oh, d.ow, dy_ld, dtype=tdt, device="cuda")
    dyg[..., :cout] = nhwc(dy, tdt)
    xg = nhwc(x, tdt)
    wd = ops.pack_conv_weight_dgrad(wt.cuda().contiguous(), dtype, tc)
    dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.float32, device="cuda")
    res = nhwc(residual, torch.float32) if residual is not None else None
    ws = torch.empty(max(1, int(ops.lib().rtsds_conv2d_tc_dgrad_workspace_bytes(d))), dtype=torch.uint8, device="cuda") if tc else None
    ops.conv2d_dgrad(d, dyg, wd, dx, F32, tc, res, ws)
    dwp = torch.zeros(cout, kh * kw, cin, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(d, xg, dyg, dwp, tc)
    gw = torch.ones(cout, cin, kh, kw, device="cuda")
    ops.unpack_conv_wgrad(dwp, gw, True)
    torch.cuda.synchronize()
    return nchw(dx), gw.cpu() - 1.0


@pytest.mark.parametrize("case", BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_conv_tc_backward(cuda, case):
    from gpu_util import bf16_round  # noqa: F401
    _, n, cin, h, w, cout, k, stride, pad, dil = case
    x, wt, g = _mk(n, cin, h, w, cout, k, seed=21)
    oh, ow = ops.conv_out_size(h, k, stride, pad, dil), ops.conv_out_size(w, k, stride, pad, dil)
    dy = torch.randn(n, cout, oh, ow, generator=g)
    res = torch.randn(n, cin, h, w, generator=g)
    dx, dw = _run_bwd("tc", x, wt, dy, stride, pad, dil, BF16, residual=res)
    rdx, rdw = _bwd_ref(x, wt, dy, stride, pad, dil, True)
    assert rel_err(dx, rdx + res) < 2e-4, rel_err(dx, rdx + res)
    assert rel_err(dw, rdw) < 2e-4, rel_err(dw, rdw)


@pytest.mark.parametrize("case", BWD_CASES[:6] + BWD_CASES[9:], ids=[c[0] for c in BWD_CASES[:6] + BWD_CASES[9:]])
def test_conv_simt_backward_fp32(cuda, case):
    _, n, cin, h, w, cout, k, stride, pad, dil = case
    x, wt, g = _mk(n, cin, h, w, cout, k, seed=22)
    oh, ow = ops.conv_out_size(h, k, stride, pad, dil), ops.conv_out_size(w, k, stride, pad, dil)
    dy = torch.randn(n, cout, oh, ow, generator=g)
    dx, dw = _run_bwd("simt", x, wt, dy, stride, pad, dil, F32)
    rdx, rdw = _bwd_ref(x, wt, dy, stride, pad, dil, False)
    assert rel_err(dx, rdx) < 1e-5 and rel_err(dw, rdw) < 1e-4


def test_conv_simt_backward_odd_channels(cuda):
    x, wt, g = _mk(1, 19, 12, 20, 19, 1, seed=23)
    dy = torch.randn(1, 19, 12, 20, generator=g)
    dx, dw = _run_bwd("simt", x, wt, dy, 1, 0, 1, F32)
    rdx, rdw = _bwd_ref(x, wt, dy, 1, 0, 1, False)
    assert rel_err(dx, rdx) < 1e-5 and rel_err(dw, rdw) < 1e-4


# ----------------------------------------------------------------------------- taps-as-N form of skinny-output convs
class _MiniPlan:
    def __init__(self, dtype):
        self.dt = dtype
        self.use_tc = dtype == BF16
        self.pack_steps, self._keep, self._ws = [], [], 0
        self.ws = None

    def buf(self, *shape, dtype=None):
        t = torch.empty(shape, dtype=ops.torch_dtype(self.dt) if dtype is None else dtype, device="cuda")
        self._keep.append(t)
        return t

    def zeros(self, *shape, dtype=None):
        t = torch.zeros(shape, dtype=ops.torch_dtype(self.dt) if dtype is None else dtype, device="cuda")
        self._keep.append(t)
        return t

    def note_ws(self, b):
        self._ws = max(self._ws, b)


@pytest.mark.parametrize("dtype,tol", [(F32, 1e-4), (BF16, 2e-2)])
@pytest.mark.parametrize("n,h,w,cin,c,k,dil", [(2, 8, 12, 256, 19, 3, 1), (1, 23, 40, 1024, 19, 3, 1), (1, 17, 9, 512, 7, 3, 2)])
def test_tapn_conv_forward_backward(cuda, dtype, tol, n, h, w, cin, c, k, dil):
    """csrc/tapn.cu + rtsds_b200/tapn.py against F.conv2d autograd: output, BN statistics, input and weight gradients."""
    from rtsds_b200.tapn import TapNConv

    g = torch.Generator().manual_seed(cin + h)
    pad = dil * (k // 2)
    conv = torch.nn.Conv2d(cin, c, k, 1, pad, dil, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(c, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5)
    conv = conv.cuda()
    x = torch.randn(n, cin, h, w, generator=g)
    dy = torch.randn(n, c, h, w, generator=g)
    rnd = bf16_round if dtype == BF16 else (lambda t: t)
    xr = rnd(x).requires_grad_(True)
    wr = rnd(conv.weight.detach().cpu()).requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, pad, dil)
    yr.backward(rnd(dy))
    plan = _MiniPlan(dtype)
    tdt = ops.torch_dtype(dtype)
    xg = nhwc(x, tdt)
    tn = TapNConv(plan, conv, xg, (n, h, w, cin), cin, train=True)
    plan.ws = torch.empty(max(plan._ws, 16), dtype=torch.uint8, device="cuda")
    for s in plan.pack_steps:
        s()
    y = torch.zeros(n, h, w, 32, device="cuda")
    stats = torch.zeros(2 * c, device="cuda")
    tn.forward(None, None, ACT_NONE, stats, y, 32)
    assert rel_err(nchw(y[..., :c]), yr.detach()) < tol
    assert rel_err(stats[:c].cpu(), yr.detach().sum((0, 2, 3))) < max(tol, 1e-3)
    assert rel_err(stats[c:].cpu(), (yr.detach() ** 2).sum((0, 2, 3))) < max(tol, 1e-3)
    dyg = torch.zeros(n, h, w, 64, dtype=tdt, device="cuda")
    dyg[..., :c] = nhwc(dy, tdt)
    tn.scatter(dyg, 64, dtype)
    gw = torch.zeros(c, cin, k, k, device="cuda")
    tn.weight_grad(gw)
    dx = torch.full((n, h, w, cin), float("nan"), dtype=tdt, device="cuda")
    tn.input_grad(dx, cin, dtype, False)
    assert rel_err(gw.cpu(), wr.grad) < tol, rel_err(gw.cpu(), wr.grad)
    assert rel_err(nchw(dx), xr.grad) < tol, rel_err(nchw(dx), xr.grad)
    tn.input_grad(dx, cin, dtype, True)                       # accumulate form: dx += dL/dx
    assert rel_err(nchw(dx), 2 * xr.grad) < 2 * tol
    tn.weight_grad(gw)                                        # the wgrad scratch was left clean
    assert rel_err(gw.cpu(), 2 * wr.grad) < 2 * tol


# ----------------------------------------------------------------------------- persistent / halo variants (many tiles)
# (name, n, cin, h, w, cout, k, stride, pad, dil): >= 2 waves of 128-pixel tiles so that the persistent kernel (resident
# weights) and, for 3x3 stride-1, its halo form (one TMA box per 8x16 tile, taps as windows into it) are selected
BIG_CASES = [
    ("halo_64", 2, 64, 160, 192, 64, 3, 1, 1, 1),
    ("halo_128_odd", 3, 128, 101, 131, 128, 3, 1, 1, 1),
    ("halo_dil2", 2, 64, 130, 170, 64, 3, 1, 2, 2),
    ("halo_dil4", 2, 64, 129, 161, 128, 3, 1, 4, 4),
    ("persist_1x1", 2, 128, 160, 192, 64, 1, 1, 0, 1),
    ("persist_s2", 2, 64, 256, 320, 128, 3, 2, 1, 1),
]


@pytest.mark.parametrize("case", BIG_CASES, ids=[c[0] for c in BIG_CASES])
def test_conv_tc_many_tiles_forward_and_dgrad(cuda, case):
    _, n, cin, h, w, cout, k, stride, pad, dil = case
    x, wt, g = _mk(n, cin, h, w, cout, k, seed=5)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g)
    y, stats, _ = run_conv("tc", x, wt, stride=stride, pad=pad, dil=dil, out_dtype=F32, out_ld=ops.cout_pad(cout), want_stats=True)
    ref, raw = conv_ref(x, wt, stride=stride, pad=pad, dil=dil)
    assert y.shape == ref.shape and rel_err(y, ref) < 2e-4, rel_err(y, ref)
    assert rel_err(stats[:cout], raw.sum((0, 2, 3))) < 2e-3 and rel_err(stats[cout:], (raw * raw).sum((0, 2, 3))) < 2e-3
    res = torch.randn_like(ref)
    y2, _, _ = run_conv("tc", x, wt, stride=stride, pad=pad, dil=dil, scale=scale, shift=shift, residual=res, act=ACT_RELU)
    ref2, _ = conv_ref(x, wt, stride=stride, pad=pad, dil=dil, scale=scale, shift=shift, residual=res, act=ACT_RELU)
    assert rel_err(y2, ref2) < 1e-2
    oh, ow = ref.shape[-2:]
    dy = torch.randn(n, cout, oh, ow, generator=g)
    dx, dw = _run_bwd("tc", x, wt, dy, stride, pad, dil, BF16)
    rdx, rdw = _bwd_ref(x, wt, dy, stride, pad, dil, True)
    assert rel_err(dx, rdx) < 2e-4, rel_err(dx, rdx)
    assert rel_err(dw, rdw) < 2e-4, rel_err(dw, rdw)


def test_conv_tc_a_tile_multicast_switch(cuda):
    """RTSDS_MC=1 (read once per process): pairs of N tiles of one M tile form a cluster and TMA-multicast each other's
    half of every A tile.  Measured slower than independent CTAs on B200, so it is off by default;
    the path stays correct: the tensor-core forward / split-K / backward cases re-run in a child process with it on."""
    import subprocess
    import sys

    env = dict(os.environ, RTSDS_MC="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "raw_fp32_out or split_k or block_n_variants or epilogue_bn or test_conv_tc_backward"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
