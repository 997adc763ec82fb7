"""Helpers shared by the -m gpu parity tests (CUDA path vs the CPU oracle)."""
import torch
import torch.nn.functional as F

from rtsds_b200 import ops
from rtsds_b200.ops import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F32


def nhwc(t: torch.Tensor, dtype) -> torch.Tensor:
    """NCHW fp32 (CPU) -> contiguous NHWC on cuda:0 in dtype."""
    return t.permute(0, 2, 3, 1).contiguous().to(device="cuda", dtype=dtype)


def nchw(t: torch.Tensor) -> torch.Tensor:
    """NHWC (cuda) -> NCHW fp32 on CPU."""
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def act_ref(t, act, slope):
    if act == ACT_RELU:
        return F.relu(t)
    if act == ACT_LRELU:
        return F.leaky_relu(t, slope)
    return t


def run_conv(kind, x_nchw, w, *, stride=1, pad=1, dil=1, scale=None, shift=None, residual=None, act=ACT_NONE, slope=0.0,
             dtype=BF16, out_dtype=None, out_ld=None, in_ld=None, x_off=0, split_k=0, want_stats=False):
    """Run the CUDA conv (kind 'tc' or 'simt') and return (y NCHW fp32 CPU, stats or None)."""
    tdt = ops.torch_dtype(dtype)
    out_dtype = dtype if out_dtype is None else out_dtype
    n, cin, h, wd = x_nchw.shape
    cout, _, kh, kw = w.shape
    in_ld_eff = cin if in_ld is None else in_ld
    xg = torch.zeros(n, h, wd, in_ld_eff, dtype=tdt, device="cuda")
    xg[..., x_off:x_off + cin] = nhwc(x_nchw, tdt)
    wp = ops.pack_conv_weight(w.cuda().contiguous(), dtype)
    out_ld_eff = cout if out_ld is None else out_ld
    d = ops.make_conv_desc(n, h, wd, cin, in_ld_eff, cout, out_ld_eff, kh, stride, pad, dil, act=act, slope=slope,
                           in_dtype=dtype, out_dtype=out_dtype, res_ld=cout if residual is not None else 0,
                           split_k=split_k, kw=kw)
    y = torch.full((n, d.oh, d.ow, out_ld_eff), float("nan"), dtype=ops.torch_dtype(out_dtype), device="cuda")
    sc = scale.cuda() if scale is not None else None
    sh = shift.cuda() if shift is not None else None
    res = nhwc(residual, ops.torch_dtype(out_dtype)) if residual is not None else None
    stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda") if want_stats else None
    xp = xg.data_ptr() + x_off * xg.element_size()
    if kind == "tc":
        ws = torch.empty(int(ops.lib().rtsds_conv2d_tc_workspace_bytes(d)), dtype=torch.uint8, device="cuda")
        ops.conv2d_tc(d, xp, wp, y, sc, sh, res, stats, ws)
    else:
        ops.conv2d_simt(d, xp, wp, y, sc, sh, res, stats)
    torch.cuda.synchronize()
    return nchw(y[..., :cout]), (stats.cpu() if stats is not None else None), y


def conv_ref(x, w, *, stride=1, pad=1, dil=1, scale=None, shift=None, residual=None, act=ACT_NONE, slope=0.0,
             round_inputs=True):
    """CPU oracle for one conv layer: torch fp32 on (optionally bf16-rounded) operands."""
    if round_inputs:
        x, w = bf16_round(x), bf16_round(w)
        if residual is not None:
            residual = bf16_round(residual)
    raw = F.conv2d(x, w, None, stride=stride, padding=pad, dilation=dil)
    y = raw
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual
    return act_ref(y, act, slope), raw


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (the BASELINE.json 'rel' tolerance)."""
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-12)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 — the metric for bf16 GRADIENTS: a bf16 forward flips the (Leaky)ReLU mask of
    activations within round-off of zero, which changes individual gradient elements by a whole term
    (max-abs error ~ 1/sqrt(fan)) while the gradient as a vector stays within round-off."""
    return (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
