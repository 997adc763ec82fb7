"""Diagnostic for the tcgen05 conv kernel (not collected by pytest): structured operands that
reveal row / K-permutation problems in the TMA->smem->UMMA descriptor chain."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from rtsds_b200 import ops
from rtsds_b200.ops import F32
from gpu_util import run_conv, conv_ref, rel_err

torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)

def show(tag, y, ref):
    e = rel_err(y, ref)
    print(f"[{tag}] rel_err={e:.3e} nan={torch.isnan(y).sum().item()}")
    if not (e < 1e-3):
        yy = y[0].reshape(y.shape[1], -1).t()      # [pixels, cout]
        rr = ref[0].reshape(ref.shape[1], -1).t()
        print(" got rows 0..3, cols 0..15:\n", yy[:4, :16])
        print(" ref rows 0..3, cols 0..15:\n", rr[:4, :16])
        print(" got rows 8..9:\n", yy[8:10, :16]); print(" ref rows 8..9:\n", rr[8:10, :16])
        bad = ((yy - rr).abs() > 1e-2 * rr.abs().max()).nonzero()
        print(" #bad", len(bad), "first bad idx", bad[:8].tolist())

# 1) column identity: X[p,k] = k ; W = I(32x64) -> Y[p,co] = co
x = torch.arange(64.).view(1, 64, 1, 1).expand(1, 64, 8, 16).contiguous()
w = torch.zeros(32, 64, 1, 1); w[torch.arange(32), torch.arange(32)] = 1
y, _, _ = run_conv("tc", x, w, pad=0, out_dtype=F32, out_ld=32); show("K-identity", y, conv_ref(x, w, pad=0)[0])
# 2) row identity: X[p,k] = p for all k; W picks k=co
x = torch.arange(128.).view(1, 1, 8, 16).expand(1, 64, 8, 16).contiguous()
y, _, _ = run_conv("tc", x, w, pad=0, out_dtype=F32, out_ld=32); show("row-identity", y, conv_ref(x, w, pad=0)[0])
# 3) random GEMM
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 64, 8, 16, generator=g); w = torch.randn(32, 64, 1, 1, generator=g)
y, _, _ = run_conv("tc", x, w, pad=0, out_dtype=F32, out_ld=32); show("random-gemm", y, conv_ref(x, w, pad=0)[0])
# 4) two k-blocks, cout 128
x = torch.randn(1, 128, 8, 16, generator=g); w = torch.randn(128, 128, 1, 1, generator=g)
y, _, _ = run_conv("tc", x, w, pad=0, out_dtype=F32, out_ld=128); show("2kblk-n128", y, conv_ref(x, w, pad=0)[0])
# 5) 3x3 pad 1 (TMA OOB fill)
x = torch.randn(1, 64, 8, 16, generator=g); w = torch.randn(64, 64, 3, 3, generator=g)
y, _, _ = run_conv("tc", x, w, pad=1, out_dtype=F32, out_ld=64); show("3x3", y, conv_ref(x, w, pad=1)[0])
# 6) stride 2 (parity maps)
x = torch.randn(1, 64, 16, 32, generator=g)
y, _, _ = run_conv("tc", x, w, stride=2, pad=1, out_dtype=F32, out_ld=64); show("3x3s2", y, conv_ref(x, w, stride=2, pad=1)[0])
print("debug_tc done")
