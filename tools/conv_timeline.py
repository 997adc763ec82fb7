"""Where the time of one batch-1 tensor-core conv launch goes: per-CTA time stamps written by conv_tc_kernel itself
(rtsds_debug_conv_trace), for the ResNet-18 layer shapes at 512x1024.

  python tools/conv_timeline.py            # prints one line per shape: mean over CTAs of every phase, in microseconds

Each shape is launched 4 times back to back (programmatic dependent launch between them, as inside the frame's CUDA
graph); the stamps of the LAST launch are read.  clock64 deltas are per-SM exact; `span` is max(exit) - min(entry) over
all CTAs on the global timer (1 us granularity unless a profiler raised it)."""
import ctypes
import os
import sys

os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rtsds_b200 import ops  # noqa: E402
from rtsds_b200.ops import F16  # noqa: E402

SHAPES = [(1, 128, 256, 64, 64, 3, 1), (1, 64, 128, 128, 128, 3, 1), (1, 32, 64, 256, 256, 3, 1), (1, 16, 32, 512, 512, 3, 1),
          (1, 64, 128, 128, 256, 3, 2), (1, 32, 64, 256, 512, 3, 2)]
MHZ = 1965.0

trace = torch.zeros(32 * 4096, dtype=torch.int64, device="cuda")
for n, h, w, cin, cout, k, st in SHAPES:
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, w, cin, generator=gen).to("cuda", torch.float16)
    wt = (torch.randn(cout, cin, k, k, generator=gen) * 0.05).cuda()
    wpk = ops.pack_conv_weight(wt, F16)
    # as the second conv of a BasicBlock runs it: folded BatchNorm, residual, ReLU
    d = ops.make_conv_desc(n, h, w, cin, cin, cout, cout, k, st, 1, 1, act=ops.ACT_RELU, in_dtype=F16, out_dtype=F16, res_ld=cout)
    y = torch.zeros(n, d.oh, d.ow, cout, dtype=torch.float16, device="cuda")
    res = torch.randn(n, d.oh, d.ow, cout, generator=gen).to("cuda", torch.float16)
    sc, sh = torch.rand(cout, generator=gen).cuda() + 0.5, torch.randn(cout, generator=gen).cuda()
    ws = torch.empty(max(int(ops.lib().rtsds_conv2d_tc_workspace_bytes(d)), 16), dtype=torch.uint8, device="cuda")
    run = lambda: ops.conv2d_tc(d, x, wpk, y, sc, sh, res, None, ws)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    trace.zero_()
    ops.lib().rtsds_debug_conv_trace(ctypes.c_void_p(trace.data_ptr()))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            run()
    g.replay()
    torch.cuda.synchronize()
    ops.lib().rtsds_debug_conv_trace(None)
    t = trace.view(-1, 32).cpu()
    t = t[t[:, 1] != 0].double()
    c = lambda a, b: ((t[:, b] - t[:, a]) / MHZ).mean().item()  # noqa: E731
    span = (t[:, 10].max() - t[:, 0].min()).item() / 1e3
    print(f"{cin}->{cout} k{k} s{st} {h // st}x{w // st}: ctas {t.shape[0]:4d} | setup {c(1, 2):5.2f} dep-wait {c(2, 3):5.2f} "
          f"first-operands {c(3, 5):5.2f} mma-loop {c(5, 6):5.2f} (tma-issue-done {c(3, 4):5.2f}) acc-complete {c(6, 7):5.2f} "
          f"epilogue {c(7, 8):5.2f} [tmem-ld0 {c(7, 13):4.2f} chunk0 {c(13, 14):4.2f} (scale/shift {c(13, 16):4.2f} act {c(16, 17):4.2f} pack+store {c(17, 14):4.2f}) tmem-ld1 {c(14, 15):4.2f} rest {c(15, 8):4.2f}] reduce/exit {c(8, 9):5.2f}" + (f" [sync1 {c(8, 11):4.2f} sum+store {c(11, 12):4.2f} sync2 {c(12, 9):4.2f}]" if t[:, 11].max() > 0 else "") + f" | cta total {c(1, 9):5.2f} us, after dep {c(3, 9):5.2f} us, span {span:5.1f} us",
          flush=True)
