"""Per-phase time stamps of the conv epilogue under a MULTI-WAVE training-shaped launch (non-persistent kernel,
RTSDS_NO_PERSISTENT=1 so that conv_tc_kernel — the traced one — runs): where a 32-column epilogue chunk spends its time."""
import ctypes
import os
import sys

os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")
os.environ["RTSDS_NO_PERSISTENT"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rtsds_b200 import ops  # noqa: E402
from rtsds_b200.ops import BF16  # noqa: E402

SHAPES = [(2, 65, 129, 256, 1024, 1, 1), (2, 65, 129, 1024, 256, 1, 1), (8, 90, 160, 128, 128, 3, 1), (8, 45, 80, 256, 256, 3, 1)]
MHZ = 1965.0
trace = torch.zeros(32 * 8192, dtype=torch.int64, device="cuda")
for n, h, w, cin, cout, k, st in SHAPES:
    for with_stats in (False, True):
        gen = torch.Generator().manual_seed(0)
        x = torch.randn(n, h, w, cin, generator=gen).to("cuda", torch.bfloat16)
        wt = (torch.randn(cout, cin, k, k, generator=gen) * 0.05).cuda()
        wpk = ops.pack_conv_weight(wt, BF16)
        d = ops.make_conv_desc(n, h, w, cin, cin, cout, cout, k, st, k // 2, 1, in_dtype=BF16, out_dtype=BF16)
        y = torch.zeros(n, d.oh, d.ow, cout, dtype=torch.bfloat16, device="cuda")
        stats = torch.zeros(2 * cout, device="cuda") if with_stats else None
        ws = torch.empty(max(int(ops.lib().rtsds_conv2d_tc_workspace_bytes(d)), 16), dtype=torch.uint8, device="cuda")
        run = lambda: ops.conv2d_tc(d, x, wpk, y, None, None, None, stats, ws)  # noqa: E731
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        trace.zero_()
        ops.lib().rtsds_debug_conv_trace(ctypes.c_void_p(trace.data_ptr()))
        run()
        torch.cuda.synchronize()
        ops.lib().rtsds_debug_conv_trace(None)
        t = trace.view(-1, 32).cpu()
        t = t[t[:, 1] != 0].double()
        c = lambda a, b: ((t[:, b] - t[:, a]) / MHZ).mean().item()  # noqa: E731
        span = (t[:, 10].max() - t[:, 0].min()).item() / 1e3
        print(f"{cin}->{cout} k{k} {n}x{h}x{w} stats={int(with_stats)}: ctas {t.shape[0]:4d} | setup {c(1, 2):5.2f} dep {c(2, 3):5.2f} first-operands {c(3, 5):5.2f} "
              f"mma-loop {c(5, 6):5.2f} acc-complete {c(6, 7):5.2f} epilogue {c(7, 8):5.2f} [tmem-ld0 {c(7, 13):4.2f} chunk0 {c(13, 14):4.2f} "
              f"(stats+scale/shift {c(13, 16):4.2f} act {c(16, 17):4.2f} pack+store {c(17, 14):4.2f}) tmem-ld1 {c(14, 15):4.2f} rest {c(15, 8):4.2f}] exit {c(8, 9):5.2f} "
              f"| cta total {c(1, 9):5.2f} us, span {span:5.1f} us", flush=True)
