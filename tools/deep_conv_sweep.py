"""Sweep tile width / split-K of the tensor-core conv on the batch-1 deep layers (layer3 / layer4 of ResNet-18 at
512x1024): warm-L2 time per launch (20 graph-replayed repeats), the configuration the auto-planner picks marked."""
import os
import sys

os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rtsds_b200 import ops  # noqa: E402
from rtsds_b200.ops import F16  # noqa: E402

SHAPES = [(1, 32, 64, 256, 256, 3, 1), (1, 16, 32, 512, 512, 3, 1), (1, 64, 128, 128, 256, 3, 2), (1, 32, 64, 256, 512, 3, 2),
          (1, 64, 128, 128, 128, 3, 1), (1, 128, 256, 64, 64, 3, 1), (1, 128, 256, 128, 256, 3, 2), (1, 64, 128, 1024, 171, 1, 1),
          (1, 256, 512, 64, 128, 3, 2)]


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return sorted(ts)[2]


for n, h, w, cin, cout, k, st in SHAPES:
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, w, cin, generator=gen).to("cuda", torch.float16)
    wt = (torch.randn(cout, cin, k, k, generator=gen) * 0.05).cuda()
    wpk = ops.pack_conv_weight(wt, F16)
    res = []
    for bn in (0, 64, 128):
        for sp in (0, 1, 2, 4, 8):
            ld = (cout + 7) // 8 * 8
            d = ops.make_conv_desc(n, h, w, cin, cin, cout, ld, k, st, k // 2, 1, in_dtype=F16, out_dtype=F16, split_k=sp)
            y = torch.zeros(n, d.oh, d.ow, ld, dtype=torch.float16, device="cuda")
            ops.lib().rtsds_conv2d_tc_tune(bn, 0)
            ws = torch.empty(max(int(ops.lib().rtsds_conv2d_tc_workspace_bytes(d)), 16), dtype=torch.uint8, device="cuda")
            try:
                t = timed(lambda: ops.conv2d_tc(d, x, wpk, y, None, None, None, None, ws))
            except Exception as e:  # noqa: BLE001
                t = float("nan")
            res.append((t, bn, sp))
    ops.lib().rtsds_conv2d_tc_tune(0, 0)
    gf = 2.0 * n * (h // st) * (w // st) * cout * cin * k * k / 1e9
    auto = [r for r in res if r[1] == 0 and r[2] == 0][0][0]
    best = min(r for r in res if r[0] == r[0])
    print(f"{cin}->{cout} k{k} s{st} {h // st}x{w // st}: auto {auto:.2f} us ({gf / auto * 1e3:.0f} TF/s) best {best[0]:.2f} us bn={best[1]} split={best[2]} | " +
          " ".join(f"{bn}/{sp}:{t:.1f}" for t, bn, sp in res), flush=True)
