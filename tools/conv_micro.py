"""Micro-benchmark of the tensor-core conv entry points on the shapes of the training step (b=8, 720x1280) or of
batch-1 inference; prints CUDA-event timings, usable under ncu.  python tools_conv_micro.py [--set train|infer] [--ops fwd,dgrad,wgrad] [--reps 5]"""
import argparse
import os
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rtsds_b200 import ops  # noqa: E402
from rtsds_b200.ops import BF16, F32  # noqa: E402

SHAPES = {
    "train": [(8, 180, 320, 64, 64, 3, 1), (8, 90, 160, 128, 128, 3, 1), (8, 45, 80, 256, 256, 3, 1), (8, 23, 40, 512, 512, 3, 1),
              (8, 360, 640, 64, 128, 3, 2), (8, 180, 320, 128, 256, 3, 2), (8, 90, 160, 1024, 19, 3, 1)],
    "infer": [(1, 128, 256, 64, 64, 3, 1), (1, 64, 128, 128, 128, 3, 1), (1, 32, 64, 256, 256, 3, 1), (1, 16, 32, 512, 512, 3, 1),
              (1, 256, 512, 64, 128, 3, 2), (1, 128, 256, 128, 256, 3, 2), (1, 64, 128, 1024, 19, 3, 1)],
    "deeplab": [(2, 65, 129, 1024, 256, 1, 1), (2, 65, 129, 256, 256, 3, 1, 2), (2, 65, 129, 256, 1024, 1, 1), (2, 65, 129, 2048, 512, 1, 1),
                (2, 65, 129, 512, 512, 3, 1, 4), (2, 65, 129, 512, 2048, 1, 1), (2, 65, 129, 2048, 19, 3, 1, 12)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="train")
    ap.add_argument("--ops", default="fwd,dgrad,wgrad")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    which = args.ops.split(",")
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for si, shp in enumerate(SHAPES[args.set]):
        if args.only >= 0 and si != args.only:
            continue
        n, h, w, cin, cout, k, st = shp[:7]
        dil = shp[7] if len(shp) > 7 else 1
        pad = dil * (k // 2)
        g = torch.Generator().manual_seed(si)
        x = torch.randn(n, h, w, cin, generator=g).to("cuda", torch.bfloat16)
        wt = (torch.randn(cout, cin, k, k, generator=g) * 0.05).cuda()
        out_f32 = cout < 32
        old = 32 if out_f32 else cout
        d = ops.make_conv_desc(n, h, w, cin, cin, cout, old, k, st, pad, dil, in_dtype=BF16, out_dtype=F32 if out_f32 else BF16)
        y = torch.zeros(n, d.oh, d.ow, old, dtype=torch.float32 if out_f32 else torch.bfloat16, device="cuda")
        wpk = ops.pack_conv_weight(wt, BF16)
        wdg = ops.pack_conv_weight_dgrad(wt, BF16, True)
        ck = ops.dgrad_ck(cout, True)
        dy = torch.zeros(n, d.oh, d.ow, ck, dtype=torch.bfloat16, device="cuda")
        dy[..., :cout] = torch.randn(n, d.oh, d.ow, cout, generator=g).to("cuda", torch.bfloat16)
        dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device="cuda")
        dw = torch.zeros(cout * k * k * cin, dtype=torch.float32, device="cuda")
        wsb = max(int(ops.lib().rtsds_conv2d_tc_workspace_bytes(d)), int(ops.lib().rtsds_conv2d_tc_dgrad_workspace_bytes(d)), 16)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        dd = ops.make_conv_desc(n, h, w, cin, cin, cout, ck, k, st, pad, dil, in_dtype=BF16, out_dtype=BF16)
        stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
        fns = {"fwd": lambda: ops.conv2d_tc(d, x, wpk, y, None, None, None, None, ws),
               "fwds": lambda: ops.conv2d_tc(d, x, wpk, y, None, None, None, stats, ws),      # + train-mode BatchNorm sums
               "dgrad": lambda: ops.conv2d_dgrad(dd, dy, wdg, dx, BF16, True, None, ws),
               "wgrad": lambda: ops.conv2d_wgrad(dd, x, dy, dw, True)}
        gf = 2.0 * n * d.oh * d.ow * cout * cin * k * k / 1e9
        res = []
        for name in which:
            fn = fns[name]
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                flush.add_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            t = sorted(ts)[len(ts) // 2]
            res.append(f"{name} {t:7.1f} us {gf / t * 1e3:6.0f} TF/s")
        print(f"[{si}] n{n} {h}x{w} {cin}->{cout} k{k} s{st} d{dil}  {gf:6.1f} GF | " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
