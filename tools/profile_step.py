"""In-situ per-entry-point GPU timing of one training step (CUDA events around every C-ABI call; valid when the
step is GPU-bound so that the CPU runs ahead of the device).  python tools_profile_step.py [--batch 8] [--workload train|adversarial|deeplab]"""
import argparse
import collections
import os
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rtsds_b200 import _lib, ops  # noqa: E402


class Proxy:
    def __init__(self, real):
        self.real, self.records, self.on = real, [], False

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if not self.on or name in ("rtsds_last_error_string", "rtsds_launch_count", "rtsds_conv_cout_pad", "rtsds_check_device",
                                   "rtsds_conv2d_tc_workspace_bytes", "rtsds_conv2d_tc_dgrad_workspace_bytes", "rtsds_resize_ce_fused_supported"):
            return fn

        def wrapped(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            tag = name
            if name.startswith("rtsds_conv2d_tc"):
                d = a[0]._obj if hasattr(a[0], "_obj") else None
                if d is not None:
                    tag = f"{name} {d.cin}->{d.cout} k{d.kh} s{d.stride} d{d.dil} {d.h}x{d.w}"
            self.records.append((tag, e0, e1))
            return r

        return wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--workload", default="train")
    ap.add_argument("--detail", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    real = _lib.lib()
    proxy = Proxy(real)
    _lib._lib = proxy
    import bench
    from rtsds_b200.bisenet_autograd import bisenet_fused_ce

    g = torch.Generator().manual_seed(1)
    if args.workload == "train":
        model = bench.make_model(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        x = torch.randn(args.batch, 3, 720, 1280, generator=g).to(dev)
        y = torch.randint(0, 20, (args.batch, 720, 1280), generator=g).to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            loss, _, _ = bisenet_fused_ce(model, x, y, 19)
            loss.backward()
            opt.step()
    else:
        from models.deeplabv2.deeplabv2 import get_deeplab_v2
        from rtsds_b200.deeplab_engine import deeplab_fused_ce

        model = get_deeplab_v2(19, pretrain=False)
        for mod in model.modules():
            if isinstance(mod, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(mod.weight, mode="fan_in", nonlinearity="relu")
        model = model.to(dev).train()
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-3)
        x = torch.randn(args.batch, 3, 512, 1024, generator=g).to(dev)
        y = torch.randint(0, 20, (args.batch, 512, 1024), generator=g).to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            loss, _, _ = deeplab_fused_ce(model, x, y, 19)
            loss.backward()
            opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    proxy.on = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    proxy.on = False
    total = e0.elapsed_time(e1) * 1e3
    agg = collections.defaultdict(lambda: [0, 0.0])
    for tag, a, b in proxy.records:
        key = tag if args.detail else tag.split(" ")[0]
        agg[key][0] += 1
        agg[key][1] += a.elapsed_time(b) * 1e3
    s = sum(v[1] for v in agg.values())
    print(f"step {total:.0f} us (with event overhead); inside C-ABI calls {s:.0f} us; calls {len(proxy.records)}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{v[1]:9.1f} us {v[0]:4d} {100 * v[1] / total:5.1f}%  {k}")


if __name__ == "__main__":
    main()
