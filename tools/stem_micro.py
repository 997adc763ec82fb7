"""Times the fused stem kernels (forward + weight gradient) at the training and inference sizes."""
import os, sys
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtsds_b200 import ops

def run(n, h, w, reps=3):
    x = torch.randn(n, 3, h, w, device="cuda")
    w7 = torch.randn(64, 3, 7, 7, device="cuda") * 0.1
    w3 = torch.randn(64, 3, 3, 3, device="cuda") * 0.2
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    wpk = ops.stem_pack_weights(w7, w3)
    ycp = torch.empty(n, oh, ow, 64, dtype=torch.bfloat16, device="cuda")
    ysp = torch.empty_like(ycp)
    st7 = torch.zeros(128, device="cuda"); st3 = torch.zeros(128, device="cuda")
    ws = torch.zeros(128 * 192, device="cuda")
    g7 = torch.zeros(64, 3, 7, 7, device="cuda"); g3 = torch.zeros(64, 3, 3, 3, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fn in (("fwd", lambda: ops.stem_pair_tc_fwd(x, wpk, ycp, ysp, None, None, False, st7, st3)),
                     ("wgrad", lambda: ops.stem_pair_tc_wgrad(x, ycp, ysp, ws, g7, g3))):
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"n={n} {h}x{w} {name:6s} {ts[len(ts) // 2]:8.1f} us", flush=True)

if __name__ == "__main__":
    run(8, 720, 1280)
    if len(sys.argv) < 2: run(1, 512, 1024)
