"""Times the BatchNorm backward kernels at the training step's map sizes (graph-replayed, L2 flushed by size)."""
import os, sys
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtsds_b200._lib import check, lib
from rtsds_b200.ops import BF16

def _p(t):
    return t.data_ptr()

def run(n_pix, c, mode, reps=20):
    raw = torch.randn(n_pix, c, device="cuda").bfloat16()
    dy = torch.randn(n_pix, c, device="cuda").bfloat16()
    y = torch.relu(raw)
    mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
    fsc = torch.ones(c, device="cuda"); fsh = torch.zeros(c, device="cuda")
    sums = torch.empty(2 * c, device="cuda")
    d_raw = torch.empty_like(raw); dgam = torch.zeros(c, device="cuda"); dbet = torch.zeros(c, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def red():
        if mode == "rawmask":
            check(lib().rtsds_bn_bwd_reduce_rawmask(_p(dy), c, _p(raw), c, _p(mean), _p(invstd), _p(fsc), _p(fsh), n_pix, c, BF16, _p(sums), st), "r")
        else:
            check(lib().rtsds_bn_bwd_reduce(_p(dy), c, _p(y), c, _p(raw), c, _p(mean), _p(invstd), n_pix, c, 1, BF16, _p(sums), st), "r")
    def app():
        if mode == "rawmask":
            check(lib().rtsds_bn_bwd_apply_rawmask(_p(dy), c, _p(raw), c, _p(mean), _p(invstd), _p(fsc), _p(sums), _p(fsc), _p(fsh), n_pix, c, BF16, _p(d_raw), c, BF16, None, 0, _p(dgam), _p(dbet), st), "a")
        else:
            check(lib().rtsds_bn_bwd_apply(_p(dy), c, _p(y), c, _p(raw), c, _p(mean), _p(invstd), _p(fsc), _p(sums), n_pix, c, 1, BF16, _p(d_raw), c, BF16, None, 0, _p(dgam), _p(dbet), st), "a")
    out = []
    for fn, nrd, nwr in ((red, 2 + (mode == "y"), 0), (app, 2 + (mode == "y"), 1)):
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        t = ts[len(ts) // 2]
        gb = n_pix * c * 2 * (nrd + nwr) / 1e9
        out.append(f"{t:7.1f} us {gb / t * 1e6:7.0f} GB/s")
    print(f"n_pix={n_pix:9d} c={c:4d} {mode:8s} reduce {out[0]}   apply {out[1]}", flush=True)

if __name__ == "__main__":
    for n_pix, c in ((8 * 360 * 640, 64), (8 * 180 * 320, 64), (8 * 90 * 160, 128), (8 * 45 * 80, 256), (8 * 23 * 40, 512), (8 * 90 * 160, 256)):
        for mode in ("rawmask", "y"):
            run(n_pix, c, mode)
