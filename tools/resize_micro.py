"""Times rtsds_resize_to_nchw_bwd (adjoint of the logits writer) at the training sizes."""
import os, sys
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtsds_b200._lib import check, lib

for n, c, h, w, oh, ow in ((8, 19, 90, 160, 720, 1280), (4, 19, 64, 128, 512, 1024)):
    dout = torch.randn(n, c, oh, ow, device="cuda")
    dz = torch.empty(n, h, w, 32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().rtsds_resize_to_nchw_bwd(dout.data_ptr(), n, c, oh, ow, h, w, dz.data_ptr(), 32, st), "r")
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    gb = dout.numel() * 4 / 1e9
    print(f"{n}x{c}x{oh}x{ow}: {ts[5]:8.1f} us  {gb / ts[5] * 1e6:7.0f} GB/s", flush=True)
