import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rtsds_b200 import ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for co, ci in ((512, 512), (512, 256), (256, 256), (128, 128), (64, 64)):
    dw = torch.randn(co * 9 * ci, device="cuda")
    g = torch.zeros(co, ci, 3, 3, device="cuda")
    for acc in (True, False):
        us = t(lambda: ops.unpack_conv_wgrad(dw, g, acc))
        print(co, ci, "accumulate" if acc else "assign", "%.1f us" % us, "%.0f GB/s" % (co * ci * 9 * 4 * (4 if acc else 3) / us / 1e3))
