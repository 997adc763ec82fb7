import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rtsds_b200 import ops, _lib
from rtsds_b200.ops import BF16
print("lib", _lib.LIB_PATH, os.path.getmtime(_lib.LIB_PATH), os.path.getsize(_lib.LIB_PATH), "env", os.environ.get("RTSDS_DEBUG_WG_NOEPI"))
n, h, w, cin, cout, k = 8, 23, 40, 512, 512, 3
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, w, cin, generator=g).to("cuda", torch.bfloat16)
dy = torch.randn(n, h, w, cout, generator=g).to("cuda", torch.bfloat16)
dw = torch.zeros(cout * 9 * cin, device="cuda")
dd = ops.make_conv_desc(n, h, w, cin, cin, cout, cout, k, 1, 1, 1, in_dtype=BF16, out_dtype=BF16)
wt = (torch.randn(cout, cin, k, k, generator=g) * 0.05).cuda()
wpk = ops.pack_conv_weight(wt, BF16)
y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
ws = torch.empty(max(int(ops.lib().rtsds_conv2d_tc_workspace_bytes(dd)), 16), dtype=torch.uint8, device="cuda")
for name, fn in (("wgrad", lambda: ops.conv2d_wgrad(dd, x, dy, dw, True)), ("fwd", lambda: ops.conv2d_tc(dd, x, wpk, y, None, None, None, None, ws)),
                 ("unpack", lambda: ops.unpack_conv_wgrad(dw, wt, True))):
    fn(); torch.cuda.synchronize()
    N = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(N): fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host issue {1e6*(t1-t0)/N:.1f} us/call, gpu {1e3*e0.elapsed_time(e1)/N:.1f} us/call, wall incl sync {1e6*(t2-t0)/N:.1f} us/call")
