import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rtsds_b200 import ops
from rtsds_b200.ops import BF16, F32
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n, h, w in ((1, 512, 1024), (8, 720, 1280)):
    x = torch.randn(n, 3, h, w, device="cuda")
    w7 = torch.randn(64, 3, 7, 7, device="cuda") * 0.1
    w3 = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
    oh, ow, pshape = ops.stem_s2d_shape(n, h, w)
    P = torch.empty(pshape, dtype=torch.bfloat16, device="cuda")
    w2 = torch.empty(64, 64, 4, 1, device="cuda"); ops.stem_s2d_weight(w7, w2); wpk = ops.pack_conv_weight(w2, BF16)
    y = torch.empty(n, oh, ow, 64, dtype=torch.bfloat16, device="cuda")
    dy = torch.randn(n, oh, ow, 64, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(64 * 4 * 64, device="cuda")
    stats = torch.zeros(128, device="cuda")
    t_pack = timeit(lambda: ops.stem_s2d_pack(x, P))
    t_conv = timeit(lambda: ops.stem_s2d_conv_fwd(P, n, oh, ow, wpk, 64, y, 64, BF16, stats=stats))
    t_wg = timeit(lambda: ops.stem_s2d_conv_wgrad(P, n, oh, ow, dy, 64, 64, dw))
    wpair = ops.stem_pack_weights(w7, w3)
    ycp = torch.empty_like(y); ysp = torch.empty_like(y)
    t_pair = timeit(lambda: ops.stem_pair_tc_fwd(x, wpair, ycp, ysp, None, None, False, stats[:128], stats[:128]))
    dws = torch.zeros(128 * 192, device="cuda"); g7 = torch.zeros_like(w7); g3 = torch.zeros_like(w3)
    t_pairwg = timeit(lambda: ops.stem_pair_tc_wgrad(x, dy, dy, dws, g7, g3))
    print(f"n={n} {h}x{w}: s2d pack {t_pack:.1f} us, conv fwd {t_conv:.1f} us (x2 stems), wgrad {t_wg:.1f} us (x2) | fused pair fwd {t_pair:.1f} us, wgrad {t_pairwg:.1f} us")
