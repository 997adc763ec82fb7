"""Phase breakdown of arm_gate_resize_kernel at batch 1 (per-block clock64 stamps, rtsds_debug_arm_trace)."""
import ctypes
import os
import sys

os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from models.bisenet.build_bisenet import BiSeNet  # noqa: E402
from rtsds_b200 import ops  # noqa: E402
from rtsds_b200.ops import F16  # noqa: E402

m = BiSeNet(19, "resnet18").cuda().eval()
a1, a2 = m.attention_refinement_module1, m.attention_refinement_module2
g = torch.Generator().manual_seed(0)
n = 1
f3 = torch.randn(n, 32, 64, 256, generator=g).to("cuda", torch.float16)
f4 = torch.randn(n, 16, 32, 512, generator=g).to("cuda", torch.float16)
cat = torch.empty(n, 64, 128, 1024, dtype=torch.float16, device="cuda")
for parts3, parts4 in ((1, 1), (32, 16)):
    p3, p4 = torch.randn(n, parts3, 256, generator=g).cuda(), torch.randn(n, parts4, 512, generator=g).cuda()
    run = lambda: ops.arm_gate_resize(ops.arm_side(f3, p3, a1, 32, 64, 256, 256, parts=parts3),  # noqa: E731
                                      ops.arm_side(f4, p4, a2, 16, 32, 512, 512, mul_pooled=True, parts=parts4), F16, n, 64, 128, cat, 1024)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    trace = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    ops.lib().rtsds_debug_arm_trace(ctypes.c_void_p(trace.data_ptr()))
    run()
    torch.cuda.synchronize()
    ops.lib().rtsds_debug_arm_trace(None)
    t = trace.view(-1, 8).cpu()
    t = t[t[:, 0] != 0].double()
    c = lambda a, b, rows=slice(None): ((t[rows, b] - t[rows, a]) / 1965.0).mean().item()  # noqa: E731
    nb = t.shape[0]
    st = (t[:, 4] - t[:, 4].min()) / 1e3
    en = (t[:, 5] - t[:, 4].min()) / 1e3
    print(f"block start times (us after the first): median {st.median():.1f} p90 {st.quantile(0.9):.1f} max {st.max():.1f}; last end {en.max():.1f}")
    b0 = nb * 8 // 24
    for name, rows in (("ARM1 blocks (c=256)", slice(0, b0)), ("ARM2 blocks (c=512)", slice(b0, nb))):
        print(f"parts {parts3}/{parts4} {name}: {t[rows].shape[0]} blocks | pooled {c(0, 1, rows):5.2f} gates {c(1, 2, rows):5.2f} stream {c(2, 3, rows):5.2f} total {c(0, 3, rows):5.2f} us")

# ---- ffm_head_resize: same stamps (start, attention + weights ready, z rows ready, rows written)
ffm = m.feature_fusion_module
feat = torch.randn(n, 64, 128, 32, generator=g).cuda()
out = torch.empty(n, 19, 512, 1024, dtype=torch.float32, device="cuda")
for parts in (1, 64):
    pf = torch.randn(n, parts, 19, generator=g).cuda()
    run = lambda: ops.ffm_head_resize(feat, 32, pf, n, 64, 128, 19, ffm.conv1, ffm.conv2, m.conv, out, parts=parts)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    trace = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    ops.lib().rtsds_debug_arm_trace(ctypes.c_void_p(trace.data_ptr()))
    run()
    torch.cuda.synchronize()
    ops.lib().rtsds_debug_arm_trace(None)
    t = trace.view(-1, 8).cpu()
    t = t[t[:, 0] != 0].double()
    c = lambda a, b: ((t[:, b] - t[:, a]) / 1965.0).mean().item()  # noqa: E731
    print(f"ffm_head_resize parts {parts}: {t.shape[0]} blocks | prologue {c(0, 1):5.2f} z-rows {c(1, 2):5.2f} write {c(2, 3):5.2f} total {c(0, 3):5.2f} us")
