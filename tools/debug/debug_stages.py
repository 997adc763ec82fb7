"""Diagnostic (not collected by pytest): compare the bf16 CUDA plan's intermediate buffers with the
bf16-emulating oracle stage by stage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import bisenet_bf16, bisenet_ref, weights
from models.bisenet.build_bisenet import BiSeNet
from gpu_util import nchw, rel_err

seed, n, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 0, 2, 64, 96
if len(sys.argv) > 4: n, h, w = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sd = weights.bisenet_r18_state(seed)
x = torch.randn(n, 3, h, w, generator=torch.Generator().manual_seed(1000 + seed))
with torch.no_grad():
    emu = bisenet_bf16.bisenet_eval_bf16(x, sd, True)
    ref = bisenet_ref.bisenet_forward(x, weights.clone_state(sd), False)
m = BiSeNet(19, "resnet18"); m.load_state_dict(weights.clone_state(sd)); m = m.cuda().eval(); m.rtsds_cuda_graph = False
out = m(x.cuda())
plan = next(iter(m._rtsds_plans.values()))
got = dict(sx=nchw(plan.cat[..., :256]), cat=nchw(plan.cat), f3=nchw(plan.f3), f4=nchw(plan.f4), feat=nchw(plan.feat[..., :19]),
           z=nchw(plan.z[..., :19]), out=out.cpu())
for k in ("sx", "f3", "f4", "cat", "feat", "z", "out"):
    print(f"{k:5s} cuda-vs-bf16emu rel={rel_err(got[k], emu[k]):.3e}")
print("out   cuda-vs-fp32 rel=%.3e   emu-vs-fp32 rel=%.3e" % (rel_err(got["out"], ref), rel_err(emu["out"], ref)))
print("argmax agree cuda/fp32 %.5f  emu/fp32 %.5f  cuda/emu %.5f" % ((got["out"].argmax(1) == ref.argmax(1)).float().mean(),
      (emu["out"].argmax(1) == ref.argmax(1)).float().mean(), (got["out"].argmax(1) == emu["out"].argmax(1)).float().mean()))
