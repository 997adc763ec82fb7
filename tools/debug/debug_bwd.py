"""Diagnostic: per-parameter gradient comparison between precision modes, in backward order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import weights
from models.bisenet.build_bisenet import BiSeNet
from rtsds_b200.bisenet_autograd import bisenet_fused_ce

def run(prec, x, y, seed):
    m = BiSeNet(19, "resnet18"); m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(seed))); m.rtsds_precision = prec
    m = m.cuda().train()
    loss, _, _ = bisenet_fused_ce(m, x.cuda(), y.cuda(), 19)
    loss.backward()
    return loss.item(), {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}

g = torch.Generator().manual_seed(1003)
n, h, w = 2, 128, 256
x = torch.randn(n, 3, h, w, generator=g); y = torch.randint(0, 20, (n, h, w), generator=g)
a = sys.argv[1] if len(sys.argv) > 1 else "bf16"
b = sys.argv[2] if len(sys.argv) > 2 else "bf16_simt"
la, ga = run(a, x, y, 3); lb, gb = run(b, x, y, 3)
print("loss", a, la, b, lb)
names = list(ga.keys())[::-1]
for k in names:
    e = ((ga[k].double() - gb[k].double()).norm() / gb[k].double().norm().clamp_min(1e-30)).item()
    print(f"{e:10.4f}  |{a}|={ga[k].norm().item():.3e} |{b}|={gb[k].norm().item():.3e}  {k}")
