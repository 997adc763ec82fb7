import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from models.bisenet.build_bisenet import BiSeNet
from oracle import weights
m = BiSeNet(19, "resnet18"); m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(0))); m = m.cuda()
x = torch.randn(2, 3, 128, 192).cuda(); y = torch.randint(0, 20, (2, 128, 192)).cuda()
opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
m.train()
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    outs = m(x)
    loss = sum(torch.nn.functional.cross_entropy(t, y, ignore_index=19) for t in outs)
    loss.backward(); opt.step()
from rtsds_b200.bisenet_autograd import bisenet_fused_ce
loss, pred, stats = bisenet_fused_ce(m, x, y, 19); loss.backward()
m.eval(); m.rtsds_cuda_graph = False
with torch.no_grad():
    o = m(x)
torch.cuda.synchronize()
print("done", float(loss), o.shape)
