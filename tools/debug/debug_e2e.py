import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench, bench_train
from rtsds_b200.bisenet_autograd import bisenet_fused_ce
from rtsds_b200.serving import AsyncScalarReader, DevicePrefetcher
dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
b, n_sets, steps = 8, 4, 12
g = torch.Generator().manual_seed(0)
host_x = torch.randn(n_sets, b, 3, 720, 1280, generator=g).pin_memory()
host_y = torch.randint(0, 20, (n_sets, b, 720, 1280), generator=g).pin_memory()
dev_x, dev_y = host_x.to(dev), host_y.to(dev)
def step(x, y):
    opt.zero_grad(set_to_none=True)
    loss, pred, stats = bisenet_fused_ce(model, x, y, 19)
    loss.backward(); opt.step()
    return loss
for i in range(3): step(dev_x[i], dev_y[i])
torch.cuda.synchronize()
def run(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{name}: {1e3*(time.perf_counter()-t)/steps:.2f} ms/step", flush=True)
def a():
    for i in range(steps): step(dev_x[i % n_sets], dev_y[i % n_sets])
def b_():
    for i in range(steps): step(dev_x[i % n_sets], dev_y[i % n_sets]).item()
def c():
    r = AsyncScalarReader(dev)
    for i in range(steps): r.push(step(dev_x[i % n_sets], dev_y[i % n_sets]))
    r.drain()
def d():
    for sx, sy in DevicePrefetcher(((host_x[i % n_sets], host_y[i % n_sets]) for i in range(steps)), dev): step(sx, sy)
def e():
    r = AsyncScalarReader(dev)
    for sx, sy in DevicePrefetcher(((host_x[i % n_sets], host_y[i % n_sets]) for i in range(steps)), dev): r.push(step(sx, sy))
    r.drain()
def f():
    sx, sy = torch.empty_like(dev_x[0]), torch.empty_like(dev_y[0])
    for i in range(steps):
        sx.copy_(host_x[i % n_sets], non_blocking=True); sy.copy_(host_y[i % n_sets], non_blocking=True)
        step(sx, sy)
for name, fn in (("device only", a), ("device + item", b_), ("device + async reader", c), ("prefetch", d), ("prefetch + async reader", e), ("serial copies", f), ("prefetch again", d)):
    run(name, fn)
