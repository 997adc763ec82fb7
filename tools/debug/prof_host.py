import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from rtsds_b200.serving import PipelinedSegmenter
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
host = [torch.randn(1, 3, 512, 1024).pin_memory() for _ in range(8)]
pipe = PipelinedSegmenter(model, 1, 512, 1024, depth=4, lanes=2)
for i in range(50):
    pipe.submit(host[i % 8])
pipe.drain()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(2000):
    pipe.submit(host[i % 8])
pipe.drain()
print("wall per frame us", (time.perf_counter() - t0) / 2000 * 1e6)
pr = cProfile.Profile()
pr.enable()
for i in range(2000):
    pipe.submit(host[i % 8])
pipe.drain()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
