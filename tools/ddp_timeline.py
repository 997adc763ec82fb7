"""Per-bucket timeline of the gradient all-reduce inside one data-parallel training step:
  RTSDS_DDP_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29520 tools/ddp_timeline.py [batch]
Every rank runs BASELINE configs[2]'s step (b=8 720x1280, 3 x CE, FusedAdam); rank 0 prints, for the median of 10 steps:
bucket, size, when its gradients were final, when its average was visible to the compute stream, and when backward ended
(all in ms since backward started)."""
import json
import os
import statistics
import sys

os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")
os.environ["RTSDS_DDP_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench  # noqa: E402
from rtsds_b200 import ddp  # noqa: E402
from rtsds_b200.bisenet_autograd import bisenet_fused_ce  # noqa: E402
from rtsds_b200.optim import FusedAdam  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", local)
model = bench.make_model(dev).train()
model.rtsds_ddp = True
ddp.broadcast_module(model, 0)
opt = FusedAdam(model.parameters(), lr=1e-4)
g = torch.Generator().manual_seed(42 + rank)
x = torch.randn(batch, 3, 720, 1280, generator=g).to(dev)
y = torch.randint(0, 20, (batch, 720, 1280), generator=g).to(dev)
runs = []
for i in range(15):
    opt.zero_grad(set_to_none=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loss, _, _ = bisenet_fused_ce(model, x, y, 19)
    loss.backward()
    opt.step()
    b.record()
    torch.cuda.synchronize()
    if i >= 5:
        rows, end = ddp.last_timeline()
        runs.append((rows, end, a.elapsed_time(b)))
if rank == 0:
    med = lambda v: round(statistics.median(v), 3)  # noqa: E731
    out = {"n_gpus": world, "per_gpu_batch": batch, "step_ms": med([r[2] for r in runs]), "backward_end_ms": med([r[1] for r in runs]),
           "buckets": [{"bucket": runs[0][0][k][0], "MB": round(runs[0][0][k][1], 2), "ready_ms": med([r[0][k][2] for r in runs]),
                        "averaged_ms": med([r[0][k][3] for r in runs])} for k in range(len(runs[0][0]))],
           "note": "ms since backward started on rank 0; averaged_ms = when the compute stream could first see the bucket's average "
                   "(waits are issued in bucket order after backward ends, so it is >= backward_end_ms by construction); the "
                   "exposed communication is max(averaged_ms) - backward_end_ms"}
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
