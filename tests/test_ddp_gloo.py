"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU needed)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtsds_b200 import ddp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from models.bisenet.build_bisenet import BiSeNet

        torch.manual_seed(100 + rank)                       # different init per rank on purpose
        m = BiSeNet(19, "resnet18")
        ddp.broadcast_module(m, 0)
        w0 = m.conv.weight.detach().clone()
        named = list(m.named_parameters())
        ranges = ddp.param_buckets(named, ddp.BISENET_GROUPS)
        total = sum(p.numel() for _, p in named)
        flat = torch.full((total,), float(rank + 1))
        red = ddp.BucketedAllReduce(flat, ranges)
        for g in ("head", "layer4", "layer3"):              # signalled during "backward"
            red.ready(g)
        red.ready("head")                                   # idempotent
        out = red.finish()                                  # the rest is issued here
        hist = torch.full((361,), rank + 1, dtype=torch.int64)
        ddp.allreduce_confusion(hist)
        q.put((rank, float(w0.sum()), float(out.min()), float(out.max()), int(hist[0]), sorted(ranges.items())))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_broadcast_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, lo0, hi0, h0, rg0), (r1, w1, lo1, hi1, h1, rg1) = res
    assert w0 == w1                                          # weights broadcast from rank 0
    assert lo0 == hi0 == lo1 == hi1 == 1.5                   # mean of 1 and 2 on every element of every bucket
    assert h0 == h1 == 3
    assert rg0 == rg1
    names = [g for g, _ in rg0]
    assert set(names) == {"spatial", "layer1", "layer2", "layer3", "layer4", "head"}
    spans = sorted(r for _, r in rg0)
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == 12581672


def test_param_buckets_rejects_holes():
    with pytest.raises(ValueError):
        ddp.param_buckets([("a.w", torch.zeros(2)), ("b.w", torch.zeros(2)), ("a.b", torch.zeros(2))], (("a.", "A"), ("b.", "B")))
    with pytest.raises(ValueError):
        ddp.param_buckets([("c.w", torch.zeros(2))], (("a.", "A"),))


def test_deeplab_and_discriminator_buckets_are_contiguous():
    """Bucket layout of the other two trainable models: DeepLabV2 (5 buckets, reverse-topological) and the
    discriminators (one bucket)."""
    from models.deeplabv2.deeplabv2 import get_deeplab_v2
    from rtsds_b200.deeplab_engine import DEEPLAB_GROUPS

    m = get_deeplab_v2(19, pretrain=False)
    named = list(m.named_parameters())
    ranges = ddp.param_buckets(named, DEEPLAB_GROUPS)
    assert set(ranges) == {"layer1", "layer2", "layer3", "layer4", "head"}
    total = sum(p.numel() for _, p in named)
    spans = sorted(ranges.values())
    assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert ranges["head"][1] - ranges["head"][0] == sum(p.numel() for p in m.layer6.parameters())
