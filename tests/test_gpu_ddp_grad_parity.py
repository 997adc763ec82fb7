"""SURVEY §4 "distributed" / §8(e): two data-parallel ranks, per-rank BatchNorm (nn.DataParallel semantics,
utils.py:104-105), gradients exchanged through the real collective path (rtsds_b200/ddp.py bucketed all-reduce overlapped
with the hand-written backward).  After backward every rank must hold

    local loss normalisation   :  mean over ranks of the oracle's per-shard gradients
    rtsds_loss_norm = "global" :  gradient of sum over ranks of sum(-log p) / valid pixels of ALL ranks  (what the
                                  reference's DataParallel computes on the gathered batch)

with the oracle (oracle/bisenet_ref.py, CPU autograd) run shard by shard.  Two processes: NCCL on two GPUs when the box has
them (gpurun --gpus 2), otherwise both ranks share cuda:0 over gloo — the same BucketedAllReduce code either way."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_PER, H, W = 2, 128, 192


def _data():
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2 * N_PER, 3, H, W, generator=g)
    y = torch.randint(0, 20, (2 * N_PER, H, W), generator=g)
    y[N_PER:, : H // 3] = 19                    # rank 1 sees far fewer valid pixels: the two normalisations differ
    return x, y


def _oracle(x, y, sd0, mode, skip):
    """Expected gradient on every rank after the all-reduce, for the given loss normalisation."""
    from oracle import bisenet_ref, weights

    total, n_valid = None, float((y != 19).sum())
    for r in range(2):
        sd = weights.clone_state(sd0)
        leaves = {k: v.requires_grad_(True) for k, v in sd.items()
                  if v.dtype.is_floating_point and "running" not in k and k.startswith("context_path.features.") | (not k.startswith("context_path."))}
        xs, ys = x[r * N_PER:(r + 1) * N_PER], y[r * N_PER:(r + 1) * N_PER]
        outs = bisenet_ref.bisenet_forward(xs, sd, train=True)
        if mode == "local":
            loss = sum(bisenet_ref.ce_loss(t, ys, 19) for t in outs) / 2.0                     # AVG over the two ranks
        else:
            loss = sum(torch.nn.functional.cross_entropy(t, ys, ignore_index=19, reduction="sum") for t in outs) / n_valid
        loss.backward()
        g = {k: v.grad for k, v in leaves.items() if v.grad is not None and k not in skip}
        total = g if total is None else {k: total[k] + g[k] for k in g}
    return total


def _worker(rank, world, port, ngpu, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RTSDS_ALLOW_RANDOM_INIT="1")
        import torch.distributed as dist

        dev = rank if ngpu >= 2 else 0
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl" if ngpu >= 2 else "gloo", rank=rank, world_size=world)
        from models.bisenet.build_bisenet import BiSeNet
        from oracle import weights
        from rtsds_b200 import ddp
        from rtsds_b200.bisenet_autograd import bisenet_fused_ce

        torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
        x, y = _data()
        sd0 = weights.bisenet_r18_state(21)
        m = BiSeNet(19, "resnet18")
        m.load_state_dict(weights.clone_state(sd0))
        m.rtsds_precision = "fp32"
        m = m.cuda().train()
        m.rtsds_ddp = True
        ddp.broadcast_module(m, 0)
        xs, ys = x[rank * N_PER:(rank + 1) * N_PER].cuda(), y[rank * N_PER:(rank + 1) * N_PER].cuda()
        got = {}
        for mode in ("local", "global", "stock"):
            m.load_state_dict(weights.clone_state(sd0))          # same running statistics for every pass
            m.rtsds_ddp = True
            m.rtsds_loss_norm = "global" if mode == "global" else "local"
            m.zero_grad(set_to_none=True)
            if mode == "stock":                                  # the reference's call site: criterion on the returned logits
                outs = m(xs)
                loss = sum(torch.nn.functional.cross_entropy(t, ys, ignore_index=19) for t in outs)
            else:
                loss, _, _ = bisenet_fused_ce(m, xs, ys, 19)
            loss.backward()
            torch.cuda.synchronize()
            got[mode] = ({k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}, float(loss))
        # every rank holds the same reduced gradient
        for mode in got:
            chk = torch.tensor([sum(float(g.double().sum()) for g in got[mode][0].values())], dtype=torch.float64, device="cuda")
            both = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(both, chk)
            assert abs(float(both[0]) - float(both[1])) <= 1e-9 * max(1.0, abs(float(both[0]))), (mode, both)
        res = {"backend": dist.get_backend(), "ngpu": ngpu}
        if rank == 0:
            skip = {"attention_refinement_module1.conv.bias", "attention_refinement_module2.conv.bias"}   # analytically zero
            for mode, ref_mode in (("local", "local"), ("stock", "local"), ("global", "global")):
                ref = _oracle(x, y, sd0, ref_mode, skip)
                worst = ("", 0.0)
                for k, rg in ref.items():
                    e = float((got[mode][0][k].double() - rg.double()).norm() / rg.double().norm().clamp_min(1e-30))
                    if e > worst[1]:
                        worst = (k, e)
                res[mode] = {"worst_param": worst[0], "worst_rel_l2": worst[1], "loss": got[mode][1]}
            # the two normalisations really differ on this data (otherwise the global test proves nothing)
            d = max(float((got["local"][0][k] - got["global"][0][k]).abs().max()) for k in got["local"][0])
            res["local_vs_global_max_abs_diff"] = d
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", res))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, "error", traceback.format_exc()[-3000:]))


def test_two_rank_gradients_match_the_per_shard_oracle(cuda):
    from parity_log import record

    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ngpu, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank, status, payload in out:
        assert status == "ok", f"rank {rank}:\\n{payload}"
    res = out[0][2]
    record(f"ddp/two_rank_gradient_parity/{res['backend']}", **{k: (v if not isinstance(v, dict) else str(v)) for k, v in res.items()})
    for mode in ("local", "stock", "global"):
        # fp32 check mode at 128x192: 48-sample layer4 / 2-sample ARM BatchNorm statistics amplify the fp32 atomics-order
        # noise of the statistics; 5e-3 .. 2e-2 observed on single BatchNorm weights, everything else <= 5e-3
        assert res[mode]["worst_rel_l2"] <= 4e-2, (mode, res[mode])
    assert res["local_vs_global_max_abs_diff"] > 1e-6
