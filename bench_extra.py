"""Secondary workloads of bench.py (BASELINE.json configs[3] and configs[4]):

  --workload adversarial   one iteration of train.py:adversarial_train per step: BiSeNet-R18 generator on a synthetic
                           GTA5-shaped 720x1280 source batch (3 x CE) and a Cityscapes-shaped 512x1024 target batch
                           (adversarial BCE through the frozen discriminator), discriminator on both detached
                           predictions, Adam on both; data-parallel over the ranks.
  --workload deeplab       DeepLabV2-ResNet101 (dilated, output stride 8) supervised training at 512x1024, batch 2 per
                           GPU, plus batch-1 eval FPS.
Same timing rules as the main workloads: W warm-up steps, K timed steps between barriers, CUDA events, max over ranks,
inputs rotating over sets larger than L2, clocks sampled during the timed region."""
from __future__ import annotations

import json

import torch

G_TRAIN_GFLOP_720 = 267.5
G_TRAIN_GFLOP_512 = 150.9
DEEPLAB_TRAIN_GFLOP_512 = 2090.0
DEEPLAB_FWD_GFLOP_512 = 747.3


def _line(metric, value, unit, world, K, W, ms, config, clocks, e2e, launches, roofline, extra=None):
    d = {"metric": metric, "value": round(value, 3), "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
         "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
         "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * K),
         "launches_per_step": int(launches), "roofline": roofline, "cpu_baseline": None}
    if extra:
        d.update(extra)
    return d


def run_adversarial(args, rank, world, local, emit=True):
    import bench
    from models.bisenet.build_bisenet import BiSeNet
    from models.domain_shift.adversarial.model import DomainDiscriminator, TinyDomainDiscriminator
    from rtsds_b200 import ddp, ops
    from rtsds_b200.train_steps import adversarial_step

    dev = torch.device("cuda", local)
    b = args.batch if args.batch else 4
    torch.manual_seed(42)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = BiSeNet(19, "resnet18").to(dev).train()
    dis = (DomainDiscriminator if args.disc == "full" else TinyDomainDiscriminator)(19).to(dev).train()
    gen.rtsds_ddp = dis.rtsds_ddp = world > 1
    ddp.broadcast_module(gen, 0)
    ddp.broadcast_module(dis, 0)
    from rtsds_b200.optim import FusedAdam, FusedSGD  # noqa: F401

    gopt = FusedAdam(gen.parameters(), lr=1e-4)
    dopt = FusedAdam(dis.parameters(), lr=1e-4, weight_decay=1e-4)
    ce, bce = torch.nn.CrossEntropyLoss(ignore_index=19), torch.nn.BCEWithLogitsLoss()
    n_sets = 3
    g = torch.Generator().manual_seed(42 + rank)
    hs = torch.randn(n_sets, b, 3, 720, 1280, generator=g).pin_memory()
    hl = torch.randint(0, 20, (n_sets, b, 720, 1280), generator=g).pin_memory()
    ht = torch.randn(n_sets, b, 3, 512, 1024, generator=g).pin_memory()
    ds, dl, dt_ = hs.to(dev), hl.to(dev), ht.to(dev)
    K, W = args.steps, args.warmup

    def step(s, l, t):
        return adversarial_step(gen, dis, gopt, dopt, s, l, t, ce, bce, 0.1, 100, fused=not args.stock)

    for i in range(W):
        step(ds[i % n_sets], dl[i % n_sets], dt_[i % n_sets])
    torch.cuda.synchronize()
    c0 = ops.launch_count()
    step(ds[0], dl[0], dt_[0])
    torch.cuda.synchronize()
    launches = ops.launch_count() - c0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bench.barrier(world)
    with bench.ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            out = step(ds[i % n_sets], dl[i % n_sets], dt_[i % n_sets])
        e1.record()
        bench.barrier(world)
    ms = bench.max_over_ranks(e0.elapsed_time(e1), world)
    losses = {k: float(v.item()) for k, v in out.items()}
    # end to end: pinned host batches -> device, step, the four losses read back (train.py:214,234,253,264)
    ss, sl, st = torch.empty_like(ds[0]), torch.empty_like(dl[0]), torch.empty_like(dt_[0])
    bench.barrier(world)
    e0.record()
    for i in range(K):
        ss.copy_(hs[i % n_sets], non_blocking=True)
        sl.copy_(hl[i % n_sets], non_blocking=True)
        st.copy_(ht[i % n_sets], non_blocking=True)
        out = step(ss, sl, st)
        _ = [out[k].item() for k in ("loss_gen_source", "loss_adversarial", "loss_disc_source", "loss_disc_target")]
    e1.record()
    bench.barrier(world)
    ms_e2e = bench.max_over_ranks(e0.elapsed_time(e1), world)
    if rank != 0:
        return
    pk = bench.peaks()
    it_s = K / (ms / 1e3)
    img_s = world * b * it_s                          # source images (and as many target images) per second
    tfl = b * it_s * (G_TRAIN_GFLOP_720 + G_TRAIN_GFLOP_512) / 1e3
    line = (_line(
        "BiSeNet-R18 + discriminator adversarial training throughput (source images/s; each step also trains on as many target images)",
        img_s, "images/s", world, K, W, ms / K,
        {"workload": "adversarial_bisenet_r18 src 3x720x1280 + tgt 3x512x1024 (BASELINE.json configs[4])", "per_gpu_batch": b,
         "discriminator": args.disc, "lambda": 0.1, "optimizers": "Adam 1e-4 (G), Adam 1e-4 wd 1e-4 (D)",
         "path": "stock call sequence" if args.stock else "fused CE / fused softmax->D / fused BCE", "parallelism": f"dp{world}",
         "l2": "3 rotating input sets per rank"},
        clk.summary(),
        {"value": round(world * b * K / (ms_e2e / 1e3), 3), "unit": "images/s", "h2d_bytes_per_step": b * (3 * 720 * 1280 * 4 + 720 * 1280 * 8 + 3 * 512 * 1024 * 4),
         "d2h_bytes_per_step": 16, "ms_per_step": round(ms_e2e / K, 3)},
        launches,
        {"bound": "tensor", "achieved": round(tfl, 2), "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
         "frac": round(tfl / pk["bf16_tflops_sustained"], 4), "traffic": None, "peak_source": pk["source"],
         "kernel": "whole iteration, generator conv FLOPs only (2 fwd+bwd), per GPU"},
        {"final_losses": {k: round(v, 6) for k, v in losses.items()}}))
    if emit:
        bench.emit(line)
    return line


def run_deeplab(args, rank, world, local, emit=True):
    import bench
    from models.deeplabv2.deeplabv2 import get_deeplab_v2
    from rtsds_b200 import ddp, ops
    from rtsds_b200.deeplab_engine import deeplab_fused_ce

    dev = torch.device("cuda", local)
    b = args.batch if args.batch else 2
    torch.manual_seed(42)
    model = get_deeplab_v2(19, pretrain=False)
    # He-style conv scale instead of the reference's N(0, 0.01) so that activations stay O(1) through 101 layers
    for mod in model.modules():
        if isinstance(mod, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(mod.weight, mode="fan_in", nonlinearity="relu")
    model = model.to(dev).train()
    model.rtsds_ddp = world > 1
    ddp.broadcast_module(model, 0)
    from rtsds_b200.optim import FusedSGD

    opt = FusedSGD([p for p in model.parameters() if p.requires_grad], lr=1e-3, momentum=0.9)
    n_sets = 4
    g = torch.Generator().manual_seed(42 + rank)
    hx = torch.randn(n_sets, b, 3, 512, 1024, generator=g).pin_memory()
    hy = torch.randint(0, 20, (n_sets, b, 512, 1024), generator=g).pin_memory()
    dx, dy = hx.to(dev), hy.to(dev)
    K, W = args.steps, args.warmup

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss, pred, stats = deeplab_fused_ce(model, x, y, 19)
        loss.backward()
        opt.step()
        return loss

    for i in range(W):
        step(dx[i % n_sets], dy[i % n_sets])
    torch.cuda.synchronize()
    c0 = ops.launch_count()
    step(dx[0], dy[0])
    torch.cuda.synchronize()
    launches = ops.launch_count() - c0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bench.barrier(world)
    with bench.ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            loss = step(dx[i % n_sets], dy[i % n_sets])
        e1.record()
        bench.barrier(world)
    ms = bench.max_over_ranks(e0.elapsed_time(e1), world)
    final_loss = loss.item()
    sx, sy = torch.empty_like(dx[0]), torch.empty_like(dy[0])
    bench.barrier(world)
    e0.record()
    for i in range(K):
        sx.copy_(hx[i % n_sets], non_blocking=True)
        sy.copy_(hy[i % n_sets], non_blocking=True)
        _ = step(sx, sy).item()
    e1.record()
    bench.barrier(world)
    ms_e2e = bench.max_over_ranks(e0.elapsed_time(e1), world)
    # eval b=1 FPS (device-resident, CUDA graph)
    model.eval()
    ex = torch.randn(8, 1, 3, 512, 1024, generator=g).to(dev)
    with torch.no_grad():
        for i in range(3):
            model(ex[i])
        bench.barrier(world)
        e0.record()
        for i in range(20):
            model(ex[i % 8])
        e1.record()
        bench.barrier(world)
    ms_eval = bench.max_over_ranks(e0.elapsed_time(e1), world) / 20
    if rank != 0:
        return
    pk = bench.peaks()
    img_s = world * b * K / (ms / 1e3)
    tfl = img_s / world * DEEPLAB_TRAIN_GFLOP_512 / 1e3
    line = (_line(
        "DeepLabV2-R101 512x1024 data-parallel training throughput", img_s, "images/s", world, K, W, ms / K,
        {"workload": "deeplabv2_r101_train_3x512x1024 (BASELINE.json configs[3])", "per_gpu_batch": b, "optimizer": "SGD momentum 0.9",
         "loss": "CE(ignore_index=19), fused resize+CE", "parallelism": f"dp{world}", "l2": "4 rotating input sets per rank"},
        clk.summary(),
        {"value": round(world * b * K / (ms_e2e / 1e3), 3), "unit": "images/s", "h2d_bytes_per_step": b * (3 * 512 * 1024 * 4 + 512 * 1024 * 8),
         "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 3)},
        launches,
        {"bound": "tensor", "achieved": round(tfl, 2), "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
         "frac": round(tfl / pk["bf16_tflops_sustained"], 4), "traffic": None, "peak_source": pk["source"],
         "kernel": "whole training step (conv FLOPs only), per GPU"},
        {"final_loss": round(final_loss, 4),
         "eval_b1": {"fps": round(world * 1e3 / ms_eval, 2), "ms": round(ms_eval, 3), "tflops": round(DEEPLAB_FWD_GFLOP_512 / ms_eval, 1),
                     "frac_of_bf16_peak": round(DEEPLAB_FWD_GFLOP_512 / ms_eval / pk["bf16_tflops"], 4)}}))
    if emit:
        bench.emit(line)
    return line
