"""Training workload of bench.py (BASELINE.json configs[2]): BiSeNet-ResNet18 supervised training on
synthetic GTA5-shaped 720x1280 batches, data-parallel (one process per GPU, NCCL gradient all-reduce),
Adam lr 1e-4, 3 x CrossEntropyLoss(ignore_index=19), followed by an on-device fast_hist/mIoU validation
pass at 512x1024.  A step is what the reference's train() loop body does (train.py:68-106): poly LR,
zero_grad, forward, 3 x CE, backward, optimizer.step(), pixel accuracy."""
from __future__ import annotations

import json
import os
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
import statistics
import time

import torch

TRAIN_GFLOP_PER_IMG_720 = 267.5      # fwd + dgrad + wgrad conv FLOPs at 720x1280
TRAIN_CONV_MB_PER_IMG_720 = 3 * 408.0


def poly_lr(optimizer, init_lr, it, max_iter, power=0.9):
    """utils.poly_lr_scheduler (utils.py:33-48): only param_groups[0] is updated."""
    lr = init_lr * (1 - it / max_iter) ** power
    optimizer.param_groups[0]["lr"] = lr
    return lr


def measure_train(args, rank, world, local, steps, warmup, batch, h=720, w=1280, solo=False, e2e=True):
    """solo: this rank alone, no collective at all (the N = 1 figure measured inside a multi-GPU run; pass world=1)."""
    import bench
    from rtsds_b200 import ops
    from rtsds_b200.bisenet_autograd import bisenet_fused_ce

    dev = torch.device("cuda", local)
    model = bench.make_model(dev, getattr(args, "context", "resnet18")).train()
    model.rtsds_ddp = world > 1
    from rtsds_b200 import ddp

    if not solo:
        ddp.broadcast_module(model, 0)
    # Adam lr 1e-4 as main.py:116 builds it; the step is the library's own fused multi-tensor kernel, which also refreshes
    # the plan's packed bf16 conv operands (rtsds_b200/optim.py).  RTSDS_BENCH_TORCH_ADAM=1: torch's fused Adam
    if os.environ.get("RTSDS_BENCH_TORCH_ADAM"):
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    else:
        from rtsds_b200.optim import FusedAdam

        opt = FusedAdam(model.parameters(), lr=1e-4)
    n_sets = 4                                           # 4 x b x 11 MB images: far larger than the 126 MB L2 for b >= 4
    g = torch.Generator().manual_seed(42 + rank)
    host_x = torch.randn(n_sets, batch, 3, h, w, generator=g).pin_memory()
    host_y = torch.randint(0, 20, (n_sets, batch, h, w), generator=g).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)
    max_iter = 10 * (steps + warmup)

    def step(i, x, y):
        poly_lr(opt, 1e-4, i, max_iter)
        opt.zero_grad(set_to_none=True)
        loss, pred, stats = bisenet_fused_ce(model, x, y, 19)
        loss.backward()
        opt.step()
        return loss, stats

    for i in range(warmup):
        step(i, dev_x[i % n_sets], dev_y[i % n_sets])
    torch.cuda.synchronize()
    c0 = ops.launch_count()
    step(warmup, dev_x[0], dev_y[0])
    torch.cuda.synchronize()
    launches = ops.launch_count() - c0

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bench.barrier(world)
    with bench.ClockSampler(local) as clk:
        e0.record()
        for i in range(steps):
            loss, stats = step(warmup + i, dev_x[i % n_sets], dev_y[i % n_sets])
        e1.record()
        bench.barrier(world)
    ms = bench.max_over_ranks(e0.elapsed_time(e1), world)
    last_loss = loss.item()

    # end to end: pinned host batch -> device, step, loss read back every step (train.py:71-72,99)
    # the copies of batch i+1 overlap step i (rtsds_b200.serving.DevicePrefetcher: what a pin_memory DataLoader + prefetch
    # does); every batch still crosses PCIe inside the timed region and the loss is read back every step
    from rtsds_b200.serving import AsyncScalarReader, DevicePrefetcher

    if not e2e:
        return dict(ms=ms, ms_e2e=float("nan"), launches=launches, loss=last_loss, clocks=clk.summary(), miou=float("nan"),
                    h2d=batch * (3 * h * w * 4 + h * w * 8), val_fps=None)
    # The batches cross PCIe the way a dataset holds them — uint8 images [b,3,h,w] and uint8 label maps [b,h,w], 4 bytes per
    # pixel instead of the 20 of fp32 + int64 — and are converted on the device (rtsds_b200/input_pipeline.py, SURVEY N3:
    # read_image(...).float() -> Normalize, .long() -> IntRangeTransformer(0, 19) of main.py:68-76).  At 8 ranks the fp32 + int64
    # form (147 MB per step per rank) saturated the host side; RTSDS_BENCH_F32_IO=1 times that form instead.
    from rtsds_b200.input_pipeline import DeviceInputPipeline

    u8_io = not os.environ.get("RTSDS_BENCH_F32_IO")
    if u8_io:
        host_x = torch.randint(0, 256, (n_sets, batch, 3, h, w), dtype=torch.uint8, generator=g).pin_memory()
        host_y = torch.randint(0, 20, (n_sets, batch, h, w), dtype=torch.uint8, generator=g).pin_memory()
        pipe = DeviceInputPipeline(None, (123.675, 116.28, 103.53), (58.395, 57.12, 57.375))
        x_f32 = torch.empty(batch, 3, h, w, dtype=torch.float32, device=dev)
        y_i64 = torch.empty(batch, h, w, dtype=torch.int64, device=dev)

        def prep(sx, sy):
            return pipe.images(sx, x_f32), pipe.labels(sy, (0, 19), y_i64)
    else:
        def prep(sx, sy):
            return sx, sy
    h2d = batch * (3 * h * w + h * w) if u8_io else batch * (3 * h * w * 4 + h * w * 8)
    for sx, sy in DevicePrefetcher(((host_x[i % n_sets], host_y[i % n_sets]) for i in range(4)), dev):   # warm-up: staging buffers
        step(warmup + steps, *prep(sx, sy))
    batches = ((host_x[i % n_sets], host_y[i % n_sets]) for i in range(steps))
    reader = AsyncScalarReader(dev)
    n_read = 0
    bench.barrier(world)
    t_wall = time.perf_counter()
    e0.record()
    for i, (sx, sy) in enumerate(DevicePrefetcher(batches, dev)):
        loss, stats = step(warmup + steps + i, *prep(sx, sy))
        n_read += reader.push(loss) is not None          # every step's loss crosses to the host, one step late
    n_read += len(reader.drain())
    assert n_read == steps
    e1.record()
    bench.barrier(world)
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    ms_e2e = max(bench.max_over_ranks(e0.elapsed_time(e1), world), bench.max_over_ranks(wall_ms, world))

    # validation pass: eval forward at 512x1024 + fused argmax/fast_hist on device, matrix all-reduced once
    model.eval()
    hist = torch.zeros(19 * 19, dtype=torch.int64, device=dev)
    vx = torch.randn(1, 3, 512, 1024, generator=g).to(dev)
    vy = torch.randint(0, 20, (1, 512, 1024), generator=g).to(dev)
    n_val = 64
    with torch.no_grad():
        for _ in range(4):
            out = model(vx)
            ops.argmax_hist(out, vy, hist, None)
        hist.zero_()
        bench.barrier(world)
        e0.record()
        for _ in range(n_val):                       # validation.py:41-55 with the argmax + fast_hist fused on the device
            out = model(vx)
            ops.argmax_hist(out, vy, hist, None)
        if not solo:
            ddp.allreduce_confusion(hist)            # one 19x19 int64 all-reduce at the end of validation
        e1.record()
        bench.barrier(world)
    val_ms = bench.max_over_ranks(e0.elapsed_time(e1), world)
    hh = hist.cpu().numpy().reshape(19, 19).astype("float64")
    import numpy as np

    iou = np.diag(hh) / (hh.sum(1) + hh.sum(0) - np.diag(hh) + 1e-5)        # utils.per_class_iou
    return dict(ms=ms, ms_e2e=ms_e2e, launches=launches, loss=last_loss, clocks=clk.summary(), miou=float(np.nanmean(iou)),
                h2d=h2d, val_fps=world * n_val / (val_ms / 1e3))


def train_summary(r, world, batch, K):
    """Compact object attached to the default (inference) bench line."""
    import bench

    pk = bench.peaks()
    img_s = world * batch * K / (r["ms"] / 1e3)
    per_gpu = img_s / world
    return {"metric": "BiSeNet-R18 720x1280 data-parallel training throughput", "value": round(img_s, 2), "unit": "images/s",
            "per_gpu_batch": batch, "steps": K, "ms_per_step": round(r["ms"] / K, 3),
            "e2e_images_per_s": round(world * batch * K / (r["ms_e2e"] / 1e3), 2), "launches_per_step": int(r["launches"]),
            "tflops_per_gpu": round(per_gpu * TRAIN_GFLOP_PER_IMG_720 / 1e3, 2),
            "frac_of_bf16_sustained_peak": round(per_gpu * TRAIN_GFLOP_PER_IMG_720 / 1e3 / pk["bf16_tflops_sustained"], 4),
            "algorithmic_gbs_per_gpu": round(per_gpu * TRAIN_CONV_MB_PER_IMG_720 / 1e3, 1),
            "frac_of_hbm_peak": round(per_gpu * TRAIN_CONV_MB_PER_IMG_720 / 1e3 / pk["hbm_gbs"], 4),
            "final_loss": round(r["loss"], 4), "parallelism": f"dp{world}, NCCL bucketed all-reduce overlapped with backward",
            "validation": None if r.get("val_fps") is None else {
                "what": "BASELINE configs[2] second half: eval forward at 512x1024 + on-device argmax + fast_hist (validation.py:41-55), "
                        "confusion matrix all-reduced once; frames/s over all ranks",
                "frames_per_s": round(r["val_fps"], 1), "miou_random_weights": round(r["miou"], 5)}}


def run_train(args, rank, world, local):
    import bench

    batch = args.batch
    r = measure_train(args, rank, world, local, args.steps, args.warmup, batch)
    if rank != 0:
        return
    pk = bench.peaks()
    K = args.steps
    img_s = world * batch * K / (r["ms"] / 1e3)
    img_s_e2e = world * batch * K / (r["ms_e2e"] / 1e3)
    per_gpu = img_s / world
    line = {
        "metric": "BiSeNet-R18 720x1280 data-parallel training throughput", "value": round(img_s, 2), "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": round(r["ms"] / K, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "bisenet_r18_train_3x720x1280 (BASELINE.json configs[2])", "per_gpu_batch": batch,
                   "global_batch": batch * world, "optimizer": "Adam lr 1e-4 (rtsds_b200.optim.FusedAdam: one kernel = update + bf16 operand repack), poly LR", "loss": "3 x CE(ignore_index=19), fused resize+CE",
                   "parallelism": f"dp{world}", "l2": "4 rotating input sets per rank, each larger than L2 for b>=4"},
        "clocks": r["clocks"],
        "e2e": {"value": round(img_s_e2e, 2), "unit": "images/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                "ms_per_step": round(r["ms_e2e"] / K, 3)},
        "gpu_launches": int(r["launches"] * K), "launches_per_step": int(r["launches"]),
        "roofline": {"bound": "tensor", "achieved": round(per_gpu * TRAIN_GFLOP_PER_IMG_720 / 1e3, 2), "peak": pk["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": round(per_gpu * TRAIN_GFLOP_PER_IMG_720 / 1e3 / pk["bf16_tflops_sustained"], 4),
                     "traffic": None, "peak_source": pk["source"], "kernel": "whole training step (conv FLOPs only), per GPU"},
        "whole_step_hbm": {"algorithmic_gbs": round(per_gpu * TRAIN_CONV_MB_PER_IMG_720 / 1e3, 1),
                           "frac_of_hbm_peak": round(per_gpu * TRAIN_CONV_MB_PER_IMG_720 / 1e3 / pk["hbm_gbs"], 4)},
        "final_loss": round(r["loss"], 4), "val_miou_random_weights": round(r["miou"], 5),
        "cpu_baseline": None,
    }
    if getattr(args, "context", "resnet18") == "resnet101":      # the Bottleneck context path (build_bisenet.py:95-102)
        line["metric"] = "BiSeNet-R101 720x1280 data-parallel training throughput"
        line["config"]["workload"] = "bisenet_r101_train_3x720x1280 (SURVEY N4: context_path='resnet101')"
        line["roofline"] = None                                  # the FLOP constants above are the ResNet-18 model's
        line["whole_step_hbm"] = None
    bench.emit(line)
