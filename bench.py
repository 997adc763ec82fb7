#!/usr/bin/env python
"""Benchmark of the RTSDS hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train|adversarial|deeplab] [--impl reference]

Default workload (BASELINE.json configs[1]): BiSeNet-ResNet18 eval inference, batch 1, 3x512x1024,
metric = frames per second.  One JSON line is printed by rank 0.
  value     whole-job FPS with the input already resident in HBM (K steps, CUDA events, max over ranks)
  e2e       same metric through the drop-in module with HOST buffers: pinned H2D of the image,
            model(image), argmax, D2H of the prediction map inside the timed region
            (the validation.py:41-54 call pattern)
  roofline  aggregated tcgen05 implicit-GEMM conv launches of one forward: algorithmic FLOPs / summed
            CUDA-event durations vs the measured bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port of the reference CPU path, timed on this box's host cores
--impl reference times only that CPU path (the reference ships pure PyTorch; it cannot be pip-installed:
there is no setup.py/pyproject, and /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
os.environ.setdefault("RTSDS_ALLOW_RANDOM_INIT", "1")   # synthetic benchmark: seeded random-init backbone (no hub cache offline)
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, NUM_CLASSES = 512, 1024, 19
README_ITERS = 1000                # README.md:161-177
EVAL_DTYPE = "fp16"                # eval-mode operand / activation type (fp32 accumulation); training: bf16
FWD_GFLOP_PER_IMG = 51.26          # conv FLOPs (2*MAC) of one eval forward at 512x1024
FWD_CONV_MB_PER_IMG = 236.0        # ideal bf16 conv traffic
LOGITS_MB_PER_IMG = 39.8           # fp32 [19,512,1024] API-boundary write
CONV_DRAM_BYTES_PER_FORWARD = 120935168   # dram__bytes_read+write summed over the 22 conv launches of one forward (profiles/r02_conv_tc_infer_ncu_full.csv)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            p.update({k: float(d[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
            p["source"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, world):
    if world == 1:
        return v
    import torch.distributed as dist

    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_model(device, context="resnet18"):
    """Random-init weights of the BiSeNet architecture (no checkpoints offline)."""
    from models.bisenet.build_bisenet import BiSeNet

    torch.manual_seed(42)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = BiSeNet(NUM_CLASSES, context)
    # non-trivial BatchNorm running statistics so the folded epilogues do real work
    g = torch.Generator().manual_seed(7)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
    return m.to(device).eval()


def flush_l2(buf):
    buf.add_(1.0)   # 256 MB read+write: evicts the 126 MB L2


def tc_conv_profile(model, x, reps=20):
    """Average device time of every tensor-core conv launch of one forward.  Each launch is re-issued `reps` times back to
    back inside its own CUDA graph and timed with CUDA events around the replay (so host launch gaps are not counted, as
    in the CUDA-graph forward the model really runs); inputs of different layers are different buffers, weights are
    L2-resident as in steady-state serving."""
    from rtsds_b200 import ops

    plan = next(iter(model._rtsds_plans.values()))
    real, real_gap = ops.conv2d_tc, ops.conv2d_tc_gap
    calls = []

    def record(d, xx, w, y, *a, **k):
        calls.append((real, d, xx, w, y, a, k))
        real(d, xx, w, y, *a, **k)

    def record_gap(d, xx, w, y, *a, **k):       # the two convs whose epilogue also accumulates the ARM global pools
        calls.append((real_gap, d, xx, w, y, a, k))
        real_gap(d, xx, w, y, *a, **k)

    ops.conv2d_tc, ops.conv2d_tc_gap = record, record_gap
    try:
        plan.run_pre(x)
        plan.run_mid()
        torch.cuda.synchronize()
    finally:
        ops.conv2d_tc, ops.conv2d_tc_gap = real, real_gap
    rows = []
    for fn, d, xx, w, y, a, k in calls:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn(d, xx, w, y, *a, **k)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        flops = 2.0 * d.n * d.oh * d.ow * d.cout * d.cin * d.kh * d.kw
        byts = 2.0 * (d.n * d.h * d.w * d.cin + d.cout * d.cin * d.kh * d.kw) + d.n * d.oh * d.ow * d.cout * (2 if d.out_dtype == 1 else 4)
        rows.append((flops, byts, f"{d.cin}->{d.cout} k{d.kh} s{d.stride} {d.oh}x{d.ow}", statistics.median(ts)))
        del g
    plan.run_pre(x)          # leave the plan's buffers in a consistent state
    plan.run_mid()
    torch.cuda.synchronize()
    return rows


def run_infer(args, rank, world, local):
    from rtsds_b200 import ops

    dev = torch.device("cuda", local)
    r101 = args.context == "resnet101"          # eval only; the training half of the line is skipped
    model = make_model(dev, args.context)
    n_inputs = 32                                   # 32 x 6.29 MB = 201 MB > 126 MB L2
    g = torch.Generator().manual_seed(1234 + rank)
    host = torch.randn(n_inputs, 1, 3, H, W, generator=g).pin_memory()
    dev_in = host.to(dev)
    K, Wm = args.steps, args.warmup

    cuda_graph = bool(model.rtsds_cuda_graph)
    with torch.no_grad():
        for i in range(Wm):
            model(dev_in[i % n_inputs])
        plan = next(iter(model._rtsds_plans.values()))
        # kernels per step: eager launches + kernels captured in the CUDA graph
        c0 = ops.launch_count()
        plan.run_pre(dev_in[0]); plan.run_mid(); plan.logits(plan.z)
        torch.cuda.synchronize()
        launches_per_step = ops.launch_count() - c0

        # ---- device-resident throughput: K steps back to back ----
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(world)
        with ClockSampler(local) as clk:
            e0.record()
            for i in range(K):
                out = model(dev_in[i % n_inputs])
            e1.record()
            barrier(world)
        total_ms = max_over_ranks(e0.elapsed_time(e1), world)
        fps = world * K / (total_ms / 1e3)

        # ---- the same K frames over `lanes` compute streams (one execution plan / CUDA graph per lane): neighbouring
        # batch-1 frames overlap on the GPU, every frame is still an independent batch-1 forward ----
        lanes = max(1, args.lanes)
        lane_fps = None
        if lanes > 1:
            main_s = torch.cuda.current_stream(dev)
            streams = [torch.cuda.Stream(dev) for _ in range(lanes)]

            def lane_run(i):
                l = i % lanes
                with torch.cuda.stream(streams[l]):
                    model.rtsds_lane = l
                    model(dev_in[i % n_inputs])
                model.rtsds_lane = 0

            for st in streams:
                st.wait_stream(main_s)
            for i in range(4 * lanes):
                lane_run(i)
            torch.cuda.synchronize()
            barrier(world)
            e0.record()
            for st in streams:
                st.wait_event(e0)
            for i in range(K):
                lane_run(i)
            for st in streams:
                main_s.wait_stream(st)
            e1.record()
            barrier(world)
            lane_ms = max_over_ranks(e0.elapsed_time(e1), world)
            lane_fps = world * K / (lane_ms / 1e3)

        # ---- README protocol (README.md:157-177): per-iteration latency with a sync, mean/std of latency and 1/latency ----
        # ALWAYS 1000 iterations, whatever --steps says: this loop is BASELINE.json's metric and the line's `value`
        lat = []
        barrier(world)
        with ClockSampler(local) as clk_readme:
            for i in range(README_ITERS):
                t0 = time.perf_counter()
                out = model(dev_in[i % n_inputs])
                torch.cuda.synchronize()
                lat.append(time.perf_counter() - t0)
        fps_i = [1.0 / t for t in lat]
        readme_ms = max_over_ranks(1e3 * sum(lat) / len(lat), world)

        # ---- cold-L2 latency: flush between iterations, each iteration timed by its own event pair ----
        flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        cold = []
        for i in range(min(K, 50)):
            flush_l2(flush)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = model(dev_in[i % n_inputs]); b.record()
            torch.cuda.synchronize()
            cold.append(a.elapsed_time(b))
        del flush

        # ---- end to end with host buffers: H2D image, forward, argmax, D2H predictions ----
        pred_dev = torch.empty(1, H, W, dtype=torch.int64, device=dev)
        pred_host = torch.empty(1, H, W, dtype=torch.int64).pin_memory()
        stage = torch.empty(1, 3, H, W, dtype=torch.float32, device=dev)

        def e2e_step(i):
            stage.copy_(host[i % n_inputs], non_blocking=True)
            o = model(stage)
            ops.argmax_hist(o, None, None, pred_dev)
            pred_host.copy_(pred_dev, non_blocking=True)

        for i in range(3):
            e2e_step(i)
        barrier(world)
        e0.record()
        for i in range(K):
            e2e_step(i)
        e1.record()
        barrier(world)
        e2e_serial_ms = max_over_ranks(e0.elapsed_time(e1), world)

        # the same work as a steady-state loop: copies of neighbouring frames overlap the forward (rtsds_b200/serving.py);
        # every frame is still copied in from pinned host memory and its prediction map copied back, inside the timed region
        from rtsds_b200.serving import PipelinedSegmenter

        def pipe_run(pipe, frames):
            checksum = 0
            for i in range(6):
                pipe.submit(frames[i % n_inputs])
            pipe.drain()
            barrier(world)
            t_wall = time.perf_counter()
            e0.record()
            for i in range(K):
                r = pipe.submit(frames[i % n_inputs])
                if r is not None:
                    checksum += int(r[0, 0, 0])
            for r in pipe.drain():
                checksum += int(r[0, 0, 0])
            e1.record()
            barrier(world)
            wall_ms = 1e3 * (time.perf_counter() - t_wall)
            return max(max_over_ranks(e0.elapsed_time(e1), world), max_over_ranks(wall_ms, world))

        depth = 3 if lanes == 1 else lanes + 2
        e2e_f32_ms = pipe_run(PipelinedSegmenter(model, 1, H, W, depth=depth, lanes=lanes), host)
        # the same loop on RAW uint8 frames (what a camera / read_image delivers; the reference converts and normalises on
        # the CPU, datasets/cityscapes.py:66 + main.py:68-71): normalisation inside the stem kernel, uint8 class map back
        model.rtsds_input_norm = ((123.675, 116.28, 103.53), (58.395, 57.12, 57.375))
        host_u8 = torch.randint(0, 256, (n_inputs, 1, 3, H, W), dtype=torch.uint8, generator=g).pin_memory()
        e2e_ms = pipe_run(PipelinedSegmenter(model, 1, H, W, depth=depth, lanes=lanes, uint8_io=True), host_u8)
        e2e_fps = world * K / (e2e_ms / 1e3)
        # one frame at a time through the same uint8 path: H2D, forward, argmax, D2H, synchronise
        u8_dev = torch.empty(1, 3, H, W, dtype=torch.uint8, device=dev)
        p8_dev = torch.empty(1, H, W, dtype=torch.uint8, device=dev)
        p8_host = torch.empty(1, H, W, dtype=torch.uint8).pin_memory()
        ser = []
        for i in range(min(K, 200) + 5):
            t0 = time.perf_counter()
            u8_dev.copy_(host_u8[i % n_inputs], non_blocking=True)
            ops.argmax_hist(model(u8_dev), None, None, p8_dev)
            p8_host.copy_(p8_dev, non_blocking=True)
            torch.cuda.synchronize()
            ser.append(time.perf_counter() - t0)
        e2e_u8_serial_ms = 1e3 * statistics.mean(ser[5:])

        rows = tc_conv_profile(model, dev_in[0]) if rank == 0 else []

    # ---- the other half of BASELINE.json's metric: data-parallel training images/s (configs[2]) ----
    model_state = {k: v.detach().cpu() for k, v in model.state_dict().items()} if rank == 0 else None
    train, other = None, {}
    if not args.no_train and not r101:
        import bench_train

        del dev_in, host, model
        torch.cuda.empty_cache()
        tk = max(100, K)                                   # >= 100 timed steps (1 s) whatever --steps says
        tr = bench_train.measure_train(args, rank, world, local, tk, 10, args.batch)
        train = bench_train.train_summary(tr, world, args.batch, tk)
        if not args.no_extra:
            # BASELINE configs[3] / configs[4] on the same box in the same run, so that a driver record of them exists
            import bench_extra
            import copy

            a2 = copy.copy(args)
            a2.batch, a2.steps, a2.warmup = 0, 30, 5
            torch.cuda.empty_cache()
            dl = bench_extra.run_deeplab(a2, rank, world, local, emit=False)
            torch.cuda.empty_cache()
            adv = bench_extra.run_adversarial(a2, rank, world, local, emit=False)
            torch.cuda.empty_cache()
            if rank == 0:
                keep = ("metric", "value", "unit", "steps", "ms_per_step", "config", "e2e", "launches_per_step", "roofline",
                        "final_loss", "final_losses", "eval_b1")
                other = {"config4_deeplabv2": {k: dl[k] for k in keep if k in dl},
                         "config5_adversarial": {k: adv[k] for k in keep if k in adv}}
    comparator = None
    if rank == 0 and not args.no_comparator and not r101:
        import bench_comparator

        torch.cuda.empty_cache()
        comparator = bench_comparator.run_all(torch.device("cuda", local), model_state, train_batch=args.batch,
                                              full=not args.no_extra)
        torch.cuda.empty_cache()
    barrier(world)

    if rank != 0:
        return
    pk = peaks()
    tc_flops = sum(r[0] for r in rows)
    tc_ms = sum(r[3] for r in rows)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    single_fps, single_ms = fps, total_ms / K
    # headline: BASELINE.json's metric verbatim — one frame at a time, every iteration synchronised, 1000 iterations
    fps, step_ms = world * 1e3 / readme_ms, readme_ms
    hbm_achieved = (FWD_CONV_MB_PER_IMG + LOGITS_MB_PER_IMG) * 1e6 / (step_ms * 1e-3) / 1e9
    cpu = cpu_baseline(args)
    line = {
        "metric": f"BiSeNet-{'R101' if r101 else 'R18'} 512x1024 inference FPS (batch 1)", "value": round(fps, 2), "unit": "frames/s",
        "n_gpus": world, "steps": K, "timed_iterations": README_ITERS, "warmup": Wm, "ms_per_step": round(step_ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": EVAL_DTYPE, "data": "synthetic",
        "config": {"workload": ("bisenet_r101_eval_b1_3x512x1024 (config.yaml backbone: resnet101, SURVEY N4)" if r101 else
                                "bisenet_r18_eval_b1_3x512x1024 (BASELINE.json configs[1])"), "num_classes": NUM_CLASSES,
                   "weights": "random-init, seeded", "parallelism": "replicas only" if world > 1 else "single GPU",
                   "timed_region": f"value = README.md:161-177 protocol: {README_ITERS} iterations (regardless of --steps), one frame "
                                   "at a time on one stream, torch.cuda.synchronize() and wall clock around every iteration; "
                                   "device_timed = --steps frames back to back between CUDA events",
                   "arithmetic": "tcgen05 kind::f16, fp16 operands / activations (IEEE half), fp32 TMEM accumulation; training runs bf16",
                   "l2": "inputs rotate over 32 distinct images (201 MB > 126 MB L2); weights stay L2-resident as in "
                         "steady-state serving; latency_cold_l2_ms flushes L2 before every iteration",
                   "cuda_graph": cuda_graph,
                   "streams": f"{lanes} concurrent batch-1 streams, one execution plan + CUDA graph each (single_stream = 1)"},
        "clocks": clk_readme.summary(),
        "device_timed": {"value": round(single_fps, 2), "unit": "frames/s", "steps": K, "ms_per_step": round(single_ms, 4),
                         "note": "single stream, one frame at a time back to back, CUDA events, no per-iteration sync",
                         "clocks": clk.summary()},
        "multi_stream": None if lane_fps is None else {
            "value": round(lane_fps, 2), "unit": "frames/s", "streams": lanes, "ms_per_step": round(lane_ms / K, 4),
            "note": f"{lanes} concurrent batch-1 streams (one execution plan + CUDA graph each): throughput serving, NOT the headline"},
        "e2e": {"value": round(e2e_fps, 2), "unit": "frames/s", "h2d_bytes_per_step": 3 * H * W, "d2h_bytes_per_step": H * W,
                "ms_per_step": round(e2e_ms / K, 4),
                "how": f"PipelinedSegmenter(uint8_io=True): every frame is copied from pinned host memory as RAW uint8 [1,3,{H},{W}] "
                       f"(normalised inside the stem kernel), forward + argmax on {lanes} compute streams, the uint8 class map copied "
                       f"back to pinned host memory; {depth} frames in flight, wall clock incl. final drain",
                "one_frame_at_a_time_fps": round(world * 1e3 / e2e_u8_serial_ms, 2),
                "fp32_in_int64_out": {"value": round(world * K / (e2e_f32_ms / 1e3), 2), "h2d_bytes_per_step": 3 * H * W * 4,
                                      "d2h_bytes_per_step": H * W * 8, "serial_fps": round(world * K / (e2e_serial_ms / 1e3), 2),
                                      "note": "the reference's own boundary types: fp32 image in, torch.argmax's int64 map out"}},
        "gpu_launches": int(launches_per_step * K),
        "launches_per_step": int(launches_per_step),
        "readme_protocol": {"iterations": len(lat), "mean_latency_ms": round(1e3 * statistics.mean(lat), 4),
                            "median_latency_ms": round(1e3 * statistics.median(lat), 4),
                            "std_latency_ms": round(1e3 * statistics.pstdev(lat), 4), "mean_fps": round(statistics.mean(fps_i), 2),
                            "std_fps": round(statistics.pstdev(fps_i), 2)},
        "latency_cold_l2_ms": round(statistics.median(cold), 4),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": round(achieved / pk["bf16_tflops"], 4), "traffic": None if r101 else CONV_DRAM_BYTES_PER_FORWARD, "peak_source": pk["source"],
                     "traffic_source": "ncu --set full of the conv launches of one forward (cold L2), profiles/r02_conv_tc_infer_ncu_full.csv",
                     "kernel": "conv_tc_kernel (22 launches/forward, aggregated; each launch timed as the average of 20 back-to-back graph-replayed repeats, split-K finish kernels included)", "flops_per_step": tc_flops,
                     "kernel_ms_per_step": round(tc_ms, 4)},
        "whole_step": None if r101 else {"tflops": round(FWD_GFLOP_PER_IMG / step_ms, 2), "frac_of_bf16_peak": round(FWD_GFLOP_PER_IMG / step_ms / pk["bf16_tflops"], 4),
                       "algorithmic_gbs": round(hbm_achieved, 1), "frac_of_hbm_peak": round(hbm_achieved / pk["hbm_gbs"], 4)},
        "conv_layers": [{"layer": r[2], "ms": round(r[3], 4), "tflops": round(r[0] / (r[3] * 1e-3) / 1e12, 1),
                         "gbs": round(r[1] / (r[3] * 1e-3) / 1e9, 1)} for r in rows],
        "train": train,
        "other_configs": other,
        "gpu_comparator": comparator,
        "cpu_baseline": cpu,
    }
    emit(line)


def run_multi(args, rank, world, local):
    """--gpus N > 1.  Batch-1 inference does not shard (replicas only); the path that shards — by batch, with ONE exchange
    per step, the NCCL gradient all-reduce — is data-parallel TRAINING (BASELINE configs[2]), so that is the headline here:
    images/s over all ranks, max(--steps, 100) timed steps between barriers, CUDA events, max over ranks.  The same step
    is also timed on rank 0 ALONE in this run (no collective) so the scaling of this very box can be read off the line;
    inference replicas are reported beside it."""
    import bench_train

    K = max(args.steps, 100)
    Wm = max(args.warmup, 10)
    solo = None
    if rank == 0:
        solo = bench_train.measure_train(args, 0, 1, local, K, Wm, args.batch, solo=True, e2e=False)
    torch.cuda.empty_cache()
    barrier(world)
    r = bench_train.measure_train(args, rank, world, local, K, Wm, args.batch)
    torch.cuda.empty_cache()

    # ---- inference replicas (no collective): device-timed batch-1 FPS summed over ranks ----
    dev = torch.device("cuda", local)
    model = make_model(dev)
    g = torch.Generator().manual_seed(1234 + rank)
    dev_in = torch.randn(32, 1, 3, H, W, generator=g).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i in range(20):
            model(dev_in[i % 32])
        barrier(world)
        e0.record()
        for i in range(500):
            model(dev_in[i % 32])
        e1.record()
        barrier(world)
    rep_ms = max_over_ranks(e0.elapsed_time(e1), world)
    if rank != 0:
        return
    pk = peaks()
    b = args.batch
    img_s = world * b * K / (r["ms"] / 1e3)
    img_s_e2e = world * b * K / (r["ms_e2e"] / 1e3)
    solo_img_s = b * K / (solo["ms"] / 1e3)
    per_gpu = img_s / world
    G, MB = bench_train.TRAIN_GFLOP_PER_IMG_720, bench_train.TRAIN_CONV_MB_PER_IMG_720
    line = {
        "metric": "BiSeNet-R18 720x1280 data-parallel training throughput", "value": round(img_s, 2), "unit": "images/s",
        "n_gpus": world, "steps": K, "steps_requested": args.steps, "warmup": Wm, "ms_per_step": round(r["ms"] / K, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "bisenet_r18_train_3x720x1280 (BASELINE.json configs[2]) — the headline under --gpus N > 1, "
                               "because batch-1 inference (configs[1], the N = 1 headline) only replicates",
                   "per_gpu_batch": b, "global_batch": b * world,
                   "optimizer": "Adam lr 1e-4, poly LR on param_groups[0]", "loss": "3 x CE(ignore_index=19), fused resize+CE",
                   "parallelism": f"dp{world}: one process per GPU, NCCL all-reduce (AVG) of the flat fp32 gradient in buckets, overlapped with backward; per-rank BatchNorm",
                   "l2": "4 rotating input sets per rank, each larger than L2"},
        "clocks": r["clocks"],
        "single_gpu_same_run": {"value": round(solo_img_s, 2), "unit": "images/s", "ms_per_step": round(solo["ms"] / K, 3),
                                "note": "the identical step on rank 0 alone, no collective, measured before the N-GPU run on this box"},
        "e2e": {"value": round(img_s_e2e, 2), "unit": "images/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                "ms_per_step": round(r["ms_e2e"] / K, 3)},
        "gpu_launches": int(r["launches"] * K), "launches_per_step": int(r["launches"]),
        "roofline": {"bound": "tensor", "achieved": round(per_gpu * G / 1e3, 2), "peak": pk["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": round(per_gpu * G / 1e3 / pk["bf16_tflops_sustained"], 4), "traffic": None,
                     "peak_source": pk["source"], "kernel": "whole training step (conv FLOPs only), per GPU"},
        "whole_step_hbm": {"algorithmic_gbs": round(per_gpu * MB / 1e3, 1), "frac_of_hbm_peak": round(per_gpu * MB / 1e3 / pk["hbm_gbs"], 4)},
        "final_loss": round(r["loss"], 4),
        "validation": {"frames_per_s": round(r["val_fps"], 1), "miou_random_weights": round(r["miou"], 5)},
        "inference_replicas": {"value": round(world * 500 / (rep_ms / 1e3), 1), "unit": "frames/s", "ms_per_frame": round(rep_ms / 500, 4),
                               "note": "batch-1 eval forward on every GPU independently (replicas only, no collective), device-timed"},
        "cpu_baseline": None,
    }
    emit(line)


def cpu_reference_fps(seconds_budget=12.0, max_iters=60, context="resnet18"):
    """Oracle port of the reference CPU path (BiSeNet eval, b=1, 512x1024) on all host threads."""
    from oracle import bisenet_ref, weights

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = weights.bisenet_r101_state(42) if context == "resnet101" else weights.bisenet_r18_state(42)
    x = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        bisenet_ref.bisenet_forward(x, sd, train=False)      # warm-up
        ts = []
        t_start = time.perf_counter()
        while len(ts) < max_iters and time.perf_counter() - t_start < seconds_budget:
            t0 = time.perf_counter()
            bisenet_ref.bisenet_forward(x, sd, train=False)
            ts.append(time.perf_counter() - t0)
    return len(ts) / sum(ts), threads, len(ts), ts


def cpu_baseline(args):
    fps, threads, iters, _ = cpu_reference_fps(context=args.context)
    return {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{iters} eval forwards of the oracle port (torch CPU fp32, README protocol) at b=1 3x512x1024"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    K, Wm = args.steps, args.warmup
    iters = max(3, min(K, 40))
    fps, threads, n, ts = cpu_reference_fps(seconds_budget=60.0, max_iters=iters, context=args.context)
    r101 = args.context == "resnet101"
    line = {
        "impl": "reference", "metric": f"BiSeNet-{'R101' if r101 else 'R18'} 512x1024 inference FPS (batch 1)", "value": round(fps, 3),
        "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(1e3 / fps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("bisenet_r101_eval_b1_3x512x1024 (config.yaml backbone: resnet101, SURVEY N4)" if r101 else
                                "bisenet_r18_eval_b1_3x512x1024 (BASELINE.json configs[1])"), "num_classes": NUM_CLASSES,
                   "note": "reference CPU path (oracle port; the reference is not pip-installable and /root/reference is absent on the GPU box)"},
        "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{n} timed eval forwards at b=1 3x512x1024 on {threads} host threads"},
        "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line: dict) -> None:
    """Write the result line to the process's original stdout (see main(); the descriptor travels in the environment
    because bench_train / bench_extra import this file as a second module object next to __main__)."""
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(int(os.environ.get("RTSDS_BENCH_STDOUT_FD", "1")), data)


def main():
    # stdout carries exactly ONE JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its
    # "NCCL version ..." banner straight to stdout when the box exports NCCL_DEBUG) and emit() writes the line to the real one
    sys.stdout.flush()
    os.environ["RTSDS_BENCH_STDOUT_FD"] = str(os.dup(1))
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="rtsds_b200", choices=["rtsds_b200", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "adversarial", "deeplab"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 8 train, 4 adversarial, 2 deeplab)")
    ap.add_argument("--disc", default="tiny", choices=["tiny", "full"], help="discriminator of the adversarial workload")
    ap.add_argument("--stock", action="store_true", help="adversarial workload: the reference's exact call sequence instead of the fused fast paths")
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput part of the default run")
    ap.add_argument("--no-extra", action="store_true", help="skip the DeepLabV2 / adversarial extras and the long comparator legs")
    ap.add_argument("--no-comparator", action="store_true", help="skip the cuDNN comparator leg")
    ap.add_argument("--context", default="resnet18", choices=["resnet18", "resnet101"], help="BiSeNet context path (inference and --workload train)")
    ap.add_argument("--lanes", type=int, default=3, help="inference: concurrent batch-1 streams of the multi-stream / pipelined measurements")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload in ("infer", "train") and not args.batch:
        args.batch = 8
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rtsds_b200 has no CPU fallback (use --impl reference for the CPU path)")
    rank, world, local = dist_setup(args.gpus)
    try:
        if args.workload == "infer" and world > 1:
            run_multi(args, rank, world, local)
        elif args.workload == "infer":
            run_infer(args, rank, world, local)
        elif args.workload == "adversarial":
            import bench_extra

            bench_extra.run_adversarial(args, rank, world, local)
        elif args.workload == "deeplab":
            import bench_extra

            bench_extra.run_deeplab(args, rank, world, local)
        else:
            from bench_train import run_train

            run_train(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
