"""GPU comparator leg of bench.py: "the Blackwell kernel to beat on the same box" (SURVEY §2.1:100-103, §8(d) config 2).

The reference ships no native code: on a GPU every FLOP of its hot path is an ATen call that PyTorch dispatches to
cuDNN / cuBLAS.  This leg runs exactly that computation — the functional restatement of the reference modules under
oracle/ (the same F.conv2d / F.batch_norm / F.interpolate / F.cross_entropy calls, hence the same library kernels the
reference's nn.Modules launch) — moved to the GPU UNMODIFIED, in the two ways a user of the reference could run it:

  fp32_tf32        as shipped: fp32 NCHW tensors, cuDNN convolutions with TF32 allowed (torch's default)
  cl_bf16_autocast channels_last + torch.autocast(bfloat16): the fair-fight tensor-core configuration

each eagerly (one Python-dispatched kernel launch per op, the way train.py / validation.py run) and — inference only —
under CUDA-graph replay (no host launch overhead).  torch.backends.cudnn.benchmark = True (cuDNN picks its fastest
algorithm per shape; best case for the comparator).  Nothing of rtsds_b200 runs in this leg.

The oracle is test infrastructure; this file and bench.py's cpu_baseline / --impl reference legs are the only
non-test code that executes it, always as the thing COMPARED AGAINST, never as the product path.
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn.functional as F

H, W = 512, 1024


def _to_dev(sd, dev, channels_last):
    out = {}
    for k, v in sd.items():
        t = v.detach().to(dev)
        if channels_last and t.dim() == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        out[k] = t
    return out


def _ctx(variant):
    if variant == "cl_bf16_autocast":
        return torch.autocast("cuda", dtype=torch.bfloat16)
    import contextlib

    return contextlib.nullcontext()


def _prep(x, variant):
    return x.contiguous(memory_format=torch.channels_last) if variant == "cl_bf16_autocast" else x


def _time_loop(fn, min_iters, budget_s):
    """Device time per call: CUDA events around back-to-back calls (>= min_iters, up to budget_s seconds)."""
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while n < min_iters or (time.perf_counter() - t0 < budget_s and n < 50 * min_iters):
        fn()
        n += 1
        if n >= min_iters and n % 8 == 0:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, n


def _readme_loop(fn, iters):
    lat = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        lat.append(time.perf_counter() - t0)
    return sum(lat) / len(lat)


def infer(dev, sd, x, budget_s=1.0):
    """BASELINE configs[1]: eval forward, b=1 3x512x1024.  -> {variant: {eager_fps, eager_readme_fps, graph_fps}}"""
    from oracle import bisenet_ref

    res = {}
    for variant in ("fp32_tf32", "cl_bf16_autocast"):
        sdd = _to_dev(sd, dev, variant == "cl_bf16_autocast")
        xs = _prep(x.to(dev), variant)

        def fwd():
            with torch.no_grad(), _ctx(variant):
                return bisenet_ref.bisenet_forward(xs, sdd, train=False)

        ms, n = _time_loop(fwd, 20, budget_s)
        readme_ms = 1e3 * _readme_loop(fwd, 100)
        graph_ms = None
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    fwd()
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fwd()
            graph_ms, _ = _time_loop(g.replay, 50, budget_s)
            del g, out
        except Exception as e:      # capture can fail for an op; the eager numbers stand
            graph_ms = None
            res.setdefault("notes", []).append(f"{variant}: graph capture failed: {type(e).__name__}")
        res[variant] = {"eager_fps": round(1e3 / ms, 1), "eager_ms": round(ms, 4), "eager_iters": n,
                        "eager_readme_protocol_fps": round(1e3 / readme_ms, 1),
                        "graph_fps": round(1e3 / graph_ms, 1) if graph_ms else None,
                        "graph_ms": round(graph_ms, 4) if graph_ms else None}
    return res


def _leaves(sdd, prefixes):
    ls = []
    for k, v in sdd.items():
        if v.dtype.is_floating_point and "running" not in k and k.startswith(prefixes):
            v.requires_grad_(True)
            ls.append(v)
    return ls


G_PREFIXES = ("context_path.features", "saptial", "attention", "supervision", "feature_fusion", "conv.")


def train(dev, sd, batch, h, w, budget_s=2.0, min_iters=10):
    """BASELINE configs[2] step (train.py:68-106): zero_grad, forward, 3 x CE(ignore 19), backward, Adam step, argmax."""
    from oracle import bisenet_ref

    g = torch.Generator().manual_seed(7)
    x = torch.randn(batch, 3, h, w, generator=g)
    y = torch.randint(0, 20, (batch, h, w), generator=g).to(dev)
    res = {}
    for variant in ("fp32_tf32", "cl_bf16_autocast"):
        sdd = _to_dev(sd, dev, variant == "cl_bf16_autocast")
        leaves = _leaves(sdd, G_PREFIXES)
        opt = torch.optim.Adam(leaves, lr=1e-4, fused=True)
        xs = _prep(x.to(dev), variant)

        def step():
            opt.zero_grad(set_to_none=True)
            with _ctx(variant):
                outs = bisenet_ref.bisenet_forward(xs, sdd, train=True)
                loss = sum(F.cross_entropy(t.float(), y, ignore_index=19) for t in outs)
            loss.backward()
            opt.step()
            return outs[0].max(1)[1]

        ms, n = _time_loop(step, min_iters, budget_s)
        res[variant] = {"images_per_s": round(batch * 1e3 / ms, 1), "ms_per_step": round(ms, 3), "iters": n}
        del sdd, leaves, opt
        torch.cuda.empty_cache()
    return res


def config1_parity(dev, sd, x, ref_eval, ref_train):
    """How far the reference's OWN GPU path is from its CPU path on BASELINE config 1 (same weights, same input): the
    yardstick for north_star's tolerances (rel 2e-2, argmax 99.9 %)."""
    from oracle import bisenet_ref

    def cmp(a, b):
        a = a.float().cpu()
        return {"rel": round((a - b).abs().max().item() / b.abs().max().item(), 6),
                "argmax_agree": round((a.argmax(1) == b.argmax(1)).float().mean().item(), 6)}

    res = {}
    for variant in ("fp32_tf32", "cl_bf16_autocast"):
        sdd = _to_dev(sd, dev, variant == "cl_bf16_autocast")
        xs = _prep(x.to(dev), variant)
        with torch.no_grad(), _ctx(variant):
            ev = bisenet_ref.bisenet_forward(xs, sdd, train=False)
            tr = bisenet_ref.bisenet_forward(xs, sdd, train=True)
        res[variant] = {"eval": cmp(ev, ref_eval), "train_main_head": cmp(tr[0], ref_train[0])}
    return res


def deeplab(dev, budget_s=2.0):
    """BASELINE configs[3]: DeepLabV2-R101 train b=2 (CE on out[0]) and eval b=1 at 512x1024."""
    from oracle import deeplab_ref, weights

    sd = weights.deeplab_state(42)
    g = torch.Generator().manual_seed(7)
    x2 = torch.randn(2, 3, H, W, generator=g)
    y2 = torch.randint(0, 20, (2, H, W), generator=g).to(dev)
    res = {}
    for variant in ("fp32_tf32", "cl_bf16_autocast"):
        sdd = _to_dev(sd, dev, variant == "cl_bf16_autocast")
        # the reference freezes the BatchNorm affine parameters (deeplabv2.py:15-27): conv weights / biases train
        leaves = [v.requires_grad_(True) for k, v in sdd.items()
                  if v.dtype.is_floating_point and (v.dim() == 4 or (k.startswith("layer6") and k.endswith(".bias")))]
        opt = torch.optim.SGD(leaves, lr=1e-3, momentum=0.9)
        xs2 = _prep(x2.to(dev), variant)
        xs1 = xs2[:1].contiguous(memory_format=torch.channels_last) if variant == "cl_bf16_autocast" else xs2[:1].contiguous()

        def step():
            opt.zero_grad(set_to_none=True)
            with _ctx(variant):
                out = deeplab_ref.deeplab_forward(xs2, sdd, train=True)
                out = out[0] if isinstance(out, (tuple, list)) else out
                loss = F.cross_entropy(out.float(), y2, ignore_index=19)
            loss.backward()
            opt.step()

        def fwd():
            with torch.no_grad(), _ctx(variant):
                return deeplab_ref.deeplab_forward(xs1, sdd, train=False)

        ms_t, n_t = _time_loop(step, 5, budget_s)
        ms_e, n_e = _time_loop(fwd, 10, budget_s / 2)
        res[variant] = {"train_images_per_s": round(2 * 1e3 / ms_t, 2), "train_ms_per_step": round(ms_t, 2), "train_iters": n_t,
                        "eval_fps": round(1e3 / ms_e, 1), "eval_ms": round(ms_e, 3)}
        del sdd, leaves, opt
        torch.cuda.empty_cache()
    return res


def adversarial(dev, sd, batch, budget_s=2.0):
    """BASELINE configs[4]: one iteration of train.py:adversarial_train (:177-270) — 2 generator fwd+bwd, 3 discriminator
    fwd / bwd, both Adam steps — with the TinyDomainDiscriminator (config.yaml default)."""
    from oracle import adversarial_ref, weights

    dsd = weights.discriminator_state(42, tiny=True)
    g = torch.Generator().manual_seed(7)
    xs_ = torch.randn(batch, 3, 720, 1280, generator=g)
    ys_ = torch.randint(0, 20, (batch, 720, 1280), generator=g).to(dev)
    xt_ = torch.randn(batch, 3, H, W, generator=g)
    res = {}
    for variant in ("fp32_tf32", "cl_bf16_autocast"):
        gsd = _to_dev(sd, dev, variant == "cl_bf16_autocast")
        ddd = _to_dev(dsd, dev, variant == "cl_bf16_autocast")
        gl = _leaves(gsd, G_PREFIXES)
        dl = [v.requires_grad_(True) for k, v in ddd.items() if v.dtype.is_floating_point]
        gopt = torch.optim.Adam(gl, lr=1e-4, fused=True)
        dopt = torch.optim.Adam(dl, lr=1e-4, weight_decay=1e-4, fused=True)
        xs, xt = _prep(xs_.to(dev), variant), _prep(xt_.to(dev), variant)

        def step():
            gopt.zero_grad(set_to_none=True)
            dopt.zero_grad(set_to_none=True)
            with _ctx(variant):
                adversarial_ref.adversarial_iteration(gsd, ddd, xs, ys_, xt, 19, 0.1, 100)
            gopt.step()
            dopt.step()

        ms, n = _time_loop(step, 5, budget_s)
        res[variant] = {"source_images_per_s": round(batch * 1e3 / ms, 1), "ms_per_iteration": round(ms, 2), "iters": n}
        del gsd, ddd, gl, dl, gopt, dopt
        torch.cuda.empty_cache()
    return res


def run_all(dev, model_state, train_batch=8, adv_batch=4, full=True):
    """Everything above for BASELINE configs 1-5 on `dev`; a failure of one leg is reported, not raised."""
    torch.backends.cudnn.benchmark = True
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"what": "the reference's computation (functional restatement: the same ATen ops, hence cuDNN/cuBLAS kernels, its "
                   "nn.Modules dispatch) moved to this GPU unmodified; none of rtsds_b200's kernels run here",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "cudnn_benchmark": True,
           "tf32_convs": bool(torch.backends.cudnn.allow_tf32)}
    sd = {k: v.detach().cpu() for k, v in model_state.items()}
    g = torch.Generator().manual_seed(1234)
    x2 = torch.randn(2, 3, H, W, generator=g)

    def leg(name, fn):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
        out.setdefault("seconds", {})[name] = round(time.perf_counter() - t0, 1)

    leg("config2_infer_b1_512x1024", lambda: infer(dev, sd, x2[:1]))
    leg("config3_train_720x1280", lambda: dict(per_gpu_batch=train_batch, **train(dev, sd, train_batch, 720, 1280)))
    if full:
        leg("config1_train_step_b2_512x1024", lambda: dict(per_gpu_batch=2, **train(dev, sd, 2, H, W, budget_s=1.0)))

        def parity():
            from oracle import bisenet_ref

            with torch.no_grad():
                re_ = bisenet_ref.bisenet_forward(x2, {k: v.clone() for k, v in sd.items()}, train=False)
                rt_ = bisenet_ref.bisenet_forward(x2, {k: v.clone() for k, v in sd.items()}, train=True)
            return config1_parity(dev, sd, x2, re_, rt_)

        leg("config1_parity_of_the_reference_gpu_path_vs_its_cpu_path", parity)
        leg("config4_deeplabv2_512x1024", lambda: deeplab(dev))
        leg("config5_adversarial_tiny_d", lambda: dict(per_gpu_batch=adv_batch, **adversarial(dev, sd, adv_batch)))
    return out
