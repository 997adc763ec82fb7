"""End-to-end parity of the CUDA BiSeNet path (drop-in models.bisenet.build_bisenet.BiSeNet)
against (a) golden vectors produced by the REAL reference and (b) the CPU oracle on the
same seeded inputs and weights.  Tolerances are BASELINE.json's: logits rel <= 2e-2 in
bf16, <= 1e-4 in the fp32 check mode, argmax agreement >= 99.9 %."""
import os

import numpy as np
import pytest
import torch

from oracle import bisenet_ref, weights

from gpu_util import rel_err

pytestmark = pytest.mark.gpu
SUB = 3


def _input(seed, n, h, w):
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(n, 3, h, w, generator=g)
    y = torch.randint(0, 20, (n, h, w), generator=g)
    return x, y


def _model(seed, precision):
    from models.bisenet.build_bisenet import BiSeNet

    m = BiSeNet(19, "resnet18")
    m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(seed)))
    m.rtsds_precision = precision
    return m.cuda()


@pytest.mark.parametrize("name", ["bisenet_64x96", "bisenet_72x104"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_eval_forward_vs_reference_golden(cuda, golden_dir, name, precision, tol):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    x, _ = _input(seed, n, h, w)
    m = _model(seed, precision).eval()
    out = m(x.cuda())
    assert out.shape == (n, 19, h, w) and out.dtype == torch.float32
    got = out[..., ::SUB, ::SUB].cpu()
    ref = torch.from_numpy(gold["eval_result"])
    assert rel_err(got, ref) < tol, rel_err(got, ref)
    agree = (out.argmax(1)[..., ::SUB, ::SUB].cpu().numpy() == gold["eval_argmax"]).mean()
    assert agree >= (0.999 if precision == "fp32" else 0.97), agree   # tiny maps: few hundred pixels
    # graph replay and eager execution give the same answer; second call reuses the plan
    out2 = m(x.cuda())
    assert torch.equal(out, out2)
    m.rtsds_cuda_graph = False
    m.__dict__.pop("_rtsds_plans")
    out3 = m(x.cuda())
    assert torch.equal(out, out3)


@pytest.mark.parametrize("name", ["bisenet_64x96", "bisenet_72x104"])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_train_forward_vs_reference_golden(cuda, golden_dir, name, precision, tol):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    n, h, w = (int(v) for v in gold["shape"])
    seed = int(gold["seed"][0])
    x, _ = _input(seed, n, h, w)
    m = _model(seed, precision).train()
    res, s1, s2 = m(x.cuda())
    for k, t in (("train_result", res), ("train_sup1", s1), ("train_sup2", s2)):
        assert t.shape == (n, 19, h, w)
        e = rel_err(t[..., ::SUB, ::SUB].cpu(), torch.from_numpy(gold[k]))
        assert e < tol, (k, e)
    bufs = dict(m.named_buffers())
    for k in gold.files:
        if k.startswith("buf:"):
            e = rel_err(bufs[k[4:]].cpu(), torch.from_numpy(gold[k]))
            assert e < (1e-4 if precision == "fp32" else 2e-2), (k, e)
    assert int(bufs["saptial_path.convblock1.bn.num_batches_tracked"]) == 1


@pytest.mark.parametrize("n,h,w", [(1, 512, 1024), (2, 256, 512)])
def test_eval_forward_full_size_vs_oracle(cuda, n, h, w):
    """BASELINE config 1/2 shapes: bf16 CUDA path vs the CPU oracle on the same seeded weights."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.bisenet_r18_state(42)
    x, _ = _input(42, n, h, w)
    with torch.no_grad():
        ref = bisenet_ref.bisenet_forward(x, weights.clone_state(sd), train=False)
    for precision, tol, agree_min in (("fp32", 1e-4, 0.9999), ("bf16", 2e-2, 0.999)):
        m = _model(42, precision).eval()
        out = m(x.cuda()).cpu()
        e = rel_err(out, ref)
        agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
        assert e < tol, (precision, e)
        assert agree >= agree_min, (precision, agree)


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors — never compute on the CPU."""
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import RtsdsError

    m = BiSeNet(19, "resnet18").eval()
    with pytest.raises(RtsdsError):
        m(torch.zeros(1, 3, 64, 64))
