"""The reference's OWN loops (train.py:train, train.py:adversarial_train, validation.py:val / val_GTA5), imported
unmodified from /root/reference, driven against THIS repo's drop-in `models/` and `utils.py`.

Runs only where the reference tree exists (the build container) and without a GPU: RTSDS_DRYRUN=1 records the
kernel launches instead of executing them, so tensors hold garbage — what is checked is that every call site
(`model(inputs)` tuple / tensor contract, `criterion(out, target)`, `loss += ...`, `.backward()`,
`optimizer.step()`, `.max(1)`, `F.softmax` -> discriminator, `requires_grad` freezing, `.detach()`,
`utils.fast_hist` on numpy arrays, `utils.poly_lr_scheduler`, `utils.tabular_print`) resolves and runs."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("RTSDS_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="reference tree not present")


class _Loader(list):
    """A list of batches is all the loops need from a DataLoader (len + iteration)."""


@pytest.fixture()
def ref_loops(monkeypatch):
    monkeypatch.setenv("RTSDS_DRYRUN", "1")
    cb = types.ModuleType("callbacks")

    class Callback:                                    # the 9 hooks of callbacks.py:1-30
        def __getattr__(self, name):
            if name.startswith("on_"):
                return lambda *a, **k: None
            raise AttributeError(name)

    cb.Callback = Callback
    monkeypatch.setitem(sys.modules, "callbacks", cb)
    for name in ("train", "validation"):
        sys.modules.pop(name, None)
    import utils as our_utils                          # this repo's drop-in, NOT the reference's utils.py

    monkeypatch.setitem(sys.modules, "utils", our_utils)
    # fast_hist would need a GPU: in this host-only test count with the oracle instead (call-site shape check only)
    from oracle import metrics_ref

    monkeypatch.setattr(our_utils, "fast_hist", lambda a, b, n: metrics_ref.fast_hist(np.nan_to_num(a).astype(np.int64), b, n))
    sys.path.append(REF)                               # appended: `models`, `utils` keep resolving to this repo
    try:
        val_mod = importlib.import_module("validation")
        train_mod = importlib.import_module("train")
    finally:
        sys.path.remove(REF)
    assert train_mod.__file__.startswith(REF) and val_mod.__file__.startswith(REF)
    yield train_mod, val_mod
    for name in ("train", "validation"):
        sys.modules.pop(name, None)


def _batches(n_batches, n, h, w):
    g = torch.Generator().manual_seed(0)
    return _Loader((torch.randn(n, 3, h, w, generator=g), torch.randint(0, 19, (n, 1, h, w), generator=g)) for _ in range(n_batches))


def test_reference_train_and_val_loops_run_on_the_drop_in_modules(ref_loops, capsys):
    train_mod, val_mod = ref_loops
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import _lib

    model = BiSeNet(19, "resnet18")
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    crit = torch.nn.CrossEntropyLoss(ignore_index=19)
    cb = sys.modules["callbacks"].Callback()
    _lib.lib().calls.clear()
    out = train_mod.train(0, model, _batches(2, 2, 64, 96), crit, opt, 1e-4, 100, 0.9, 1, "cpu", [cb])
    assert out is model
    calls = _lib.lib().calls
    assert calls.count("rtsds_conv2d_tc_wgrad") >= 2 * 20 and calls.count("rtsds_resize_to_nchw") == 2 * 3
    miou = val_mod.val(0, model, _batches(2, 1, 64, 96), 19, "cpu", [cb])
    assert np.isfinite(miou) or np.isnan(miou)
    miou2, table = val_mod.val_GTA5(0, model, _batches(1, 1, 64, 96), 19, [f"c{i}" for i in range(19)], [cb], "cpu")
    assert len(table) == 19


def test_reference_adversarial_loop_runs_on_the_drop_in_modules(ref_loops, monkeypatch, tmp_path):
    train_mod, _ = ref_loops
    from models.bisenet.build_bisenet import BiSeNet
    from models.domain_shift.adversarial.model import TinyDomainDiscriminator
    from rtsds_b200 import _lib

    monkeypatch.chdir(tmp_path)                         # best_generator.pth is written to the cwd (train.py:310-314)
    gen, dis = BiSeNet(19, "resnet18"), TinyDomainDiscriminator(19)
    gopt = torch.optim.Adam(gen.parameters(), lr=1e-4)
    dopt = torch.optim.Adam(dis.parameters(), lr=1e-4, weight_decay=1e-4)
    _lib.lib().calls.clear()
    train_mod.adversarial_train(2, 1, gen, dis, gopt, dopt, _batches(1, 2, 96, 128), _batches(1, 2, 64, 96),
                                torch.nn.CrossEntropyLoss(ignore_index=19), torch.nn.BCEWithLogitsLoss(), 0.1,
                                1e-4, 0.9, 0.9, 1e-4, 1, 19, [f"c{i}" for i in range(19)], _batches(1, 1, 64, 96), 1, "cpu", 10,
                                [sys.modules["callbacks"].Callback()])
    calls = _lib.lib().calls
    # per iteration: 3 discriminator forwards (target-for-G, source, target), 3 backward passes of it, one of
    # which (D frozen) only produces the input gradient for the generator
    assert calls.count("rtsds_disc_cls_fwd") == 2 * 3 and calls.count("rtsds_s2d_bwd") == 2 * 1
    assert calls.count("rtsds_s2d_weight_grad") == 2 * 2
    assert all(p.requires_grad for p in dis.parameters())


def test_reference_adversarial_loop_2_runs_on_the_drop_in_modules(ref_loops, monkeypatch, tmp_path):
    """train.py:322-500: the variant that trains D on target predictions and resizes every prediction to the
    target size with F.adaptive_avg_pool2d -- stock torch ops on the drop-in modules' outputs, generator forwards under
    torch.no_grad() in train mode, D output .requires_grad_(True)."""
    train_mod, _ = ref_loops
    if not hasattr(train_mod, "adversarial_train_2"):
        pytest.skip("reference has no adversarial_train_2")
    from models.bisenet.build_bisenet import BiSeNet
    from models.domain_shift.adversarial.model import TinyDomainDiscriminator
    from rtsds_b200 import _lib

    monkeypatch.chdir(tmp_path)
    gen, dis = BiSeNet(19, "resnet18"), TinyDomainDiscriminator(19)
    gopt = torch.optim.Adam(gen.parameters(), lr=1e-4)
    dopt = torch.optim.Adam(dis.parameters(), lr=1e-4, weight_decay=1e-4)
    _lib.lib().calls.clear()
    train_mod.adversarial_train_2(2, 1, gen, dis, gopt, dopt, _batches(1, 2, 96, 128), _batches(1, 2, 64, 96),
                                  torch.nn.CrossEntropyLoss(ignore_index=19), torch.nn.BCEWithLogitsLoss(), 0.1,
                                  1e-4, 0.9, 0.9, 1e-4, 1, 19, [f"c{i}" for i in range(19)], _batches(1, 1, 64, 96), 1, "cpu", 10,
                                  [sys.modules["callbacks"].Callback()])
    calls = _lib.lib().calls
    assert calls.count("rtsds_disc_cls_fwd") >= 3          # one D forward for the generator loss, two for D's own step
    assert calls.count("rtsds_conv2d_tc_wgrad") >= 20      # the generator's backward ran through the hand-written path
    assert all(p.requires_grad for p in dis.parameters())


def test_loop_restatements_used_on_the_gpu_box_follow_the_real_loops(ref_loops):
    """tests/ref_loop_stubs.py (what the -m gpu tests run, because /root/reference does not exist on the GPU box) against the
    REAL train.py:train / validation.py:val: identical sequence of library calls and identical callback invocations."""
    train_mod, val_mod = ref_loops
    import ref_loop_stubs
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import _lib

    class Rec:
        def __init__(self):
            self.calls = []

        def __getattr__(self, name):
            if name.startswith("on_"):
                return lambda *a, **k: self.calls.append((name, a[0] if a and isinstance(a[0], int) else None,
                                                          sorted(a[1].keys()) if len(a) > 1 and isinstance(a[1], dict) else None))
            raise AttributeError(name)

    seqs = []
    for train_fn, val_fn in ((train_mod.train, val_mod.val), (ref_loop_stubs.train, ref_loop_stubs.val)):
        torch.manual_seed(0)
        model = BiSeNet(19, "resnet18")
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        rec = Rec()
        _lib.lib().calls.clear()
        train_fn(0, model, _batches(2, 2, 64, 96), torch.nn.CrossEntropyLoss(ignore_index=19), opt, 1e-4, 100, 0.9, 1, "cpu", [rec])
        val_fn(0, model, _batches(2, 1, 64, 96), 19, "cpu", [rec])
        seqs.append((list(_lib.lib().calls), rec.calls, opt.param_groups[0]["lr"]))
    assert seqs[0][0] == seqs[1][0]            # kernel-launch sequence
    assert seqs[0][1] == seqs[1][1]            # callback names, batch indices, dict keys
    assert seqs[0][2] == seqs[1][2]            # poly LR applied identically
