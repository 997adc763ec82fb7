"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the
ctypes table covers the header, the product never imports the oracle, the drop-in modules keep the
reference's parameter tree, and the host-side planning logic issues the expected launch sequence."""
import collections
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "rtsds_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(rtsds_[a-z0-9_]+)\s*\(", hdr))


def test_library_builds_loads_and_exports_header_symbols():
    from rtsds_b200 import _lib
    from rtsds_b200.build import build

    path = build()
    handle = ctypes.CDLL(str(path))
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/rtsds_b200.h but not exported"
    assert syms == set(_lib.SIGNATURES), syms ^ set(_lib.SIGNATURES)
    handle.rtsds_abi_version.restype = ctypes.c_int
    assert handle.rtsds_abi_version() == _lib.ABI_VERSION == 3
    handle.rtsds_conv_cout_pad.restype = ctypes.c_int
    assert [handle.rtsds_conv_cout_pad(c) for c in (1, 19, 32, 33, 64, 65, 128, 129, 512)] == [32, 32, 32, 64, 64, 128, 128, 256, 512]


def test_sass_contains_tcgen05_and_tma():
    """The conv kernel must be a real Blackwell kernel: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    from rtsds_b200.build import LIB

    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"):
        assert mnemonic in out, mnemonic
    assert "HMMA.16816" not in out and "HGMMA" not in out   # no legacy mma.sync / wgmma path


def test_product_never_imports_oracle_or_has_cpu_fallback():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for base in ("rtsds_b200", "models"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not pat.search(src), f"{f} imports the oracle"
    for f in ("utils.py", "validation.py", "train.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not pat.search(open(p).read()), f"{f} imports the oracle"


def test_cpu_tensor_is_rejected_not_computed():
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import RtsdsError

    m = BiSeNet(19, "resnet18").eval()
    with pytest.raises(RtsdsError):
        m(torch.zeros(1, 3, 64, 64))


def test_bisenet_parameter_tree_matches_reference_contract():
    from models.bisenet.build_bisenet import BiSeNet
    from oracle import weights

    m = BiSeNet(19, "resnet18")
    sd = m.state_dict()
    ref = weights.bisenet_r18_state(0)             # key set verified against the real reference (gen_golden)
    assert sorted(sd.keys()) == sorted(ref.keys()) and len(sd) == 290
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in sd)
    assert sum(p.numel() for p in m.parameters()) == 12581672
    assert sd["context_path.conv1.weight"].data_ptr() == sd["context_path.features.conv1.weight"].data_ptr()
    assert len(m.mul_lr) == 7 and m.mul_lr[0] is m.saptial_path
    # init_weight (reference :130-139): BN of the non-backbone modules is gamma=1, beta=0
    assert torch.all(m.saptial_path.convblock1.bn.weight == 1) and torch.all(m.feature_fusion_module.convblock.bn.bias == 0)
    assert m.load_state_dict(ref).missing_keys == []


def test_plan_launch_sequence_dry_run(monkeypatch):
    """RTSDS_DRYRUN records launches without a GPU: 22 tensor-core convs (2 of them with the ARM global pool fused into the
    epilogue), 1 fused stem pair, ONE kernel for both ARM gates + gated resizes, ONE for FFM attention + final conv + x8
    resize ... per eval forward: 27 launches of this library (34 in round 1), no memset, no atomics."""
    monkeypatch.setenv("RTSDS_DRYRUN", "1")
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import _lib

    m = BiSeNet(19, "resnet18").eval()
    m.rtsds_cuda_graph = False
    out = m(torch.zeros(1, 3, 512, 1024))
    assert out.shape == (1, 19, 512, 1024)
    c = collections.Counter(_lib.lib().calls)
    assert c["rtsds_conv2d_tc_fwd"] == 20 and c["rtsds_conv2d_tc_fwd_gap"] == 2
    assert c["rtsds_stem_pair_tc_fwd"] == 1 and c["rtsds_maxpool3x3s2_fwd"] == 1
    assert c["rtsds_stem_conv_fwd"] == 0          # both stems run as one tensor-core kernel
    assert c["rtsds_arm_gate_resize"] == 1 and c["rtsds_arm_gate"] == 0 and c["rtsds_gate_resize_nhwc"] == 0 and c["rtsds_global_avgpool"] == 0
    assert c["rtsds_tapn_gather"] == 1 and c["rtsds_ffm_head_resize"] == 1 and c["rtsds_ffm_head"] == 0 and c["rtsds_resize_to_nchw"] == 0
    assert c["rtsds_bn_fold"] == 24
    forward_calls = [k for k in _lib.lib().calls if not k.startswith(("rtsds_pack", "rtsds_bn_fold", "rtsds_stem_pack", "rtsds_tapn_weights",
                                                                      "rtsds_conv_cout_pad", "rtsds_conv2d_tc_workspace", "rtsds_check",
                                                                      "rtsds_scale_packed", "rtsds_launch_count", "rtsds_conv2d_tc_gap_parts", "rtsds_tapn_gather_parts"))]
    assert len(forward_calls) == 27, (len(forward_calls), collections.Counter(forward_calls))
    _lib.lib().calls.clear()
    m(torch.zeros(1, 3, 512, 1024))                 # weights unchanged: no repack on the second call
    c = collections.Counter(_lib.lib().calls)
    assert c["rtsds_pack_conv_weight"] == 0 and c["rtsds_conv2d_tc_fwd"] == 20
    # 720x1280 (GTA5) gives the odd 45x80 / 23x40 feature maps and a x8 head
    m.train()
    _lib.lib().calls.clear()
    outs = m(torch.zeros(2, 3, 720, 1280))
    assert [tuple(o.shape) for o in outs] == [(2, 19, 720, 1280)] * 3
    c = collections.Counter(_lib.lib().calls)
    assert c["rtsds_conv2d_tc_fwd"] == 24 and c["rtsds_bn_finalize"] + c["rtsds_bn_finalize_apply"] == 24 and c["rtsds_resize_to_nchw"] == 3
    m.rtsds_precision = "fp32"
    _lib.lib().calls.clear()
    m.eval()(torch.zeros(1, 3, 64, 96))
    c = collections.Counter(_lib.lib().calls)
    assert c["rtsds_conv2d_simt_fwd"] == 22 and c["rtsds_conv2d_tc_fwd"] == 0


def test_eval_forward_refuses_input_gradients_dry_run(monkeypatch):
    """the reference's eval forward is an autograd graph; the eval plan here is inference only, so a
    caller who asks for gradients w.r.t. the input gets an error, not a silently detached tensor.  Under no_grad (what
    validation.py does) and for parameters that merely require grad it runs."""
    monkeypatch.setenv("RTSDS_DRYRUN", "1")
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import RtsdsError

    m = BiSeNet(19, "resnet18").eval()
    m.rtsds_cuda_graph = False
    x = torch.zeros(1, 3, 64, 64, requires_grad=True)
    with pytest.raises(RtsdsError):
        m(x)
    with torch.no_grad():
        assert m(x).shape == (1, 19, 64, 64)
    out = m(torch.zeros(1, 3, 64, 64))
    assert out.shape == (1, 19, 64, 64) and not out.requires_grad


def test_weights_epoch_and_plan_slots_dry_run(monkeypatch):
    """Host logic without a GPU: (1) a backward pass makes every plan re-pack its operands at the next forward even
    when no tensor version counter moved (torch's fused optimizers); (2) a second train forward of the same shape that
    is alive together with the first gets its own plan, and the slot is reused once its backward has run."""
    monkeypatch.setenv("RTSDS_DRYRUN", "1")
    from models.bisenet.build_bisenet import BiSeNet
    from rtsds_b200 import _lib, weights_epoch

    m = BiSeNet(19, "resnet18").train()
    x = torch.zeros(2, 3, 64, 96)
    y = torch.zeros(2, 64, 96, dtype=torch.int64)
    ce = lambda outs: sum(torch.nn.functional.cross_entropy(t, y, ignore_index=19) for t in outs)
    calls = _lib.lib().calls
    calls.clear()
    outs = m(x)
    assert calls.count("rtsds_pack_conv_weights_batch") == 1
    e0 = weights_epoch.value()
    ce(outs).backward()
    assert weights_epoch.value() == e0 + 1
    calls.clear()
    outs = m(x)                                       # nothing touched the parameters, but a backward ran: re-pack
    assert calls.count("rtsds_pack_conv_weights_batch") == 1 and len(m._rtsds_train_plans) == 1
    calls.clear()
    with torch.no_grad():
        m(x)                                          # first plan still awaits its backward: second slot, own operands
    assert len(m._rtsds_train_plans) == 2 and calls.count("rtsds_pack_conv_weights_batch") == 1
    outs2 = m(x)                                      # the no-grad forward left slot 2 free: reused, no third plan
    assert len(m._rtsds_train_plans) == 2
    (ce(outs) + ce(outs2)).backward()                 # both pending backward passes run on their own activations
    assert weights_epoch.value() == e0 + 3
    m.eval()
    calls.clear()
    m.rtsds_cuda_graph = False
    m(x); m(x)
    assert calls.count("rtsds_bn_fold") == 24          # eval plan folded its BatchNorms once, not on the second call


def test_geometry_helpers():
    from rtsds_b200 import ops

    assert ops.conv_out_size(720, 3, 2, 1) == 360 and ops.conv_out_size(45, 3, 2, 1) == 23
    assert ops.conv_out_size(65, 3, 1, 6, 6) == 65
    for i in (7, 8, 64, 65, 129, 256, 257):
        for ceil in (False, True):
            ref = torch.nn.functional.max_pool2d(torch.zeros(1, 1, i, i), 3, 2, 1, ceil_mode=ceil).shape[-1]
            assert ops.maxpool_out_size(i, ceil) == ref
