"""Stand-alone forwards of the reference's sub-modules (rtsds_b200/module_ops.py) against the CPU oracle's functions for
the same modules, with weights copied by state_dict.  fp32 check mode: 1e-4; bf16: 3e-2."""
import pytest
import torch

from oracle import bisenet_ref, deeplab_ref, weights

from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _bisenet(precision):
    from models.bisenet.build_bisenet import BiSeNet

    m = BiSeNet(19, "resnet18")
    m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(11)))
    for mod in m.modules():
        mod.rtsds_precision = precision
    return m.cuda()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
@pytest.mark.parametrize("train", [False, True])
def test_bisenet_submodules_standalone(cuda, precision, tol, train):
    m = _bisenet(precision)
    m.train(train)
    sd = weights.clone_state(weights.bisenet_r18_state(11))
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 96, 128, generator=g)
    with torch.no_grad():
        # ConvBlock and Spatial_path (build_bisenet.py:16-18, :28-32)
        got = m.saptial_path.convblock1(x.cuda()).cpu()
        ref = bisenet_ref.conv_block(x, weights.clone_state(sd), "saptial_path.convblock1", 2, train)
        assert rel_err(got, ref) < tol
        sx = m.saptial_path(x.cuda()).cpu()
        sx_ref = bisenet_ref.spatial_path(x, weights.clone_state(sd), train)
        assert sx.shape == sx_ref.shape and rel_err(sx, sx_ref) < tol
        # context path (build_contextpath.py:18-29)
        f3, f4, tail = (t.cpu() for t in m.context_path(x.cuda()))
        r3, r4, rt = bisenet_ref.context_path_r18(x, weights.clone_state(sd), train)
        assert f3.shape == r3.shape and f4.shape == r4.shape and tail.shape == rt.shape
        ctol = tol if precision == "fp32" else 0.15       # 17 stacked bf16 layers on a 3x4 map
        assert rel_err(f3, r3) < ctol and rel_err(f4, r4) < ctol and rel_err(tail, rt) < ctol
        # one BasicBlock
        xb = torch.randn(2, 64, 24, 32, generator=g)
        got = m.context_path.layer2[0](xb.cuda()).cpu()
        ref = bisenet_ref.basic_block(xb, weights.clone_state(sd), "context_path.features.layer2.0", 2, train)
        assert rel_err(got, ref) < tol
        # AttentionRefinementModule (:44-53) and FeatureFusionModule (:71-81)
        xa = torch.randn(2, 256, 6, 8, generator=g)
        got = m.attention_refinement_module1(xa.cuda()).cpu()
        ref = bisenet_ref.arm(xa, weights.clone_state(sd), "attention_refinement_module1", train)
        assert rel_err(got, ref) < tol
        a, b = torch.randn(2, 256, 12, 16, generator=g), torch.randn(2, 768, 12, 16, generator=g)
        got = m.feature_fusion_module(a.cuda(), b.cuda()).cpu()
        ref = bisenet_ref.ffm(a, b, weights.clone_state(sd), train)
        assert got.shape == ref.shape and rel_err(got, ref) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_deeplab_submodules_and_upsampler_standalone(cuda, precision, tol):
    from models.deeplabv2.deeplabv2 import get_deeplab_v2
    from models.domain_shift.adversarial.model import UpSampler

    m = get_deeplab_v2(19, pretrain=False)
    sd = weights.deeplab_state(12)
    m.load_state_dict(weights.clone_state(sd))
    for mod in m.modules():
        mod.rtsds_precision = precision
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(8)
    with torch.no_grad():
        x = torch.randn(1, 512, 17, 21, generator=g)
        got = m.layer3[0](x.cuda()).cpu()                                    # stride 1, dilation 2, with downsample
        ref = deeplab_ref.bottleneck(x, weights.clone_state(sd), "layer3.0", 1, 2, False)
        assert got.shape == ref.shape and rel_err(got, ref) < tol
        x = torch.randn(1, 256, 33, 41, generator=g)
        got = m.layer2[0](x.cuda()).cpu()                                    # stride 2 on conv1
        ref = deeplab_ref.bottleneck(x, weights.clone_state(sd), "layer2.0", 2, 1, False)
        assert got.shape == ref.shape and rel_err(got, ref) < tol
        x = torch.randn(1, 2048, 33, 41, generator=g)
        got = m.layer6(x.cuda()).cpu()                                       # ASPP
        ref = deeplab_ref.classifier(x, weights.clone_state(sd))
        assert got.shape == ref.shape and rel_err(got, ref) < tol
        up = UpSampler(19)
        up.rtsds_precision = precision
        up = up.cuda()
        x = torch.randn(2, 19, 9, 13, generator=g)
        got = up(x.cuda()).cpu()
        ref = torch.nn.functional.conv2d(torch.nn.functional.interpolate(x, scale_factor=8, mode="bilinear"),
                                         up.conv.weight.detach().cpu(), up.conv.bias.detach().cpu())
        assert got.shape == ref.shape and rel_err(got, ref) < (1e-4 if precision == "fp32" else 2e-2)


def test_submodule_forward_refuses_autograd(cuda):
    from rtsds_b200 import RtsdsError

    m = _bisenet("fp32").train()
    with pytest.raises(RtsdsError):
        m.saptial_path(torch.zeros(2, 3, 64, 64, device="cuda"))
