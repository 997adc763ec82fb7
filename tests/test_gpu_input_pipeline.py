"""device-side input pipeline (csrc/input.cu, rtsds_b200/input_pipeline.py) against the torchvision
transforms the reference composes at main.py:60-108 — read_image(...).float() -> Resize(antialias=True) -> Normalize,
and read_image(...).long() -> Resize(antialias=True) -> IntRangeTransformer — evaluated by torchvision on the CPU."""
import pytest
import torch
from torchvision import transforms

from rtsds_b200 import ops
from rtsds_b200.input_pipeline import MEAN, STD, DeviceInputPipeline

pytestmark = pytest.mark.gpu


class IntRangeTransformer:               # utils.py:67-75 restated (the oracle side of this test)
    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def __call__(self, t):
        return torch.clamp(t, self.lo, self.hi).long()


@pytest.mark.parametrize("src,size", [((1024, 2048), (512, 1024)),       # Cityscapes: exact 2x down (config.yaml image_size)
                                      ((1052, 1914), (720, 1280)),       # GTA5: non-integer factor
                                      ((96, 130), (96, 130)),            # identity
                                      ((61, 77), (128, 200)),            # up-scaling
                                      ((700, 333), (64, 96))])           # strong anisotropic down-scaling
def test_image_pipeline_matches_torchvision(cuda, src, size):
    g = torch.Generator().manual_seed(src[0])
    u8 = torch.randint(0, 256, (2, 3, *src), dtype=torch.uint8, generator=g)
    tf = transforms.Compose([transforms.Resize(list(size), antialias=True), transforms.Normalize(mean=list(MEAN), std=list(STD))])
    ref = torch.stack([tf(img.float()) for img in u8])                   # the Dataset applies it per sample
    out = DeviceInputPipeline(size).images(u8.cuda()).cpu()
    assert out.shape == ref.shape and out.dtype == torch.float32
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-5, err


@pytest.mark.parametrize("src,size,clamp", [((1024, 2048), (512, 1024), (0, 19)), ((1052, 1914), (720, 1280), None),
                                            ((64, 96), (64, 96), (0, 19)), ((333, 517), (100, 150), (0, 19))])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_label_pipeline_matches_torchvision(cuda, src, size, clamp, dtype):
    g = torch.Generator().manual_seed(7)
    lab = torch.randint(0, 34, (2, 1, *src), generator=g)
    lab[:, :, : src[0] // 4] = 255                                      # a void band, as in the Cityscapes label ids
    tf = [transforms.Resize(list(size), antialias=True)]
    if clamp is not None:
        tf.append(IntRangeTransformer(*clamp))
    tf = transforms.Compose(tf)
    ref = torch.stack([tf(l.long()) for l in lab]).long().squeeze(1)     # train.py:72 .squeeze(1)
    out = DeviceInputPipeline(size).labels(lab.to(dtype).cuda(), clamp=clamp).cpu()
    assert out.dtype == torch.int64 and out.shape == ref.shape
    # fp32 accumulation order differs from ATen's separable passes by a few 1e-5 grey levels: a resized value within that
    # distance of k + 0.5 rounds the other way (never by more than one class id)
    diff = (out - ref).abs()
    assert diff.max().item() <= 1 and (diff != 0).float().mean().item() <= 3e-4, (diff.max().item(), (diff != 0).float().mean().item())


def test_model_accepts_raw_uint8_frames(cuda):
    """model(uint8 frame) == model(Normalize(frame.float())): the forward converts + normalises raw frames on the device."""
    from models.bisenet.build_bisenet import BiSeNet
    from oracle import weights

    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (1, 3, 256, 384), dtype=torch.uint8, generator=g).cuda()
    m = BiSeNet(19, "resnet18")
    m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(9)))
    m = m.cuda().eval()
    m.rtsds_input_norm = ((123.7, 116.3, 103.5), (58.4, 57.1, 57.4))     # ImageNet statistics on the 0..255 scale
    x = DeviceInputPipeline(None, *m.rtsds_input_norm).images(u8)
    a = m(u8)
    b = m(x)
    # same fp16 patch values; the only run-to-run difference is the order of the fp32 atomics of the global average pools
    assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item()
    # uint8 class map
    p8 = torch.empty(1, 256, 384, dtype=torch.uint8, device="cuda")
    p64 = torch.empty(1, 256, 384, dtype=torch.int64, device="cuda")
    ops.argmax_hist(a, None, None, p8)
    ops.argmax_hist(a, None, None, p64)
    assert torch.equal(p8.long(), p64) and torch.equal(p64, a.argmax(1))
    # the other modes convert on the device first
    m.rtsds_precision = "fp32"
    c = m(u8)
    assert (c - m(x)).abs().max().item() <= 1e-5 * c.abs().max().item()


def test_pipelined_segmenter_uint8_io(cuda):
    from models.bisenet.build_bisenet import BiSeNet
    from oracle import weights
    from rtsds_b200.serving import PipelinedSegmenter

    m = BiSeNet(19, "resnet18")
    m.load_state_dict(weights.clone_state(weights.bisenet_r18_state(9)))
    m = m.cuda().eval()
    m.rtsds_input_norm = ((123.7, 116.3, 103.5), (58.4, 57.1, 57.4))
    g = torch.Generator().manual_seed(4)
    frames = [torch.randint(0, 256, (1, 3, 128, 192), dtype=torch.uint8, generator=g).pin_memory() for _ in range(7)]
    want = [m(f.cuda()).argmax(1).cpu() for f in frames]
    pipe = PipelinedSegmenter(m, 1, 128, 192, depth=3, lanes=2, uint8_io=True)
    got = []
    for f in frames:
        r = pipe.submit(f)
        if r is not None:
            got.append(r.clone())
    got += [r.clone() for r in pipe.drain()]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.dtype == torch.uint8 and torch.equal(a.long(), b)
