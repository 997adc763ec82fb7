"""BASELINE.json configs[0] verbatim (SURVEY 8c/8d "Config 1"): torch.manual_seed(42) -> BiSeNet(19, 'resnet18')
(the constructor's own random init), data generator seed 1234, x ~ N(0,1) [2,3,512,1024], labels uniform in [0,19].
tests/golden/config1.npz holds what the REAL reference produced for exactly this (oracle/gen_golden.py config1)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1.npz")
S_LOGIT, S_ARGMAX = 16, 4       # sub-sampling of the stored logits / argmax maps


def inputs():
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 512, 1024, generator=g)
    y = torch.randint(0, 20, (2, 512, 1024), generator=g)
    return x, y


def seeded_model():
    """The drop-in BiSeNet under the reference's seeding; same RNG draws as the reference constructor
    (tests/test_config1_cpu.py pins the weights against the golden checksums)."""
    from models.bisenet.build_bisenet import BiSeNet

    torch.manual_seed(42)
    return BiSeNet(19, "resnet18")


def golden():
    return np.load(GOLDEN)


def state_clone(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
