"""The example generator step (examples/generator_step.py) runs end to end on the GPU: the three fused loss terms sit
inside autocast + GradScaler like the reference's step, every generator parameter receives a finite gradient, and
the FFT term really contributes to it (upstream it is a constant)."""

import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_example():
    spec = importlib.util.spec_from_file_location("generator_step", os.path.join(ROOT, "examples", "generator_step.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_generator_step_trains_through_the_fused_losses():
    import tfc_gan_b200 as tfc
    ex = load_example()
    torch.manual_seed(0)
    G, D = ex.SmallUNet(16).cuda(), ex.SmallPatchDiscriminator(16).cuda()
    losses = (tfc.SpectralLoss(grid=4, weight=1 / 100, input_scale=255.0), tfc.PatchTripletLoss(grid=4))
    n = 4
    real_A = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1)
    real_B = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1)
    B_tf = (real_B + 0.1 * torch.randn_like(real_B)).clamp(-1, 1)
    T_B = tfc.vectorize_temps(real_B * 0.5 + 0.5)[:, 0]
    grads = {}
    for use_fft in (True, False):
        opt = torch.optim.SGD(G.parameters(), lr=0.0)        # lr 0: both runs see the same weights
        scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
        import numpy as np
        np.random.seed(5)                                    # same triplet negatives in both runs
        logs = ex.generator_step(G, D, opt, scaler, real_A, real_B, T_B, B_tf, losses, use_fft=use_fft)
        assert all(torch.isfinite(v).item() for v in logs.values())
        g = torch.cat([p.grad.flatten().float() for p in G.parameters()])
        assert torch.isfinite(g).all() and float(g.abs().sum()) > 0
        grads[use_fft] = g / 1024.0
    assert float(logs["fft"]) == 0.0
    diff = (grads[True] - grads[False]).norm() / grads[True].norm()
    assert float(diff) > 1e-3, "the FFT loss does not reach the generator"
