"""Seeded configuration sweep on the GPU: random shapes / options / dtypes / views for every entry point, each compared
with the fp64 oracle at BASELINE.json's tolerances (loss rel <= 1e-4, gradient L2-rel <= 1e-3; 16-bit I/O: output
rounding of the gradient)."""

import numpy as np
import pytest
import torch

import oracle
import tfc_gan_b200 as tfc
from oracle import regional as oreg
from oracle import temperature as otemp
from oracle import triplet as otri
from util import l2rel

pytestmark = pytest.mark.gpu


def draw(seed):
    rs = np.random.RandomState(1000 + seed)
    side = int(rs.choice([64, 128, 256]))
    grid = int(rs.choice([g for g in (1, 2, 4) if side // g >= 16]))
    n = int(rs.randint(1, 6))
    c = int(rs.choice([1, 3]))
    dtype = torch.float16 if rs.rand() < 0.3 else torch.float32
    view = rs.rand() < 0.35
    kind = rs.choice(["uniform", "tanh", "smooth"])
    return rs, side, grid, n, c, dtype, view, kind


def tensors(rs, n, c, side, kind, dtype, view):
    def one():
        if kind == "uniform":
            a = rs.uniform(-1, 1, (n, c, side, side))
        elif kind == "tanh":
            a = np.tanh(rs.normal(size=(n, c, side, side)))
        else:
            a = rs.normal(size=(n, c, side, side))
            a = np.cumsum(np.cumsum(a, -1), -2)
            a = a / np.abs(a).max()
        t = torch.from_numpy(a.astype(np.float32)).cuda().to(dtype)
        if view:  # a strided window of a larger allocation (row pitch 2 * side, 16-byte aligned offset)
            big = torch.zeros(n, c, side + 8, 2 * side, device="cuda", dtype=dtype)
            big[:, :, 4:4 + side, 8:8 + side] = t
            t = big[:, :, 4:4 + side, 8:8 + side]
        return t
    return one(), one()


@pytest.mark.parametrize("seed", range(16))
def test_fft_loss_sweep(seed):
    rs, side, grid, n, c, dtype, view, kind = draw(seed)
    opt = dict(channels=str(rs.choice(["luma", "rgb"])), distance=str(rs.choice(["l1", "mse"])), use_phase=bool(rs.rand() < 0.7),
               patch_reduce=str(rs.choice(["mean", "sum"])))
    if opt["distance"] == "mse" and opt["use_phase"]:
        opt["use_phase"] = False  # the squared phase distance amplifies branch-cut noise beyond any fixed tolerance
    f, r = tensors(rs, n, c, side, kind, dtype, view)
    F = f.detach().requires_grad_(True)
    loss = tfc.spectral_loss(F, r, grid=grid, weight=0.05, input_scale=100.0, **opt)
    loss.backward()
    l, _, _, g = oracle.spectral_loss_and_grad_r1(f.double().cpu().numpy(), r.double().cpu().numpy(), grid=grid, weight=0.05,
                                                  input_scale=100.0, **opt)
    assert float(loss) == pytest.approx(l, rel=1e-4), (side, grid, n, c, dtype, view, kind, opt)
    assert l2rel(F.grad.double().cpu().numpy(), g) <= (1e-3 if dtype == torch.float32 else 4e-3), (side, grid, n, c, dtype, opt)


@pytest.mark.parametrize("seed", range(8))
def test_patch_triplet_sweep(seed):
    rs, side, grid, n, c, dtype, view, kind = draw(100 + seed)
    f, r = tensors(rs, n, c, side, kind, dtype, view)
    neg = [int(k) for k in rs.randint(grid * grid, size=grid * grid)]
    margin = float(rs.choice([0.2, 1.0, 3.0]))
    F = f.detach().requires_grad_(True)
    loss = tfc.patch_triplet_loss(F, r, neg, grid=grid, margin=margin, weight=2.0)
    loss.backward()
    wl, _, _, g = otri.patch_triplet_loss_and_grad(f.double().cpu().numpy(), r.double().cpu().numpy(), neg, grid=grid, margin=margin,
                                                   weight=2.0)
    assert float(loss) == pytest.approx(wl, rel=1e-4), (side, grid, n, c, dtype, view)
    if np.abs(g).max() > 0:
        assert l2rel(F.grad.double().cpu().numpy(), g) <= (1e-3 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("seed", range(4))
def test_temperature_and_regional_sweep(seed):
    rs = np.random.RandomState(2000 + seed)
    n = int(rs.randint(1, 5))
    dtype = torch.float16 if seed % 2 else torch.float32
    mk = lambda: torch.from_numpy(rs.uniform(0, 1, (n, 3, 256, 256)).astype(np.float32)).cuda().to(dtype)
    f, r, j = mk(), mk(), mk()
    F = f.detach().requires_grad_(True)
    loss = tfc.temperature_triplet_loss(F, r, j, weight=10.0, input_scale=255.0)
    loss.backward()
    wl, _, _, g = otemp.temperature_triplet(f.double().cpu().numpy(), r.double().cpu().numpy(), j.double().cpu().numpy(), quantize=False,
                                            weight=10.0, input_scale=255.0)
    assert float(loss) == pytest.approx(wl, rel=1e-4)
    assert l2rel(F.grad.double().cpu().numpy(), g) <= (1e-3 if dtype == torch.float32 else 4e-3)
    F2 = (f * 2 - 1).detach().requires_grad_(True)
    loss = tfc.regional_spectral_loss(F2, r * 2 - 1, weight=0.01, input_scale=255.0)
    loss.backward()
    l, _, _, g = oreg.regional_loss_and_grad_r1(F2.detach().double().cpu().numpy(), (r * 2 - 1).double().cpu().numpy(), weight=0.01,
                                                input_scale=255.0)
    assert float(loss) == pytest.approx(l, rel=1e-4)
    assert l2rel(F2.grad.double().cpu().numpy(), g) <= (1e-3 if dtype == torch.float32 else 4e-3)
