"""GPU tests of the drop-in module path (``-m gpu``): the expected ``grad_output`` folded into the producing launch,
arbitrary ``grad_output`` still honoured, gradient accumulation into a shared buffer, the loader-quadrant pointer
path, and the spectra workspace sizing."""

import numpy as np
import pytest
import torch

import oracle
import tfc_gan_b200 as tfc
from tfc_gan_b200 import compat
from util import l2rel

pytestmark = pytest.mark.gpu


def _pair(n, side=256, seed=3, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    f = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g).to(dtype)
    r = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g).to(dtype)
    return f, r


@pytest.mark.parametrize("grid", [4, 2, 1])
def test_folded_grad_scale_equals_plain_backward(grid):
    """``SpectralLoss(grad_scaler=..., loss_multiplier=...)`` produces ``unit * scale * multiplier`` directly and the
    backward pass launches only the scalar check (no pass over the tensor); the values equal the unfolded path."""
    f, r = _pair(6, seed=grid)
    scaler = torch.amp.GradScaler("cuda", init_scale=4096.0)
    plain = tfc.SpectralLoss(grid=grid, input_scale=255.0)
    fa = f.clone().requires_grad_(True)
    scaler.scale(0.01 * plain(fa, r)).backward()
    folded = tfc.SpectralLoss(grid=grid, input_scale=255.0, grad_scaler=scaler, loss_multiplier=0.01)
    fb = f.clone().requires_grad_(True)
    loss = folded(fb, r)
    tfc.reset_launch_count()
    scaler.scale(0.01 * loss).backward()
    assert tfc.launch_count() == 1  # tfcfft_grad_rescale: compares two scalars and exits
    assert l2rel(fb.grad.cpu().numpy(), fa.grad.cpu().numpy()) <= 2e-6
    # and the unit gradient against the oracle
    l, _, _, g = oracle.spectral_loss_and_grad_r1(f.cpu().numpy(), r.cpu().numpy(), grid=grid, input_scale=255.0)
    assert loss.item() == pytest.approx(l, rel=1e-4)
    assert l2rel(fb.grad.cpu().numpy() / (4096.0 * 0.01), g) <= 1e-3


def test_arbitrary_grad_output_and_retained_graph():
    f, r = _pair(4, seed=11)
    want = tfc.spectral_loss_and_grad(f, r, grid=4, input_scale=255.0)[2]
    for expected in (None, 7.0):  # nothing folded / a wrong guess folded: both must give grad_output * unit
        fa = f.clone().requires_grad_(True)
        loss = tfc.spectral_loss(fa, r, grid=4, input_scale=255.0, grad_scale=expected)
        (loss * 3.5).backward(retain_graph=True)
        assert l2rel(fa.grad.cpu().numpy(), 3.5 * want.cpu().numpy()) <= 2e-6
        fa.grad = None
        (loss * -2.0).backward()  # second pass through the retained graph: the buffer is produced again
        assert l2rel(fa.grad.cpu().numpy(), -2.0 * want.cpu().numpy()) <= 2e-6


def test_non_leaf_input_and_gradient_flow():
    """fake is a generator output in the training script: the returned buffer feeds the previous node."""
    f, r = _pair(3, seed=12)
    w = torch.full((1,), 0.5, device="cuda", requires_grad=True)
    loss = tfc.spectral_loss(f * w, r, grid=4, input_scale=255.0)
    loss.backward()
    unit = tfc.spectral_loss_and_grad(f * 0.5, r, grid=4, input_scale=255.0)[2]
    assert w.grad.item() == pytest.approx(float((unit * f).sum()), rel=1e-4)


@pytest.mark.parametrize("grid", [4, 2, 1])
def test_accumulate_into_shared_gradient_buffer(grid):
    f, r = _pair(5, seed=20 + grid)
    a = tfc.spectral_loss_and_grad(f, r, grid=grid, input_scale=255.0)[2]
    base = torch.randn_like(f)
    acc = base.clone()
    tfc.spectral_loss_and_grad(f, r, grid=grid, input_scale=255.0, accumulate_into=acc)
    assert l2rel((acc - base).cpu().numpy(), a.cpu().numpy()) <= 1e-5
    with pytest.raises(ValueError):
        tfc.spectral_loss_and_grad(f, r, grid=grid, accumulate_into=acc[:2])


def test_loader_quadrants_take_the_pointer_path(monkeypatch):
    """``fft_loss(fake_B, B1..B4)`` with separately allocated quadrants (``datasets_temp.py:76-118``): four base
    pointers, no concatenation kernel -- value and gradient equal the single-tensor call."""
    f, r = _pair(7, seed=31)
    quads = [r[:, :, y:y + 128, x:x + 128].contiguous() for y in (0, 128) for x in (0, 128)]
    fa = f.clone().requires_grad_(True)
    la = tfc.spectral_loss(fa, r, grid=2, patch_reduce="sum", input_scale=255.0)
    la.backward()
    fb = f.clone().requires_grad_(True)

    def no_cat(*a, **k):
        raise AssertionError("the quadrant path must not concatenate")

    monkeypatch.setattr(torch, "cat", no_cat)
    lb = compat.fft_loss(fb, *quads)
    mean4 = compat.patch4_fft_loss(f, *quads).item()
    monkeypatch.undo()
    lb.backward()
    assert lb.item() == la.item()
    assert torch.equal(fa.grad, fb.grad)
    # mean convention (TFCGAN_multigpu_patchFFT.py:498-511)
    assert mean4 == pytest.approx(la.item() / 4, rel=1e-6)


def test_spectra_workspace_is_sized_for_its_own_geometry():
    """fft_components on more than 111 images of 256 x 256 failed with a fresh workspace."""
    from tfc_gan_b200 import functional as F

    F._WORKSPACES.clear()
    x = torch.empty(120, 3, 256, 256, device="cuda").uniform_(-1, 1)
    amp, pha = compat.fft_components(x)
    assert amp.shape == (120, 1, 256, 129) and torch.isfinite(amp).all()
    F._WORKSPACES.clear()
    xr = x[:40].clone().requires_grad_(True)
    a, p = tfc.spectral_components(xr, channels="rgb", input_scale=255.0)  # 120 tiles of 256 x 256
    (a.sum() + p.sum()).backward()
    assert torch.isfinite(xr.grad).all()
    ref = torch.fft.rfft2(x[:2].double().mul(torch.tensor(oracle.LUMA_WEIGHTS, device="cuda").view(1, 3, 1, 1)).sum(1) * 255.0)
    got = torch.fft.ifftshift(amp[:2, 0].double(), dim=(-2, -1))
    assert l2rel(got.cpu().numpy(), ref.abs().cpu().numpy()) <= 1e-5


def test_temperature_accumulate_into_is_validated():
    f, r = _pair(2, seed=41)
    n = torch.rand_like(f)
    with pytest.raises(ValueError):
        tfc.temperature_triplet_loss_and_grad(f, r, n, accumulate_into=torch.zeros(1, 3, 256, 256, device="cuda"))
    with pytest.raises(ValueError):
        tfc.temperature_triplet_loss_and_grad(f, r, n, accumulate_into=torch.zeros_like(f, dtype=torch.float16))
    # 1-channel images with precomputed temperatures: say so explicitly
    f1, n1 = f[:, :1].contiguous(), n[:, :1].contiguous()
    tb = tfc.vectorize_temps(r)
    out_a, _ = tfc.temperature_triplet_loss_and_grad(f1, tb, n1, positive_is_temperatures=True, input_scale=255.0)
    out_b, _ = tfc.temperature_triplet_loss_and_grad(f1, tb[:, 0], n1, input_scale=255.0)  # [N,H,W]: temperatures by shape
    assert out_a[0].item() == out_b[0].item()


@pytest.mark.parametrize("side,grid,channels", [(256, 4, "luma"), (256, 2, "luma"), (256, 1, "luma"), (256, 1, "rgb"),
                                                (256, 2, "rgb"), (512, 1, "luma"), (512, 4, "luma")])
def test_fake_equals_real_gives_exact_zero(side, grid, channels):
    """Known answer: identical inputs -> loss 0 and gradient EXACTLY 0 (sign(0) = 0 in the reference's
    L1Loss).  The 64 x 64 engine detects the tile-level equality while it folds the pixels to luma; the sub-tile engine
    (128 / 256 / 512 tiles) has its forward launch leave one flag per load unit and the combine launch zero-fills the
    planes of a tile whose units are all flagged."""
    n = 9 if side == 256 else 3
    g = torch.Generator(device="cuda").manual_seed(51)
    f = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g)
    kw = dict(grid=grid, channels=channels, weight=0.01, input_scale=255.0)
    loss, terms, grad = tfc.spectral_loss_and_grad(f, f.clone(), **kw)
    assert loss.item() == 0.0 and float(terms.abs().max()) == 0.0
    assert float(grad.abs().max()) == 0.0
    # mixed batch: only the identical images get the exact zero, the others the value they have on their own
    r = f.clone()
    r[::2] = torch.rand_like(r[::2])
    loss, _, grad = tfc.spectral_loss_and_grad(f, r, **kw)
    assert loss.item() > 0
    assert float(grad[1::2].abs().max()) == 0.0 and float(grad[::2].abs().max()) > 0.0
    alone, _, g_alone = tfc.spectral_loss_and_grad(f[::2].contiguous(), r[::2].contiguous(), **kw)
    k = f[::2].shape[0]
    assert loss.item() == pytest.approx(alone.item() * k / n, rel=1e-5)
    assert torch.allclose(grad[::2] * (n / k), g_alone, rtol=1e-4, atol=1e-5 * float(g_alone.abs().max()))
    # one differing pixel in one tile is enough to switch the whole tile back to the normal path
    r2 = f.clone()
    r2[0, 1, side - 1, side - 1] += 0.25
    loss, _, grad = tfc.spectral_loss_and_grad(f, r2, **kw)
    assert loss.item() > 0 and float(grad[0].abs().max()) > 0.0 and float(grad[1:].abs().max()) == 0.0
    # accumulate mode: an identical batch adds exactly nothing
    acc = torch.full_like(f, 0.5)
    tfc.spectral_loss_and_grad(f, f.clone(), accumulate_into=acc, **kw)
    assert bool((acc == 0.5).all())


def test_triplet_on_spectra_variant():
    """``triplet_patches`` of ``..._debiased_V5.py:386-443``: TripletMarginLoss on amplitude / phase spectra of the
    quadrants + the pixel-space patch triplet, against the same composition on the fp64 oracle spectra."""
    f, r = _pair(3, seed=61)
    quads = [r[:, :, y:y + 128, x:x + 128].contiguous() for y in (0, 128) for x in (0, 128)]
    neg = [2, 0, 3, 3]
    fa = f.clone().requires_grad_(True)
    amp, pha, patch = compat.triplet_patches(fa, *quads, negatives=neg)
    (amp + pha + patch).backward()
    fo = f.double().cpu().requires_grad_(True)
    crit = torch.nn.TripletMarginLoss(margin=1.0, p=2)
    fq = [fo[:, :, y:y + 128, x:x + 128] for y in (0, 128) for x in (0, 128)]
    rq = [q.double().cpu() for q in quads]
    sf = [oracle.fft_components_r1(q, input_scale=255.0) for q in fq]
    sr = [oracle.fft_components_r1(q, input_scale=255.0) for q in rq]
    a = sum(crit(sf[i][0], sr[i][0], sr[neg[i]][0]) for i in range(4)) / 4
    p = sum(crit(sf[i][1], sr[i][1], sr[neg[i]][1]) for i in range(4)) / 4
    t = sum(crit(fq[i], rq[i], rq[neg[i]]) for i in range(4)) / 4
    (a + p + t).backward()
    assert amp.item() == pytest.approx(a.item(), rel=1e-4)
    assert pha.item() == pytest.approx(p.item(), rel=1e-4)
    assert patch.item() == pytest.approx(t.item(), rel=1e-4)
    assert l2rel(fa.grad.cpu().numpy(), fo.grad.numpy()) <= 2e-3


@pytest.mark.parametrize("chunk", [0, 3])
def test_combined_patch16_plus_global_matches_oracle(chunk):
    """BASELINE config 5: patch-16 + global FFT loss on the same 512 x 512 tensors through ``multi_grid_loss[_and_grad]``
    (one summed gradient, the second grid adds into the first one's buffer; optional walk in chunks of the batch)
    against the sum of the two fp64 oracle evaluations; the autograd entry gives the same loss and gradient."""
    rs = np.random.RandomState(41)
    fake = rs.uniform(-1, 1, (5, 3, 512, 512)).astype(np.float32)
    real = rs.uniform(-1, 1, (5, 3, 512, 512)).astype(np.float32)
    want_l, want_g, want_terms = 0.0, 0.0, []
    for grid in (4, 1):
        l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=grid, weight=0.01, input_scale=255.0)
        want_l += l
        want_g = want_g + g
        want_terms.append((a, p))
    f = torch.from_numpy(fake).cuda()
    r = torch.from_numpy(real).cuda()
    tfc.reset_launch_count()
    loss, per_cfg, grad = tfc.multi_grid_loss_and_grad(f, r, grids=(4, 1), chunk=chunk, weight=0.01, input_scale=255.0)
    torch.cuda.synchronize()
    assert tfc.launch_count() > 0
    assert loss.item() == pytest.approx(want_l, rel=1e-4)
    np.testing.assert_allclose(per_cfg.cpu().numpy(), np.array(want_terms), rtol=1e-4)
    assert np.linalg.norm(grad.cpu().numpy() - want_g) / np.linalg.norm(want_g) <= 1e-3
    fa = f.clone().requires_grad_(True)
    l2 = tfc.multi_grid_loss(fa, r, grids=(4, 1), chunk=chunk, weight=0.01, input_scale=255.0)
    l2.backward()
    assert l2.item() == pytest.approx(want_l, rel=1e-4)
    assert np.linalg.norm(fa.grad.cpu().numpy() - want_g) / np.linalg.norm(want_g) <= 1e-3


_QUAD_SCRIPT = r"""
import sys
import numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import oracle
import tfc_gan_b200 as tfc
from util import l2rel
g = torch.Generator(device="cuda").manual_seed(77)
for ch, n in (("luma", 5), ("rgb", 2)):
    f = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    r = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    r[1] = f[1]                                   # one identical image: exact zeros
    tfc.reset_launch_count()
    loss, terms, grad = tfc.spectral_loss_and_grad(f, r, grid=1, channels=ch, weight=0.01, input_scale=255.0)
    fwd = tfc.spectral_loss(f, r, grid=1, channels=ch, weight=0.01, input_scale=255.0)
    l, a, p, go = oracle.spectral_loss_and_grad_r1(f.cpu().numpy(), r.cpu().numpy(), grid=1, weight=0.01, input_scale=255.0, channels=ch)
    assert abs(loss.item() - l) <= 1e-4 * abs(l), (ch, loss.item(), l)
    assert abs(fwd.item() - l) <= 1e-4 * abs(l), (ch, fwd.item(), l)
    assert l2rel(grad.cpu().numpy(), go) <= 1e-3, (ch, l2rel(grad.cpu().numpy(), go))
    assert float(grad[1].abs().max()) == 0.0
print("quad ok")
"""


def test_quad_combine_opt_in_matches_oracle():
    """The opt-in quad combine launch of 256 x 256 tiles (``TFCFFT_COMBINE_QUAD=1``, read once per process: run in a
    subprocess) against the fp64 oracle, luma and per-channel, with an identical image in the batch."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    env = dict(os.environ, TFCFFT_COMBINE_QUAD="1")
    p = subprocess.run([sys.executable, "-c", _QUAD_SCRIPT.format(root=root, tests=here)], env=env, capture_output=True, text=True, timeout=280)
    assert p.returncode == 0 and "quad ok" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
