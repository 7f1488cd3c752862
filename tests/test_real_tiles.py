"""Parity on the 20 real 256 x 256 face tiles of ``/root/reference/TFC-STN/samples/*.png`` (SURVEY.md sections 4 / 8c) --
natural images instead of synthetic noise.  The tiles are read in place (``tests/golden/real_tiles.py``) and every test
skips when they are not available; the golden numbers (``golden_real_tiles.json``) are the outputs of the reference's
own functions on those tiles (``make_golden_real_tiles.py``).

* CPU (``-m "not gpu"``): the oracle's R0 reproduces the reference numbers; the kernels' own arithmetic (serial CPU
  emulation) reproduces them in reference-as-shipped mode and matches the fp64 oracle's loss / gradient in R1 mode.
* GPU (``-m gpu``): the CUDA path, through the public API, against the same numbers.
"""

import json
import os

import numpy as np
import pytest
import torch

import oracle
from real_tiles import load_pairs
from util import l2rel, robust_grad_error

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_real_tiles.json")))
KEYS = {4: "p16", 2: "p4", 1: "glob"}
LOSS_TOL, GRAD_TOL = 1e-4, 1e-3


def _pairs(dtype):
    got = load_pairs(dtype)
    if got is None:
        pytest.skip("real face tiles not available (no /root/reference and no TFC_SAMPLES_DIR)")
    case = next(c for c in GOLD["cases"] if c["dtype"] == dtype)
    fake, real = got
    chk = float(np.abs(fake.astype(np.float64)).sum() + np.abs(real.astype(np.float64)).sum())
    assert chk == pytest.approx(case["checksum"], rel=1e-9), "the tiles differ from the ones the golden values were made from"
    return fake, real, case


@pytest.mark.parametrize("dtype", ["float32", "float16"])
@pytest.mark.parametrize("grid", [4, 2, 1])
def test_oracle_r0_reproduces_reference_on_real_tiles(dtype, grid):
    fake, real, case = _pairs(dtype)
    loss, _, _ = oracle.spectral_loss_r0(fake, real, grid)
    assert float(loss) == pytest.approx(case[KEYS[grid]], rel=2e-6)
    if grid == 2:
        s, _, _ = oracle.spectral_loss_r0(fake, real, 2, patch_reduce="sum")
        assert float(s) == pytest.approx(case["p4sum"], rel=2e-6)


@pytest.mark.parametrize("grid", [4, 2, 1])
def test_emulated_kernels_on_real_tiles(grid):
    """The kernels' arithmetic (CPU emulation of the same templates), 6 of the 20 pairs to keep the CPU suite short."""
    import tfc_gan_b200._lib as L
    from util import emulate

    fake, real, _ = _pairs("float32")
    sel = [0, 3, 7, 10, 14, 19]
    f, r = np.ascontiguousarray(fake[sel]), np.ascontiguousarray(real[sel])
    # reference-as-shipped mode against the oracle's R0 (itself pinned to the reference above)
    rc, out, _, _ = emulate(f, r, grid, L.QUANTIZE_U8, grad=False)
    assert rc == 0
    want, _, _ = oracle.spectral_loss_r0(f, r, grid)
    assert out[0] == pytest.approx(float(want), rel=LOSS_TOL)
    # differentiable mode against the fp64 oracle
    rc, out, _, g = emulate(f, r, grid, 0, weight=0.01, input_scale=255.0)
    assert rc == 0
    l, a, p, go = oracle.spectral_loss_and_grad_r1(f, r, grid=grid, weight=0.01, input_scale=255.0)
    assert out[0] == pytest.approx(l, rel=LOSS_TOL)
    _check_gradient(g, go, f, r, grid, "luma")


def _check_gradient(got, want, fake, real, grid, channels):
    """Natural image pairs contain bins whose amplitude / phase difference is below fp32 transform noise; the L1 loss
    gives them an arbitrary sign in any fp32 evaluation (torch.fft fp32 included: on these tiles its own plain error
    against fp64 is 3e-3 at grid=4) and ONE such bin moves the plain norm by ~1e-2.  The comparison therefore masks
    the bins the fp64 oracle marks as marginal (``util.robust_grad_error``) and bounds how many there are."""
    # smooth natural images at 256 x 256 keep most high-frequency bins BELOW the fp32 noise floor of a transform whose
    # DC term is 10^5 times larger: their phases are rounding noise in any fp32 implementation, so the mask is wide
    # there (a few per cent of the bins at kappa = 64); what is left must agree to the contract tolerance
    for kappa in (16.0, 64.0):
        err, masked, plain = robust_grad_error(got, want, fake, real, grid, channels, 255.0, kappa=kappa)
        if err <= GRAD_TOL:
            break
    assert err <= GRAD_TOL, f"unmasked error {err:.2e} (plain {plain:.2e}, masked {masked:.2%})"
    assert masked <= 0.12
    assert plain <= 3e-2  # a handful of flipped marginal bins at most


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_cuda_as_shipped_mode_on_real_tiles(dtype):
    """R0 mode on the GPU against the numbers the reference's own functions produced on the same tiles."""
    import tfc_gan_b200 as tfc

    fake, real, case = _pairs(dtype)
    f, r = torch.from_numpy(fake).cuda(), torch.from_numpy(real).cuda()
    for grid in (4, 2, 1):
        loss = tfc.spectral_loss(f, r, grid=grid, quantize=True)
        assert loss.item() == pytest.approx(case[KEYS[grid]], rel=LOSS_TOL)
    assert tfc.spectral_loss(f, r, grid=2, quantize=True, patch_reduce="sum").item() == pytest.approx(case["p4sum"], rel=LOSS_TOL)
    # the compat entry points a training script calls
    from tfc_gan_b200 import compat

    with compat.reference_mode():
        assert compat.calculate_ffts(*compat.make_16_patches(f), *compat.make_16_patches(r)).item() == pytest.approx(case["p16"], rel=LOSS_TOL)
        quads = [r[:, :, y:y + 128, x:x + 128].contiguous() for y in (0, 128) for x in (0, 128)]
        assert compat.fft_loss(f, *quads).item() == pytest.approx(case["p4sum"], rel=LOSS_TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [4, 2, 1])
@pytest.mark.parametrize("channels", ["luma", "rgb"])
def test_cuda_gradient_on_real_tiles(grid, channels):
    """R1 mode: loss and gradient against the fp64 oracle on natural images (smooth spectra, a few dominant bins)."""
    import tfc_gan_b200 as tfc

    fake, real, _ = _pairs("float32")
    f = torch.from_numpy(fake).cuda().requires_grad_(True)
    r = torch.from_numpy(real).cuda()
    loss = tfc.spectral_loss(f, r, grid=grid, channels=channels, weight=0.01, input_scale=255.0)
    loss.backward()
    l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=grid, channels=channels, weight=0.01, input_scale=255.0)
    assert loss.item() == pytest.approx(l, rel=LOSS_TOL)
    _check_gradient(f.grad.cpu().numpy(), g, fake, real, grid, channels)
