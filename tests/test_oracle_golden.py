"""Pins the oracle to the reference (CPU, no GPU needed).

* R0 (``oracle/r0_literal.py``) must reproduce what the reference's own code produced
  (``tests/golden/golden_cases.json`` / ``golden_arrays.npz``, written by
  ``tests/golden/make_golden.py`` from ``/root/reference``).
* R1 fed the quantised luma must reproduce R0 (cross-pin).
* The closed form the CUDA kernels implement must agree with autograd on R1.
"""

import itertools
import json
import os

import numpy as np
import pytest
import torch

import oracle
from inputs import make_gray_pairs, make_pair

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_cases.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_arrays.npz"))
LOSS_CASES = [c for c in GOLD["cases"] if "loss" in c]


@pytest.mark.parametrize("case", LOSS_CASES, ids=[c["name"] for c in LOSS_CASES])
def test_r0_matches_reference_loss(case):
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    loss, amp, pha = oracle.spectral_loss_r0(fake, real, case["grid"], case["patch_reduce"])
    loss = float(loss) * case.get("weight", 1.0)
    assert loss == pytest.approx(case["loss"], rel=2e-6)
    if "amp" in case:
        assert float(amp) == pytest.approx(case["amp"], rel=2e-6)
        assert float(pha) == pytest.approx(case["pha"], rel=2e-6)


def test_r0_spectra_bit_exact():
    case = next(c for c in GOLD["cases"] if c["name"] == "p16_uniform_11_float32")
    fake, _ = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    amp, pha = oracle.fft_components_r0(fake[:, :, 64:128, 64:128])  # B6 = row 1, col 1
    assert amp.shape == (2, 1, 64, 33)
    np.testing.assert_array_equal(amp, ARR["p16_uniform_11_float32_amp_B6"])
    np.testing.assert_array_equal(pha, ARR["p16_uniform_11_float32_pha_B6"])


def test_r0_luma_and_make_spectra_bit_exact():
    fake, _ = make_pair("unit", 61, (1, 3, 256, 256), "float32")
    g = oracle.gray_u8(fake)[0]
    np.testing.assert_array_equal(g, ARR["luma_unit_61"])
    np.testing.assert_array_equal(oracle.make_spectra_r0(g).astype(np.float32), ARR["make_spectra_unit_61"])


@pytest.mark.parametrize("metric", ["mse", "mae"])
def test_r0_mag_metric(metric):
    case = next(c for c in GOLD["cases"] if c["name"] == f"mag_{metric}_71")
    reals, fakes = make_gray_pairs(case["seed"], case["n"], case["side"])
    values, skipped = oracle.mag_mse_r0(reals, fakes, metric)
    assert skipped == [1]  # the constant image has -inf in its log spectrum
    np.testing.assert_allclose(values, case["values"], rtol=2e-6)


def test_quantize_wraps_like_numpy_cast():
    x = np.array([-1.0, -0.999, -0.5, -0.004, 0.0, 0.003, 0.5, 0.999, 1.0], np.float32)
    np.testing.assert_array_equal(oracle.quantize_u8(x), [1, 2, 129, 255, 0, 0, 127, 254, 255])
    for dt in (np.float16, np.float32):
        a = np.random.RandomState(0).uniform(-1.1, 1.1, 20001).astype(dt)
        np.testing.assert_array_equal(oracle.quantize_u8(a), (a * 255).astype(np.uint8))


@pytest.mark.parametrize("grid", [1, 2, 4])
def test_r1_on_quantised_luma_reproduces_r0(grid):
    fake, real = make_pair("uniform", 5, (2, 3, 256, 256), "float32")
    l0, a0, p0 = oracle.spectral_loss_r0(fake, real, grid)
    l1, a1, p1 = oracle.spectral_loss_r1(torch.from_numpy(fake), torch.from_numpy(real), grid=grid, quantize=True)
    assert float(l1) == pytest.approx(float(l0), rel=1e-6)
    assert float(a1) == pytest.approx(float(a0), rel=1e-6)
    assert float(p1) == pytest.approx(float(p0), rel=1e-6)


def test_survey_recorded_values():
    """SURVEY.md §8c recorded R0/R1 numbers for torch.manual_seed(0) inputs."""
    torch.manual_seed(0)
    fake = torch.rand(8, 3, 256, 256) * 2 - 1
    real = torch.rand(8, 3, 256, 256) * 2 - 1
    assert float(oracle.spectral_loss_r0(fake, real, 4)[0]) == pytest.approx(818.144348, rel=1e-6)
    l, a, p = oracle.spectral_loss_r1(fake, real, grid=4, channels="rgb")
    assert float(l) == pytest.approx(10.640495575, rel=1e-9)
    assert float(a) == pytest.approx(19.185863605, rel=1e-9)
    assert float(p) == pytest.approx(2.095127545, rel=1e-9)


MODES = list(itertools.product([1, 2], ["luma", "rgb"], [True, False], ["l1", "mse"], [False, True], ["half", "full"]))


@pytest.mark.parametrize("grid,channels,use_phase,distance,logmag,spectrum", MODES)
def test_closed_form_matches_autograd(grid, channels, use_phase, distance, logmag, spectrum):
    fake, real = make_pair("uniform", 3, (2, 3, 16, 16), "float64")
    kw = dict(grid=grid, channels=channels, use_phase=use_phase, distance=distance, patch_reduce="sum",
              log_magnitude=logmag, spectrum=spectrum, weight=0.7, input_scale=3.0)
    l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, **kw)
    l2, a2, p2, g2 = oracle.spectral_grad_analytic(fake, real, **kw)
    assert l2 == pytest.approx(l, rel=1e-12)
    assert a2 == pytest.approx(a, rel=1e-12)
    assert p2 == pytest.approx(p, rel=1e-12, abs=1e-15)
    assert np.linalg.norm(g - g2) <= 1e-10 * np.linalg.norm(g)


def test_r1_gradcheck_small():
    fake, real = make_pair("uniform", 9, (1, 3, 8, 8), "float64")
    fk = torch.from_numpy(fake).requires_grad_(True)
    rl = torch.from_numpy(real)
    fn = lambda x: oracle.spectral_loss_r1(x, rl, grid=1, channels="luma", distance="mse")[0]
    assert torch.autograd.gradcheck(fn, (fk,), eps=1e-6, atol=1e-6, rtol=1e-4)


def test_known_answers():
    p = 16
    # delta image -> flat amplitude == input_scale; phase 0
    x = np.zeros((1, 1, p, p)); x[0, 0, 0, 0] = 1.0
    a, ph = oracle.fft_components_r1(torch.from_numpy(x), input_scale=2.5)
    assert torch.allclose(a, torch.full_like(a, 2.5)) and torch.allclose(ph, torch.zeros_like(ph))
    # single cosine at (u,v) -> one half-plane bin of p^2/2 (its mirror is in the dropped half)
    yy, xx = np.mgrid[0:p, 0:p]
    c = np.cos(2 * np.pi * (3 * yy + 2 * xx) / p)[None, None]
    a, _ = oracle.fft_components_r1(torch.from_numpy(c), shift=False)
    assert a[0, 0, 3, 2].item() == pytest.approx(p * p / 2)
    assert (a > 1e-9).sum().item() == 1
    # Parseval with Hermitian weights
    z = np.random.RandomState(1).normal(size=(1, 1, p, p))
    a, _ = oracle.fft_components_r1(torch.from_numpy(z), shift=False)
    w = torch.full((p // 2 + 1,), 2.0); w[0] = 1; w[-1] = 1
    assert ((a ** 2) * w).sum().item() == pytest.approx(p * p * (z ** 2).sum())
    # sum over patches == g^2 * mean; tile-permutation invariance; fake == real -> 0
    fake, real = make_pair("uniform", 2, (2, 3, 32, 32), "float64")
    tf, tr = torch.from_numpy(fake), torch.from_numpy(real)
    lm = oracle.spectral_loss_r1(tf, tr, grid=4)[0]
    ls = oracle.spectral_loss_r1(tf, tr, grid=4, patch_reduce="sum")[0]
    assert float(ls) == pytest.approx(16 * float(lm))
    perm = lambda t: torch.cat([t[..., 16:, :], t[..., :16, :]], -2)
    assert float(oracle.spectral_loss_r1(perm(tf), perm(tr), grid=4)[0]) == pytest.approx(float(lm))
    l0, _, _, g0 = oracle.spectral_loss_and_grad_r1(real, real, grid=4)
    assert l0 == 0.0 and not g0.any()
    # amplitude term is linear in input_scale, phase term invariant
    _, a1, p1 = oracle.spectral_loss_r1(tf, tr, grid=2)
    _, a3, p3 = oracle.spectral_loss_r1(tf, tr, grid=2, input_scale=3.0)
    assert float(a3) == pytest.approx(3 * float(a1)) and float(p3) == pytest.approx(float(p1))
