"""Host-side check of the kernels' algebra: ``libtfcfft_emu.so`` executes the very templates the
sm_100a kernels instantiate (index maps of the digit-reversed in-place FFT, Hermitian un-mixing, bin
ownership, reduction layout), serially on the CPU, and must agree with the oracle.  This is test
infrastructure, not a product fallback."""

import json
import os

import numpy as np
import pytest
import torch

import oracle
from inputs import make_gray_pairs, make_pair
from util import L, emulate, flags_of, l2rel

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_cases.json")))

CASES = [
    # side, grid, options
    (64, 4, dict()),
    (64, 2, dict(channels="rgb")),
    (64, 1, dict(distance="mse", patch_reduce="sum")),
    (128, 1, dict(use_phase=False)),
    (128, 2, dict(spectrum="full", channels="rgb")),
    (256, 4, dict()),
    (256, 2, dict(distance="mse")),
    (256, 1, dict()),
    (256, 1, dict(channels="rgb", use_phase=False, log_magnitude=True, spectrum="full", distance="mse")),
    (512, 1, dict()),
    (512, 1, dict(distance="mse", use_phase=False)),
    (512, 1, dict(spectrum="full", log_magnitude=True)),
    (512, 1, dict(force_split=True)),
    (256, 4, dict(force_split=True)),
    (256, 2, dict(force_split=True, channels="rgb")),
    (256, 1, dict(force_split=True)),
    (256, 2, dict(force_generic=True)),
    (256, 1, dict(channels="rgb", distance="mse")),
]


@pytest.mark.parametrize("side,grid,opt", CASES, ids=[f"{s}-g{g}-{'-'.join(f'{k}={v}' for k, v in o.items()) or 'default'}" for s, g, o in CASES])
def test_emulation_matches_r1(side, grid, opt):
    n = 2 if side <= 256 else 1
    # seed 7 puts one 256 x 256 single-channel bin 1.6e-7 (relative) from the phase branch cut at +-pi, which fp32
    # cannot resolve and the squared phase distance amplifies; that case uses another seed
    seed = 5 if (side, grid, opt.get("channels"), opt.get("distance")) == (256, 1, "rgb", "mse") and opt.get("use_phase", True) else 7
    fake, real = make_pair("uniform", seed, (n, 3, side, side), "float32")
    rc, out, per, g = emulate(fake, real, grid, flags_of(**opt), weight=0.7, input_scale=3.0)
    assert rc == 0
    okw = {k: v for k, v in opt.items() if not k.startswith("force_")}
    l, a, p, gr = oracle.spectral_loss_and_grad_r1(fake, real, grid=grid, weight=0.7, input_scale=3.0, **okw)
    assert out[0] == pytest.approx(l, rel=1e-5)
    assert out[1] == pytest.approx(a, rel=1e-5)
    assert out[2] == pytest.approx(p, rel=1e-5, abs=1e-12)
    assert l2rel(g, gr) <= 1e-3
    assert per[:, 0].mean() == pytest.approx(a, rel=1e-5)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_emulation_half_inputs(dtype):
    fake, real = make_pair("tanh", 3, (2, 3, 128, 128), "float32")
    tf = torch.from_numpy(fake).to(getattr(torch, dtype))
    tr = torch.from_numpy(real).to(getattr(torch, dtype))
    if dtype == "float16":
        af, ar, code = tf.numpy(), tr.numpy(), L.F16
    else:
        af, ar, code = tf.view(torch.int16).numpy().view(np.uint16), tr.view(torch.int16).numpy().view(np.uint16), L.BF16
    rc, out, _, g = emulate(af, ar, 2, 0, input_scale=255.0, dtype_code=code)
    assert rc == 0
    l, a, p, gr = oracle.spectral_loss_and_grad_r1(tf.double().numpy(), tr.double().numpy(), grid=2, input_scale=255.0)
    assert out[0] == pytest.approx(l, rel=1e-5)
    g32 = torch.from_numpy(g.view(np.int16)).view(torch.bfloat16).float().numpy() if dtype == "bfloat16" else g.astype(np.float32)
    assert l2rel(g32, gr) <= (1e-2 if dtype == "bfloat16" else 2e-3)  # output rounding of the 16-bit gradient


LOSS_CASES = [c for c in GOLD["cases"] if "loss" in c]


@pytest.mark.parametrize("case", LOSS_CASES, ids=[c["name"] for c in LOSS_CASES])
def test_emulation_quantised_mode_matches_reference_golden(case):
    """QUANTIZE_U8 (reference-as-shipped input path) against what the reference's own code produced."""
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    flags = flags_of(quantize=True, patch_reduce=case["patch_reduce"])
    rc, out, _, _ = emulate(fake, real, case["grid"], flags, weight=case.get("weight", 1.0), grad=False)
    assert rc == 0
    assert out[0] == pytest.approx(case["loss"], rel=1e-4)
    if "amp" in case:
        assert out[1] == pytest.approx(case["amp"], rel=1e-4)
        assert out[2] == pytest.approx(case["pha"], rel=1e-4)


def test_emulation_mag_mse_metric_matches_reference_golden():
    case = next(c for c in GOLD["cases"] if c["name"] == "mag_mse_71")
    reals, fakes = make_gray_pairs(case["seed"], case["n"], case["side"])
    r = np.stack(reals)[:, None]
    f = np.stack(fakes)[:, None]
    flags = flags_of(use_phase=False, distance="mse", log_magnitude=True, spectrum="full")
    rc, out, per, _ = emulate(f, r, 1, flags, grad=False)
    assert rc == 0
    vals = per[:, 0]
    assert not np.isfinite(vals[1])  # the constant image: log(0) -> skipped by the reference
    np.testing.assert_allclose(vals[[0, 2, 3]], case["values"], rtol=1e-4)
    assert out[3] == 1.0  # non-finite flag


def test_emulation_rejects_gradient_for_quantised_inputs():
    fake, real = make_pair("uniform", 1, (1, 3, 64, 64), "float32")
    rc, *_ = emulate(fake, real, 1, flags_of(quantize=True), grad=True)
    assert rc == -9


def test_emulation_strided_views():
    """Reference-style patch views (``B[:, :, 64:128, 64:128]``) are consumed in place."""
    fake, real = make_pair("uniform", 4, (2, 3, 256, 256), "float32")
    fv, rv = fake[:, :, 64:128, 64:128], real[:, :, 128:192, 0:64]
    rc, out, _, g = emulate(fv, rv, 1, 0, grad=False)
    assert rc == 0
    l, *_ = oracle.spectral_loss_r1(torch.from_numpy(fv.copy()), torch.from_numpy(rv.copy()), grid=1)
    assert out[0] == pytest.approx(float(l), rel=1e-5)


@pytest.mark.parametrize("opt", [dict(), dict(channels="rgb", distance="mse"), dict(use_phase=False)])
def test_packed_pair_path_matches_generic_path(opt):
    """The packed 64x64 tile-pair fast path and the generic resident path are two implementations of the
    same algebra; odd tile counts exercise the duplicated last lane."""
    fake, real = make_pair("tanh", 17, (3, 3, 64, 64), "float32")  # grid=1: 3 (luma) or 9 (rgb) tiles, odd
    rc, o1, p1, g1 = emulate(fake, real, 1, flags_of(use_pair=True, **opt), input_scale=255.0)
    rc2, o2, p2, g2 = emulate(fake, real, 1, flags_of(force_generic=True, **opt), input_scale=255.0)
    assert rc == 0 and rc2 == 0
    np.testing.assert_allclose(o1[:3], o2[:3], rtol=2e-6)
    np.testing.assert_allclose(p1, p2, rtol=2e-6)
    assert l2rel(g1, g2) <= 2e-5


@pytest.mark.parametrize("side", [64, 256])
@pytest.mark.parametrize("opt", [dict(), dict(channels="rgb"), dict(spectrum="full", log_magnitude=True)])
def test_emulated_spectra_match_fft_components(side, opt):
    """tfcfft_spectra == differentiable fft_components (amp / phase, fftshift-ed half plane or full log plane)."""
    from util import emulate_spectra, emulate_spectra_bwd

    x, y = make_pair("tanh", 23, (2, 3, side, side), "float32")
    rc, ax, px, ay, py = emulate_spectra(x, flags_of(**opt), input_scale=255.0, shift=True, y=y)
    assert rc == 0
    full = opt.get("spectrum") == "full"
    for arr, (a_emu, p_emu) in ((x, (ax, px)), (y, (ay, py))):
        t = oracle.r1_differentiable._prepare(torch.from_numpy(arr), opt.get("channels", "luma"), 255.0, False, torch.float64)
        F = torch.fft.fft2(t) if full else torch.fft.rfft2(t)
        F = torch.fft.fftshift(F, dim=(-2, -1))
        amp = F.abs().log() if opt.get("log_magnitude") else F.abs()
        np.testing.assert_allclose(a_emu, amp.numpy(), rtol=2e-4, atol=2e-3 if not opt.get("log_magnitude") else 2e-4)
        # phase of tiny bins is rounding noise: compare where the amplitude is not negligible
        big = F.abs().numpy() > 1e-3 * F.abs().numpy().max()
        dphi = np.angle(np.exp(1j * (p_emu - torch.angle(F).numpy())))
        assert np.abs(dphi[big]).max() < 2e-3
    # backward against autograd for a random cotangent
    rs = np.random.RandomState(0)
    ga = rs.normal(size=ax.shape).astype(np.float32)
    gp = (rs.normal(size=px.shape) * 0.0 if opt.get("log_magnitude") else rs.normal(size=px.shape)).astype(np.float32)
    rc, g = emulate_spectra_bwd(x, ga, gp, flags_of(**opt), input_scale=255.0, shift=True)
    assert rc == 0
    xt = torch.from_numpy(x).double().requires_grad_(True)
    t = oracle.r1_differentiable._prepare(xt, opt.get("channels", "luma"), 255.0, False, torch.float64)
    F = torch.fft.fftshift(torch.fft.fft2(t) if full else torch.fft.rfft2(t), dim=(-2, -1))
    amp = F.abs().log() if opt.get("log_magnitude") else F.abs()
    loss = (amp * torch.from_numpy(ga).double()).sum() + (torch.angle(F) * torch.from_numpy(gp).double()).sum()
    loss.backward()
    assert l2rel(g, xt.grad.numpy()) <= 1e-3


def test_emulated_spectra_quantised_match_reference_golden():
    from util import emulate_spectra

    case = next(c for c in GOLD["cases"] if c["name"] == "p16_uniform_11_float32")
    fake, _ = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    patch = np.ascontiguousarray(fake[:, :, 64:128, 64:128])  # B6
    rc, amp, pha, _, _ = emulate_spectra(patch, flags_of(quantize=True), shift=True)
    assert rc == 0
    gold = np.load(os.path.join(HERE, "golden", "golden_arrays.npz"))
    np.testing.assert_allclose(amp, gold["p16_uniform_11_float32_amp_B6"], rtol=1e-5, atol=0.05)
    ga = gold["p16_uniform_11_float32_amp_B6"]
    dphi = np.angle(np.exp(1j * (pha - gold["p16_uniform_11_float32_pha_B6"])))
    assert np.abs(dphi[ga > 1.0]).max() < 1e-3


@pytest.mark.parametrize("opt", [dict(), dict(channels="rgb", distance="mse"), dict(use_phase=False)])
def test_thread_per_line_path_matches_oracle_and_pair_path(opt):
    """The register-resident 64-point line kernel (USE_LINE) against R1 and against the packed pair path."""
    fake, real = make_pair("tanh", 19, (3, 3, 128, 128), "float32")  # grid=2 -> 64x64 tiles
    rc, o1, p1, g1 = emulate(fake, real, 2, flags_of(**opt), input_scale=255.0)
    rc2, o2, p2, g2 = emulate(fake, real, 2, flags_of(use_pair=True, **opt), input_scale=255.0)
    assert rc == 0 and rc2 == 0
    l, a, p, gr = oracle.spectral_loss_and_grad_r1(fake, real, grid=2, input_scale=255.0, **opt)
    assert o1[0] == pytest.approx(l, rel=1e-5)
    assert l2rel(g1, gr) <= 1e-3
    np.testing.assert_allclose(o1[:3], o2[:3], rtol=2e-6)
    assert l2rel(g1, g2) <= 2e-5


@pytest.mark.parametrize("cf,cr", [(-0.5, 0.5), (0.5, -0.5), (-0.5, -0.25), (0.5, 0.25)])
@pytest.mark.parametrize("grid", [4, 1])
def test_real_axis_phase_signs(cf, cr, grid):
    """Self-conjugate bins are exactly real: the phase difference there is 0 or +-pi and its sign decides the sign of
    the gradient.  Constant images of either sign plus a little noise exercise every (F, R) sign combination of the
    single-arctangent phase difference (pair_tile.cuh: phase_delta) at the DC / Nyquist bins."""
    rs = np.random.RandomState(5)
    shape = (2, 3, 64 * grid if grid == 4 else 64, 64 * grid if grid == 4 else 64)
    fake = (cf + 0.05 * rs.uniform(-1, 1, shape)).astype(np.float32)
    real = (cr + 0.05 * rs.uniform(-1, 1, shape)).astype(np.float32)
    rc, out, _, g = emulate(fake, real, grid, 0, weight=1.0, input_scale=255.0)
    assert rc == 0
    l, a, p, go = oracle.spectral_loss_and_grad_r1(fake, real, grid=grid, weight=1.0, input_scale=255.0)
    assert out[0] == pytest.approx(l, rel=1e-4)
    assert out[2] == pytest.approx(p, rel=1e-4)
    assert l2rel(g, go) <= 1e-3


@pytest.mark.parametrize("opt", [dict(), dict(channels="rgb"), dict(distance="mse"), dict(use_phase=False)],
                         ids=["default", "rgb", "mse", "amp-only"])
def test_half_line_engine_matches_oracle(opt):
    """The half-line engine (two threads per line, 32-point core, line_tile.cuh: line2_fft_pass) executed serially:
    same loss and gradient as the fp64 oracle, and the thread-per-line engine to rounding."""
    fake, real = make_pair("tanh", 321, (2, 3, 256, 256), "float32")
    base = flags_of(**opt)
    rc, out_h, _, g_h = emulate(fake, real, 4, base | L.USE_HALFLINE, weight=0.01, input_scale=255.0)
    assert rc == 0
    rc, out_l, _, g_l = emulate(fake, real, 4, base, weight=0.01, input_scale=255.0)
    assert rc == 0
    l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=4, weight=0.01, input_scale=255.0, **opt)
    assert out_h[0] == pytest.approx(l, rel=1e-4)
    assert l2rel(g_h, g) <= 1e-3
    assert out_h[0] == pytest.approx(out_l[0], rel=2e-6)
    assert l2rel(g_h, g_l) <= 1e-4
