"""Patch triplet loss on the CPU: the oracle against the golden vectors the reference's own
lines produced (``tests/golden/make_golden_triplet.py``), the kernel arithmetic (serial emulation,
``libtfcfft_emu.so``) against the oracle, and the host-side argument checks of the C entry point."""

import ctypes
import json
import os

import numpy as np
import pytest
import torch

import tfc_gan_b200 as tfc
from inputs import make_pair
from oracle import triplet as otri
from oracle import temperature as otemp
from oracle import regional as oreg
from util import emulate_regional, emulate_regional_spectra, emulate_temperature, emulate_triplet, l2rel

L = tfc._lib
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_triplet.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_triplet.npz"))
TRI = [c for c in GOLD["cases"] if "op" not in c]
REG = [c for c in GOLD["cases"] if c.get("op") == "regional"]
TEMP = [c for c in GOLD["cases"] if c.get("op") == "temperature"]


@pytest.mark.parametrize("case", TRI, ids=[c["name"] for c in TRI])
def test_oracle_matches_reference_lines(case):
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    _, loss, _, grad = otri.patch_triplet_loss_and_grad(fake, real, case["negatives"], grid=case["grid"])
    tol = 1e-12 if case["dtype"] == "float64" else 2e-7  # the fp32 reference runs carry fp32 rounding
    assert loss == pytest.approx(case["loss"], rel=tol)
    assert np.sqrt((grad * grad).sum()) == pytest.approx(case["grad_l2"], rel=tol)
    assert np.abs(grad).sum() == pytest.approx(case["grad_abs_sum"], rel=tol)
    ref = ARR[case["name"] + "_grad_n0_c1_rows60_70"]
    assert np.abs(grad[0, 1, 60:70, :] - ref).max() <= 2e-7 * np.abs(grad).max() + 1e-12


def test_negative_draw_replays_the_reference_stream():
    for case in TRI:
        np.random.seed(case["numpy_seed"])
        assert otri.draw_negatives(case["grid"] ** 2) == case["negatives"]
        np.random.seed(case["numpy_seed"])
        assert tfc.compat.draw_negatives(case["grid"] ** 2) == case["negatives"]


def test_oracle_against_torch_autograd_small():
    rs = np.random.RandomState(3)
    fake, real = rs.normal(size=(2, 3, 32, 32)), rs.normal(size=(2, 3, 32, 32))
    neg = [3, 0, 1, 1]
    f = torch.from_numpy(fake).requires_grad_(True)
    r = torch.from_numpy(real)
    crit = torch.nn.TripletMarginLoss(margin=0.7, p=2)
    tl = lambda t, i: t[:, :, (i // 2) * 16:(i // 2 + 1) * 16, (i % 2) * 16:(i % 2 + 1) * 16]
    loss = sum(crit(tl(f, i), tl(r, i), tl(r, neg[i])) for i in range(4)) / 4
    loss.backward()
    wl, l, _, g = otri.patch_triplet_loss_and_grad(fake, real, neg, grid=2, margin=0.7, weight=2.5)
    assert l == pytest.approx(float(loss), rel=1e-12)
    assert wl == pytest.approx(2.5 * float(loss), rel=1e-12)
    assert l2rel(g, 2.5 * f.grad.numpy()) <= 1e-12


EMU_CASES = [(64, 4, "float32"), (64, 2, "float32"), (128, 1, "float32"), (256, 4, "float32"), (256, 2, "float16"), (256, 1, "float32"),
             (512, 1, "float32"), (512, 4, "float32")]


@pytest.mark.parametrize("side,grid,dtype", EMU_CASES, ids=[f"{s}-g{g}-{d}" for s, g, d in EMU_CASES])
def test_emulated_kernel_arithmetic_matches_oracle(side, grid, dtype):
    n = 2 if side <= 256 else 1
    fake, real = make_pair("tanh", 5, (n, 3, side, side), dtype)
    rs = np.random.RandomState(side + grid)
    neg = [int(k) for k in rs.randint(grid * grid, size=grid * grid)]
    rc, out, g = emulate_triplet(fake, real, grid, neg, margin=1.0, weight=0.5)
    assert rc == 0
    wl, l, act, gr = otri.patch_triplet_loss_and_grad(fake, real, neg, grid=grid, weight=0.5)
    assert out[0] == pytest.approx(wl, rel=1e-5)
    assert out[1] == pytest.approx(l, rel=1e-5)
    assert out[2] == pytest.approx(act, abs=1e-6)
    assert l2rel(g.astype(np.float64), gr) <= (2e-3 if dtype == "float16" else 1e-5)


def test_emulated_accumulate_and_forward_only():
    fake, real = make_pair("uniform", 9, (1, 3, 64, 64), "float32")
    neg = [1, 0, 3, 2]
    base = np.full_like(fake, 0.25)
    rc, _, g = emulate_triplet(fake, real, 2, neg, accumulate_into=base.copy())
    assert rc == 0
    _, _, _, gr = otri.patch_triplet_loss_and_grad(fake, real, neg, grid=2)
    assert l2rel(g - 0.25, gr) <= 1e-4
    rc, out, g = emulate_triplet(fake, real, 2, neg, grad=False)
    assert rc == 0 and g is None and out[1] > 0


def test_degenerate_rows():
    # fake == real and the negative is the patch itself: d_ap = d_an = eps*sqrt(P) -> hinge = margin, gradient 0
    x, _ = make_pair("uniform", 2, (1, 1, 32, 32), "float32")
    rc, out, g = emulate_triplet(x, x.copy(), 2, [0, 1, 2, 3], margin=1.0)
    assert rc == 0
    assert out[1] == pytest.approx(1.0, rel=1e-6) and out[2] == pytest.approx(1.0)
    assert np.abs(g).max() <= 1e-9
    # eps = 0 on identical tensors: distances are exactly zero, the gradient must stay finite (zero)
    rc, out, g = emulate_triplet(x, x.copy(), 2, [0, 1, 2, 3], margin=1.0, eps=0.0)
    assert rc == 0 and np.isfinite(g).all() and np.abs(g).max() == 0.0
    # a margin so negative that no row is active
    rc, out, g = emulate_triplet(x, -x, 2, [3, 2, 1, 0], margin=-1e6)
    assert out[1] == 0.0 and out[2] == 0.0 and np.abs(g).max() == 0.0


def desc(shape=(2, 3, 256, 256), grid=4, flags=0, dtype=L.F32):
    st = (shape[1] * shape[2] * shape[3], shape[2] * shape[3], shape[3], 1)
    return L.make_desc(dtype, grid, flags, shape, st, st, st, 1.0, 1.0)


def call(d, neg, fake=256, real=256, out=256, grad=None, ws=256, ws_bytes=None):
    lib = L.load()
    n = (ctypes.c_int32 * len(neg))(*neg) if neg is not None else None
    nb = lib.tfcfft_triplet_workspace_bytes() if ws_bytes is None else ws_bytes
    return lib.tfcfft_patch_triplet(ctypes.byref(d), fake, real, n, 1.0, 1e-6, out, grad, ws, nb, None)


def test_c_entry_point_rejects_bad_arguments_without_touching_the_gpu():
    # every call below fails in host-side validation (fake pointers are never dereferenced)
    assert L.load().tfcfft_triplet_workspace_bytes() >= 256
    ok16 = list(range(16))
    assert call(desc(shape=(2, 3, 256, 128)), ok16) == -4        # not square
    assert call(desc(grid=3), list(range(9))) == -4               # unsupported grid
    assert call(desc(), ok16[:15] + [16]) == -4                   # negative index out of range
    assert call(desc(), None) == -1                               # no negatives
    assert call(desc(flags=L.CHANNELS_RGB), ok16) == -7           # FFT-loss flags do not apply
    assert call(desc(), ok16, fake=None) == -1
    assert call(desc(), ok16, fake=260) == -6                     # base not aligned to 4 elements
    assert call(desc(), ok16, ws=None) == -8
    assert call(desc(), ok16, ws_bytes=128) == -8
    assert call(desc(dtype=L.U8), ok16, grad=256) == -9           # no gradient for integer inputs
    d = desc()
    d.n = 0
    assert call(d, ok16) == -10


def test_python_layer_validates_negatives_and_device():
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError):
        tfc.patch_triplet_loss(x, x, [0] * 16, grid=4)            # CPU tensors: no fallback
    with pytest.raises(TypeError):
        tfc.compat.patch_triplet([x] * 3, [x] * 3)


# ---- temperature triplet --------------------------------------------------------------------
LUT32 = np.linspace(24, 38, num=256).astype(np.float32)


def temp_inputs(case):
    f, r = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    j, _ = make_pair(case["kind"], case["seed"] + 100, (case["n"], 3, 256, 256), case["dtype"])
    return f, r, j


@pytest.mark.parametrize("case", TEMP, ids=[c["name"] for c in TEMP])
def test_temperature_oracle_matches_reference_functions(case):
    f, r, j = temp_inputs(case)
    t = otemp.vectorize_temps_r0(f)
    assert float(t.astype(np.float64).sum()) == case["temps_sum"]                     # bit-exact table gather
    assert np.array_equal(t[0, 0, 100:104], ARR[case["name"] + "_TFB_n0_rows100_104"])
    tb = otemp.vectorize_temps_r0(r)
    assert float(tb.astype(np.float64).sum()) == case["tb_sum"]
    wl, _, _, g = otemp.temperature_triplet(f, tb, j, positive_is_temps=True, weight=case["lambda_t"])
    assert g is None and wl == pytest.approx(case["loss"], rel=3e-7)


@pytest.mark.parametrize("case", TEMP, ids=[c["name"] for c in TEMP])
def test_temperature_emulated_kernel_matches_reference_and_oracle(case):
    f, r, j = temp_inputs(case)
    tb = otemp.vectorize_temps_r0(r)
    # as shipped: quantise + table, positive = the loader's temperatures, forward only
    rc, out, _ = emulate_temperature(f, tb, j, LUT32, flags=L.QUANTIZE_U8 | L.TEMPS_POSITIVE, weight=10.0, grad=False)
    assert rc == 0 and out[0] == pytest.approx(case["loss"], rel=1e-5)
    # same thing with the positive given as an image
    rc, out2, _ = emulate_temperature(f, r, j, LUT32, flags=L.QUANTIZE_U8, weight=10.0, grad=False)
    assert rc == 0 and out2[0] == pytest.approx(case["loss"], rel=1e-5)
    # the gradient is refused in quantised mode, like the reference has none
    rc, _, _ = emulate_temperature(f, r, j, LUT32, flags=L.QUANTIZE_U8)
    assert rc == -9


@pytest.mark.parametrize("side,dtype", [(64, "float32"), (256, "float32"), (256, "float16"), (512, "float32")])
def test_temperature_differentiable_variant_matches_oracle(side, dtype):
    f, r = make_pair("unit", 61, (2, 3, side, side), dtype)
    j, _ = make_pair("unit", 62, (2, 3, side, side), dtype)
    rc, out, g = emulate_temperature(f, r, j, LUT32, weight=10.0, input_scale=255.0)
    assert rc == 0
    wl, l, act, gr = otemp.temperature_triplet(f, r, j, quantize=False, weight=10.0, input_scale=255.0)
    assert out[0] == pytest.approx(wl, rel=1e-5) and out[1] == pytest.approx(l, rel=1e-5) and out[2] == pytest.approx(act, abs=1e-6)
    assert np.abs(g[:, 1:]).max() == 0.0                         # only the red channel receives a gradient
    assert l2rel(g.astype(np.float64), gr) <= (3e-3 if dtype == "float16" else 1e-5)


def test_temperature_c_entry_point_argument_checks():
    lib = L.load()
    st = (3 * 65536, 65536, 256, 1)
    d = L.make_desc(L.F32, 1, 0, (2, 3, 256, 256), st, st, st, 1.0, 255.0)
    lut = (ctypes.c_float * 256)(*LUT32)
    ns = (ctypes.c_int64 * 4)(*st)
    nb = lib.tfcfft_triplet_workspace_bytes()
    f = lib.tfcfft_temperature_triplet
    assert f(ctypes.byref(d), 256, 256, 256, ns, None, 1.0, 1e-6, 256, None, 256, nb, None) == -1     # no table
    assert f(ctypes.byref(d), 256, 256, None, ns, lut, 1.0, 1e-6, 256, None, 256, nb, None) == -1    # no negative
    assert f(ctypes.byref(d), 256, 256, 258, ns, lut, 1.0, 1e-6, 256, None, 256, nb, None) == -6     # misaligned negative
    assert f(ctypes.byref(d), 256, 256, 256, ns, lut, 1.0, 1e-6, 256, None, 256, 16, None) == -8
    d.flags = L.CHANNELS_RGB
    assert f(ctypes.byref(d), 256, 256, 256, ns, lut, 1.0, 1e-6, 256, None, 256, nb, None) == -7
    d.flags = L.QUANTIZE_U8
    assert f(ctypes.byref(d), 256, 256, 256, ns, lut, 1.0, 1e-6, 256, 256, 256, nb, None) == -9      # no gradient as shipped
    d.flags = 0
    d.h = d.w = 96
    assert f(ctypes.byref(d), 256, 256, 256, ns, lut, 1.0, 1e-6, 256, None, 256, nb, None) == -4
    assert lib.tfcfft_vectorize_temps(ctypes.byref(d), 256, lut, 256, None) == -4


# ---- regional FFT loss on the 100 x 256 bands ------------------------------------------------
@pytest.mark.parametrize("case", REG, ids=[c["name"] for c in REG])
def test_regional_oracle_and_emulated_kernel_match_the_reference_function(case):
    f, r = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    l, _, _ = oreg.regional_loss_r0(f, r)
    assert float(l) == pytest.approx(case["loss"], rel=2e-6)
    rc, out, _, _ = emulate_regional(f, r, flags=L.QUANTIZE_U8, grad=False)
    assert rc == 0 and out[0] == pytest.approx(case["loss"], rel=1e-5)
    rc, _, _, _ = emulate_regional(f, r, flags=L.QUANTIZE_U8)     # as shipped there is no gradient
    assert rc == -9


REG_OPTS = [dict(), dict(channels="rgb"), dict(distance="mse", use_phase=False), dict(use_phase=False)]


@pytest.mark.parametrize("opt", REG_OPTS, ids=["default", "rgb", "mse-amp", "amp-only"])
def test_regional_emulated_kernel_matches_r1(opt):
    f, r = make_pair("tanh", 7, (2, 3, 256, 256), "float32")
    flags = tfc.SpectralConfig(grid=1, **opt).flags()
    rc, out, per, g = emulate_regional(f, r, flags=flags, weight=0.7, input_scale=3.0)
    assert rc == 0
    l, a, p, gr = oreg.regional_loss_and_grad_r1(f, r, weight=0.7, input_scale=3.0, **opt)
    assert out[0] == pytest.approx(l, rel=1e-5) and out[1] == pytest.approx(a, rel=1e-5)
    assert out[2] == pytest.approx(p, rel=1e-5, abs=1e-12)
    assert l2rel(g, gr) <= 1e-3
    assert np.abs(g[:, :, 200:]).max() == 0.0                      # rows outside the two bands get no gradient
    assert per[:, 0].mean() == pytest.approx(a, rel=1e-5)


def test_regional_single_channel_half_and_argument_checks():
    f, r = make_pair("uniform", 9, (1, 1, 256, 256), "float16")
    rc, out, _, g = emulate_regional(f, r, input_scale=255.0)
    l, _, _, gr = oreg.regional_loss_and_grad_r1(f.astype(np.float64), r.astype(np.float64), input_scale=255.0)
    assert rc == 0 and out[0] == pytest.approx(l, rel=1e-5) and l2rel(g.astype(np.float64), gr) <= 3e-3
    lib = L.load()
    st = (3 * 128 * 128, 128 * 128, 128, 1)
    d = L.make_desc(L.F32, 1, 0, (2, 3, 128, 128), st, st, st, 1.0, 1.0)
    assert lib.tfcfft_regional_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.tfcfft_regional_loss(ctypes.byref(d), 256, 256, 256, None, None, 256, 1 << 20, None) == -4   # bands need 256 x 256
    st = (3 * 65536, 65536, 256, 1)
    d = L.make_desc(L.F32, 1, L.LOG_MAGNITUDE, (2, 3, 256, 256), st, st, st, 1.0, 1.0)
    assert lib.tfcfft_regional_loss(ctypes.byref(d), 256, 256, 256, None, None, 256, 1 << 20, None) == -7
    d = L.make_desc(L.F32, 1, 0, (2, 3, 256, 256), st, st, st, 1.0, 1.0)
    assert lib.tfcfft_regional_workspace_bytes(ctypes.byref(d)) >= 256
    assert lib.tfcfft_regional_loss(ctypes.byref(d), 256, 256, 256, None, None, None, 0, None) == -8


def test_regional_spectra_emulated_match_reference_components_and_autograd():
    from oracle.r0_literal import components_r0, gray_u8
    f, _ = make_pair("unit", 91, (2, 3, 256, 256), "float32")
    # as shipped (quantised, fftshift-ed): amplitudes against FFT_Components.make_components on each band
    rc, amp, pha = emulate_regional_spectra(f, flags=L.QUANTIZE_U8, shift=True)
    assert rc == 0
    g = gray_u8(f)
    for t in range(2):
        for b, (lo, hi) in enumerate(((0, 100), (100, 200))):
            a, p = components_r0(g[t, lo:hi])
            assert np.abs(amp[t, 0, b] - a).max() <= 2e-6 * a.max()
            big = a > 1e-3 * a.max()                                    # phases of negligible bins are noise
            assert np.abs(np.angle(np.exp(1j * (pha[t, 0, b] - p))))[big].max() <= 1e-3
    # differentiable variant: d/dx of sum(w_a * amp + w_p * pha) against torch autograd (fp64)
    rs = np.random.RandomState(0)
    wa = rs.normal(size=(2, 1, 2, 100, 129)).astype(np.float32)
    wp = (1e-3 * rs.normal(size=(2, 1, 2, 100, 129))).astype(np.float32)
    rc, gx = emulate_regional_spectra(f, input_scale=2.0, shift=False, grad_amp=wa, grad_pha=wp)
    assert rc == 0
    x = torch.from_numpy(f).double().requires_grad_(True)
    w = torch.tensor([19595.0, 38470.0, 7471.0], dtype=torch.float64).view(1, 3, 1, 1) / 65536.0
    lum = (x * 2.0 * w).sum(1)
    tot = 0.0
    for b, (lo, hi) in enumerate(((0, 100), (100, 200))):
        F = torch.fft.rfft2(lum[:, lo:hi])
        tot = tot + (torch.from_numpy(wa[:, 0, b]).double() * F.abs()).sum() + (torch.from_numpy(wp[:, 0, b]).double() * torch.angle(F)).sum()
    tot.backward()
    assert l2rel(gx.astype(np.float64), x.grad.numpy()) <= 1e-3
