"""GPU parity tests (``-m gpu``): the CUDA path, called through the C ABI via the public Python
surface, against the oracle on the same seeded inputs, against the committed golden vectors produced
by the reference's own code, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Tolerances are BASELINE.json's: loss rel <= 1e-4, gradient L2-rel <= 1e-3 (fp64 oracle).
"""

import json
import os

import numpy as np
import pytest
import torch

import oracle
import tfc_gan_b200 as tfc
from inputs import make_gray_pairs, make_pair
from util import l2rel

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_cases.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_arrays.npz"))
LOSS_TOL, GRAD_TOL = 1e-4, 1e-3


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def run_cuda(fake, real, **opt):
    f = cu(fake).requires_grad_(True)
    r = cu(real)
    loss, terms = tfc.spectral_loss(f, r, return_terms=True, **opt)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), terms.cpu().numpy(), f.grad.cpu().numpy()


CASES = [
    (256, 4, dict()),                                   # config 1 / 3: patch-FFT-16
    (256, 4, dict(channels="rgb")),
    (256, 2, dict()),                                   # config 4: patch-FFT-4
    (256, 2, dict(patch_reduce="sum")),                 # fft_loss convention
    (256, 1, dict()),                                   # config 2: global
    (256, 1, dict(channels="rgb")),
    (256, 4, dict(distance="mse")),
    (256, 4, dict(use_phase=False, channels="rgb", distance="mse")),
    (256, 1, dict(use_phase=False, log_magnitude=True, spectrum="full", distance="mse")),  # MagMSE algebra
    (512, 4, dict()),                                   # config 5: 512^2 patch-16
    (512, 1, dict()),                                   # config 5: 512^2 global
    (128, 4, dict()),
    (64, 4, dict(spectrum="full")),
    (64, 1, dict()),
    (256, 4, dict(force_generic=True)),                 # generic resident kernel
    (256, 4, dict(use_pair=True)),                      # packed pair kernel (the default is the thread-per-line kernel)
    (256, 4, dict(use_pair=True, channels="rgb", distance="mse")),   # pair kernel on single-channel tiles
    (256, 4, dict(force_split=True)),                   # split kernels on a size the resident path also covers
    (256, 2, dict(force_split=True, channels="rgb")),
    (256, 1, dict(force_split=True)),                   # the split kernels at their native size
    (256, 2, dict(force_generic=True)),                 # generic resident kernel at 128 x 128
    (256, 1, dict(distance="mse", channels="rgb")),     # sub-tile path (D = 4), other modes
    (256, 2, dict(use_phase=False, patch_reduce="sum")),  # sub-tile path (D = 2)
]


@pytest.mark.parametrize("side,grid,opt", CASES, ids=[f"{s}-g{g}-{'-'.join(f'{k}={v}' for k, v in o.items()) or 'default'}" for s, g, o in CASES])
@pytest.mark.parametrize("kind", ["uniform", "tanh"])
def test_loss_and_gradient_match_oracle(side, grid, opt, kind):
    n = 3 if side <= 256 else 2
    fake, real = make_pair(kind, 101, (n, 3, side, side), "float32")
    loss, terms, grad = run_cuda(fake, real, grid=grid, weight=0.01, input_scale=255.0, **opt)
    okw = {k: v for k, v in opt.items() if not (k.startswith("force_") or k.startswith("use_l") or k == "use_pair")}
    l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=grid, weight=0.01, input_scale=255.0, **okw)
    assert loss == pytest.approx(l, rel=LOSS_TOL)
    assert terms[0] == pytest.approx(a, rel=LOSS_TOL)
    assert terms[1] == pytest.approx(p, rel=LOSS_TOL, abs=1e-12)
    assert l2rel(grad, g) <= GRAD_TOL


def test_phase_dominated_gradient():
    """input_scale = 1 on [-1,1] data: |F| is small, the phase term dominates the gradient -- the hardest
    case for the fast atan2 / rsqrt of the packed path."""
    fake, real = make_pair("uniform", 55, (2, 3, 256, 256), "float32")
    for opt in (dict(), dict(use_pair=True), dict(force_generic=True)):
        loss, terms, grad = run_cuda(fake, real, grid=4, **opt)
        l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=4)
        assert loss == pytest.approx(l, rel=LOSS_TOL)
        assert terms[1] == pytest.approx(p, rel=LOSS_TOL)
        assert l2rel(grad, g) <= GRAD_TOL


def test_odd_tile_count_and_single_channel():
    fake, real = make_pair("tanh", 56, (3, 1, 64, 64), "float32")  # 3 tiles: the last pair is half empty
    loss, terms, grad = run_cuda(fake, real, grid=1, input_scale=255.0)
    l, a, p, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=1, input_scale=255.0)
    assert loss == pytest.approx(l, rel=LOSS_TOL)
    assert l2rel(grad, g) <= GRAD_TOL


def test_image_like_inputs():
    fake, real = make_pair("lowpass", 5, (2, 3, 256, 256), "float32")
    for grid in (1, 2, 4):
        loss, _, grad = run_cuda(fake, real, grid=grid, input_scale=255.0)
        l, _, _, g = oracle.spectral_loss_and_grad_r1(fake, real, grid=grid, input_scale=255.0)
        assert loss == pytest.approx(l, rel=LOSS_TOL)
        assert l2rel(grad, g) <= GRAD_TOL


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("grid", [4, 1])
def test_half_precision_io(dtype, grid):
    """The reference feeds HalfTensor (``...patchFFT_16P.py:524-530``); values are exact in fp32 inside."""
    fake, real = make_pair("tanh", 9, (2, 3, 256, 256), "float32")
    f = cu(fake, dtype).requires_grad_(True)
    r = cu(real, dtype)
    loss = tfc.spectral_loss(f, r, grid=grid, input_scale=255.0)
    loss.backward()
    l, _, _, g = oracle.spectral_loss_and_grad_r1(f.detach().double().cpu().numpy(), r.double().cpu().numpy(),
                                                  grid=grid, input_scale=255.0)
    assert loss.item() == pytest.approx(l, rel=LOSS_TOL)
    assert f.grad.dtype == dtype
    # the gradient is rounded to the 16-bit storage type on the way out: against the oracle gradient ROUNDED THE SAME
    # WAY the contract tolerance (1e-3) holds -- the kernel's only extra error is that one rounding (fp16 gradients are
    # stored with a power-of-two prescale that backward divides out, so they do not go subnormal)
    want = torch.from_numpy(g).to(dtype).float().numpy()
    assert l2rel(f.grad.float().cpu().numpy(), want) <= 1e-3
    assert l2rel(f.grad.float().cpu().numpy(), g) <= (2e-3 if dtype == torch.float16 else 1e-2)


LOSS_CASES = [c for c in GOLD["cases"] if "loss" in c]


@pytest.mark.parametrize("case", LOSS_CASES, ids=[c["name"] for c in LOSS_CASES])
def test_quantised_mode_matches_reference_golden(case):
    """``quantize=True`` reproduces the numbers the reference's own functions produced
    (``tests/golden/make_golden.py``)."""
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    loss, terms = tfc.spectral_loss(cu(fake), cu(real), grid=case["grid"], patch_reduce=case["patch_reduce"],
                                    weight=case.get("weight", 1.0), quantize=True, return_terms=True)
    assert loss.item() == pytest.approx(case["loss"], rel=LOSS_TOL)
    if "amp" in case:
        assert terms[0].item() == pytest.approx(case["amp"], rel=LOSS_TOL)
        assert terms[1].item() == pytest.approx(case["pha"], rel=LOSS_TOL)


def test_quantised_mode_has_no_gradient_like_the_reference():
    fake, real = make_pair("uniform", 1, (1, 3, 256, 256), "float32")
    f = cu(fake).requires_grad_(True)
    loss = tfc.spectral_loss(f, cu(real), grid=4, quantize=True)
    assert not loss.requires_grad or f.grad is None
    assert loss.item() == pytest.approx(float(oracle.spectral_loss_r0(fake, real, 4)[0]), rel=LOSS_TOL)


def test_compat_signatures():
    from tfc_gan_b200 import compat

    fake, real = make_pair("uniform", 31, (2, 3, 256, 256), "float32")
    f, r = cu(fake).requires_grad_(True), cu(real)
    # calculate_ffts(fake_B1..16, B1..16) on views of a common tensor -> one fused launch
    tfc.reset_launch_count()
    loss16 = compat.calculate_ffts(*compat.make_16_patches(f), *compat.make_16_patches(r))
    assert tfc.launch_count() == 1
    l16 = oracle.spectral_loss_r1(torch.from_numpy(fake), torch.from_numpy(real), grid=4, input_scale=255.0)[0]
    assert loss16.item() == pytest.approx(float(l16), rel=LOSS_TOL)
    loss16.backward()
    assert f.grad is not None and torch.isfinite(f.grad).all()
    # separately allocated patches (the loader's B1..B4) are assembled first
    quads = [q.contiguous() for q in compat.make_4_patches(r)]
    ls = compat.fft_loss(f, *quads)
    lm = compat.patch4_fft_loss(f, *quads)
    assert ls.item() == pytest.approx(4 * lm.item(), rel=1e-6)
    l4 = oracle.spectral_loss_r1(torch.from_numpy(fake), torch.from_numpy(real), grid=2, input_scale=255.0)[0]
    assert lm.item() == pytest.approx(float(l4), rel=LOSS_TOL)
    lg = compat.global_fft_loss(f, r)
    assert compat.global_fourier_loss(r, f).item() == pytest.approx(0.01 * lg.item(), rel=1e-5)
    # reference-as-shipped mode against the golden vectors
    compat.set_mode("r0")
    try:
        case = next(c for c in GOLD["cases"] if c["name"] == "p16_uniform_11_float32")
        gf, gr = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
        v = compat.calculate_ffts(*compat.make_16_patches(cu(gf)), *compat.make_16_patches(cu(gr)))
        assert v.item() == pytest.approx(case["loss"], rel=LOSS_TOL)
    finally:
        compat.set_mode("r1")


def test_patch_views_pass_through_without_copy():
    """``B[:, :, 64:128, 64:128]`` (``make_16_patches``) is consumed with its strides."""
    fake, real = make_pair("uniform", 4, (2, 3, 256, 256), "float32")
    f, r = cu(fake), cu(real)
    fv, rv = f[:, :, 64:128, 64:128], r[:, :, 128:192, 0:64]
    loss = tfc.spectral_loss(fv, rv, grid=1)
    l = oracle.spectral_loss_r1(fv.cpu(), rv.cpu(), grid=1)[0]
    assert loss.item() == pytest.approx(float(l), rel=LOSS_TOL)
    # unaligned view (odd column offset) falls back to a contiguous copy, same value
    fo, ro = f[:, :, 0:64, 1:65], r[:, :, 0:64, 1:65]
    lo = oracle.spectral_loss_r1(fo.cpu(), ro.cpu(), grid=1)[0]
    assert tfc.spectral_loss(fo, ro, grid=1).item() == pytest.approx(float(lo), rel=LOSS_TOL)


def test_mag_mse_metric_matches_reference_golden():
    from tfc_gan_b200 import compat

    for metric in ("mse", "mae"):
        case = next(c for c in GOLD["cases"] if c["name"] == f"mag_{metric}_71")
        reals, fakes = make_gray_pairs(case["seed"], case["n"], case["side"])
        values, mean = compat.mse_spec(cu(np.stack(reals)), cu(np.stack(fakes)), metric)
        np.testing.assert_allclose(values.cpu().numpy(), case["values"], rtol=LOSS_TOL)
        assert mean.item() == pytest.approx(np.mean(case["values"]), rel=LOSS_TOL)


def test_backward_scales_by_grad_output():
    fake, real = make_pair("uniform", 8, (2, 3, 256, 256), "float32")
    f1 = cu(fake).requires_grad_(True)
    tfc.spectral_loss(f1, cu(real), grid=4).backward()
    f2 = cu(fake).requires_grad_(True)
    (tfc.spectral_loss(f2, cu(real), grid=4) * 0.01 * 65536.0).backward()  # weight x GradScaler scale
    torch.testing.assert_close(f2.grad, f1.grad * (0.01 * 65536.0), rtol=1e-6, atol=0)
    l, t, g = tfc.spectral_loss_and_grad(cu(fake), cu(real), grid=4)
    torch.testing.assert_close(g, f1.grad, rtol=0, atol=0)


@pytest.mark.parametrize("grid", [4, 2, 1])
def test_run_to_run_bit_stable(grid):
    """Many work units per persistent CTA (double-buffer hand-offs, drain): any race in the FULL / DONE protocol
    would show up as run-to-run differences.  Also pins the deterministic reduction."""
    g = torch.Generator(device="cuda").manual_seed(12)
    f = torch.empty(96, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    r = torch.empty(96, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    outs = [tfc.spectral_loss_and_grad(f, r, grid=grid) for _ in range(4)]
    for l, t, gr in outs[1:]:
        assert l.item() == outs[0][0].item()
        assert torch.equal(gr, outs[0][2])


@pytest.mark.parametrize("grid,n", [(1, 64), (2, 96)])
def test_back_to_back_calls_do_not_race_through_the_workspace(grid, n):
    # The three launches of the sub-tile path are chained with programmatic dependent launch and share one L2
    # workspace across calls: 40 un-synchronised calls on alternating inputs must reproduce, bit for bit, what each
    # input gives in isolation.
    g = torch.Generator(device="cuda").manual_seed(grid)
    ins = [(torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g),
            torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)) for _ in range(2)]
    cfg = tfc.SpectralConfig(grid=grid, weight=0.01, input_scale=255.0)
    ref = []
    for f, r in ins:
        loss, terms, grad = tfc.spectral_loss_and_grad(f, r, config=cfg)
        torch.cuda.synchronize()
        ref.append((loss.clone(), terms.clone(), grad.clone()))
    outs = [tfc.spectral_loss_and_grad(*ins[i % 2], config=cfg) for i in range(40)]
    torch.cuda.synchronize()
    for i, (loss, terms, grad) in enumerate(outs):
        assert torch.equal(loss, ref[i % 2][0]) and torch.equal(terms, ref[i % 2][1])
        assert torch.equal(grad, ref[i % 2][2])


def test_fast_paths_agree_with_generic_kernels_at_scale():
    """Packed pair kernel / sub-tile pipeline vs the generic resident / split kernels on a batch that gives every
    persistent CTA several units."""
    g = torch.Generator(device="cuda").manual_seed(13)
    f = torch.empty(48, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    r = torch.empty(48, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    for grid, ref_opt in ((4, dict(force_generic=True)), (4, dict(use_pair=True)), (2, dict(force_generic=True)),
                          (1, dict(force_split=True))):
        l1, t1, g1 = tfc.spectral_loss_and_grad(f, r, grid=grid, input_scale=255.0)
        l2, t2, g2 = tfc.spectral_loss_and_grad(f, r, grid=grid, input_scale=255.0, **ref_opt)
        assert l1.item() == pytest.approx(l2.item(), rel=2e-6)
        torch.testing.assert_close(t1, t2, rtol=2e-6, atol=0)
        assert l2rel(g1.cpu().numpy(), g2.cpu().numpy()) <= 2e-5


def test_finite_difference_directions():
    """Directional derivatives of the smooth variant (amplitude-only, squared distance); the phase term
    jumps by 2*pi when a bin crosses the negative real axis, so central differences do not apply to it."""
    fake, real = make_pair("tanh", 21, (1, 3, 128, 128), "float32")
    f, r = cu(fake).double(), cu(real).double()
    opt = dict(grid=2, distance="mse", use_phase=False, input_scale=4.0)
    _, _, g = tfc.spectral_loss_and_grad(f.float(), r.float(), **opt)
    rs = np.random.RandomState(0)
    for _ in range(16):
        d = torch.from_numpy(rs.normal(size=fake.shape)).cuda()
        eps = 1e-4
        lp = oracle.spectral_loss_r1((f + eps * d).cpu(), r.cpu(), **opt)[0]
        lm = oracle.spectral_loss_r1((f - eps * d).cpu(), r.cpu(), **opt)[0]
        fd = float(lp - lm) / (2 * eps)
        assert float((g.double() * d).sum()) == pytest.approx(fd, rel=1e-3)


# ---- properties at BASELINE.json's full sizes (no oracle needed) --------------------------------

@pytest.mark.parametrize("n,grid", [(64, 1), (256, 4), (128, 2)])
def test_full_size_properties(n, grid):
    g = torch.Generator(device="cuda").manual_seed(1234)
    fake = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    real = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    loss, terms, grad = tfc.spectral_loss_and_grad(fake, real, grid=grid)
    assert torch.isfinite(loss) and torch.isfinite(grad).all()
    # batch mean == mean of per-image terms == mean of two half-batch evaluations
    per = tfc.spectral_terms_per_image(fake, real, grid=grid)
    torch.testing.assert_close(per.mean(0), terms, rtol=1e-5, atol=0)
    la, ta, ga = tfc.spectral_loss_and_grad(fake[: n // 2], real[: n // 2], grid=grid)
    lb, tb, gb = tfc.spectral_loss_and_grad(fake[n // 2 :], real[n // 2 :], grid=grid)
    assert 0.5 * (la.item() + lb.item()) == pytest.approx(loss.item(), rel=1e-5)
    torch.testing.assert_close(torch.cat([ga, gb]) * 0.5, grad, rtol=1e-5, atol=1e-12)
    # sum over patches == g^2 * mean; amplitude term linear in input_scale, phase term invariant
    ls, ts, _ = tfc.spectral_loss_and_grad(fake, real, grid=grid, patch_reduce="sum")
    assert ls.item() == pytest.approx(grid * grid * loss.item(), rel=1e-5)
    _, t3, _ = tfc.spectral_loss_and_grad(fake, real, grid=grid, input_scale=3.0)
    assert t3[0].item() == pytest.approx(3 * terms[0].item(), rel=1e-5)
    assert t3[1].item() == pytest.approx(terms[1].item(), rel=1e-5)
    # permuting tiles of both images leaves the loss unchanged
    if grid > 1:
        p = 256 // grid
        roll = lambda t: torch.roll(t, shifts=(p, p), dims=(2, 3))
        lr, _, _ = tfc.spectral_loss_and_grad(roll(fake), roll(real), grid=grid)
        assert lr.item() == pytest.approx(loss.item(), rel=1e-5)
    # fake == real: the packed transform cannot be bit-symmetric, but the loss is at rounding level
    l0, t0, _ = tfc.spectral_loss_and_grad(real, real, grid=grid, distance="mse", use_phase=False)
    assert l0.item() <= 1e-9 * ts[0].item() + 1e-9


def test_errors_raise():
    x = torch.zeros(2, 3, 256, 256, device="cuda")
    with pytest.raises(RuntimeError, match="shape"):
        tfc.spectral_loss(x, x, grid=3)
    with pytest.raises(ValueError):
        tfc.spectral_loss(x, x[:1], grid=4)
    with pytest.raises(ValueError):
        tfc.spectral_loss(x[:0], x[:0], grid=4)
    with pytest.raises(RuntimeError, match="shape"):
        tfc.spectral_loss(torch.zeros(1, 3, 256, 128, device="cuda"), torch.zeros(1, 3, 256, 128, device="cuda"), grid=1)


@pytest.mark.parametrize("side", [64, 128, 256])
def test_fft_components_materialised_and_differentiable(side):
    """compat.fft_components: (AMP, PHA) [N,1,p,p/2+1], fftshift-ed, with a gradient."""
    from tfc_gan_b200 import compat

    x, _ = make_pair("tanh", 29, (2, 3, side, side), "float32")
    xt = cu(x).requires_grad_(True)
    amp, pha = compat.fft_components(xt)
    assert amp.shape == (2, 1, side, side // 2 + 1) and pha.shape == amp.shape
    ra, rp = oracle.fft_components_r1(torch.from_numpy(x), input_scale=255.0, shift=True)
    np.testing.assert_allclose(amp.detach().cpu().numpy(), ra.numpy(), rtol=2e-4, atol=2e-3)
    big = ra.numpy() > 1e-3 * ra.numpy().max()
    dphi = np.angle(np.exp(1j * (pha.detach().cpu().numpy() - rp.numpy())))
    assert np.abs(dphi[big]).max() < 2e-3
    # a downstream loss on the spectra back-propagates to x (e.g. the triplet-on-spectra variants)
    w = torch.linspace(0.5, 1.5, amp.numel(), device="cuda").reshape(amp.shape)
    (amp * w).sum().backward()
    xd = torch.from_numpy(x).double().requires_grad_(True)
    a64, _ = oracle.fft_components_r1(xd, input_scale=255.0, shift=True)
    (a64 * w.cpu().double()).sum().backward()
    assert l2rel(xt.grad.cpu().numpy(), xd.grad.numpy()) <= GRAD_TOL


def test_fft_components_r0_mode_matches_reference_spectra():
    from tfc_gan_b200 import compat

    case = next(c for c in GOLD["cases"] if c["name"] == "p16_uniform_11_float32")
    fake, _ = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    compat.set_mode("r0")
    try:
        amp, pha = compat.fft_components(compat.make_16_patches(cu(fake))[5])  # a strided view: patch B6
        spec = compat.sample_spectra(cu(fake)[:1])
    finally:
        compat.set_mode("r1")
    ga, gp = ARR["p16_uniform_11_float32_amp_B6"], ARR["p16_uniform_11_float32_pha_B6"]
    np.testing.assert_allclose(amp.cpu().numpy(), ga, rtol=1e-5, atol=0.05)
    dphi = np.angle(np.exp(1j * (pha.cpu().numpy() - gp)))
    assert np.abs(dphi[ga > 1.0]).max() < 1e-3
    ref = oracle.make_spectra_r0(oracle.gray_u8(fake[:1])[0])
    np.testing.assert_allclose(spec[0, 0].cpu().numpy(), ref, rtol=1e-4, atol=1e-3)
