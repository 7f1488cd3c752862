"""GPU parity tests of the patch triplet loss: the CUDA kernel, through the C ABI / Python surface,
against the golden vectors produced by the reference's own lines, against the fp64 oracle, and through
size-independent properties at full batch size.  Tolerances: loss rel <= 1e-4, gradient L2-rel <= 1e-3."""

import json
import os

import numpy as np
import pytest
import torch

import tfc_gan_b200 as tfc
from inputs import make_pair
from oracle import triplet as otri
from util import l2rel

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_triplet.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_triplet.npz"))
TRI = [c for c in GOLD["cases"] if "op" not in c]
REG = [c for c in GOLD["cases"] if c.get("op") == "regional"]
TEMP = [c for c in GOLD["cases"] if c.get("op") == "temperature"]
LOSS_TOL, GRAD_TOL = 1e-4, 1e-3


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("case", TRI, ids=[c["name"] for c in TRI])
def test_matches_reference_golden(case):
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), "float32")
    f = cu(fake).requires_grad_(True)
    loss = tfc.patch_triplet_loss(f, cu(real), case["negatives"], grid=case["grid"])
    loss.backward()
    g = f.grad.double().cpu().numpy()
    assert float(loss) == pytest.approx(case["loss"], rel=LOSS_TOL)
    assert np.sqrt((g * g).sum()) == pytest.approx(case["grad_l2"], rel=GRAD_TOL)
    ref = ARR[case["name"] + "_grad_n0_c1_rows60_70"]
    assert l2rel(g[0, 1, 60:70, :], ref) <= GRAD_TOL


def test_compat_block_replays_the_reference_draws():
    case = TRI[0]
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), "float32")
    f, r = cu(fake), cu(real)
    np.random.seed(case["numpy_seed"])  # the reference script's NumPy stream
    loss = tfc.compat.patch_triplet(tfc.compat.make_16_patches(f), tfc.compat.make_16_patches(r))
    assert float(loss) == pytest.approx(case["loss"], rel=LOSS_TOL)
    case = next(c for c in TRI if c["grid"] == 2)
    fake, real = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), "float32")
    f, r = cu(fake), cu(real)
    quads = [q.contiguous() for q in tfc.compat.make_4_patches(r)]  # the loader hands out separate quadrants
    np.random.seed(case["numpy_seed"])
    loss = tfc.compat.patch_triplet(tfc.compat.make_4_patches(f), quads)
    assert float(loss) == pytest.approx(case["loss"], rel=LOSS_TOL)


CASES = [(256, 4, 3, torch.float32), (256, 2, 2, torch.float32), (512, 4, 1, torch.float32),
         (512, 2, 1, torch.float32), (512, 1, 1, torch.float32), (64, 4, 5, torch.float32), (128, 2, 3, torch.float16),
         (256, 4, 2, torch.bfloat16)]


@pytest.mark.parametrize("side,grid,n,dtype", CASES, ids=[f"{s}-g{g}-n{n}-{str(d).split('.')[-1]}" for s, g, n, d in CASES])
def test_loss_and_gradient_match_oracle(side, grid, n, dtype):
    fake, real = make_pair("tanh", 17 + side + grid, (n, 3, side, side), "float32")
    f, r = cu(fake, dtype), cu(real, dtype)
    rs = np.random.RandomState(grid * 1000 + side)
    neg = [int(k) for k in rs.randint(grid * grid, size=grid * grid)]
    out, g = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=grid, margin=1.0, weight=0.3)
    wl, l, act, gr = otri.patch_triplet_loss_and_grad(f.double().cpu().numpy(), r.double().cpu().numpy(), neg, grid=grid, weight=0.3)
    out = out.cpu().numpy()
    assert out[0] == pytest.approx(wl, rel=LOSS_TOL)
    assert out[1] == pytest.approx(l, rel=LOSS_TOL)
    assert out[2] == pytest.approx(act, abs=2e-5)
    tol = GRAD_TOL if dtype == torch.float32 else (2e-3 if dtype == torch.float16 else 1e-2)  # output rounding of 16-bit gradients
    assert l2rel(g.double().cpu().numpy(), gr) <= tol
    assert g.dtype == dtype


def test_global_grid_is_the_degenerate_case():
    # grid = 1: the only possible negative is the positive itself -> every hinge equals the margin, gradient exactly 0
    fake, real = make_pair("tanh", 19, (2, 3, 256, 256), "float32")
    out, g = tfc.patch_triplet_loss_and_grad(cu(fake), cu(real), [0], grid=1, margin=1.0)
    assert float(out[1]) == pytest.approx(1.0, rel=1e-6) and float(out[2]) == 1.0
    assert float(g.abs().max()) == 0.0


def test_single_channel_views_and_odd_batch():
    fake, real = make_pair("uniform", 5, (3, 3, 256, 256), "float32")
    big_f, big_r = cu(fake), cu(real)
    f, r = big_f[:, 1:2, 128:, 128:], big_r[:, 0:1, :128, 128:]  # strided views, one channel, 128 x 128, grid 2
    neg = [2, 2, 0, 1]
    out, g = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=2)
    _, l, _, gr = otri.patch_triplet_loss_and_grad(f.cpu().numpy(), r.cpu().numpy(), neg, grid=2)
    assert float(out[1]) == pytest.approx(l, rel=LOSS_TOL)
    assert l2rel(g.cpu().numpy(), gr) <= GRAD_TOL


def test_accumulates_into_the_fft_gradient_in_the_same_pass():
    fake, real = make_pair("tanh", 23, (4, 3, 256, 256), "float32")
    f, r = cu(fake), cu(real)
    neg = list(np.random.RandomState(1).randint(16, size=16))
    _, _, g_fft = tfc.spectral_loss_and_grad(f, r, grid=4, weight=0.01, input_scale=255.0)
    _, g_tri = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=4)
    both = g_fft.clone()
    out, same = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=4, accumulate_into=both)
    assert same.data_ptr() == both.data_ptr()
    assert l2rel(both.cpu().numpy(), (g_fft + g_tri).cpu().numpy()) <= 1e-6


def test_backward_scales_by_grad_output_and_module_front_end():
    fake, real = make_pair("uniform", 29, (2, 3, 256, 256), "float32")
    r = cu(real)
    crit = tfc.PatchTripletLoss(grid=4, margin=1.0)
    f1 = cu(fake).requires_grad_(True)
    np.random.seed(7)
    (crit(f1, r) * 1.0).backward()
    neg = crit.last_negatives
    f2 = cu(fake).requires_grad_(True)
    (crit(f2, r, negatives=neg) * 655.36).backward()  # weight x GradScaler-like factor
    assert l2rel(f2.grad.cpu().numpy(), 655.36 * f1.grad.cpu().numpy()) <= 1e-6
    # half-precision training tensors: gradient comes back in the input dtype
    f3 = cu(fake, torch.float16).requires_grad_(True)
    crit(f3, r.half(), negatives=neg).backward()
    assert f3.grad.dtype == torch.float16 and l2rel(f3.grad.float().cpu().numpy(), f1.grad.cpu().numpy()) <= 5e-3


@pytest.mark.parametrize("grid", [4, 2])
def test_run_to_run_bit_stable(grid):
    fake, real = make_pair("uniform", 31, (32, 3, 256, 256), "float32")
    f, r = cu(fake), cu(real)
    neg = list(np.random.RandomState(grid).randint(grid * grid, size=grid * grid))
    outs, grads = [], []
    for _ in range(3):
        o, g = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=grid)
        outs.append(o.cpu().numpy().tobytes())
        grads.append(g.cpu().numpy().tobytes())
    assert outs[0] == outs[1] == outs[2]
    assert grads[0] == grads[1] == grads[2]


def test_full_size_properties():
    g = torch.Generator(device="cuda").manual_seed(3)
    f = torch.empty(256, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    r = torch.empty(256, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    ident = list(range(16))
    # a patch that draws itself as the negative: d_ap == d_an for every row -> loss == margin, gradient == 0
    out, grad = tfc.patch_triplet_loss_and_grad(f, r, ident, grid=4, margin=0.75)
    assert float(out[1]) == pytest.approx(0.75, rel=1e-6) and float(out[2]) == 1.0
    assert float(grad.abs().max()) == 0.0
    # linear in weight; permuting the batch leaves the loss unchanged (mean over samples)
    neg = list(np.random.RandomState(11).randint(16, size=16))
    o1, g1 = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=4, weight=1.0)
    o2, g2 = tfc.patch_triplet_loss_and_grad(f, r, neg, grid=4, weight=4.0)
    assert float(o2[0]) == pytest.approx(4.0 * float(o1[0]), rel=1e-6)
    assert torch.allclose(g2, 4.0 * g1, rtol=1e-6, atol=0)
    perm = torch.randperm(256, device="cuda", generator=g)
    o3, g3 = tfc.patch_triplet_loss_and_grad(f[perm].contiguous(), r[perm].contiguous(), neg, grid=4)
    assert float(o3[1]) == pytest.approx(float(o1[1]), rel=1e-5)
    assert torch.equal(g3, g1[perm])
    # against torch's own op on the GPU for one patch (anchor = fake patch 5, negative = real patch neg[5])
    crit = torch.nn.TripletMarginLoss(margin=1.0, p=2)
    t = lambda x, i: x[:, :, (i // 4) * 64:(i // 4 + 1) * 64, (i % 4) * 64:(i % 4 + 1) * 64]
    ref = sum(crit(t(f, i), t(r, i), t(r, neg[i])) for i in range(16)) / 16
    assert float(o1[1]) == pytest.approx(float(ref), rel=1e-5)


def test_errors_raise():
    f = torch.zeros(2, 3, 256, 256, device="cuda")
    with pytest.raises(ValueError):
        tfc.patch_triplet_loss(f, f, [0] * 15, grid=4)
    with pytest.raises(ValueError):
        tfc.patch_triplet_loss(f, f, [16] + [0] * 15, grid=4)
    with pytest.raises(RuntimeError):
        tfc.patch_triplet_loss(f[:, :, :, :128], f[:, :, :, :128], [0] * 16, grid=4)  # not square


# ---- temperature triplet --------------------------------------------------------------------
from oracle import temperature as otemp  # noqa: E402


def temp_inputs(case):
    f, r = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    j, _ = make_pair(case["kind"], case["seed"] + 100, (case["n"], 3, 256, 256), case["dtype"])
    return f, r, j


@pytest.mark.parametrize("case", TEMP, ids=[c["name"] for c in TEMP])
def test_temperature_block_matches_reference_golden(case):
    f, r, j = temp_inputs(case)
    F, R, J = cu(f), cu(r), cu(j)
    # vectorize_temps: bit-exact against the reference's own function output
    t = tfc.compat.vectorize_temps(F)
    assert t.shape == (case["n"], 1, 256, 256) and t.dtype == torch.float32
    assert float(t.double().sum()) == case["temps_sum"]
    assert np.array_equal(t[0, 0, 100:104].cpu().numpy(), ARR[case["name"] + "_TFB_n0_rows100_104"])
    tb = tfc.vectorize_temps(R)[:, 0]                         # what the loader hands out: [N, H, W] temperatures
    assert float(tb.double().sum()) == case["tb_sum"]
    tfc.compat.set_mode("r0")
    try:
        loss = tfc.compat.temperature_loss(F, tb, J)          # the reference block, lambda_t = 10
    finally:
        tfc.compat.set_mode("r1")
    assert float(loss) == pytest.approx(case["loss"], rel=LOSS_TOL)
    assert not loss.requires_grad


@pytest.mark.parametrize("side,n,dtype", [(256, 3, torch.float32), (512, 1, torch.float32), (64, 4, torch.float32), (256, 2, torch.float16)])
def test_temperature_differentiable_variant(side, n, dtype):
    f, r = make_pair("unit", 71 + side, (n, 3, side, side), "float32")
    j, _ = make_pair("unit", 72 + side, (n, 3, side, side), "float32")
    F, R, J = cu(f, dtype).requires_grad_(True), cu(r, dtype), cu(j, dtype)
    loss = tfc.temperature_triplet_loss(F, R, J, weight=10.0, input_scale=255.0)
    (loss * 3.0).backward()
    wl, _, _, gr = otemp.temperature_triplet(F.detach().double().cpu().numpy(), R.double().cpu().numpy(), J.double().cpu().numpy(),
                                             quantize=False, weight=10.0, input_scale=255.0)
    assert float(loss) == pytest.approx(wl, rel=LOSS_TOL)
    g = F.grad.double().cpu().numpy()
    assert np.abs(g[:, 1:]).max() == 0.0
    assert l2rel(g, 3.0 * gr) <= (GRAD_TOL if dtype == torch.float32 else 3e-3)
    # positive given as the loader's temperature tensor instead of an image: same numbers
    tb = (24.0 + (38.0 - 24.0) / 255.0 * 255.0 * R.float()[:, 0]).contiguous()
    out, _ = tfc.temperature_triplet_loss_and_grad(F.detach(), tb, J, weight=10.0, input_scale=255.0)
    assert float(out[0]) == pytest.approx(wl, rel=1e-4)


def test_temperature_full_size_and_determinism():
    g = torch.Generator(device="cuda").manual_seed(5)
    F = torch.empty(256, 3, 256, 256, device="cuda").uniform_(0, 1, generator=g)
    R = torch.empty(256, 3, 256, 256, device="cuda").uniform_(0, 1, generator=g)
    J = torch.empty(256, 3, 256, 256, device="cuda").uniform_(0, 1, generator=g)
    o1, g1 = tfc.temperature_triplet_loss_and_grad(F, R, J, weight=10.0)
    o2, g2 = tfc.temperature_triplet_loss_and_grad(F, R, J, weight=10.0)
    assert torch.equal(o1, o2) and torch.equal(g1, g2)
    # torch's own op on the materialised (linear) temperatures
    lin = lambda x: 24.0 + 14.0 * x[:, 0:1]
    ref = 10.0 * torch.nn.TripletMarginLoss(margin=1.0, p=2)(lin(F), lin(R), lin(J))
    assert float(o1[0]) == pytest.approx(float(ref), rel=1e-5)
    # negative == positive -> every hinge equals the margin, zero gradient
    o3, g3 = tfc.temperature_triplet_loss_and_grad(F, R, R, weight=1.0, margin=0.5)
    assert float(o3[1]) == pytest.approx(0.5, rel=1e-6) and float(g3.abs().max()) == 0.0



# ---- regional FFT loss on the 100 x 256 bands ------------------------------------------------
from oracle import regional as oreg  # noqa: E402


@pytest.mark.parametrize("case", REG, ids=[c["name"] for c in REG])
def test_regional_loss_matches_reference_golden(case):
    f, r = make_pair(case["kind"], case["seed"], (case["n"], 3, 256, 256), case["dtype"])
    tfc.compat.set_mode("r0")
    try:
        loss = tfc.compat.regional_fft_loss(cu(f), cu(r))
    finally:
        tfc.compat.set_mode("r1")
    assert float(loss) == pytest.approx(case["loss"], rel=LOSS_TOL) and not loss.requires_grad


@pytest.mark.parametrize("opt", [dict(), dict(channels="rgb"), dict(distance="mse", use_phase=False)], ids=["default", "rgb", "mse-amp"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_regional_loss_and_gradient_match_oracle(opt, dtype):
    f, r = make_pair("tanh", 81, (3, 3, 256, 256), "float32")
    F, R = cu(f, dtype).requires_grad_(True), cu(r, dtype)
    loss, terms = tfc.regional_spectral_loss(F, R, return_terms=True, weight=0.01, input_scale=255.0, **opt)
    (loss * 2.0).backward()
    l, a, p, gr = oreg.regional_loss_and_grad_r1(F.detach().double().cpu().numpy(), R.double().cpu().numpy(), weight=0.01,
                                                 input_scale=255.0, **opt)
    assert float(loss) == pytest.approx(l, rel=LOSS_TOL)
    assert float(terms[0]) == pytest.approx(a, rel=LOSS_TOL)
    g = F.grad.double().cpu().numpy()
    assert l2rel(g, 2.0 * gr) <= (GRAD_TOL if dtype == torch.float32 else 3e-3)
    assert np.abs(g[:, :, 200:]).max() == 0.0


def test_regional_full_batch_properties():
    g = torch.Generator(device="cuda").manual_seed(9)
    F = torch.empty(64, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    R = torch.empty(64, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
    l1, t1, g1 = tfc.regional_spectral_loss_and_grad(F, R, weight=0.01, input_scale=255.0)
    l2, t2, g2 = tfc.regional_spectral_loss_and_grad(F, R, weight=0.01, input_scale=255.0)
    assert torch.equal(l1, l2) and torch.equal(g1, g2)                           # deterministic
    l0, _, g0 = tfc.regional_spectral_loss_and_grad(F, F.clone(), weight=0.01, input_scale=255.0)
    assert 0.0 <= float(l0) <= 1e-5 * float(l1)                                   # identical inputs: zero up to fp32 un-mixing noise
    # the loss only sees rows 0..199: changing the rest changes nothing
    F2 = F.clone()
    F2[:, :, 200:] = 0.123
    l3, _, _ = tfc.regional_spectral_loss_and_grad(F2, R, weight=0.01, input_scale=255.0)
    assert torch.equal(l1, l3)
    # torch.fft on the GPU as an independent yard-stick
    w = torch.tensor([19595.0, 38470.0, 7471.0], device="cuda").view(1, 3, 1, 1) / 65536.0
    lum = lambda x: (x * 255.0 * w).sum(1)
    ref = 0.0
    for lo, hi in ((0, 100), (100, 200)):
        a, b = torch.fft.rfft2(lum(F)[:, lo:hi]), torch.fft.rfft2(lum(R)[:, lo:hi])
        ref = ref + (a.abs() - b.abs()).abs().mean() + (torch.angle(a) - torch.angle(b)).abs().mean()
    assert float(l1) == pytest.approx(0.01 * 0.5 * float(ref), rel=2e-4)


def test_regional_components_differentiable_and_kl_variant_composes():
    f, r = make_pair("unit", 93, (3, 3, 256, 256), "float32")
    F, R = cu(f).requires_grad_(True), cu(r)
    (Ah, Ph), (Ae, Pe) = tfc.compat.regional_components(F)
    assert Ah.shape == (3, 1, 100, 129) and Pe.shape == (3, 1, 100, 129)
    # the L1 loss rebuilt from the materialised spectra equals the fused kernel's
    (Ahr, Phr), (Aer, Per) = tfc.compat.regional_components(R)
    rebuilt = 0.5 * ((Ah - Ahr).abs().mean() + (Ae - Aer).abs().mean() + (Ph - Phr).abs().mean() + (Pe - Per).abs().mean())
    fused = tfc.compat.regional_fft_loss(F, R)
    assert float(rebuilt) == pytest.approx(float(fused), rel=1e-4)
    # the KL variant's criterion (log_softmax over the batch + KLDivLoss, ..._withregion_FFT_KL.py:84, 401-413) through autograd
    kl = torch.nn.KLDivLoss(reduction="mean", log_target=True)
    ls = lambda t: torch.nn.functional.log_softmax(t, dim=0)
    loss = kl(ls(Ah), ls(Ahr)) + kl(ls(Ae), ls(Aer))
    loss.backward()
    g = F.grad
    assert torch.isfinite(g).all() and float(g.abs().sum()) > 0 and float(g[:, :, 200:].abs().max()) == 0.0
    # against torch.fft autograd for the same criterion
    X = cu(f).double().requires_grad_(True)
    w = torch.tensor([19595.0, 38470.0, 7471.0], device="cuda", dtype=torch.float64).view(1, 3, 1, 1) / 65536.0
    lumx, lumr = (X * 255.0 * w).sum(1, keepdim=True), (R.double() * 255.0 * w).sum(1, keepdim=True)
    ref = 0.0
    for lo, hi in ((0, 100), (100, 200)):
        ref = ref + kl(ls(torch.fft.rfft2(lumx[:, :, lo:hi]).abs()), ls(torch.fft.rfft2(lumr[:, :, lo:hi]).abs()))
    ref.backward()
    assert float(loss) == pytest.approx(float(ref), rel=1e-3, abs=1e-9)
    assert l2rel(g.double().cpu().numpy(), X.grad.cpu().numpy()) <= 5e-3
