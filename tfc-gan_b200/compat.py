"""Drop-in replacements that keep the reference's call signatures.

The reference has no operator API: the FFT helpers are module-level functions copy-pasted into
every training script, reading a global ``opt``.  These functions take the same
positional arguments and return the same kind of value, but shapes come from the tensors and the
work runs in the fused CUDA path -- with a gradient, which the reference lacks.

Two modes (``set_mode``):

* ``"r1"`` (default): differentiable.  Luma with Pillow's coefficients, no uint8 quantisation,
  ``input_scale = 255`` so magnitudes are on the reference's 8-bit scale.
* ``"r0"``: the reference as shipped -- uint8 wrap + integer luma (``...patchFFT_16P.py:300``), forward
  only; the result carries no gradient, exactly like the reference's detached loss.
"""

from __future__ import annotations

import torch

import numpy as np

from .functional import SpectralConfig, patch_triplet_loss, regional_components as _regional_components, regional_spectral_loss, temperature_triplet_loss, vectorize_temps as _vectorize_temps, spectral_components, spectral_loss, spectral_terms_per_image

_MODE = {"mode": "r1", "input_scale": 255.0}


def set_mode(mode: str = "r1", input_scale: float = 255.0) -> None:
    if mode not in ("r0", "r1"):
        raise ValueError("mode must be 'r0' or 'r1'")
    _MODE["mode"], _MODE["input_scale"] = mode, float(input_scale)


import contextlib


@contextlib.contextmanager
def reference_mode():
    """``with compat.reference_mode():`` -- the reference as shipped (``"r0"``) inside the block."""
    old = dict(_MODE)
    set_mode("r0")
    try:
        yield
    finally:
        _MODE.update(old)


def _cfg(grid, patch_reduce="mean", weight=1.0):
    if _MODE["mode"] == "r0":
        return SpectralConfig(grid=grid, patch_reduce=patch_reduce, weight=weight, quantize=True)
    return SpectralConfig(grid=grid, patch_reduce=patch_reduce, weight=weight, input_scale=_MODE["input_scale"])


def fft_components(thermal_tensor, patch=True):
    """``fft_components(x, patch=True)`` (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:293-319``; global
    variants ``TFCGAN_multigpu_globalFFT.py:266-284``): ``[N,3,p,p]`` -> ``(AMP, PHA)`` each ``[N,1,p,p/2+1]`` fp32,
    fftshift-ed like the reference.  Shapes come from the tensor (``patch`` is accepted and ignored; upstream it
    only selected hard-coded reshape constants).  Differentiable in ``"r1"`` mode."""
    if _MODE["mode"] == "r0":
        return spectral_components(thermal_tensor, quantize=True)
    return spectral_components(thermal_tensor, input_scale=_MODE["input_scale"])


def sample_spectra(thermal_tensor):
    """``sample_spectra`` / ``FFT_Components.make_spectra`` (``...patchFFT_16P.py:284-289,378-388``):
    ``log|fftshift(fft2(L))|`` of every image, ``[N,1,H,W]`` (``-inf`` where a bin is exactly zero)."""
    kw = dict(quantize=True) if _MODE["mode"] == "r0" else dict(input_scale=_MODE["input_scale"])
    amp, _ = spectral_components(thermal_tensor, spectrum="full", log_magnitude=True, **kw)
    return amp


def make_16_patches(B):
    """``make_16_patches`` (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:227-253``): 16 row-major
    views ``B1..B16`` of side ``H/4`` (B1..B4 = top row, left to right)."""
    p = B.shape[2] // 4
    return tuple(B[:, :, y * p:(y + 1) * p, x * p:(x + 1) * p] for y in range(4) for x in range(4))


def make_4_patches(B):
    """The quadrant slicing of ``TFCGAN_multigpu_patchFFT.py:468-471`` (TL, TR, BL, BR)."""
    p = B.shape[2] // 2
    return tuple(B[:, :, y * p:(y + 1) * p, x * p:(x + 1) * p] for y in range(2) for x in range(2))


def _common_base(patches, g):
    """If ``patches`` are the row-major ``g x g`` tiling views of one NCHW tensor, return it."""
    b = patches[0]._base if patches[0]._base is not None else None
    if b is None or b.dim() != 4:
        return None
    p = patches[0].shape[-1]
    if b.shape[2] != g * p or b.shape[3] != g * p or b.shape[:2] != patches[0].shape[:2]:
        return None
    for i, t in enumerate(patches):
        y, x = divmod(i, g)
        if t._base is not b or t.shape != patches[0].shape or t.stride() != b.stride():
            return None
        if t.storage_offset() != b.storage_offset() + y * p * b.stride(2) + x * p * b.stride(3):
            return None
    return b


def _assemble(patches, g):
    b = _common_base(patches, g)
    if b is not None:
        return b
    rows = [torch.cat(list(patches[y * g:(y + 1) * g]), dim=-1) for y in range(g)]
    return torch.cat(rows, dim=-2)


def calculate_ffts(*patches):
    """``calculate_ffts(fake_B1..fake_B16, B1..B16)`` (``...patchFFT_16P.py:323-375``) -> scalar
    ``1/2 (1/16 sum L1(A) + 1/16 sum L1(P))``.  Views of a common tensor are routed to one fused
    launch on that tensor; separately allocated patches are concatenated first."""
    if len(patches) != 32:
        raise TypeError(f"calculate_ffts expects 32 tensors, got {len(patches)}")
    fake = _assemble(patches[:16], 4)
    real = _assemble(patches[16:], 4)
    return spectral_loss(fake, real, config=_cfg(4))


def _loss_quadrants(fake_B, quads, reduce):
    """4-patch loss on the loader's quadrant tensors: views of one tensor go to the plain launch, separately allocated
    quadrants are handed to the kernels as four base pointers (``tfcfft_loss_quads``) -- no concatenation copy."""
    b = _common_base(quads, 2)
    if b is not None:
        return spectral_loss(fake_B, b, config=_cfg(2, reduce))
    return spectral_loss(fake_B, quads[0], config=_cfg(2, reduce), real_quadrants=quads)


def fft_loss(fake_B, B1, B2, B3, B4):
    """``fft_loss`` (``TFCGAN_multigpu_patchFFT_experiment.py:317-339``): 4-patch loss with the patch
    terms SUMMED; the real quadrants arrive as separate tensors from the loader
    (``datasets_temp.py:76-118``)."""
    return _loss_quadrants(fake_B, (B1, B2, B3, B4), "sum")


def patch4_fft_loss(fake_B, B1, B2, B3, B4):
    """The inline 4-patch block of ``TFCGAN_multigpu_patchFFT.py:498-511`` (patch terms averaged)."""
    return _loss_quadrants(fake_B, (B1, B2, B3, B4), "mean")


def draw_negatives(patch_num: int):
    """The reference's negative sampling, call for call: one ``np.random.randint(patch_num, size=1).item()`` per patch,
    in patch order (``...patchFFT_16P.py:567-582``) -- seed NumPy the same way and the same negatives come out.  A
    patch may draw itself, exactly as upstream."""
    return [np.random.randint(patch_num, size=1).item() for _ in range(patch_num)]


def patch_triplet(fake_patches, real_patches, negatives=None, margin: float = 1.0):
    """The inline patch-triplet block of the generator step (``...patchFFT_16P.py:558-583``; 4-patch copies
    ``TFCGAN_multigpu_patchFFT.py:474-484``): ``1/n sum_i triplet_loss(fake_B_i, B_i, random_patches[k_i])`` with
    ``triplet_loss = nn.TripletMarginLoss(margin=1.0, p=2)`` (``:75``).  ``fake_patches`` / ``real_patches`` are the 4
    or 16 row-major patches (views of a common tensor are used in place); ``negatives`` defaults to
    :func:`draw_negatives`."""
    n = len(fake_patches)
    if n not in (4, 16) or len(real_patches) != n:
        raise TypeError(f"patch_triplet expects 4 or 16 fake and as many real patches, got {n} / {len(real_patches)}")
    g = 2 if n == 4 else 4
    if negatives is None:
        negatives = draw_negatives(n)
    return patch_triplet_loss(_assemble(tuple(fake_patches), g), _assemble(tuple(real_patches), g), negatives, grid=g, margin=margin)


def triplet_patches(fake_B, B1, B2, B3, B4, negatives=None, margin: float = 1.0):
    """``triplet_patches(fake_B, B1, B2, B3, B4) -> (Amp_loss, Pha_loss, Patch_loss)``
    (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_debiased_V5.py:386-443``): ``nn.TripletMarginLoss(margin=1, p=2)`` on the
    AMPLITUDE and on the PHASE spectra of the four quadrants (``criterion_amp``, ``:93-94``; anchor = fake quadrant,
    positive = real quadrant, negative = a randomly drawn real quadrant ``K_i``) plus the pixel-space patch triplet with the
    same negatives.  The spectra come from the differentiable :func:`fft_components` (upstream they are detached), the
    pixel term from the fused patch-triplet kernel; ``negatives`` defaults to the reference's four NumPy draws."""
    if negatives is None:
        negatives = draw_negatives(4)
    fq = make_4_patches(fake_B)
    rq = (B1, B2, B3, B4)
    crit = torch.nn.TripletMarginLoss(margin=margin, p=2)
    sf = [fft_components(q) for q in fq]
    sr = [fft_components(q) for q in rq]  # K_i is one of B1..B4: its spectra are already here
    amp = sum(crit(sf[i][0], sr[i][0], sr[negatives[i]][0]) for i in range(4)) / 4
    pha = sum(crit(sf[i][1], sr[i][1], sr[negatives[i]][1]) for i in range(4)) / 4
    patch = patch_triplet(fq, rq, negatives, margin=margin)
    return amp, pha, patch


def vectorize_temps(fake_B):
    """``vectorize_temps`` (``...patchFFT_16P.py:260-268``): ``[N,1,H,W]`` fp32 temperatures of the red channel."""
    return _vectorize_temps(fake_B)


def temperature_loss(fake_B, TB, B_tf, lambda_t: float = 10.0):
    """The temperature block of the generator step (``...patchFFT_16P.py:585-595``):
    ``criterion_temp(vectorize_temps(fake_B), TB, vectorize_temps(B_tf)) * lambda_t`` in one fused pass.  ``TB`` is the
    loader's ``T_B`` (``[N,H,W]``, reshaped like ``:593``); ``B_tf`` the colour-jittered real batch.  Mode ``"r0"``
    reproduces the reference (no gradient), ``"r1"`` is differentiable."""
    return temperature_triplet_loss(fake_B, TB, B_tf, quantize=(_MODE["mode"] == "r0"), weight=lambda_t,
                                    input_scale=_MODE["input_scale"])


def regional_fft_loss(fake_B, real_B):
    """``regional_fft_loss(fake_B, real_B)`` (``TFCGAN_multigpu_patchFFT_withregion_FFT.py:353-402``): hair (rows 0..99) and
    eyes (rows 100..199) bands, amplitude + phase L1, bands summed, ``1/2 (amp + pha)``."""
    if _MODE["mode"] == "r0":
        return regional_spectral_loss(fake_B, real_B, quantize=True)
    return regional_spectral_loss(fake_B, real_B, input_scale=_MODE["input_scale"])


def regional_components(thermal_tensor):
    """The four tensors ``reg_fft`` produces for one image batch in ``regional_fft_loss``
    (``TFCGAN_multigpu_patchFFT_withregion_FFT.py:358-389``): ``((A_hair, P_hair), (A_eyes, P_eyes))``, each
    ``[N,1,100,129]`` fp32 in the reference's fftshift-ed layout -- differentiable in ``"r1"`` mode, so the KL variant
    (``..._withregion_FFT_KL.py:398-414``) or any other criterion can be applied with plain torch ops."""
    kw = dict(quantize=True) if _MODE["mode"] == "r0" else dict(input_scale=_MODE["input_scale"])
    amp, pha = _regional_components(thermal_tensor, **kw)
    return (amp[:, :, 0], pha[:, :, 0]), (amp[:, :, 1], pha[:, :, 1])


def global_fft_loss(fake_B, real_B):
    """The inline global block of ``TFCGAN_multigpu_globalFFT.py:494-499``."""
    return spectral_loss(fake_B, real_B, config=_cfg(1))


def global_fourier_loss(BR, fake_B):
    """``global_fourier_loss(BR, fake_B)`` (``TFC-STN/TFCGAN_STN21_Original_NewModel3_B2A.py:461-467``):
    global loss times the 0.01 lambda; note the (real, fake) argument order."""
    return spectral_loss(fake_B, BR, config=_cfg(1, weight=0.01))


def mse_spec(real_gray, fake_gray, metric: str = "mse"):
    """``mse_spec`` (``TFC-GAN-FFT/Devcom_MagMSE.py:91-118``; ``metric="mae"``:
    ``eval/Eurecom/Eurecom_MagOther.py:90-117``) on uint8 grey images given as ``[N,H,W]`` /
    ``[N,1,H,W]`` CUDA tensors.  Returns ``(values, mean)``: per-pair scores with the pairs whose log
    spectrum is not finite dropped (the reference's ``except: continue``), and their mean."""
    r = real_gray if real_gray.dim() == 4 else real_gray[:, None]
    f = fake_gray if fake_gray.dim() == 4 else fake_gray[:, None]
    cfg = SpectralConfig(grid=1, use_phase=False, distance="mse" if metric == "mse" else "l1",
                         log_magnitude=True, spectrum="full")
    per = spectral_terms_per_image(f, r, config=cfg)[:, 0]
    values = per[torch.isfinite(per)]
    return values, values.mean()
