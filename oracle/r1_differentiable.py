"""R1 -- differentiable CPU restatement of the reference FFT-loss path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  torch (CPU) + NumPy.

R1 is R0 (``oracle/r0_literal.py``) with the two non-differentiable steps
removed, as defined in SURVEY.md §8c; it is what BASELINE.json calls "the
reference torch.fft path" and it is the parity target of the CUDA kernels:

* no uint8 quantisation: ``x' = input_scale * x`` (``quantize=True`` re-inserts
  the reference's uint8 wrap + integer luma, forward only, for the R0 cross-pin);
* ``channels="luma"``: float ITU-R 601 luma with Pillow's fixed-point
  coefficients ``(19595, 38470, 7471) / 65536`` (the float counterpart of
  ``.convert("L")``, ``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:300``), or
  ``channels="rgb"``: every channel separately (BASELINE.json wording);
* row-major ``g x g`` tiling (``make_16_patches``, ``...patchFFT_16P.py:227-253``);
* ``torch.fft.rfft2`` per tile, unnormalised, half spectrum ``p x (p/2+1)``
  (``np.fft.rfft2`` at ``...patchFFT_16P.py:278``; the ``fftshift`` at ``:279`` is a
  permutation and cannot change a mean-reduced loss, so it is dropped);
* ``abs`` / ``angle`` (``:280-281``), optional ``log`` of the magnitude and the
  full ``p x p`` spectrum (``Devcom_MagMSE.py:91-107``);
* mean |d| (``nn.L1Loss``, ``...patchFFT_16P.py:83-84``) or mean d^2 over all
  ``N*C'*g^2*K`` bins -- equal to the reference's mean of per-patch means because
  patches are equal-sized -- times ``g^2`` for the sum-over-patches convention
  (``TFCGAN_multigpu_patchFFT_experiment.py:335-336``);
* ``loss = weight * 1/2 (amp + pha)`` (``...patchFFT_16P.py:373``), or
  ``weight * amp`` when the phase term is off (the MagMSE metric has no 1/2).

:func:`spectral_grad_analytic` is an independent NumPy fp64 evaluation of the
same loss and of its gradient by the closed form the CUDA kernels use (packed
``fake + i*real`` transform, Hermitian un-mixing, spectral gradient, zero-extended
inverse transform).  ``tests/test_oracle_golden.py`` checks it against autograd.
"""

from __future__ import annotations

import numpy as np
import torch

#: Pillow's ``convert("L")`` coefficients as exact binary fractions (sum == 1).
LUMA_WEIGHTS = (19595.0 / 65536.0, 38470.0 / 65536.0, 7471.0 / 65536.0)


def _tiles(x: torch.Tensor, grid: int) -> torch.Tensor:
    """``[N,C,H,W]`` -> ``[N,C,g,g,p,p]`` row-major tiles."""
    n, c, h, w = x.shape
    if h != w or h % grid:
        raise ValueError("square images divisible by grid only")
    p = h // grid
    return x.reshape(n, c, grid, p, grid, p).permute(0, 1, 2, 4, 3, 5)


def _prepare(x, channels, input_scale, quantize, dtype):
    if quantize:
        from .r0_literal import gray_u8

        g = gray_u8(x)  # [N,H,W] uint8
        return torch.from_numpy(g.astype(np.float64)).to(dtype)[:, None]
    x = x.to(dtype) * input_scale
    if channels == "luma" and x.shape[1] == 3:
        w = torch.tensor(LUMA_WEIGHTS, dtype=dtype).view(1, 3, 1, 1)
        return (x * w).sum(1, keepdim=True)
    if channels in ("luma", "rgb"):
        return x
    raise ValueError(channels)


def _spectra(t, spectrum):
    if spectrum == "half":
        return torch.fft.rfft2(t)
    if spectrum == "full":
        return torch.fft.fft2(t)
    raise ValueError(spectrum)


def _dist(a, b, distance):
    d = a - b
    if distance == "l1":
        return d.abs()
    if distance == "mse":
        return d * d
    raise ValueError(distance)


def spectral_loss_r1(
    fake,
    real,
    *,
    grid: int = 4,
    channels: str = "luma",
    use_phase: bool = True,
    distance: str = "l1",
    patch_reduce: str = "mean",
    log_magnitude: bool = False,
    spectrum: str = "half",
    weight: float = 1.0,
    input_scale: float = 1.0,
    quantize: bool = False,
    dtype=torch.float64,
    per_image: bool = False,
):
    """Returns ``(loss, amp_term, pha_term)`` 0-dim tensors (autograd-connected to
    ``fake``); with ``per_image=True`` also the ``[N]`` per-image amp and pha terms."""
    f = _prepare(fake, channels, input_scale, quantize, dtype)
    r = _prepare(real.detach() if hasattr(real, "detach") else real, channels, input_scale, quantize, dtype)
    F = _spectra(_tiles(f, grid), spectrum)
    R = _spectra(_tiles(r, grid), spectrum)
    af, ar = F.abs(), R.abs()
    if log_magnitude:
        af, ar = af.log(), ar.log()
    red = float(grid * grid) if patch_reduce == "sum" else 1.0
    if patch_reduce not in ("mean", "sum"):
        raise ValueError(patch_reduce)
    da = _dist(af, ar, distance)
    amp_img = da.flatten(1).mean(1) * red
    amp = amp_img.mean()
    if use_phase:
        dp = _dist(torch.angle(F), torch.angle(R), distance)
        pha_img = dp.flatten(1).mean(1) * red
        pha = pha_img.mean()
        loss = weight * 0.5 * (amp + pha)
    else:
        pha_img = torch.zeros_like(amp_img)
        pha = torch.zeros((), dtype=dtype)
        loss = weight * amp
    if per_image:
        return loss, amp, pha, amp_img, pha_img
    return loss, amp, pha


def spectral_loss_and_grad_r1(fake, real, **kw):
    """Evaluate R1 and ``d loss / d fake``.  Returns ``(loss, amp, pha, grad)`` as
    Python floats and an ndarray in the evaluation dtype."""
    dtype = kw.get("dtype", torch.float64)
    fk = torch.as_tensor(fake).detach().to(dtype).clone().requires_grad_(True)
    rl = torch.as_tensor(real).detach().to(dtype)
    loss, amp, pha = spectral_loss_r1(fk, rl, **kw)
    (g,) = torch.autograd.grad(loss, fk)
    return float(loss.detach()), float(amp.detach()), float(pha.detach()), g.numpy()


def fft_components_r1(x, *, channels="luma", input_scale=1.0, shift=True, dtype=torch.float64):
    """Differentiable counterpart of ``fft_components`` (``...patchFFT_16P.py:293-319``):
    ``[N,C,p,p]`` -> ``(AMP, PHA)`` each ``[N,C',p,p/2+1]``, fftshift-ed over both axes of
    the half spectrum like the reference (``:279``) when ``shift``."""
    t = _prepare(x, channels, input_scale, False, dtype)
    F = torch.fft.rfft2(t)
    if shift:
        F = torch.fft.fftshift(F, dim=(-2, -1))
    return F.abs(), torch.angle(F)


# ---------------------------------------------------------------------------
# independent closed form (NumPy fp64) -- mirrors the CUDA kernels' algebra
# ---------------------------------------------------------------------------

def spectral_grad_analytic(
    fake,
    real,
    *,
    grid=4,
    channels="luma",
    use_phase=True,
    distance="l1",
    patch_reduce="mean",
    log_magnitude=False,
    spectrum="half",
    weight=1.0,
    input_scale=1.0,
):
    """Loss and gradient by the kernels' closed form.  Returns ``(loss, amp, pha, grad)``.

    Per tile: ``z = f + i r`` -> ``Z = fft2(z)``; ``2F(k) = Z(k) + conj Z(-k)``,
    ``2R(k) = -i (Z(k) - conj Z(-k))`` on the half plane ``kc in [0, p/2]``;
    per-bin distance with Hermitian multiplicity ``m(k)`` (1 for the half
    spectrum; for the full spectrum 2 on columns ``1..p/2-1`` and 1 on the two
    self-conjugate columns); spectral gradient
    ``G = gA * F/|F|  (or F/|F|^2 with log)  +  gP * iF/|F|^2``; ``G`` is zero-extended
    to the full plane and ``grad = Re(p^2 * ifft2(G))``; luma weights and
    ``input_scale`` applied on the way out.
    """
    fake = np.asarray(fake, dtype=np.float64)
    real = np.asarray(real, dtype=np.float64)
    n, c, h, w = fake.shape
    p = h // grid
    hp = p // 2 + 1
    lw = np.asarray(LUMA_WEIGHTS).reshape(1, 3, 1, 1)
    luma = channels == "luma" and c == 3
    f = fake * input_scale
    r = real * input_scale
    if luma:
        f = (f * lw).sum(1, keepdims=True)
        r = (r * lw).sum(1, keepdims=True)
    cp = f.shape[1]
    if spectrum == "half":
        mult = np.ones(hp)
        kbins = p * hp
    else:
        mult = np.full(hp, 2.0)
        mult[0] = 1.0
        mult[p // 2] = 1.0
        kbins = p * p
    red = float(grid * grid) if patch_reduce == "sum" else 1.0
    norm = red / (n * cp * grid * grid * kbins)
    if use_phase:
        sa = sp = 0.5 * weight * norm
    else:
        sa, sp = weight * norm, 0.0
    kr = np.arange(p)
    neg_r = (-kr) % p
    amp_sum = 0.0
    pha_sum = 0.0
    gl = np.zeros_like(f)
    for i in range(n):
        for ch in range(cp):
            for gy in range(grid):
                for gx in range(grid):
                    ys, xs = slice(gy * p, gy * p + p), slice(gx * p, gx * p + p)
                    z = f[i, ch, ys, xs] + 1j * r[i, ch, ys, xs]
                    Z = np.fft.fft2(z)
                    Zn = np.conj(Z[neg_r][:, (-np.arange(p)) % p])  # conj Z(-k)
                    F = 0.5 * (Z + Zn)[:, :hp]
                    R = (-0.5j * (Z - Zn))[:, :hp]
                    af, ar = np.abs(F), np.abs(R)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        inv = np.where(af > 0, 1.0 / af, 0.0)
                        if log_magnitude:
                            va, vb = np.log(af), np.log(ar)
                            dA_dF = F * inv * inv  # d log|F| = F/|F|^2
                        else:
                            va, vb = af, ar
                            dA_dF = F * inv
                    d = va - vb
                    if distance == "l1":
                        amp_sum += (np.abs(d) * mult).sum()
                        ga = np.sign(d)
                    else:
                        amp_sum += (d * d * mult).sum()
                        ga = 2.0 * d
                    G = sa * mult * ga * dA_dF
                    if use_phase:
                        pf = np.arctan2(F.imag, F.real)
                        pr = np.arctan2(R.imag, R.real)
                        dpp = pf - pr
                        if distance == "l1":
                            pha_sum += (np.abs(dpp) * mult).sum()
                            gp = np.sign(dpp)
                        else:
                            pha_sum += (dpp * dpp * mult).sum()
                            gp = 2.0 * dpp
                        G = G + sp * mult * gp * (1j * F) * inv * inv
                    Gfull = np.zeros((p, p), complex)
                    Gfull[:, :hp] = G
                    gl[i, ch, ys, xs] = np.real(np.fft.ifft2(Gfull)) * (p * p)
    if luma:
        grad = gl * lw * input_scale
    else:
        grad = gl * input_scale
    amp = amp_sum * norm
    pha = pha_sum * norm if use_phase else 0.0
    loss = weight * 0.5 * (amp + pha) if use_phase else weight * amp
    return loss, amp, pha, grad
