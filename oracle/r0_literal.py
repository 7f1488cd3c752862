"""R0 -- literal CPU restatement of the reference FFT-loss path *as shipped*.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pure NumPy.

What the reference does, per sample and per patch (all citations relative to
``/root/reference``):

1. ``transforms.ToPILImage()(x[t])``
   (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:300``) -- torchvision
   ``to_pil_image``: CHW tensor -> HWC NumPy (in the tensor's own float dtype,
   fp16 in the training scripts) -> ``(npimg * 255).astype(np.uint8)``.  The
   cast truncates toward zero and wraps negatives modulo 256 (SURVEY.md §0
   fact 4; re-probed for fp16/fp32/fp64 in this build).
2. ``.convert("L")`` (same line) -- Pillow's integer ITU-R 601 luma
   ``(19595 R + 38470 G + 7471 B + 0x8000) >> 16``.
3. ``FFT_Components.make_components`` (``...patchFFT_16P.py:276-282``):
   ``np.fft.rfft2`` (uint8 -> float64, unnormalised) -> ``np.fft.fftshift`` over
   both axes of the half spectrum -> ``np.abs`` and ``np.arctan2(imag, real)``.
4. ``torch.Tensor((amp, phase))`` (``:302``) -- cast to fp32; batch assembled
   to ``[N,1,p,p/2+1]`` (``:310-312``).
5. ``nn.L1Loss()`` (mean, fp32) per patch on amplitude and on phase
   (``:83-84,360-371``); patches averaged with ``1/16`` (``:360``), ``0.25``
   (``TFCGAN_multigpu_patchFFT.py:509-510``) or summed
   (``TFCGAN_multigpu_patchFFT_experiment.py:335-336``); global has one tile
   (``TFCGAN_multigpu_globalFFT.py:494-499``); ``loss_FFT = 1/2 (amp + pha)``.

The offline metric ``mse_spec`` (``TFC-GAN-FFT/Devcom_MagMSE.py:91-118``) is
restated in :func:`mag_mse_r0`.
"""

from __future__ import annotations

import numpy as np

try:  # the reference metric uses scipy.fft (Devcom_MagMSE.py:12)
    from scipy.fft import fft2 as _sp_fft2, fftshift as _sp_fftshift
except Exception:  # pragma: no cover - scipy is present in this image
    _sp_fft2, _sp_fftshift = np.fft.fft2, np.fft.fftshift


def _as_numpy(x):
    """torch tensor or ndarray -> ndarray, keeping the float dtype (fp16 stays fp16)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu()
        if str(x.dtype) == "torch.bfloat16":  # NumPy has no bf16; the reference would raise
            x = x.float()
        x = x.numpy()
    return np.asarray(x)


def quantize_u8(x) -> np.ndarray:
    """torchvision ``to_pil_image`` float branch: ``(npimg * 255).astype(np.uint8)``.

    The product is formed in the array's own dtype (fp16 product is rounded to
    fp16 first), then truncated toward zero and wrapped modulo 256
    (-1.0 -> 1, -0.5 -> 129, 0.999 -> 254).  Follows
    ``TFCGAN_multigpu_patchFFT_16P.py:300``.
    """
    x = _as_numpy(x)
    if x.dtype == np.uint8:
        return x
    v = x * 255  # stays in x.dtype, like the reference
    return (np.trunc(v.astype(np.float64)).astype(np.int64) & 0xFF).astype(np.uint8)


def luma_u8(rgb_u8: np.ndarray) -> np.ndarray:
    """Pillow ``convert("L")`` on ``[..., 3, H, W]`` uint8 (``...patchFFT_16P.py:300``)."""
    r = rgb_u8[..., 0, :, :].astype(np.int64)
    g = rgb_u8[..., 1, :, :].astype(np.int64)
    b = rgb_u8[..., 2, :, :].astype(np.int64)
    return ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)


def gray_u8(x) -> np.ndarray:
    """``[N,C,H,W]`` float (C = 3 or 1) -> ``[N,H,W]`` uint8 grey, steps 1+2 above."""
    q = quantize_u8(x)
    if q.shape[1] == 3:
        return luma_u8(q)
    if q.shape[1] == 1:
        return q[:, 0]
    raise ValueError("expected 1 or 3 channels")


def components_r0(img_u8: np.ndarray):
    """``FFT_Components.make_components`` (``...patchFFT_16P.py:276-282``), float64."""
    f_result = np.fft.rfft2(np.asarray(img_u8))
    fshift = np.fft.fftshift(f_result)
    return np.abs(fshift), np.arctan2(fshift.imag, fshift.real)


def make_spectra_r0(img_u8: np.ndarray) -> np.ndarray:
    """``FFT_Components.make_spectra`` (``...patchFFT_16P.py:284-289``); ``-inf`` possible."""
    with np.errstate(divide="ignore"):
        return np.log(np.abs(np.fft.fftshift(np.fft.fft2(np.asarray(img_u8)))))


def fft_components_r0(x):
    """``fft_components`` (``...patchFFT_16P.py:293-319``): ``[N,C,p,p]`` -> two fp32
    ``[N,1,p,p/2+1]`` arrays in the reference's fftshift-ed layout."""
    g = gray_u8(x)
    n, p, _ = g.shape
    amp = np.empty((n, 1, p, p // 2 + 1), np.float32)
    pha = np.empty_like(amp)
    for t in range(n):
        a, ph = components_r0(g[t])
        amp[t, 0] = a.astype(np.float32)
        pha[t, 0] = ph.astype(np.float32)
    return amp, pha


def _l1_mean_f32(a: np.ndarray, b: np.ndarray) -> np.float32:
    """``nn.L1Loss()`` on fp32 tensors: fp32 difference, mean."""
    d = np.abs(a.astype(np.float32) - b.astype(np.float32))
    return np.float32(d.mean(dtype=np.float64))


def spectral_loss_r0(fake, real, grid: int = 4, patch_reduce: str = "mean"):
    """The as-shipped loss.  Returns ``(loss_FFT, loss_Amp, loss_Pha)`` as fp32.

    grid=4 -> ``calculate_ffts`` (``...patchFFT_16P.py:323-375``);
    grid=2, mean -> inline block ``TFCGAN_multigpu_patchFFT.py:498-511``;
    grid=2, sum  -> ``fft_loss`` (``..._experiment.py:317-339``);
    grid=1 -> inline block ``TFCGAN_multigpu_globalFFT.py:494-499``.
    Tiles are row-major (``make_16_patches``, ``...patchFFT_16P.py:227-253``).
    """
    fake, real = _as_numpy(fake), _as_numpy(real)
    n, c, h, w = fake.shape
    if h != w or h % grid:
        raise ValueError("square images divisible by grid only")
    p = h // grid
    la = np.float32(0.0)
    lp = np.float32(0.0)
    for gy in range(grid):
        for gx in range(grid):
            sl = (slice(None), slice(None), slice(gy * p, gy * p + p), slice(gx * p, gx * p + p))
            af, pf = fft_components_r0(fake[sl])
            ar, pr = fft_components_r0(real[sl])
            la = np.float32(la + _l1_mean_f32(af, ar))
            lp = np.float32(lp + _l1_mean_f32(pf, pr))
    if patch_reduce == "mean":
        la = np.float32(la / (grid * grid))
        lp = np.float32(lp / (grid * grid))
    elif patch_reduce != "sum":
        raise ValueError(patch_reduce)
    return np.float32(0.5 * (la + lp)), la, lp


def mag_mse_r0(real_gray, fake_gray, metric: str = "mse"):
    """``mse_spec`` (``TFC-GAN-FFT/Devcom_MagMSE.py:91-118``) on lists of uint8 grey images.

    Per pair: float32 -> ``log(abs(fftshift(fft2(.))))`` on the full spectrum
    (complex64) -> mean squared error (``sklearn.metrics.mean_squared_error`` ==
    plain mean of squared differences); ``metric="mae"`` restates
    ``eval/Eurecom/Eurecom_MagOther.py:90-117``.  A pair whose log spectrum
    contains ``-inf`` makes sklearn raise and is skipped by the reference's bare
    ``except: continue`` (``:108-111``).  Returns ``(values, skipped_indices)``.
    """
    values, skipped = [], []
    for i, (r, f) in enumerate(zip(real_gray, fake_gray)):
        ra = np.array(r, dtype=np.float32)
        fa = np.array(f, dtype=np.float32)
        with np.errstate(divide="ignore"):
            rm = np.log(np.abs(_sp_fftshift(_sp_fft2(ra))))
            fm = np.log(np.abs(_sp_fftshift(_sp_fft2(fa))))
        if not (np.isfinite(rm).all() and np.isfinite(fm).all()):
            skipped.append(i)
            continue
        d = rm.astype(np.float64) - fm.astype(np.float64)
        values.append(float(np.mean(d * d)) if metric == "mse" else float(np.mean(np.abs(d))))
    return values, skipped
