"""CPU restatement of the regional FFT loss on the two 100 x 256 bands.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  NumPy (R0) and torch CPU autograd (R1).

Follows ``regional_fft_loss`` (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_withregion_FFT.py:353-402``): ``hair`` = rows
``0:100``, ``eyes`` = rows ``100:img_width-56`` (= ``100:200``) of the 256-wide image (``:375-380``); per band and sample
``ToPILImage().convert("L")`` -> ``FFT_Components.make_components`` (``np.fft.rfft2`` of the 100 x 256 grey band,
fftshift, abs / arctan2; ``:250-256``) -> fp32 ``[N,1,100,129]`` (``:358-371``); ``criterion_amp`` / ``criterion_phase``
(``nn.L1Loss``) per band, the two bands SUMMED (``:398-399``), ``1/2 (amp + pha)`` (``:400``).  numpy.fft handles the
non-power-of-two length 100 (pocketfft, not vendored).  Pinned by ``tests/golden/make_golden_triplet.py``.
"""

from __future__ import annotations

import numpy as np
import torch

from .r0_literal import _l1_mean_f32, components_r0, gray_u8
from .r1_differentiable import _dist, _prepare

BANDS = ((0, 100), (100, 200))


def regional_loss_r0(fake, real):
    """The reference as shipped: ``(loss, amp, pha)`` as fp32 scalars; forward only."""
    gf, gr = gray_u8(fake), gray_u8(real)
    amp = pha = np.float32(0)
    for lo, hi in BANDS:
        af, pf, ar, pr = [], [], [], []
        for t in range(gf.shape[0]):
            a, p = components_r0(gf[t, lo:hi])
            af.append(a.astype(np.float32)); pf.append(p.astype(np.float32))
            a, p = components_r0(gr[t, lo:hi])
            ar.append(a.astype(np.float32)); pr.append(p.astype(np.float32))
        amp = np.float32(amp + _l1_mean_f32(np.stack(af), np.stack(ar)))
        pha = np.float32(pha + _l1_mean_f32(np.stack(pf), np.stack(pr)))
    return np.float32(0.5) * (amp + pha), amp, pha


def regional_loss_and_grad_r1(fake, real, *, channels="luma", use_phase=True, distance="l1", weight=1.0, input_scale=1.0,
                              quantize=False, dtype=torch.float64):
    """Differentiable restatement (R1 conventions of ``oracle/r1_differentiable.py``): ``(loss, amp, pha, grad)``."""
    f = torch.as_tensor(np.asarray(fake)) if not torch.is_tensor(fake) else fake
    r = torch.as_tensor(np.asarray(real)) if not torch.is_tensor(real) else real
    f = f.detach().clone().to(dtype).requires_grad_(not quantize)
    xf = _prepare(f if not quantize else fake, channels, input_scale, quantize, dtype)
    xr = _prepare(r.to(dtype) if not quantize else real, channels, input_scale, quantize, dtype)
    amp = torch.zeros((), dtype=dtype)
    pha = torch.zeros((), dtype=dtype)
    for lo, hi in BANDS:
        Ff, Fr = torch.fft.rfft2(xf[:, :, lo:hi]), torch.fft.rfft2(xr[:, :, lo:hi])
        amp = amp + _dist(Ff.abs(), Fr.abs(), distance).mean()
        if use_phase:
            pha = pha + _dist(torch.angle(Ff), torch.angle(Fr), distance).mean()
    loss = weight * (0.5 * (amp + pha) if use_phase else amp)
    grad = None
    if not quantize:
        loss.backward()
        grad = f.grad.numpy()
    return float(loss), float(amp), float(pha), grad
