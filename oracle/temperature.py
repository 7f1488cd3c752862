"""CPU restatement of the generator step's temperature triplet loss.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  NumPy.

Follows ``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py``:

* ``:257-258``  ``T = np.linspace(24, 38, num=256)``: temperature of each 8-bit grey level;
* ``:260-268``  ``vectorize_temps``: per sample ``ToPILImage()(x[t]).convert("RGB")`` (uint8 truncation with wrap, as
  in the FFT path) -> ``TempVector_PyTorch.make_pixel_vectors`` (``datasets_temp.py:14-35``): RED channel ->
  ``vs[np.searchsorted(ks, img)]`` = a table gather -> ``torch.Tensor`` (fp32) -> ``[N, 1, H, W]``;
* ``:80, :585-595``  ``criterion_temp = nn.TripletMarginLoss(margin=1.0, p=2)`` on (temps of fake_B, the loader's ``T_B``
  ``datasets_temp.py:65-67``, temps of the colour-jittered real batch), times ``lambda_t = 10`` (``:77``).
  ``ColorJitter`` is the caller's augmentation: its OUTPUT is an input here.

No gradient upstream (the PIL detour detaches).  The differentiable variant (R1 convention) replaces the table by
its linear law on the unquantised value, ``T(x) = lut[0] + (lut[255]-lut[0])/255 * input_scale * x``.
Pinned by ``tests/golden/make_golden_triplet.py`` (which executes the reference's own functions / lines).
"""

from __future__ import annotations

import numpy as np

from .r0_literal import quantize_u8

LUT = np.linspace(24, 38, num=256)


def vectorize_temps_r0(x, lut=LUT):
    """``[N, C, H, W]`` -> fp32 ``[N, 1, H, W]`` temperatures of the red channel (``...patchFFT_16P.py:260-268``)."""
    u8 = quantize_u8(x)[:, 0]
    return np.asarray(lut)[u8].astype(np.float32)[:, None]


def _row_triplet(a, p, n, margin, eps):
    dp = a - p + eps
    dn = a - n + eps
    dap = np.sqrt((dp * dp).sum(-1))
    dan = np.sqrt((dn * dn).sum(-1))
    hinge = margin + dap - dan
    act = hinge >= 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(act & (dap > 0), 1.0 / dap, 0.0)[..., None] * dp - np.where(act & (dan > 0), 1.0 / dan, 0.0)[..., None] * dn
    return np.where(act, hinge, 0.0), act, g


def temperature_triplet(fake, positive, negative, *, lut=LUT, quantize=True, positive_is_temps=False, margin=1.0, eps=1e-6,
                        weight=1.0, input_scale=1.0):
    """Returns ``(weight*loss, loss, active_fraction, grad_or_None)`` (float64).  ``quantize=True`` is the reference as
    shipped (no gradient); ``False`` the differentiable linear variant with ``d/d fake`` (channel 0 only)."""
    lut = np.asarray(lut, dtype=np.float64)
    if quantize:
        tf = vectorize_temps_r0(fake, lut).astype(np.float64)
        tn = vectorize_temps_r0(negative, lut).astype(np.float64)
        tp = np.asarray(positive, np.float64).reshape(tf.shape) if positive_is_temps else vectorize_temps_r0(positive, lut).astype(np.float64)
        slope = 0.0
    else:
        slope = float(np.float32(np.float32(lut[255] - lut[0]) / np.float32(255.0)) * np.float32(input_scale))
        lin = lambda x: lut[0] + slope * np.asarray(x, np.float64)[:, 0:1]
        tf, tn = lin(fake), lin(negative)
        tp = np.asarray(positive, np.float64).reshape(tf.shape) if positive_is_temps else lin(positive)
    hinge, act, g = _row_triplet(tf, tp, tn, margin, eps)
    rows = hinge.size
    loss = hinge.sum() / rows
    grad = None
    if not quantize:
        grad = np.zeros(np.asarray(fake).shape, np.float64)
        grad[:, 0:1] = g * (slope * weight / rows)
    return weight * loss, loss, act.sum() / rows, grad
