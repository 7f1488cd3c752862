"""Seeded synthetic inputs shared by the golden-vector generator and the tests.

``np.random.RandomState`` streams are frozen by NumPy's compatibility policy, so
the fixtures only need to store (kind, seed, shape, dtype), not the pixels.
"""

from __future__ import annotations

import numpy as np


def make_pair(kind: str, seed: int, shape, dtype: str = "float32"):
    """Returns ``(fake, real)`` ndarrays of ``shape`` = (N, C, H, W) in ``dtype``.

    kinds: ``uniform`` U(-1,1) (SURVEY.md §8d timing input); ``tanh`` tanh(N(0,1))
    (generator-like); ``lowpass`` 1/f-filtered noise rescaled to (-1,1) (image-like);
    ``unit`` U(0,1) (no negative wrap in R0).
    """
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(2):
        if kind == "uniform":
            a = rs.uniform(-1.0, 1.0, size=shape)
        elif kind == "unit":
            a = rs.uniform(0.0, 1.0, size=shape)
        elif kind == "tanh":
            a = np.tanh(rs.normal(size=shape))
        elif kind == "lowpass":
            n, c, h, w = shape
            white = rs.normal(size=shape)
            fy = np.fft.fftfreq(h)[:, None]
            fx = np.fft.fftfreq(w)[None, :]
            filt = 1.0 / np.sqrt(fy * fy + fx * fx + (1.0 / max(h, w)) ** 2)
            a = np.real(np.fft.ifft2(np.fft.fft2(white) * filt))
            a = a / np.abs(a).max(axis=(-1, -2), keepdims=True) * 0.98
        else:
            raise ValueError(kind)
        out.append(a.astype(dtype))
    return out[0], out[1]


def make_gray_pairs(seed: int, n: int, side: int):
    """``n`` pairs of uint8 grey images for the MagMSE metric; pair 1 (if present)
    contains a constant image whose log spectrum has ``-inf`` (skipped by the reference)."""
    rs = np.random.RandomState(seed)
    reals = [rs.randint(0, 256, size=(side, side)).astype(np.uint8) for _ in range(n)]
    fakes = [np.clip(r.astype(np.int64) + rs.randint(-40, 41, size=r.shape), 0, 255).astype(np.uint8) for r in reals]
    if n > 1:
        fakes[1] = np.full((side, side), 77, np.uint8)
    return reals, fakes
