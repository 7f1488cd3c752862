"""The only real images in the reference: ``TFC-STN/samples/{100,200,300,450}.png``, four 1280 x 256 RGB strips =
20 visible / thermal face tiles of 256 x 256.  They are read IN PLACE -- faces of real people
are not copied into this repository -- from ``$TFC_SAMPLES_DIR`` or ``/root/reference/TFC-STN/samples``; callers skip
when neither exists (the GPU box has no ``/root/reference``).
"""

from __future__ import annotations

import os

import numpy as np

STRIPS = ("100.png", "200.png", "300.png", "450.png")
# (fake tile, real tile) index pairs inside one strip; (3, 2) is the strip's fake / real thermal pair
PAIRS = ((3, 2), (1, 0), (4, 2), (0, 4), (2, 3))


def samples_dir():
    for d in (os.environ.get("TFC_SAMPLES_DIR"), "/root/reference/TFC-STN/samples"):
        if d and all(os.path.exists(os.path.join(d, s)) for s in STRIPS):
            return d
    return None


def load_pairs(dtype="float32"):
    """(fake, real): two [20, 3, 256, 256] arrays in [-1, 1] -- ToTensor + Normalize(0.5, 0.5) like the reference's
    loaders (``datasets_temp.py:44-48``) -- or None when the strips are not available."""
    d = samples_dir()
    if d is None:
        return None
    from PIL import Image

    fakes, reals = [], []
    for s in STRIPS:
        im = np.asarray(Image.open(os.path.join(d, s)).convert("RGB"), dtype=np.float32) / 255.0  # [256, 1280, 3]
        tiles = [im[:, 256 * k:256 * (k + 1), :].transpose(2, 0, 1) for k in range(5)]
        for i, j in PAIRS:
            fakes.append((tiles[i] - 0.5) / 0.5)
            reals.append((tiles[j] - 0.5) / 0.5)
    return np.ascontiguousarray(np.stack(fakes), dtype=dtype), np.ascontiguousarray(np.stack(reals), dtype=dtype)
