"""CPU restatement of the generator step's patch triplet loss.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  NumPy fp64.

Follows ``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py``:

* ``:75``       ``triplet_loss = nn.TripletMarginLoss(margin=1.0, p=2)``
* ``:227-253``  ``make_16_patches``: 16 row-major ``H/4`` tiles
* ``:558-583``  anchor = fake patch i, positive = real patch i, negative =
  ``random_patches[np.random.randint(16, size=1).item()]`` (one draw per patch, in patch order, a patch may draw
  itself); ``loss_triplet_patch = 1/16 * sum_i``.  4-patch copies: ``TFCGAN_multigpu_patchFFT.py:474-484``,
  ``TFCGAN_multigpu_globalFFT.py:470-480``.

The arithmetic lives in torch (not vendored; this image: torch 2.11): ``F.triplet_margin_loss`` =
``mean(clamp_min(margin + d(a, p) - d(a, n), 0))`` with ``d = F.pairwise_distance(x1, x2, p=2, eps=1e-6)`` =
``|| x1 - x2 + eps ||_2`` over the LAST dimension -- for ``[N, C, P, P]`` patches one distance per (n, c, patch row).
Patches are equal-sized, so the mean of per-patch means is the mean over all ``N*C*H*g`` rows.

Pinned by ``tests/golden/make_golden_triplet.py``, which executes the reference's own lines with torch on the CPU.
The gradient (w.r.t. fake; the reference back-propagates through this term) is the closed form
``[active] * ((a - p + eps)/d_ap - (a - n + eps)/d_an) / rows``.
"""

from __future__ import annotations

import numpy as np


def draw_negatives(patch_num: int):
    """The reference's sampling (``...patchFFT_16P.py:567-582``), same NumPy calls in the same order."""
    return [np.random.randint(patch_num, size=1).item() for _ in range(patch_num)]


def patch_triplet_loss_and_grad(fake, real, negatives, grid=4, margin=1.0, eps=1e-6, weight=1.0):
    """Returns ``(weight*loss, loss, active_fraction, grad)`` in float64; ``fake`` / ``real`` are ``[N,C,H,H]``."""
    f = np.asarray(fake, dtype=np.float64)
    r = np.asarray(real, dtype=np.float64)
    n, c, h, w = f.shape
    assert h == w and h % grid == 0 and len(negatives) == grid * grid
    p = h // grid
    grad = np.zeros_like(f)
    total, active = 0.0, 0.0
    for i in range(grid * grid):
        py, px = divmod(i, grid)
        ky, kx = divmod(int(negatives[i]), grid)
        a = f[:, :, py * p:(py + 1) * p, px * p:(px + 1) * p]
        pos = r[:, :, py * p:(py + 1) * p, px * p:(px + 1) * p]
        neg = r[:, :, ky * p:(ky + 1) * p, kx * p:(kx + 1) * p]
        dp = a - pos + eps
        dn = a - neg + eps
        dap = np.sqrt((dp * dp).sum(-1))
        dan = np.sqrt((dn * dn).sum(-1))
        hinge = margin + dap - dan
        act = hinge >= 0.0
        total += np.where(act, hinge, 0.0).sum()
        active += act.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            ga = np.where(act & (dap > 0), 1.0 / dap, 0.0)[..., None] * dp
            gn = np.where(act & (dan > 0), 1.0 / dan, 0.0)[..., None] * dn
        grad[:, :, py * p:(py + 1) * p, px * p:(px + 1) * p] = ga - gn
    rows = n * c * h * grid
    loss = total / rows
    return weight * loss, loss, active / rows, grad * (weight / rows)
