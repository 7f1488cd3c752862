"""``nn.Module`` front end of the FFT loss (stateless: no parameters, no buffers, so reference
checkpoints -- ``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:692-695`` -- are unaffected)."""

from __future__ import annotations

import torch
import torch.nn as nn

from .functional import SpectralConfig, patch_triplet_loss, spectral_loss


class SpectralLoss(nn.Module):
    """``SpectralLoss(grid, channels, use_phase, distance, patch_reduce, ...)(fake, real) -> scalar``.

    ``grid=4`` is the 16-patch loss (``calculate_ffts``, ``...patchFFT_16P.py:323-375``), ``grid=2`` the
    4-patch loss (``TFCGAN_multigpu_patchFFT.py:498-511``), ``grid=1`` the global loss
    (``TFCGAN_multigpu_globalFFT.py:494-499``).  ``weight`` folds the call-site factor (``1/100`` at
    ``...patchFFT_16P.py:607``) into the kernel.  After each call ``last_terms`` holds the detached
    ``(amp, pha)`` terms for logging (the reference logs ``loss_FFT.item()`` every step, ``:664``).

    ``grad_scaler`` (the script's ``torch.amp.GradScaler``, ``:518``) and ``loss_multiplier`` (the factor the script
    applies to the returned loss before ``backward``, e.g. ``1/100`` if ``weight`` is left at 1) name the
    ``grad_output`` that ``scaler.scale(loss_G).backward()`` (``:610``) will send back: it is folded into the gradient
    while it is produced, so the backward pass does not touch the tensor again.  Any other ``grad_output`` still
    gives the right gradient (one in-place rescale).
    """

    def __init__(self, grid: int = 4, channels: str = "luma", use_phase: bool = True, distance: str = "l1",
                 patch_reduce: str = "mean", log_magnitude: bool = False, spectrum: str = "half",
                 weight: float = 1.0, input_scale: float = 1.0, quantize: bool = False, grad_scaler=None,
                 loss_multiplier: float = 1.0):
        super().__init__()
        self.grad_scaler, self.loss_multiplier = grad_scaler, float(loss_multiplier)
        self.config = SpectralConfig(grid=grid, channels=channels, use_phase=use_phase, distance=distance,
                                     patch_reduce=patch_reduce, log_magnitude=log_magnitude, spectrum=spectrum,
                                     weight=weight, input_scale=input_scale, quantize=quantize)
        self.config.flags()  # validate early
        self.last_terms = None

    def forward(self, fake: torch.Tensor, real: torch.Tensor) -> torch.Tensor:
        gs = None
        if self.grad_scaler is not None or self.loss_multiplier != 1.0:
            gs = (self.grad_scaler, self.loss_multiplier)
        loss, terms = spectral_loss(fake, real, config=self.config, return_terms=True, grad_scale=gs)
        self.last_terms = terms
        return loss

    def extra_repr(self) -> str:
        c = self.config
        return (f"grid={c.grid}, channels={c.channels!r}, use_phase={c.use_phase}, distance={c.distance!r}, "
                f"patch_reduce={c.patch_reduce!r}, weight={c.weight}, input_scale={c.input_scale}")


class PatchTripletLoss(nn.Module):
    """``PatchTripletLoss(grid, margin)(fake, real[, negatives]) -> scalar``: the generator step's patch triplet term
    (``nn.TripletMarginLoss(margin=1.0, p=2)`` on every patch with a randomly drawn real patch as negative,
    ``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:75,558-583``) for all ``grid x grid`` patches in one fused
    forward + backward pass.  ``negatives`` (one patch index per patch) defaults to the reference's NumPy draw."""

    def __init__(self, grid: int = 4, margin: float = 1.0, eps: float = 1e-6, weight: float = 1.0):
        super().__init__()
        if grid not in (1, 2, 4):
            raise ValueError("grid must be 1, 2 or 4")
        self.grid, self.margin, self.eps, self.weight = grid, margin, eps, weight
        self.last_negatives = None

    def forward(self, fake: torch.Tensor, real: torch.Tensor, negatives=None) -> torch.Tensor:
        if negatives is None:
            from .compat import draw_negatives
            negatives = draw_negatives(self.grid * self.grid)
        self.last_negatives = list(negatives)
        return patch_triplet_loss(fake, real, negatives, grid=self.grid, margin=self.margin, eps=self.eps, weight=self.weight)

    def extra_repr(self) -> str:
        return f"grid={self.grid}, margin={self.margin}, eps={self.eps}, weight={self.weight}"
