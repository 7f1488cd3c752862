"""tfc-gan_b200 -- B200-native frequency-domain (FFT) loss path of TFC-GAN.

Scope: SURVEY.md §8.  ``SpectralLoss`` / ``spectral_loss`` are the clean entry points; ``compat``
keeps the reference's function names and positional signatures.  All arithmetic runs in the
hand-written sm_100a kernels behind the C ABI in ``include/tfcfft.h``.
"""

from . import compat, dist  # noqa: F401
from .functional import (  # noqa: F401
    SpectralConfig,
    launch_count,
    multi_grid_loss,
    multi_grid_loss_and_grad,
    patch_triplet_loss,
    regional_components,
    regional_spectral_loss,
    regional_spectral_loss_and_grad,
    temperature_triplet_loss,
    temperature_triplet_loss_and_grad,
    vectorize_temps,
    patch_triplet_loss_and_grad,
    reset_launch_count,
    spectral_components,
    spectral_loss,
    spectral_loss_and_grad,
    spectral_terms_per_image,
)
from .modules import PatchTripletLoss, SpectralLoss  # noqa: F401

__all__ = [
    "SpectralConfig", "SpectralLoss", "spectral_loss", "spectral_components", "spectral_loss_and_grad", "spectral_terms_per_image",
    "PatchTripletLoss", "patch_triplet_loss", "regional_components", "regional_spectral_loss", "regional_spectral_loss_and_grad", "temperature_triplet_loss", "temperature_triplet_loss_and_grad", "vectorize_temps", "patch_triplet_loss_and_grad", "launch_count", "reset_launch_count", "multi_grid_loss", "multi_grid_loss_and_grad", "compat", "dist",
]
