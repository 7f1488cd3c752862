"""CPU oracle for the TFC-GAN frequency-domain loss path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``tfc-gan_b200/``)
imports this directory.  The only legitimate callers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / the CPU arm, never as
the thing that is shipped or measured as the product.

Two restatements of the reference (``/root/reference``, read-only):

* ``r0_literal``  -- R0: the path *as shipped* (tensor -> uint8 with wrap ->
  PIL "L" integer luma -> ``np.fft.rfft2`` -> fftshift -> abs / arctan2 ->
  fp32 -> ``nn.L1Loss``), forward only (the reference has no gradient).
  Pure NumPy; does not need PIL / torchvision at run time.
* ``r1_differentiable`` -- R1: the same pipeline with the two
  non-differentiable steps (uint8 quantisation, integer luma) replaced by
  their float counterparts, in ``torch.fft`` on the CPU with autograd.  This is
  what BASELINE.json calls "the reference torch.fft path" and it is the parity
  target of the CUDA kernels (loss rel <= 1e-4, gradient L2-rel <= 1e-3 against
  the fp64 evaluation).

Three further restatements cover the "next" rows of the scope table, each with its own pin:
``triplet`` (patch ``TripletMarginLoss`` block), ``temperature`` (``vectorize_temps`` + temperature triplet) and
``regional`` (``regional_fft_loss`` on the 100 x 256 bands); their golden vectors come from
``tests/golden/make_golden_triplet.py``, which executes the reference's own lines / functions.

Parity pin: the reference ships no tests and no golden vectors,
so the pin is "outputs of the reference itself run here":
``tests/golden/make_golden.py`` extracts the reference's own function bodies
from ``/root/reference`` with ``ast`` (``FFT_Components``, ``fft_components``,
``make_16_patches``, ``calculate_ffts``, ``fft_loss``, ``global_fourier_loss``,
``mse_spec`` and the inline loss blocks), executes them unmodified on seeded
inputs and commits the results under ``tests/golden/``.  R0 must reproduce
those bit-for-bit (spectra) / to fp32 rounding (losses); R1 fed the quantised
luma must reproduce R0 (cross-pin).  ``tests/test_oracle_golden.py`` checks both.
"""

from .r0_literal import (  # noqa: F401
    quantize_u8,
    luma_u8,
    gray_u8,
    components_r0,
    fft_components_r0,
    spectral_loss_r0,
    make_spectra_r0,
    mag_mse_r0,
)
from .r1_differentiable import (  # noqa: F401
    LUMA_WEIGHTS,
    spectral_loss_r1,
    spectral_loss_and_grad_r1,
    fft_components_r1,
    spectral_grad_analytic,
)
