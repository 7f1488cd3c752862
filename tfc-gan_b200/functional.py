"""``torch.autograd.Function`` over the C-ABI library: the differentiable FFT loss.

Replaces, on the GPU and with a gradient, the reference's per-sample CPU detour
``fft_components`` / ``calculate_ffts`` (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:271-375``),
the 4-patch and global variants (``TFCGAN_multigpu_patchFFT.py:498-511``,
``TFCGAN_multigpu_globalFFT.py:494-499``) and the offline ``mse_spec`` metric
(``Devcom_MagMSE.py:91-118``).  PyTorch is used for device memory and streams only; all
arithmetic happens in ``libtfcfft.so``.  There is no CPU or cuFFT fallback: a missing library or a
non-CUDA tensor raises.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.uint8: _lib.U8}


@dataclass(frozen=True)
class SpectralConfig:
    """Options of the loss (defaults = the reference training loss on luma)."""

    grid: int = 4                # 1 global, 2 = 4-patch, 4 = 16-patch
    channels: str = "luma"       # "luma" (reference: .convert("L")) | "rgb" (per channel)
    use_phase: bool = True       # amplitude + phase (reference) or amplitude only
    distance: str = "l1"         # "l1" (nn.L1Loss) | "mse"
    patch_reduce: str = "mean"   # "mean" (calculate_ffts) | "sum" (fft_loss)
    log_magnitude: bool = False  # log|F| (Devcom_MagMSE)
    spectrum: str = "half"       # "half" rfft2 plane | "full" fft2 plane
    weight: float = 1.0
    input_scale: float = 1.0
    quantize: bool = False       # reference-as-shipped uint8 wrap + integer luma; forward only
    force_split: bool = False    # testing: route 64/128 patches through the split kernels
    force_generic: bool = False  # testing: bypass the packed 64x64 fast path
    use_line: bool = False       # testing: 64x64 tiles through the thread-per-line kernel
    use_pair: bool = False       # testing: 64x64 tiles through the packed pair kernel

    def flags(self) -> int:
        if self.channels not in ("luma", "rgb"):
            raise ValueError(f"channels must be 'luma' or 'rgb', got {self.channels!r}")
        if self.distance not in ("l1", "mse"):
            raise ValueError(f"distance must be 'l1' or 'mse', got {self.distance!r}")
        if self.patch_reduce not in ("mean", "sum"):
            raise ValueError(f"patch_reduce must be 'mean' or 'sum', got {self.patch_reduce!r}")
        if self.spectrum not in ("half", "full"):
            raise ValueError(f"spectrum must be 'half' or 'full', got {self.spectrum!r}")
        f = 0
        if self.channels == "rgb":
            f |= _lib.CHANNELS_RGB
        if not self.use_phase:
            f |= _lib.NO_PHASE
        if self.distance == "mse":
            f |= _lib.DIST_MSE
        if self.patch_reduce == "sum":
            f |= _lib.PATCH_SUM
        if self.log_magnitude:
            f |= _lib.LOG_MAGNITUDE
        if self.spectrum == "full":
            f |= _lib.FULL_SPECTRUM
        if self.quantize:
            f |= _lib.QUANTIZE_U8
        if self.force_split:
            f |= _lib.FORCE_SPLIT
        if self.force_generic:
            f |= _lib.FORCE_GENERIC
        if self.use_line:
            f |= _lib.USE_LINE
        if self.use_pair:
            f |= _lib.USE_PAIR
        return f


# one workspace per (device, stream): calls on different streams may overlap
_WORKSPACES: dict = {}


def _workspace(device: torch.device, stream_ptr: int, nbytes: int) -> torch.Tensor:
    key = (device.index, stream_ptr)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _lib.check(_lib.load().tfcfft_workspace_init(ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr)),
                   "tfcfft_workspace_init")
        _WORKSPACES[key] = ws
    return ws


def _acceptable(t: torch.Tensor) -> bool:
    st = t.stride()
    return st[3] == 1 and all(s % 4 == 0 and s >= 0 for s in st[:3]) and t.data_ptr() % (4 * t.element_size()) == 0


def _prep(fake: torch.Tensor, real: torch.Tensor):
    if fake.dim() != 4 or real.dim() != 4 or fake.shape != real.shape:
        raise ValueError(f"fake and real must be 4-D NCHW tensors of the same shape, got {tuple(fake.shape)} / {tuple(real.shape)}")
    if not (fake.is_cuda and real.is_cuda):
        raise RuntimeError("tfcfft runs on CUDA tensors only (no CPU fallback)")
    if fake.device != real.device:
        raise RuntimeError("fake and real must live on the same device")
    if fake.shape[0] == 0:
        raise ValueError("empty batch")
    if fake.dtype not in _DTYPES:
        fake = fake.float()
    if real.dtype != fake.dtype:
        if fake.dtype == torch.uint8 or real.dtype == torch.uint8:
            raise ValueError("uint8 inputs must both be uint8")
        fake, real = fake.float(), real.float()
    # views with 16-byte friendly strides (e.g. B[:, :, 0:64, 64:128]) pass straight through
    if not _acceptable(fake):
        fake = fake.contiguous()
    if not _acceptable(real):
        real = real.contiguous()
    return fake, real


def _launch(fake, real, cfg: SpectralConfig, want_grad: bool, want_per_image: bool):
    lib = _lib.load()
    dev = fake.device
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float32, device=dev)
        per = torch.empty((fake.shape[0], 2), dtype=torch.float32, device=dev) if want_per_image else None
        grad = torch.empty(fake.shape, dtype=fake.dtype, device=dev) if want_grad else None
        desc = _lib.make_desc(
            _DTYPES[fake.dtype], cfg.grid, cfg.flags(), fake.shape, fake.stride(), real.stride(),
            grad.stride() if want_grad else None, cfg.weight, cfg.input_scale,
        )
        nbytes = lib.tfcfft_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            _lib.check(lib.tfcfft_validate(ctypes.byref(desc)), "tfcfft_validate")
        ws = _workspace(dev, stream_ptr, nbytes)
        rc = lib.tfcfft_loss(
            ctypes.byref(desc), fake.data_ptr(), real.data_ptr(), out.data_ptr(),
            per.data_ptr() if want_per_image else None, grad.data_ptr() if want_grad else None,
            ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr),
        )
        if rc > 0:  # a CUDA error may have left the ticket header dirty
            _WORKSPACES.pop((dev.index, stream_ptr), None)
        _lib.check(rc, "tfcfft_loss")
    return out, per, grad


def _scale_saved_gradient(unit: torch.Tensor, grad_loss: torch.Tensor, in_dtype: torch.dtype) -> torch.Tensor:
    """``unit * grad_loss`` (the saved d loss / d fake times autograd's incoming scalar: loss weight x GradScaler scale) in
    one launch of the library's scaling kernel; the scalar stays on the device."""
    lib = _lib.load()
    dev = unit.device
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        go = grad_loss.detach().to(device=dev, dtype=torch.float32).contiguous()
        res = torch.empty_like(unit)
        _lib.check(lib.tfcfft_grad_scale(res.data_ptr(), unit.data_ptr(), _DTYPES[unit.dtype], unit.numel(), go.data_ptr(), 1.0,
                                         ctypes.c_void_p(stream_ptr)), "tfcfft_grad_scale")
    return res if res.dtype == in_dtype else res.to(in_dtype)


class _SpectralLossFn(torch.autograd.Function):
    """forward: loss and the unit gradient in ONE pass over fake / real (3 tensor passes of HBM
    traffic); backward: one scaling launch by ``grad_output`` (weight x GradScaler scale)."""

    @staticmethod
    def forward(ctx, fake, real, cfg):
        fake_p, real_p = _prep(fake.detach(), real.detach())
        want_grad = ctx.needs_input_grad[0] and not cfg.quantize and fake_p.dtype != torch.uint8
        out, _, grad = _launch(fake_p, real_p, cfg, want_grad, False)
        ctx.has_grad = want_grad
        ctx.in_dtype = fake.dtype
        if want_grad:
            ctx.save_for_backward(grad)
        terms = out[1:3]
        ctx.mark_non_differentiable(terms)
        return out[0], terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        if not ctx.has_grad:
            return None, None, None
        (unit,) = ctx.saved_tensors
        return _scale_saved_gradient(unit, grad_loss, ctx.in_dtype), None, None


def spectral_loss(fake, real, *, return_terms: bool = False, config: SpectralConfig | None = None, **options):
    """Differentiable frequency-domain loss.  ``options`` are the fields of :class:`SpectralConfig`.

    Returns a 0-dim fp32 CUDA tensor (``weight * 1/2 (amp + pha)``); with ``return_terms`` also a
    detached ``[2]`` tensor ``(amp, pha)`` for logging.  Gradient flows to ``fake`` only -- ``real``
    is data in every reference call site.
    """
    cfg = config if config is not None else SpectralConfig(**options)
    loss, terms = _SpectralLossFn.apply(fake, real, cfg)
    return (loss, terms) if return_terms else loss


@torch.no_grad()
def spectral_loss_and_grad(fake, real, *, config: SpectralConfig | None = None, **options):
    """The fused hot path without autograd: ``(loss, terms, d loss / d fake)`` in one launch."""
    cfg = config if config is not None else SpectralConfig(**options)
    fake_p, real_p = _prep(fake, real)
    out, _, grad = _launch(fake_p, real_p, cfg, True, False)
    return out[0], out[1:3], grad


@torch.no_grad()
def spectral_terms_per_image(fake, real, *, config: SpectralConfig | None = None, **options):
    """Forward only: ``[N, 2]`` per-image ``(amp, pha)`` terms (their mean over N is the batch term)."""
    cfg = config if config is not None else SpectralConfig(**options)
    fake_p, real_p = _prep(fake, real)
    _, per, _ = _launch(fake_p, real_p, cfg, False, True)
    return per


def _spectra_desc(x, cfg: SpectralConfig, grad=None):
    if cfg.grid != 1:
        raise ValueError("spectra are materialised per tensor: pass the patch itself (grid must be 1)")
    return _lib.make_desc(_DTYPES[x.dtype], 1, cfg.flags(), x.shape, x.stride(), x.stride(),
                          grad.stride() if grad is not None else None, 1.0, cfg.input_scale)


def _spectra_shape(x, cfg: SpectralConfig):
    n, c, p, _ = x.shape
    cp = 3 if (cfg.channels == "rgb" and c == 3) else 1
    return (n, cp, p, p if cfg.spectrum == "full" else p // 2 + 1)


def _prep1(x):
    if x.dim() != 4:
        raise ValueError("expected a 4-D NCHW tensor")
    if not x.is_cuda:
        raise RuntimeError("tfcfft runs on CUDA tensors only (no CPU fallback)")
    if x.dtype not in _DTYPES:
        x = x.float()
    return x if _acceptable(x) else x.contiguous()


class _SpectraFn(torch.autograd.Function):
    """amp / phase of rfft2 (or fft2) of every image: differentiable ``fft_components``."""

    @staticmethod
    def forward(ctx, x, cfg, shift):
        xp = _prep1(x.detach())
        lib = _lib.load()
        dev = xp.device
        with torch.cuda.device(dev):
            stream_ptr = torch.cuda.current_stream(dev).cuda_stream
            shape = _spectra_shape(xp, cfg)
            amp = torch.empty(shape, dtype=torch.float32, device=dev)
            pha = torch.empty(shape, dtype=torch.float32, device=dev)
            desc = _spectra_desc(xp, cfg)
            nbytes = lib.tfcfft_workspace_bytes(ctypes.byref(desc))
            if nbytes == 0:
                _lib.check(lib.tfcfft_validate(ctypes.byref(desc)), "tfcfft_validate")
            ws = _workspace(dev, stream_ptr, nbytes)
            _lib.check(lib.tfcfft_spectra(ctypes.byref(desc), xp.data_ptr(), None, amp.data_ptr(), pha.data_ptr(), None,
                                          None, int(shift), ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr)),
                       "tfcfft_spectra")
        ctx.cfg, ctx.shift, ctx.in_dtype = cfg, shift, x.dtype
        ctx.differentiable = not cfg.quantize and xp.dtype != torch.uint8
        if ctx.differentiable:
            ctx.save_for_backward(xp)
        return amp, pha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_amp, g_pha):
        if not ctx.differentiable or not ctx.needs_input_grad[0]:
            return None, None, None
        (xp,) = ctx.saved_tensors
        lib = _lib.load()
        dev = xp.device
        with torch.cuda.device(dev):
            stream_ptr = torch.cuda.current_stream(dev).cuda_stream
            g_amp = g_amp.contiguous().float()
            g_pha = g_pha.contiguous().float()
            grad = torch.empty(xp.shape, dtype=xp.dtype, device=dev)
            desc = _spectra_desc(xp, ctx.cfg, grad)
            nbytes = lib.tfcfft_workspace_bytes(ctypes.byref(desc))
            ws = _workspace(dev, stream_ptr, nbytes)
            _lib.check(lib.tfcfft_spectra_bwd(ctypes.byref(desc), xp.data_ptr(), g_amp.data_ptr(), g_pha.data_ptr(),
                                              grad.data_ptr(), int(ctx.shift), ws.data_ptr(), ws.numel(),
                                              ctypes.c_void_p(stream_ptr)),
                       "tfcfft_spectra_bwd")
        return (grad if grad.dtype == ctx.in_dtype else grad.to(ctx.in_dtype)), None, None


def spectral_components(x, *, channels: str = "luma", input_scale: float = 1.0, spectrum: str = "half",
                        log_magnitude: bool = False, quantize: bool = False, fftshift: bool = True):
    """``(AMP, PHA)`` of every image of ``x`` ``[N,C,P,P]`` -> two fp32 ``[N,C',P,P/2+1]`` tensors (``P`` wide for
    ``spectrum="full"``), optionally fftshift-ed over both axes like the reference's ``fft_components``
    (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:271-319``).  Differentiable w.r.t. ``x``."""
    cfg = SpectralConfig(grid=1, channels=channels, input_scale=input_scale, spectrum=spectrum,
                         log_magnitude=log_magnitude, quantize=quantize)
    return _SpectraFn.apply(x, cfg, bool(fftshift))


# ---- regional FFT loss ------------------------------------------------------------------------
def _launch_regional(fake, real, cfg: SpectralConfig, want_grad: bool):
    lib = _lib.load()
    dev = fake.device
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float32, device=dev)
        # rows 200..255 belong to no band: their gradient is zero
        grad = torch.zeros(fake.shape, dtype=fake.dtype, device=dev) if want_grad else None
        desc = _lib.make_desc(_DTYPES[fake.dtype], 1, cfg.flags(), fake.shape, fake.stride(), real.stride(),
                              grad.stride() if want_grad else None, cfg.weight, cfg.input_scale)
        nbytes = lib.tfcfft_regional_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            raise RuntimeError("tfcfft_regional_loss: unsupported configuration (needs [N, 1|3, 256, 256]; options: channels, "
                               "use_phase, distance, quantize)")
        ws = _workspace(dev, stream_ptr, nbytes)
        rc = lib.tfcfft_regional_loss(ctypes.byref(desc), fake.data_ptr(), real.data_ptr(), out.data_ptr(), None,
                                      grad.data_ptr() if want_grad else None, ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr))
        if rc > 0:
            _WORKSPACES.pop((dev.index, stream_ptr), None)
        _lib.check(rc, "tfcfft_regional_loss")
    return out, grad


class _RegionalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, real, cfg):
        fake_p, real_p = _prep(fake.detach(), real.detach())
        want_grad = ctx.needs_input_grad[0] and not cfg.quantize and fake_p.dtype != torch.uint8
        out, grad = _launch_regional(fake_p, real_p, cfg, want_grad)
        ctx.has_grad, ctx.in_dtype = want_grad, fake.dtype
        if want_grad:
            ctx.save_for_backward(grad)
        terms = out[1:3]
        ctx.mark_non_differentiable(terms)
        return out[0], terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        if not ctx.has_grad:
            return None, None, None
        (unit,) = ctx.saved_tensors
        return _scale_saved_gradient(unit, grad_loss, ctx.in_dtype), None, None


def regional_spectral_loss(fake, real, *, return_terms: bool = False, channels: str = "luma", use_phase: bool = True,
                           distance: str = "l1", weight: float = 1.0, input_scale: float = 1.0, quantize: bool = False):
    """The reference's ``regional_fft_loss`` (``TFCGAN_multigpu_patchFFT_withregion_FFT.py:353-402``) on the GPU, with a
    gradient: FFT amplitude + phase L1 loss on the two 100 x 256 bands (rows 0..99 "hair", 100..199 "eyes"), the two
    bands summed.  ``fake`` / ``real``: ``[N, 1|3, 256, 256]``."""
    cfg = SpectralConfig(grid=1, channels=channels, use_phase=use_phase, distance=distance, weight=weight, input_scale=input_scale,
                         quantize=quantize)
    loss, terms = _RegionalLossFn.apply(fake, real, cfg)
    return (loss, terms) if return_terms else loss


@torch.no_grad()
def regional_spectral_loss_and_grad(fake, real, **options):
    """The fused pass without autograd: ``(loss, terms, d loss / d fake)``."""
    cfg = SpectralConfig(grid=1, **options)
    fake_p, real_p = _prep(fake, real)
    out, grad = _launch_regional(fake_p, real_p, cfg, True)
    return out[0], out[1:3], grad


class _RegionalSpectraFn(torch.autograd.Function):
    """amp / phase of rfft2 of the two 100 x 256 bands of every image: differentiable ``reg_fft``."""

    @staticmethod
    def forward(ctx, x, cfg, shift):
        xp = _prep1(x.detach())
        lib = _lib.load()
        dev = xp.device
        with torch.cuda.device(dev):
            stream_ptr = torch.cuda.current_stream(dev).cuda_stream
            cp = 3 if (cfg.channels == "rgb" and xp.shape[1] == 3) else 1
            shape = (xp.shape[0], cp, 2, 100, 129)
            amp = torch.empty(shape, dtype=torch.float32, device=dev)
            pha = torch.empty(shape, dtype=torch.float32, device=dev)
            desc = _lib.make_desc(_DTYPES[xp.dtype], 1, cfg.flags(), xp.shape, xp.stride(), xp.stride(), None, 1.0, cfg.input_scale)
            nbytes = lib.tfcfft_regional_workspace_bytes(ctypes.byref(desc))
            if nbytes == 0:
                raise RuntimeError("tfcfft_regional_spectra: needs [N, 1|3, 256, 256]")
            ws = _workspace(dev, stream_ptr, nbytes)
            _lib.check(lib.tfcfft_regional_spectra(ctypes.byref(desc), xp.data_ptr(), amp.data_ptr(), pha.data_ptr(), int(shift),
                                                   ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr)), "tfcfft_regional_spectra")
        ctx.cfg, ctx.shift, ctx.in_dtype = cfg, shift, x.dtype
        ctx.differentiable = not cfg.quantize and xp.dtype != torch.uint8
        if ctx.differentiable:
            ctx.save_for_backward(xp)
        return amp, pha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_amp, g_pha):
        if not ctx.differentiable or not ctx.needs_input_grad[0]:
            return None, None, None
        (xp,) = ctx.saved_tensors
        lib = _lib.load()
        dev = xp.device
        with torch.cuda.device(dev):
            stream_ptr = torch.cuda.current_stream(dev).cuda_stream
            g_amp, g_pha = g_amp.contiguous().float(), g_pha.contiguous().float()
            grad = torch.zeros(xp.shape, dtype=xp.dtype, device=dev)  # rows 200..255 belong to no band
            cfg = ctx.cfg
            desc = _lib.make_desc(_DTYPES[xp.dtype], 1, cfg.flags(), xp.shape, xp.stride(), xp.stride(), grad.stride(), 1.0,
                                  cfg.input_scale)
            ws = _workspace(dev, stream_ptr, lib.tfcfft_regional_workspace_bytes(ctypes.byref(desc)))
            _lib.check(lib.tfcfft_regional_spectra_bwd(ctypes.byref(desc), xp.data_ptr(), g_amp.data_ptr(), g_pha.data_ptr(),
                                                       grad.data_ptr(), int(ctx.shift), ws.data_ptr(), ws.numel(),
                                                       ctypes.c_void_p(stream_ptr)), "tfcfft_regional_spectra_bwd")
        return (grad if grad.dtype == ctx.in_dtype else grad.to(ctx.in_dtype)), None, None


def regional_components(x, *, channels: str = "luma", input_scale: float = 1.0, quantize: bool = False, fftshift: bool = True):
    """``(AMP, PHA)`` of the two 100 x 256 bands of ``x`` ``[N, 1|3, 256, 256]``: fp32 ``[N, C', 2, 100, 129]`` (axis 2: hair
    rows 0..99, eyes rows 100..199), optionally fftshift-ed like the reference's ``reg_fft``
    (``TFCGAN_multigpu_patchFFT_withregion_FFT.py:358-371``).  Differentiable w.r.t. ``x``."""
    cfg = SpectralConfig(grid=1, channels=channels, input_scale=input_scale, quantize=quantize)
    return _RegionalSpectraFn.apply(x, cfg, bool(fftshift))


# ---- patch triplet loss ----------------------------------------------------------------------
def _check_negatives(negatives, grid: int):
    neg = [int(k) for k in negatives]
    if len(neg) != grid * grid or any(k < 0 or k >= grid * grid for k in neg):
        raise ValueError(f"negatives must hold {grid * grid} patch indices in [0, {grid * grid}), got {neg}")
    return neg


def _launch_triplet(fake, real, grid, negatives, margin, eps, weight, want_grad, accumulate_into=None):
    lib = _lib.load()
    dev = fake.device
    neg = _check_negatives(negatives, grid)
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float32, device=dev)
        flags = 0
        if accumulate_into is not None:
            if accumulate_into.shape != fake.shape or accumulate_into.dtype != fake.dtype or not _acceptable(accumulate_into):
                raise ValueError("accumulate_into must match fake in shape / dtype and have 16-byte friendly strides")
            grad, flags = accumulate_into, _lib.GRAD_ACCUMULATE
        else:
            grad = torch.empty(fake.shape, dtype=fake.dtype, device=dev) if want_grad else None
        desc = _lib.make_desc(_DTYPES[fake.dtype], grid, flags, fake.shape, fake.stride(), real.stride(),
                              grad.stride() if grad is not None else None, weight, 1.0)
        ws = _workspace(dev, stream_ptr, lib.tfcfft_triplet_workspace_bytes())
        rc = lib.tfcfft_patch_triplet(ctypes.byref(desc), fake.data_ptr(), real.data_ptr(), (ctypes.c_int32 * len(neg))(*neg),
                                      float(margin), float(eps), out.data_ptr(), grad.data_ptr() if grad is not None else None,
                                      ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr))
        if rc > 0:
            _WORKSPACES.pop((dev.index, stream_ptr), None)
        _lib.check(rc, "tfcfft_patch_triplet")
    return out, grad


class _PatchTripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, real, grid, negatives, margin, eps, weight):
        fake_p, real_p = _prep(fake.detach(), real.detach())
        want_grad = ctx.needs_input_grad[0] and fake_p.dtype != torch.uint8
        out, grad = _launch_triplet(fake_p, real_p, grid, negatives, margin, eps, weight, want_grad)
        ctx.has_grad, ctx.in_dtype = want_grad, fake.dtype
        if want_grad:
            ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        if not ctx.has_grad:
            return (None,) * 7
        (unit,) = ctx.saved_tensors
        return (_scale_saved_gradient(unit, grad_loss, ctx.in_dtype),) + (None,) * 6


def patch_triplet_loss(fake, real, negatives, *, grid: int = 4, margin: float = 1.0, eps: float = 1e-6, weight: float = 1.0):
    """``1/g^2 * sum_i TripletMarginLoss(margin, p=2)(fake_patch_i, real_patch_i, real_patch_{negatives[i]})`` for the
    row-major ``g x g`` patches of ``fake`` / ``real`` ``[N,C,H,H]`` (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:
    558-583``), forward and backward in one streaming pass.  0-dim fp32 CUDA tensor; gradient flows to ``fake``."""
    return _PatchTripletFn.apply(fake, real, int(grid), tuple(int(k) for k in negatives), float(margin), float(eps), float(weight))


@torch.no_grad()
def patch_triplet_loss_and_grad(fake, real, negatives, *, grid: int = 4, margin: float = 1.0, eps: float = 1e-6,
                                weight: float = 1.0, accumulate_into=None):
    """The fused pass without autograd: ``(out[4], grad)``; ``out`` = (weight*loss, loss, active-row fraction, 0).
    With ``accumulate_into`` the gradient is ADDED to that tensor (e.g. the FFT loss gradient) in the same pass."""
    fake_p, real_p = _prep(fake, real)
    return _launch_triplet(fake_p, real_p, int(grid), negatives, margin, eps, weight, True, accumulate_into)


# ---- temperature triplet loss ----------------------------------------------------------------
#: the reference's table: ``T = np.linspace(24, 38, num=256)`` (``TFCGAN_multigpu_patchFFT_16P.py:257``), as fp32
DEFAULT_TEMPERATURE_LUT = tuple(float(v) for v in __import__("numpy").linspace(24, 38, num=256).astype("float32"))


def _lut_array(lut):
    vals = DEFAULT_TEMPERATURE_LUT if lut is None else tuple(float(v) for v in lut)
    if len(vals) != 256:
        raise ValueError("the temperature table must have 256 entries")
    return (ctypes.c_float * 256)(*vals)


def _prep_img(x, like=None):
    if x.dim() != 4 or x.shape[2] != x.shape[3]:
        raise ValueError(f"expected a square 4-D NCHW tensor, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("tfcfft runs on CUDA tensors only (no CPU fallback)")
    if like is not None and x.dtype != like.dtype:
        x = x.to(like.dtype)
    if x.dtype not in _DTYPES:
        x = x.float()
    return x if _acceptable(x) else x.contiguous()


@torch.no_grad()
def vectorize_temps(x, lut=None):
    """``[N,C,H,H]`` -> fp32 ``[N,1,H,H]``: ``lut[uint8(red channel)]`` with ToPILImage's uint8 rule -- the reference's
    ``vectorize_temps`` (``TFCGAN_multigpu_patchFFT_16P.py:260-268``) without the per-sample CPU round trip."""
    lib = _lib.load()
    xp = _prep_img(x)
    dev = xp.device
    with torch.cuda.device(dev):
        out = torch.empty((xp.shape[0], 1, xp.shape[2], xp.shape[3]), dtype=torch.float32, device=dev)
        desc = _lib.make_desc(_DTYPES[xp.dtype], 1, 0, xp.shape, xp.stride(), xp.stride(), None, 1.0, 1.0)
        _lib.check(lib.tfcfft_vectorize_temps(ctypes.byref(desc), xp.data_ptr(), _lut_array(lut), out.data_ptr(),
                                              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "tfcfft_vectorize_temps")
    return out


def _launch_temperature(fake, positive, negative, lut, quantize, margin, eps, weight, input_scale, want_grad, accumulate_into=None):
    lib = _lib.load()
    dev = fake.device
    positive_is_temps = positive.shape[1] == 1 and positive.dtype == torch.float32 and fake.shape[1] != 1
    flags = (_lib.QUANTIZE_U8 if quantize else 0) | (_lib.TEMPS_POSITIVE if positive_is_temps else 0)
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float32, device=dev)
        grad = None
        if accumulate_into is not None:
            grad, flags = accumulate_into, flags | _lib.GRAD_ACCUMULATE
        elif want_grad:
            grad = torch.zeros(fake.shape, dtype=fake.dtype, device=dev)  # only channel 0 receives a gradient
        desc = _lib.make_desc(_DTYPES[fake.dtype], 1, flags, fake.shape, fake.stride(), positive.stride(),
                              grad.stride() if grad is not None else None, weight, input_scale)
        ws = _workspace(dev, stream_ptr, lib.tfcfft_triplet_workspace_bytes())
        rc = lib.tfcfft_temperature_triplet(ctypes.byref(desc), fake.data_ptr(), positive.data_ptr(), negative.data_ptr(),
                                            (ctypes.c_int64 * 4)(*negative.stride()), _lut_array(lut), float(margin), float(eps),
                                            out.data_ptr(), grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(),
                                            ctypes.c_void_p(stream_ptr))
        if rc > 0:
            _WORKSPACES.pop((dev.index, stream_ptr), None)
        _lib.check(rc, "tfcfft_temperature_triplet")
    return out, grad


def _prep_temperature(fake, positive, negative):
    fake = _prep_img(fake)
    negative = _prep_img(negative, like=fake)
    if positive.dim() == 3:
        positive = positive.reshape(positive.shape[0], 1, positive.shape[1], positive.shape[2])  # the reference's reshape (:593)
    if positive.shape[1] == 1 and fake.shape[1] != 1:
        positive = positive.float()
        positive = positive if _acceptable(positive) else positive.contiguous()
    else:
        positive = _prep_img(positive, like=fake)
    if negative.shape != fake.shape or positive.shape[0] != fake.shape[0] or positive.shape[2:] != fake.shape[2:]:
        raise ValueError("fake, positive and negative must agree in batch and image size")
    return fake, positive, negative


class _TemperatureTripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, positive, negative, lut, quantize, margin, eps, weight, input_scale):
        f, p, n = _prep_temperature(fake.detach(), positive.detach(), negative.detach())
        want_grad = ctx.needs_input_grad[0] and not quantize and f.dtype != torch.uint8
        out, grad = _launch_temperature(f, p, n, lut, quantize, margin, eps, weight, input_scale, want_grad)
        ctx.has_grad, ctx.in_dtype = want_grad, fake.dtype
        if want_grad:
            ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        if not ctx.has_grad:
            return (None,) * 9
        (unit,) = ctx.saved_tensors
        return (_scale_saved_gradient(unit, grad_loss, ctx.in_dtype),) + (None,) * 8


def temperature_triplet_loss(fake, positive, negative, *, lut=None, quantize: bool = False, margin: float = 1.0, eps: float = 1e-6,
                             weight: float = 1.0, input_scale: float = 255.0):
    """``weight * TripletMarginLoss(margin, p=2)(temps(fake), temps(positive), temps(negative))`` in one fused pass
    (``TFCGAN_multigpu_patchFFT_16P.py:585-595``; ``weight`` is the reference's ``lambda_t``).  ``positive`` is either an
    image batch or the loader's precomputed temperatures ``T_B`` (fp32 ``[N,H,W]`` / ``[N,1,H,W]``).  ``quantize=True``
    is the reference as shipped (uint8 + table, no gradient); the default is the differentiable linear variant on
    ``input_scale * x``."""
    return _TemperatureTripletFn.apply(fake, positive, negative, None if lut is None else tuple(lut), bool(quantize), float(margin),
                                       float(eps), float(weight), float(input_scale))


@torch.no_grad()
def temperature_triplet_loss_and_grad(fake, positive, negative, *, lut=None, quantize: bool = False, margin: float = 1.0,
                                      eps: float = 1e-6, weight: float = 1.0, input_scale: float = 255.0, accumulate_into=None):
    """The fused pass without autograd: ``(out[4], grad_or_None)``."""
    f, p, n = _prep_temperature(fake, positive, negative)
    return _launch_temperature(f, p, n, lut, quantize, margin, eps, weight, input_scale, not quantize, accumulate_into)


def launch_count() -> int:
    """Kernels launched by ``libtfcfft.so`` in this process since the last reset."""
    return int(_lib.load().tfcfft_launch_count())


def reset_launch_count() -> None:
    _lib.load().tfcfft_launch_count_reset()
