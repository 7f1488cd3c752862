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
    use_halfline: bool = False   # testing / A-B: 64x64 tiles on the half-line engine (two threads per line)

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
        if self.use_halfline:
            f |= _lib.USE_HALFLINE
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


def _fp16_prescale(cfg: "SpectralConfig", shape) -> float:
    """Power-of-two factor folded into an fp16 gradient when the caller does not name the real ``grad_output`` (a
    GradScaler): the 1 / (N g^2 K) normalised gradient of a batch of 32+ images is fp16-SUBNORMAL (round-1 advisor
    finding), an MSE gradient on 8-bit-scale spectra can exceed fp16's maximum.  The factor brings the estimated rms
    of the stored values to ~2^-3 (the estimate only has to be right within 2^+-15); ``backward`` divides it out again
    together with autograd's ``grad_output``."""
    import math

    n, c, h, _ = shape
    p = h // cfg.grid
    k = p * (p if cfg.spectrum == "full" else p // 2 + 1)
    cp = 3 if (cfg.channels == "rgb" and c == 3) else 1
    per_bin = abs(cfg.weight) * (1.0 if not cfg.use_phase else 0.5) / (n * cp * cfg.grid ** 2 * k)
    if cfg.patch_reduce == "sum":
        per_bin *= cfg.grid ** 2
    amp = 0.4 * abs(cfg.input_scale) * p  # typical |F| of a [-1, 1] image
    if cfg.distance == "mse":
        per_bin *= 2.0 * (1.0 if cfg.log_magnitude else amp)
    if cfg.log_magnitude:
        per_bin /= amp
    rms = per_bin * math.sqrt(2.0 * k) * abs(cfg.input_scale) * (0.6 if cp == 1 and c == 3 else 1.0)
    if not (rms > 0.0) or not math.isfinite(rms):
        return 1.0
    e = round(math.log2(0.125 / rms))
    return float(2.0 ** max(-24, min(30, e)))


def _resolve_grad_scale(grad_scale, dtype, dev, cfg=None, shape=None):
    """``grad_scale`` -> (host factor, device scalar tensor or None).  Accepts None, a float, a 0-dim / 1-element
    CUDA tensor, a ``torch.amp.GradScaler`` (its device-side scale is read at kernel time: no sync), or a tuple
    ``(GradScaler | tensor, float)`` whose float is the factor applied to the returned loss before ``backward``."""
    host, dev_t = 1.0, None
    items = grad_scale if isinstance(grad_scale, (tuple, list)) else (grad_scale,)
    for it in items:
        if it is None:
            continue
        if isinstance(it, (int, float)):
            host *= float(it)
        elif isinstance(it, torch.Tensor):
            if it.numel() != 1 or not it.is_cuda:
                raise ValueError("grad_scale tensor must be a 1-element CUDA tensor")
            dev_t = it.detach().reshape(()).to(device=dev, dtype=torch.float32)
        elif hasattr(it, "_scale") and hasattr(it, "is_enabled"):  # torch.amp.GradScaler
            if it.is_enabled():
                if it._scale is None:
                    it._lazy_init_scale_growth_tracker(dev)
                dev_t = it._scale.detach().reshape(()).to(device=dev, dtype=torch.float32)
        else:
            raise TypeError(f"unsupported grad_scale item {type(it).__name__}")
    if grad_scale is None and dtype == torch.float16 and cfg is not None:
        host = _fp16_prescale(cfg, shape)
    return host, dev_t


_DESC_CACHE: dict = {}


def _desc_for(lib, key):
    """Filled descriptor + workspace size per distinct call signature (dtype, grid, flags, shape, strides, scalars): a
    training loop repeats the same few signatures every step, and building the ctypes struct is ~12 us of host time
    against a ~75 us GPU step."""
    hit = _DESC_CACHE.get(key)
    if hit is None:
        dtype, grid, flags, shape, fst, rst, gst, weight, input_scale, gs_host, gs_ptr = key
        desc = _lib.make_desc(dtype, grid, flags, shape, fst, rst, gst, weight, input_scale,
                              grad_scale_host=gs_host, grad_scale_dev=gs_ptr)
        nbytes = lib.tfcfft_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            _lib.check(lib.tfcfft_validate(ctypes.byref(desc)), "tfcfft_validate")
        if len(_DESC_CACHE) >= 512:
            _DESC_CACHE.clear()
        hit = _DESC_CACHE[key] = (desc, ctypes.byref(desc), nbytes)
    return hit


def _launch(fake, real, cfg: SpectralConfig, want_grad: bool, want_per_image: bool, *, gs_host: float = 1.0, gs_dev=None,
            accumulate_into=None, real_quads=None, grad_buffer=None):
    dev = fake.device
    if torch.cuda.current_device() == dev.index:  # the usual case: no device-guard round trip
        return _launch_here(fake, real, cfg, want_grad, want_per_image, gs_host, gs_dev, accumulate_into, real_quads, grad_buffer)
    with torch.cuda.device(dev):
        return _launch_here(fake, real, cfg, want_grad, want_per_image, gs_host, gs_dev, accumulate_into, real_quads, grad_buffer)


def _launch_here(fake, real, cfg, want_grad, want_per_image, gs_host, gs_dev, accumulate_into, real_quads, grad_buffer):
    lib = _lib.load()
    dev = fake.device
    stream_ptr = torch.cuda.current_stream(dev).cuda_stream
    out = torch.empty(8, dtype=torch.float32, device=dev)
    per = torch.empty((fake.shape[0], 2), dtype=torch.float32, device=dev) if want_per_image else None
    flags = cfg.flags()
    if accumulate_into is not None:
        if accumulate_into.shape != fake.shape or accumulate_into.dtype != fake.dtype or accumulate_into.device != dev \
                or not _acceptable(accumulate_into):
            raise ValueError("accumulate_into must match fake in shape / dtype / device and have 16-byte friendly strides")
        grad, flags = accumulate_into, flags | _lib.GRAD_ACCUMULATE
    elif grad_buffer is not None:  # caller-owned destination, overwritten
        if grad_buffer.shape != fake.shape or grad_buffer.dtype != fake.dtype or grad_buffer.device != dev \
                or not _acceptable(grad_buffer):
            raise ValueError("grad_buffer must match fake in shape / dtype / device and have 16-byte friendly strides")
        grad = grad_buffer
    else:
        grad = torch.empty(fake.shape, dtype=fake.dtype, device=dev) if want_grad else None
    key = (_DTYPES[fake.dtype], cfg.grid, flags, tuple(fake.shape), fake.stride(), real.stride(),
           grad.stride() if grad is not None else None, cfg.weight, cfg.input_scale, gs_host,
           gs_dev.data_ptr() if gs_dev is not None else None)
    _, desc_ref, nbytes = _desc_for(lib, key)
    ws = _workspace(dev, stream_ptr, nbytes)
    if real_quads is None:
        rc = lib.tfcfft_loss(
            desc_ref, fake.data_ptr(), real.data_ptr(), out.data_ptr(),
            per.data_ptr() if want_per_image else None, grad.data_ptr() if grad is not None else None,
            ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr),
        )
    else:
        qp = (ctypes.c_void_p * 4)(*[q.data_ptr() for q in real_quads])
        rc = lib.tfcfft_loss_quads(
            desc_ref, fake.data_ptr(), qp, out.data_ptr(),
            per.data_ptr() if want_per_image else None, grad.data_ptr() if grad is not None else None,
            ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr),
        )
    if rc > 0:  # a CUDA error may have left the ticket header dirty
        _WORKSPACES.pop((dev.index, stream_ptr), None)
    _lib.check(rc, "tfcfft_loss")
    return out, per, grad


def _scale_saved_gradient(unit: torch.Tensor, grad_loss: torch.Tensor, in_dtype: torch.dtype) -> torch.Tensor:
    """``unit * grad_loss`` (the saved d loss / d fake times autograd's incoming scalar: loss weight x GradScaler scale) in
    one launch of the library's scaling kernel; the scalar stays on the device.  Used by the auxiliary losses; the FFT
    loss folds the expected scale into its producing launch instead (``_rescale_in_place``)."""
    lib = _lib.load()
    dev = unit.device
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        go = grad_loss.detach().to(device=dev, dtype=torch.float32).contiguous()
        res = torch.empty_like(unit)
        _lib.check(lib.tfcfft_grad_scale(res.data_ptr(), unit.data_ptr(), _DTYPES[unit.dtype], unit.numel(), go.data_ptr(), 1.0,
                                         ctypes.c_void_p(stream_ptr)), "tfcfft_grad_scale")
    return res if res.dtype == in_dtype else res.to(in_dtype)


def _rescale_in_place(buf: torch.Tensor, out: torch.Tensor, grad_loss: torch.Tensor) -> None:
    """``buf *= grad_loss / applied`` on the device, where ``applied`` (``out[4]``) is the scale the producing launch
    already folded in.  Equal scalars -- the expected case -- cost one tiny launch and no pass over ``buf``."""
    lib = _lib.load()
    dev = buf.device
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        go = grad_loss.detach().to(device=dev, dtype=torch.float32).reshape(())
        ws = _workspace(dev, stream_ptr, 256)
        _lib.check(lib.tfcfft_grad_rescale(buf.data_ptr(), _DTYPES[buf.dtype], buf.numel(), go.data_ptr(),
                                           out.data_ptr() + 16, ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream_ptr)),
                   "tfcfft_grad_rescale")


def _common_quads(quads, fake):
    """The loader's four real quadrants as pointer-array arguments: same dtype / shape / strides, else None."""
    q0 = quads[0]
    n, c, h, w = fake.shape
    for q in quads:
        if q.dtype != fake.dtype or q.device != fake.device or tuple(q.shape) != (n, c, h // 2, w // 2) \
                or q.stride() != q0.stride() or not _acceptable(q):
            return None
    return quads


class _SpectralLossFn(torch.autograd.Function):
    """forward: loss and d loss / d fake -- with the expected ``grad_output`` already folded in -- in ONE pass over
    fake / real (3 tensor passes of HBM traffic); backward: a scalar comparison on the device; the tensor is touched
    again (in place, no allocation) only when autograd's ``grad_output`` differs from the expected one."""

    @staticmethod
    def forward(ctx, fake, real, cfg, grad_scale, quads):
        if quads is not None:
            fake_p, _ = _prep(fake.detach(), fake.detach())
            qs = _common_quads([q.detach() for q in quads], fake_p)
            if qs is None:  # mixed layouts: one concatenation copy (the reference's tensors are always alike)
                top = torch.cat((quads[0], quads[1]), dim=3)
                bot = torch.cat((quads[2], quads[3]), dim=3)
                fake_p, real_p = _prep(fake.detach(), torch.cat((top, bot), dim=2).detach())
            else:
                real_p = qs[0]
        else:
            fake_p, real_p = _prep(fake.detach(), real.detach())
            qs = None
        want_grad = ctx.needs_input_grad[0] and not cfg.quantize and fake_p.dtype != torch.uint8
        gs_host, gs_dev = _resolve_grad_scale(grad_scale, fake_p.dtype, fake_p.device, cfg, fake_p.shape) if want_grad else (1.0, None)
        out, _, grad = _launch(fake_p, real_p, cfg, want_grad, False, gs_host=gs_host, gs_dev=gs_dev, real_quads=qs)
        ctx.has_grad = want_grad
        ctx.in_dtype = fake.dtype
        ctx.set_materialize_grads(False)  # no zero-fill launch for the non-differentiable `terms` output in backward
        if want_grad:
            ctx.buf, ctx.out = grad, out
            ctx.regen = (fake_p, real_p, cfg, gs_host, gs_dev, qs)
        terms = out[1:3]
        ctx.mark_non_differentiable(terms)
        return out[0], terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        if not ctx.has_grad or grad_loss is None:
            return None, None, None, None, None
        buf = ctx.buf
        if buf is None:  # a second backward through a retained graph: the buffer was handed out, produce it again
            fake_p, real_p, cfg, gs_host, gs_dev, qs = ctx.regen
            ctx.out, _, buf = _launch(fake_p, real_p, cfg, True, False, gs_host=gs_host, gs_dev=gs_dev, real_quads=qs)
        _rescale_in_place(buf, ctx.out, grad_loss)
        ctx.buf = None  # autograd now owns the only reference: no defensive copy when it accumulates into .grad
        return (buf if buf.dtype == ctx.in_dtype else buf.to(ctx.in_dtype)), None, None, None, None


def spectral_loss(fake, real, *, return_terms: bool = False, config: SpectralConfig | None = None, grad_scale=None,
                  real_quadrants=None, **options):
    """Differentiable frequency-domain loss.  ``options`` are the fields of :class:`SpectralConfig`.

    Returns a 0-dim fp32 CUDA tensor (``weight * 1/2 (amp + pha)``); with ``return_terms`` also a
    detached ``[2]`` tensor ``(amp, pha)`` for logging.  Gradient flows to ``fake`` only -- ``real``
    is data in every reference call site.

    ``grad_scale`` names the ``grad_output`` autograd will hand back (``scaler.scale(loss_G).backward()``,
    ``TFCGAN_multigpu_patchFFT_16P.py:607-610``): a ``GradScaler``, a device scalar, a float, or a tuple of them.  It
    is folded into the gradient in the producing launch, so ``backward`` does not touch the tensor again (3 passes
    of HBM traffic instead of 5); any other ``grad_output`` is still honoured (one in-place rescale).
    ``real_quadrants``: the loader's four separate real quadrant tensors (``grid=2``; ``real`` is then ignored).
    """
    cfg = config if config is not None else SpectralConfig(**options)
    if real_quadrants is not None and cfg.grid != 2:
        raise ValueError("real_quadrants needs grid=2")
    loss, terms = _SpectralLossFn.apply(fake, real, cfg, grad_scale,
                                        tuple(real_quadrants) if real_quadrants is not None else None)
    return (loss, terms) if return_terms else loss


@torch.no_grad()
def spectral_loss_and_grad(fake, real, *, config: SpectralConfig | None = None, grad_scale=None, accumulate_into=None,
                           **options):
    """The fused hot path without autograd: ``(loss, terms, d loss / d fake)`` in one launch.  ``grad_scale`` (float /
    device scalar / GradScaler) multiplies the gradient only; with ``accumulate_into`` the gradient is ADDED to that
    tensor (several loss terms share one buffer without an add pass)."""
    cfg = config if config is not None else SpectralConfig(**options)
    fake_p, real_p = _prep(fake, real)
    gs_host, gs_dev = (1.0, None) if grad_scale is None else _resolve_grad_scale(grad_scale, fake_p.dtype, fake_p.device)
    out, _, grad = _launch(fake_p, real_p, cfg, True, False, gs_host=gs_host, gs_dev=gs_dev, accumulate_into=accumulate_into)
    return out[0], out[1:3], grad


def _multi_launch(fake_p, real_p, cfgs, chunk, gs_host, gs_dev):
    """Several loss configurations on the same tensors, one summed gradient: the first configuration writes the
    gradient and the others add into it (``TFCFFT_GRAD_ACCUMULATE``: no add pass).  ``chunk`` (images; 0 / None = the
    whole batch, the default) optionally walks the batch in L2-sized pieces so that the later configurations find fake /
    real / the gradient piece in L2 -- fewer HBM bytes, but every launch then runs a fraction of a wave; measured on
    B200 at 512 x 512, batch 32 (``profiles/r02_d8_ab.txt``): 57 k / 69 k / 84 k images/s at chunk 8 / 16 / whole batch.
    Chunk losses are means over their own images: weighted by chunk size."""
    import dataclasses

    n = fake_p.shape[0]
    chunk = n if not chunk or chunk <= 0 else min(int(chunk), n)
    grad = torch.empty(fake_p.shape, dtype=fake_p.dtype, device=fake_p.device)
    total, terms, out = None, [], None
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        w = (hi - lo) / n
        for j, cfg in enumerate(cfgs):
            c = dataclasses.replace(cfg, weight=cfg.weight * w)
            out, _, _ = _launch(fake_p[lo:hi], real_p[lo:hi], c, True, False, gs_host=gs_host, gs_dev=gs_dev,
                                accumulate_into=grad[lo:hi] if j else None, grad_buffer=None if j else grad[lo:hi])
            total = out[0] if total is None else total + out[0]
            terms.append(out[1:3] * w)
    k = len(cfgs)
    per_cfg = torch.stack([torch.stack(terms[j::k]).sum(0) for j in range(k)])  # [len(cfgs), 2]: (amp, pha) of each grid
    return total, per_cfg, grad, out


class _MultiGridLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, real, cfgs, chunk, grad_scale):
        fake_p, real_p = _prep(fake.detach(), real.detach())
        gs_host, gs_dev = _resolve_grad_scale(grad_scale, fake_p.dtype, fake_p.device)
        total, per_cfg, grad, out = _multi_launch(fake_p, real_p, cfgs, chunk, gs_host, gs_dev)
        ctx.buf, ctx.out, ctx.in_dtype = grad, out, fake.dtype
        ctx.regen = (fake_p, real_p, cfgs, chunk, gs_host, gs_dev)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(per_cfg)
        return total, per_cfg

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _g):
        if grad_loss is None:
            return None, None, None, None, None
        buf = ctx.buf
        if buf is None:
            _, _, buf, ctx.out = _multi_launch(*ctx.regen)
        _rescale_in_place(buf, ctx.out, grad_loss)
        ctx.buf = None
        return (buf if buf.dtype == ctx.in_dtype else buf.to(ctx.in_dtype)), None, None, None, None


def _grid_configs(grids, options):
    return tuple(SpectralConfig(grid=int(g), **options) for g in grids)


def multi_grid_loss(fake, real, *, grids=(4, 1), chunk: int = 0, return_terms: bool = False, grad_scale=None, **options):
    """Sum of the FFT losses of several grids on the SAME tensors -- BASELINE config 5's "patch-FFT-16 + global-FFT
    combined loss" -- with ONE gradient tensor.  Differentiable w.r.t. ``fake``; ``options`` as for
    :func:`spectral_loss` (each grid's loss carries ``weight``).  ``chunk``: see :func:`_multi_launch` (0 = whole batch)."""
    total, per_cfg = _MultiGridLossFn.apply(fake, real, _grid_configs(grids, options), int(chunk or 0), grad_scale)
    return (total, per_cfg) if return_terms else total


@torch.no_grad()
def multi_grid_loss_and_grad(fake, real, *, grids=(4, 1), chunk: int = 0, grad_scale=None, **options):
    """The fused hot path of :func:`multi_grid_loss` without autograd: ``(loss, per-grid (amp, pha), d loss / d fake)``."""
    fake_p, real_p = _prep(fake, real)
    gs_host, gs_dev = (1.0, None) if grad_scale is None else _resolve_grad_scale(grad_scale, fake_p.dtype, fake_p.device)
    total, per_cfg, grad, _ = _multi_launch(fake_p, real_p, _grid_configs(grids, options), int(chunk or 0), gs_host, gs_dev)
    return total, per_cfg, grad


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
            nbytes = lib.tfcfft_spectra_workspace_bytes(ctypes.byref(desc))
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
            nbytes = lib.tfcfft_spectra_workspace_bytes(ctypes.byref(desc))
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
        out = torch.empty(8, dtype=torch.float32, device=dev)
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


def _launch_temperature(fake, positive, negative, lut, quantize, margin, eps, weight, input_scale, want_grad, accumulate_into=None,
                        positive_is_temps=None):
    lib = _lib.load()
    dev = fake.device
    if positive_is_temps is None:  # heuristic kept for 3-channel images; 1-channel callers say which one they pass
        positive_is_temps = positive.shape[1] == 1 and positive.dtype == torch.float32 and fake.shape[1] != 1
    flags = (_lib.QUANTIZE_U8 if quantize else 0) | (_lib.TEMPS_POSITIVE if positive_is_temps else 0)
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float32, device=dev)
        grad = None
        if accumulate_into is not None:
            if accumulate_into.shape != fake.shape or accumulate_into.dtype != fake.dtype or accumulate_into.device != dev \
                    or not _acceptable(accumulate_into):
                raise ValueError("accumulate_into must match fake in shape / dtype / device and have 16-byte friendly strides")
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


def _prep_temperature(fake, positive, negative, positive_is_temps=None):
    """Returns (fake, positive, negative, positive_is_temps).  ``positive`` holds temperatures (the loader's ``T_B``) when
    the caller says so, when it is 3-D ``[N,H,W]`` (the reference's shape, ``...patchFFT_16P.py:593``), or -- for
    multi-channel images -- when it is a 1-channel fp32 tensor."""
    fake = _prep_img(fake)
    negative = _prep_img(negative, like=fake)
    if positive_is_temps is None:
        positive_is_temps = positive.dim() == 3 or (positive.shape[1] == 1 and fake.shape[1] != 1 and positive.dtype == torch.float32)
    if positive.dim() == 3:
        positive = positive.reshape(positive.shape[0], 1, positive.shape[1], positive.shape[2])  # the reference's reshape (:593)
    if positive_is_temps:
        if positive.shape[1] != 1:
            raise ValueError("temperatures must be [N,H,W] or [N,1,H,W]")
        positive = positive.float()
        positive = positive if _acceptable(positive) else positive.contiguous()
    else:
        positive = _prep_img(positive, like=fake)
    if negative.shape != fake.shape or positive.shape[0] != fake.shape[0] or positive.shape[2:] != fake.shape[2:]:
        raise ValueError("fake, positive and negative must agree in batch and image size")
    return fake, positive, negative, bool(positive_is_temps)


class _TemperatureTripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, positive, negative, lut, quantize, margin, eps, weight, input_scale, positive_is_temps):
        f, p, n, pit = _prep_temperature(fake.detach(), positive.detach(), negative.detach(), positive_is_temps)
        want_grad = ctx.needs_input_grad[0] and not quantize and f.dtype != torch.uint8
        out, grad = _launch_temperature(f, p, n, lut, quantize, margin, eps, weight, input_scale, want_grad, positive_is_temps=pit)
        ctx.has_grad, ctx.in_dtype = want_grad, fake.dtype
        if want_grad:
            ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        if not ctx.has_grad:
            return (None,) * 10
        (unit,) = ctx.saved_tensors
        return (_scale_saved_gradient(unit, grad_loss, ctx.in_dtype),) + (None,) * 9


def temperature_triplet_loss(fake, positive, negative, *, lut=None, quantize: bool = False, margin: float = 1.0, eps: float = 1e-6,
                             weight: float = 1.0, input_scale: float = 255.0, positive_is_temperatures=None):
    """``weight * TripletMarginLoss(margin, p=2)(temps(fake), temps(positive), temps(negative))`` in one fused pass
    (``TFCGAN_multigpu_patchFFT_16P.py:585-595``; ``weight`` is the reference's ``lambda_t``).  ``positive`` is either an
    image batch or the loader's precomputed temperatures ``T_B`` (fp32 ``[N,H,W]`` / ``[N,1,H,W]``).  ``quantize=True``
    is the reference as shipped (uint8 + table, no gradient); the default is the differentiable linear variant on
    ``input_scale * x``.  ``positive_is_temperatures`` states explicitly whether ``positive`` already holds temperatures
    (needed for 1-channel images, where the shapes alone cannot tell)."""
    return _TemperatureTripletFn.apply(fake, positive, negative, None if lut is None else tuple(lut), bool(quantize), float(margin),
                                       float(eps), float(weight), float(input_scale), positive_is_temperatures)


@torch.no_grad()
def temperature_triplet_loss_and_grad(fake, positive, negative, *, lut=None, quantize: bool = False, margin: float = 1.0,
                                      eps: float = 1e-6, weight: float = 1.0, input_scale: float = 255.0, accumulate_into=None,
                                      positive_is_temperatures=None):
    """The fused pass without autograd: ``(out[4], grad_or_None)``."""
    f, p, n, pit = _prep_temperature(fake, positive, negative, positive_is_temperatures)
    return _launch_temperature(f, p, n, lut, quantize, margin, eps, weight, input_scale, not quantize, accumulate_into,
                               positive_is_temps=pit)


def launch_count() -> int:
    """Kernels launched by ``libtfcfft.so`` in this process since the last reset."""
    return int(_lib.load().tfcfft_launch_count())


def reset_launch_count() -> None:
    _lib.load().tfcfft_launch_count_reset()
