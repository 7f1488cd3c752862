"""Shared helpers for the tests."""

from __future__ import annotations

import ctypes
import os

import numpy as np

import tfc_gan_b200 as tfc

L = tfc._lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_PATH = os.path.join(ROOT, "tfc-gan_b200", "libtfcfft_emu.so")

NP_DTYPES = {"float32": L.F32, "float16": L.F16, "uint8": L.U8}


def flags_of(channels="luma", use_phase=True, distance="l1", patch_reduce="mean", log_magnitude=False,
             spectrum="half", quantize=False, force_split=False, force_generic=False, use_line=False, use_pair=False):
    return tfc.SpectralConfig(channels=channels, use_phase=use_phase, distance=distance, patch_reduce=patch_reduce,
                              log_magnitude=log_magnitude, spectrum=spectrum, quantize=quantize,
                              force_split=force_split, force_generic=force_generic, use_line=use_line, use_pair=use_pair).flags()


_EMU = None


def emu_lib():
    global _EMU
    if _EMU is None:
        _EMU = ctypes.CDLL(EMU_PATH)
        _EMU.tfcfft_emulate.restype = ctypes.c_int
        _EMU.tfcfft_emulate.argtypes = [ctypes.POINTER(L.Desc)] + [ctypes.c_void_p] * 5
    return _EMU


def _strides(a):
    return [s // a.itemsize for s in a.strides]


def emulate(fake: np.ndarray, real: np.ndarray, grid: int, flags: int, weight=1.0, input_scale=1.0, grad=True,
            dtype_code=None):
    """Runs the kernels' arithmetic serially on the CPU (libtfcfft_emu.so).  bf16 is passed as uint16
    arrays with ``dtype_code=L.BF16``."""
    n = fake.shape[0]
    code = dtype_code if dtype_code is not None else NP_DTYPES[str(fake.dtype)]
    out = np.zeros(8, np.float32)
    per = np.zeros((n, 2), np.float32)
    g = np.zeros_like(fake) if grad else None
    d = L.make_desc(code, grid, flags, fake.shape, _strides(fake), _strides(real), _strides(g) if grad else None,
                    weight, input_scale)
    rc = emu_lib().tfcfft_emulate(ctypes.byref(d), fake.ctypes.data, real.ctypes.data, out.ctypes.data,
                                  per.ctypes.data, g.ctypes.data if grad else None)
    return rc, out, per, g


def emulate_spectra(x: np.ndarray, flags: int, input_scale=1.0, shift=True, y=None):
    """CPU twin of tfcfft_spectra.  Returns (rc, amp_x, pha_x[, amp_y, pha_y])."""
    lib = emu_lib()
    lib.tfcfft_emulate_spectra.restype = ctypes.c_int
    lib.tfcfft_emulate_spectra.argtypes = [ctypes.POINTER(L.Desc)] + [ctypes.c_void_p] * 6 + [ctypes.c_int]
    n, c, p, _ = x.shape
    cp = 3 if (flags & L.CHANNELS_RGB and c == 3) else 1
    w = p if flags & L.FULL_SPECTRUM else p // 2 + 1
    outs = [np.zeros((n, cp, p, w), np.float32) for _ in range(4)]
    d = L.make_desc(NP_DTYPES[str(x.dtype)], 1, flags, x.shape, _strides(x), _strides(y if y is not None else x), None,
                    1.0, input_scale)
    rc = lib.tfcfft_emulate_spectra(ctypes.byref(d), x.ctypes.data, y.ctypes.data if y is not None else None,
                                    outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, outs[3].ctypes.data,
                                    int(shift))
    return (rc, *outs)


def emulate_spectra_bwd(x: np.ndarray, g_amp, g_pha, flags: int, input_scale=1.0, shift=True):
    lib = emu_lib()
    lib.tfcfft_emulate_spectra_bwd.restype = ctypes.c_int
    lib.tfcfft_emulate_spectra_bwd.argtypes = [ctypes.POINTER(L.Desc)] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
    g = np.zeros_like(x)
    d = L.make_desc(NP_DTYPES[str(x.dtype)], 1, flags, x.shape, _strides(x), _strides(x), _strides(g), 1.0, input_scale)
    ga = np.ascontiguousarray(g_amp, np.float32) if g_amp is not None else None
    gp = np.ascontiguousarray(g_pha, np.float32) if g_pha is not None else None
    rc = lib.tfcfft_emulate_spectra_bwd(ctypes.byref(d), x.ctypes.data, ga.ctypes.data if ga is not None else None,
                                        gp.ctypes.data if gp is not None else None, g.ctypes.data, int(shift))
    return rc, g


def l2rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def emulate_triplet(fake: np.ndarray, real: np.ndarray, grid: int, negatives, margin=1.0, eps=1e-6, weight=1.0, grad=True,
                    accumulate_into=None, dtype_code=None):
    """CPU twin of tfcfft_patch_triplet (libtfcfft_emu.so).  Returns (rc, out[4], grad)."""
    lib = emu_lib()
    lib.tfcfft_emulate_triplet.restype = ctypes.c_int
    lib.tfcfft_emulate_triplet.argtypes = [ctypes.POINTER(L.Desc), ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32),
                                           ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    code = dtype_code if dtype_code is not None else NP_DTYPES[str(fake.dtype)]
    out = np.zeros(4, np.float32)
    g = accumulate_into if accumulate_into is not None else (np.zeros_like(fake) if grad else None)
    flags = L.GRAD_ACCUMULATE if accumulate_into is not None else 0
    d = L.make_desc(code, grid, flags, fake.shape, _strides(fake), _strides(real), _strides(g) if g is not None else None, weight, 1.0)
    neg = (ctypes.c_int32 * len(negatives))(*[int(k) for k in negatives])
    rc = lib.tfcfft_emulate_triplet(ctypes.byref(d), fake.ctypes.data, real.ctypes.data, neg, margin, eps, out.ctypes.data,
                                    g.ctypes.data if g is not None else None)
    return rc, out, g


def emulate_temperature(fake, positive, negative, lut, flags=0, margin=1.0, eps=1e-6, weight=1.0, input_scale=1.0, grad=True,
                        dtype_code=None):
    """CPU twin of tfcfft_temperature_triplet.  Returns (rc, out[4], grad)."""
    lib = emu_lib()
    fn = lib.tfcfft_emulate_temperature_triplet
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(L.Desc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                   ctypes.POINTER(ctypes.c_float), ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    code = dtype_code if dtype_code is not None else NP_DTYPES[str(fake.dtype)]
    out = np.zeros(4, np.float32)
    g = np.zeros_like(fake) if grad else None
    d = L.make_desc(code, 1, flags, fake.shape, _strides(fake), _strides(positive), _strides(g) if g is not None else None, weight, input_scale)
    rc = fn(ctypes.byref(d), fake.ctypes.data, positive.ctypes.data, negative.ctypes.data, (ctypes.c_int64 * 4)(*_strides(negative)),
            (ctypes.c_float * 256)(*[float(v) for v in lut]), margin, eps, out.ctypes.data, g.ctypes.data if g is not None else None)
    return rc, out, g


def emulate_regional(fake, real, flags=0, weight=1.0, input_scale=1.0, grad=True, dtype_code=None):
    """CPU twin of tfcfft_regional_loss.  Returns (rc, out[4], per_image, grad)."""
    lib = emu_lib()
    lib.tfcfft_emulate_regional.restype = ctypes.c_int
    lib.tfcfft_emulate_regional.argtypes = [ctypes.POINTER(L.Desc)] + [ctypes.c_void_p] * 5
    code = dtype_code if dtype_code is not None else NP_DTYPES[str(fake.dtype)]
    out = np.zeros(8, np.float32)
    per = np.zeros((fake.shape[0], 2), np.float32)
    g = np.zeros_like(fake) if grad else None
    d = L.make_desc(code, 1, flags, fake.shape, _strides(fake), _strides(real), _strides(g) if grad else None, weight, input_scale)
    rc = lib.tfcfft_emulate_regional(ctypes.byref(d), fake.ctypes.data, real.ctypes.data, out.ctypes.data, per.ctypes.data,
                                     g.ctypes.data if grad else None)
    return rc, out, per, g


def emulate_regional_spectra(x, flags=0, input_scale=1.0, shift=True, grad_amp=None, grad_pha=None):
    """CPU twin of tfcfft_regional_spectra (no incoming gradients) / _bwd.  Returns (rc, amp, pha) or (rc, grad_x)."""
    lib = emu_lib()
    fn = lib.tfcfft_emulate_regional_spectra
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(L.Desc)] + [ctypes.c_void_p] * 6 + [ctypes.c_int]
    n, c = x.shape[:2]
    cp = 3 if (flags & L.CHANNELS_RGB and c == 3) else 1
    bwd = grad_amp is not None or grad_pha is not None
    g = np.zeros_like(x) if bwd else None
    d = L.make_desc(NP_DTYPES[str(x.dtype)], 1, flags, x.shape, _strides(x), _strides(x), _strides(g) if bwd else None, 1.0, input_scale)
    if bwd:
        rc = fn(ctypes.byref(d), x.ctypes.data, None, None, grad_amp.ctypes.data if grad_amp is not None else None,
                grad_pha.ctypes.data if grad_pha is not None else None, g.ctypes.data, int(shift))
        return rc, g
    amp = np.zeros((n, cp, 2, 100, 129), np.float32)
    pha = np.zeros_like(amp)
    rc = fn(ctypes.byref(d), x.ctypes.data, amp.ctypes.data, pha.ctypes.data, None, None, None, int(shift))
    return rc, amp, pha


LUMA_W = np.array([19595.0, 38470.0, 7471.0]) / 65536.0


def robust_grad_error(got, want, fake, real, grid, channels="luma", input_scale=1.0, kappa=8.0):
    """Gradient error that is aware of the L1 loss's discontinuity.

    ``d |x| / dx = sign(x)``: a half-plane bin whose amplitude (or phase) difference between ``fake`` and ``real`` is
    smaller than the rounding noise of an fp32 transform gets an arbitrary sign in ANY fp32 implementation (torch.fft
    in fp32 included), and one flipped bin of a 256 x 256 spectrum already moves the plain L2-relative error to
    ``2 / sqrt(33024) = 1.1e-2``.  Natural images (huge DC, tiny high frequencies, similar fake / real) hit this; white
    noise does not.  This metric transforms the luma-space residual back to the spectrum of every tile, masks the bins
    the fp64 oracle marks as *marginal* -- ``| |F| - |R| |`` or ``|F| * |angle F - angle R|`` below ``kappa * eps32 *
    max|F|`` of the tile -- and returns ``(error over the unmasked bins, fraction of bins masked, plain L2-rel error)``.
    """
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    n, c, h, w = got.shape
    p = h // grid
    if channels == "luma" and c == 3:
        lw = LUMA_W * input_scale
        proj = lambda g: np.tensordot(lw, g, axes=([0], [1])) / (lw @ lw)  # noqa: E731  [N,H,W]
        lum = lambda x: np.tensordot(LUMA_W * input_scale, np.asarray(x, np.float64), axes=([0], [1]))  # noqa: E731
        ge, gw, fl, rl = proj(got - want)[:, None], proj(want)[:, None], lum(fake)[:, None], lum(real)[:, None]
    else:
        ge, gw = (got - want) / input_scale, want / input_scale
        fl, rl = np.asarray(fake, np.float64) * input_scale, np.asarray(real, np.float64) * input_scale

    def tiles(x):  # [N,C',H,W] -> [N,C',g,g,p,p]
        a, b = x.shape[:2]
        return x.reshape(a, b, grid, p, grid, p).transpose(0, 1, 2, 4, 3, 5)

    E, G = np.fft.rfft2(tiles(ge)), np.fft.rfft2(tiles(gw))
    F, R = np.fft.rfft2(tiles(fl)), np.fft.rfft2(tiles(rl))
    sigma = kappa * 2.0 ** -24 * np.abs(F).max(axis=(-1, -2), keepdims=True)
    da = np.abs(np.abs(F) - np.abs(R))
    dp = np.abs(np.angle(F) - np.angle(R)) * np.minimum(np.abs(F), np.abs(R))
    mask = (da < sigma) | (dp < sigma)
    # self-conjugate columns hold k and -k as separate rows: a flip at one shows up at both
    for col in (0, p // 2):
        mk = mask[..., col]
        mask[..., col] = mk | np.roll(mk[..., ::-1], 1, axis=-1)
    wgt = np.full(p // 2 + 1, 2.0)
    wgt[0] = wgt[-1] = 1.0
    num = (np.abs(E) ** 2 * wgt * ~mask).sum()
    den = (np.abs(G) ** 2 * wgt).sum()
    return float(np.sqrt(num / den)), float(mask.mean()), l2rel(got, want)
