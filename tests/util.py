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
    out = np.zeros(4, np.float32)
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
    out = np.zeros(4, np.float32)
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
