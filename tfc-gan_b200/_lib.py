"""ctypes binding of the C-ABI library ``libtfcfft.so`` (``include/tfcfft.h``).

The product path has no fallback: if the shared library is missing or a symbol is absent the
import of the compute path raises, and any non-zero return code becomes a ``RuntimeError`` with
``tfcfft_strerror`` text.  Nothing here imports ``oracle/``.
"""

from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TFCFFT_LIB selects another build of the SAME library (A/B runs of kernel variants, tools/sessions/gpu_ab.sh)
LIB_PATH = os.environ.get("TFCFFT_LIB") or os.path.join(HERE, "libtfcfft.so")

# enum tfcfft_dtype
F32, F16, BF16, U8 = 0, 1, 2, 3

# flags
CHANNELS_RGB = 1 << 0
NO_PHASE = 1 << 1
DIST_MSE = 1 << 2
PATCH_SUM = 1 << 3
LOG_MAGNITUDE = 1 << 4
FULL_SPECTRUM = 1 << 5
QUANTIZE_U8 = 1 << 6
GRAD_ACCUMULATE = 1 << 8
TEMPS_POSITIVE = 1 << 9
USE_HALFLINE = 1 << 27
USE_PAIR = 1 << 28
USE_LINE = 1 << 29
FORCE_GENERIC = 1 << 30
FORCE_SPLIT = 1 << 31

#: every symbol ``include/tfcfft.h`` declares
EXPORTS = (
    "tfcfft_version",
    "tfcfft_strerror",
    "tfcfft_validate",
    "tfcfft_workspace_bytes",
    "tfcfft_spectra_workspace_bytes",
    "tfcfft_workspace_init",
    "tfcfft_loss",
    "tfcfft_loss_quads",
    "tfcfft_spectra",
    "tfcfft_spectra_bwd",
    "tfcfft_regional_workspace_bytes",
    "tfcfft_regional_loss",
    "tfcfft_regional_spectra",
    "tfcfft_regional_spectra_bwd",
    "tfcfft_triplet_workspace_bytes",
    "tfcfft_patch_triplet",
    "tfcfft_temperature_triplet",
    "tfcfft_vectorize_temps",
    "tfcfft_grad_scale",
    "tfcfft_grad_rescale",
    "tfcfft_debug_trace",
    "tfcfft_launch_count",
    "tfcfft_launch_count_reset",
)


class Desc(ctypes.Structure):
    """``struct tfcfft_desc``."""

    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("dtype", ctypes.c_int32),
        ("grid", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("n", ctypes.c_int64),
        ("c", ctypes.c_int64),
        ("h", ctypes.c_int64),
        ("w", ctypes.c_int64),
        ("fake_stride", ctypes.c_int64 * 4),
        ("real_stride", ctypes.c_int64 * 4),
        ("grad_stride", ctypes.c_int64 * 4),
        ("weight", ctypes.c_float),
        ("input_scale", ctypes.c_float),
        ("grad_scale_host", ctypes.c_float),
        ("reserved", ctypes.c_uint32),
        ("grad_scale_dev", ctypes.c_void_p),
    ]


def make_desc(dtype, grid, flags, shape, fake_stride, real_stride, grad_stride=None, weight=1.0, input_scale=1.0,
              grad_scale_host=1.0, grad_scale_dev=None):
    d = Desc()
    d.struct_size = ctypes.sizeof(Desc)
    d.dtype = dtype
    d.grid = grid
    d.flags = flags & 0xFFFFFFFF
    d.n, d.c, d.h, d.w = (int(v) for v in shape)
    d.fake_stride = (ctypes.c_int64 * 4)(*[int(v) for v in fake_stride])
    d.real_stride = (ctypes.c_int64 * 4)(*[int(v) for v in real_stride])
    d.grad_stride = (ctypes.c_int64 * 4)(*[int(v) for v in (grad_stride or (0, 0, 0, 0))])
    d.weight = float(weight)
    d.input_scale = float(input_scale)
    d.grad_scale_host = float(grad_scale_host)
    d.reserved = 0
    d.grad_scale_dev = grad_scale_dev  # device address (int) or None
    return d


def bind(lib):
    """Declares argument / result types of every export on a loaded ``CDLL``."""
    vp, f32p = ctypes.c_void_p, ctypes.c_void_p
    dp = ctypes.POINTER(Desc)
    lib.tfcfft_version.restype = ctypes.c_int
    lib.tfcfft_version.argtypes = []
    lib.tfcfft_strerror.restype = ctypes.c_char_p
    lib.tfcfft_strerror.argtypes = [ctypes.c_int]
    lib.tfcfft_validate.restype = ctypes.c_int
    lib.tfcfft_validate.argtypes = [dp]
    lib.tfcfft_workspace_bytes.restype = ctypes.c_size_t
    lib.tfcfft_workspace_bytes.argtypes = [dp]
    lib.tfcfft_spectra_workspace_bytes.restype = ctypes.c_size_t
    lib.tfcfft_spectra_workspace_bytes.argtypes = [dp]
    lib.tfcfft_workspace_init.restype = ctypes.c_int
    lib.tfcfft_workspace_init.argtypes = [vp, ctypes.c_size_t, vp]
    lib.tfcfft_loss.restype = ctypes.c_int
    lib.tfcfft_loss.argtypes = [dp, vp, vp, f32p, f32p, vp, vp, ctypes.c_size_t, vp]
    lib.tfcfft_loss_quads.restype = ctypes.c_int
    lib.tfcfft_loss_quads.argtypes = [dp, vp, ctypes.POINTER(ctypes.c_void_p), f32p, f32p, vp, vp, ctypes.c_size_t, vp]
    lib.tfcfft_spectra.restype = ctypes.c_int
    lib.tfcfft_spectra.argtypes = [dp, vp, vp, f32p, f32p, f32p, f32p, ctypes.c_int, vp, ctypes.c_size_t, vp]
    lib.tfcfft_spectra_bwd.restype = ctypes.c_int
    lib.tfcfft_spectra_bwd.argtypes = [dp, vp, f32p, f32p, vp, ctypes.c_int, vp, ctypes.c_size_t, vp]
    lib.tfcfft_regional_workspace_bytes.restype = ctypes.c_size_t
    lib.tfcfft_regional_workspace_bytes.argtypes = [dp]
    lib.tfcfft_regional_loss.restype = ctypes.c_int
    lib.tfcfft_regional_loss.argtypes = [dp, vp, vp, f32p, f32p, vp, vp, ctypes.c_size_t, vp]
    lib.tfcfft_regional_spectra.restype = ctypes.c_int
    lib.tfcfft_regional_spectra.argtypes = [dp, vp, f32p, f32p, ctypes.c_int, vp, ctypes.c_size_t, vp]
    lib.tfcfft_regional_spectra_bwd.restype = ctypes.c_int
    lib.tfcfft_regional_spectra_bwd.argtypes = [dp, vp, f32p, f32p, vp, ctypes.c_int, vp, ctypes.c_size_t, vp]
    lib.tfcfft_triplet_workspace_bytes.restype = ctypes.c_size_t
    lib.tfcfft_triplet_workspace_bytes.argtypes = []
    lib.tfcfft_patch_triplet.restype = ctypes.c_int
    lib.tfcfft_patch_triplet.argtypes = [dp, vp, vp, ctypes.POINTER(ctypes.c_int32), ctypes.c_float, ctypes.c_float, f32p, vp,
                                         vp, ctypes.c_size_t, vp]
    lib.tfcfft_temperature_triplet.restype = ctypes.c_int
    lib.tfcfft_temperature_triplet.argtypes = [dp, vp, vp, vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float),
                                               ctypes.c_float, ctypes.c_float, f32p, vp, vp, ctypes.c_size_t, vp]
    lib.tfcfft_vectorize_temps.restype = ctypes.c_int
    lib.tfcfft_vectorize_temps.argtypes = [dp, vp, ctypes.POINTER(ctypes.c_float), f32p, vp]
    lib.tfcfft_grad_scale.restype = ctypes.c_int
    lib.tfcfft_grad_scale.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int64, f32p, ctypes.c_float, vp]
    lib.tfcfft_grad_rescale.restype = ctypes.c_int
    lib.tfcfft_grad_rescale.argtypes = [vp, ctypes.c_int32, ctypes.c_int64, f32p, f32p, vp, ctypes.c_size_t, vp]
    lib.tfcfft_debug_trace.restype = None
    lib.tfcfft_debug_trace.argtypes = [vp]
    lib.tfcfft_launch_count.restype = ctypes.c_int64
    lib.tfcfft_launch_count.argtypes = []
    lib.tfcfft_launch_count_reset.restype = None
    lib.tfcfft_launch_count_reset.argtypes = []
    return lib


_LIB = None


def load():
    """Loads ``libtfcfft.so`` (built in-tree by ``__graft_entry__.build()``).  Raises if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        missing = [s for s in EXPORTS if not hasattr(lib, s)]
        if missing:
            raise RuntimeError(f"{LIB_PATH} lacks symbols {missing}")
        _LIB = bind(lib)
    return _LIB


def check(rc: int, what: str = "tfcfft"):
    """Raises ``RuntimeError`` for a non-zero return code (negative: argument error; positive: CUDA)."""
    if rc == 0:
        return
    msg = load().tfcfft_strerror(rc)
    raise RuntimeError(f"{what} failed with code {rc}: {msg.decode() if msg else 'unknown error'}")
