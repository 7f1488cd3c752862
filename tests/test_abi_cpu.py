"""The C-ABI library loads without a GPU, exports every symbol ``include/tfcfft.h`` declares, and
its host-side argument checking behaves as documented.  No compute calls here."""

import ctypes
import os
import re

import pytest

import tfc_gan_b200 as tfc

L = tfc._lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "tfcfft.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfcfft_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = header_functions()
    assert names, "no functions parsed from the header"
    assert sorted(L.EXPORTS) == names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name
    assert L.load().tfcfft_version() == 200


def test_struct_layout_matches_header():
    # 4x4 + 4x8 + 3x4x8 + 2x4 (weight, input_scale) + 2x4 (grad_scale_host, reserved) + 8 (grad_scale_dev) = 168 bytes
    assert ctypes.sizeof(L.Desc) == 168
    assert L.Desc.grad_scale_dev.offset == 160


def desc(**kw):
    shape = kw.pop("shape", (2, 3, 256, 256))
    st = kw.pop("stride", None) or (shape[1] * shape[2] * shape[3], shape[2] * shape[3], shape[3], 1)
    return L.make_desc(kw.pop("dtype", L.F32), kw.pop("grid", 4), kw.pop("flags", 0), shape, st, st, st, **kw)


def test_validate_accepts_reference_configurations():
    lib = L.load()
    for grid, side in [(4, 256), (2, 256), (1, 256), (4, 512), (1, 512), (4, 64), (1, 16)]:
        d = desc(grid=grid, shape=(3, 3, side, side))
        assert lib.tfcfft_validate(ctypes.byref(d)) == 0
        assert lib.tfcfft_workspace_bytes(ctypes.byref(d)) >= 256
    # a reference-style view: B[:, :, 0:64, 64:128] of a contiguous 256x256 batch
    d = desc(grid=1, shape=(2, 3, 64, 64), stride=(3 * 65536, 65536, 256, 1))
    assert lib.tfcfft_validate(ctypes.byref(d)) == 0


@pytest.mark.parametrize(
    "kw,code",
    [
        (dict(shape=(2, 3, 256, 128)), -4),            # not square
        (dict(shape=(2, 2, 256, 256)), -4),            # channels
        (dict(grid=3), -4),                            # 256 % 3
        (dict(grid=32), -4),                           # patch side 8
        (dict(shape=(2, 3, 1024, 1024), grid=1), -4),  # patch side 1024
        (dict(dtype=7), -3),
        (dict(flags=1 << 20), -7),
        (dict(flags=L.QUANTIZE_U8 | L.CHANNELS_RGB), -7),
        (dict(stride=(3 * 65536, 65536, 256, 2)), -5),
        (dict(stride=(3 * 65536 + 2, 65536, 256, 1)), -5),
        (dict(shape=(0, 3, 256, 256)), -10),
    ],
)
def test_validate_rejects(kw, code):
    lib = L.load()
    d = desc(**kw)
    assert lib.tfcfft_validate(ctypes.byref(d)) == code
    assert lib.tfcfft_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.tfcfft_strerror(code).startswith(b"tfcfft:")


def test_struct_size_skew_is_detected():
    d = desc()
    d.struct_size = 100
    assert L.load().tfcfft_validate(ctypes.byref(d)) == -2
    assert L.load().tfcfft_validate(None) == -1


def test_workspace_grows_with_split_path():
    lib = L.load()
    small = lib.tfcfft_workspace_bytes(ctypes.byref(desc(grid=4, shape=(8, 3, 256, 256))))
    big = lib.tfcfft_workspace_bytes(ctypes.byref(desc(grid=1, shape=(8, 3, 256, 256))))
    assert big >= 8 * 256 * 256 * 8 and small < 1 << 20


def test_product_path_refuses_cpu_tensors():
    import torch

    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        tfc.spectral_loss(x, x, grid=1)
    with pytest.raises(ValueError):
        tfc.SpectralLoss(channels="hsv")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tfc-gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
