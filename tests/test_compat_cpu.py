"""Host logic of the compat layer that needs no GPU: patch geometry and common-base detection."""

import torch

from tfc_gan_b200 import compat


def test_make_16_patches_geometry():
    """Row-major 64x64 views, B2 = columns 64..128 of the top row (``...patchFFT_16P.py:234-251``)."""
    B = torch.arange(2 * 3 * 256 * 256, dtype=torch.float32).reshape(2, 3, 256, 256)
    P = compat.make_16_patches(B)
    assert len(P) == 16 and all(p.shape == (2, 3, 64, 64) for p in P)
    assert torch.equal(P[1], B[:, :, 0:64, 64:128])
    assert torch.equal(P[4], B[:, :, 64:128, 0:64])
    assert torch.equal(P[15], B[:, :, 192:256, 192:256])
    assert P[5].data_ptr() == B[:, :, 64:, 64:].data_ptr()  # views, not copies
    Q = compat.make_4_patches(B)
    assert torch.equal(Q[1], B[:, :, 0:128, 128:256]) and torch.equal(Q[2], B[:, :, 128:256, 0:128])


def test_common_base_detection():
    B = torch.randn(2, 3, 256, 256)
    P = compat.make_16_patches(B)
    assert compat._common_base(P, 4) is B
    assert compat._assemble(P, 4) is B
    # separately allocated patches: concatenated back in row-major order
    C = [p.clone() for p in P]
    assert compat._common_base(C, 4) is None
    assert torch.equal(compat._assemble(C, 4), B)
    # wrong order is not mistaken for the base
    swapped = (P[1], P[0]) + P[2:]
    assert compat._common_base(swapped, 4) is None
    assert not torch.equal(compat._assemble(swapped, 4), B)
    # quadrants
    Q = compat.make_4_patches(B)
    assert compat._common_base(Q, 2) is B
