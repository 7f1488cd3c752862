#!/usr/bin/env python
"""Stage timeline of the packed 64x64 kernel (debug aid; uses tfcfft_debug_trace).  Run on the GPU box."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import tfc_gan_b200 as tfc

lib = tfc._lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ch = sys.argv[2] if len(sys.argv) > 2 else "luma"
g = torch.Generator(device="cuda").manual_seed(0)
fake = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
real = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
cfg = tfc.SpectralConfig(grid=4, channels=ch, weight=0.01, input_scale=255.0)
for _ in range(3):
    tfc.spectral_loss_and_grad(fake, real, config=cfg)
nblk = 148 * 3
buf = torch.zeros(nblk * 6 * 16, dtype=torch.int64, device="cuda")
lib.tfcfft_debug_trace(ctypes.c_void_p(buf.data_ptr()))
tfc.spectral_loss_and_grad(fake, real, config=cfg)
torch.cuda.synchronize()
lib.tfcfft_debug_trace(None)
t = buf.cpu().numpy().reshape(nblk, 6, 16)
valid = t[:, :, 15] != 0
names = ["load", "rows1", "rows2", "cols1", "cols2", "bins", "icols2", "icols1", "irows2", "store"]
d = np.diff(t[:, :, :11], axis=2).astype(np.float64)
print("blocks with data:", int(valid[:, 0].sum()), " pairs traced:", int(valid.sum()))
print("stage durations in SM cycles (mean / p10 / p90) over all traced pairs:")
tot = 0
for i, nm in enumerate(names):
    v = d[:, :, i][valid]
    tot += v.mean()
    print(f"  {nm:7s} {v.mean():9.0f} {np.percentile(v,10):9.0f} {np.percentile(v,90):9.0f}")
print(f"  total   {tot:9.0f}")
print("per pair-iteration mean stage cycles:")
for it in range(6):
    m = valid[:, it]
    if m.any():
        print(f"  it{it}: " + " ".join(f"{nm}={d[:, it, i][m].mean():.0f}" for i, nm in enumerate(names)))
gt = t[:, :, 15].astype(np.float64)
g0 = gt[valid].min()
for it in range(6):
    m = valid[:, it]
    if m.any():
        e = (gt[:, it][m] - g0) / 1e3
        print(f"pair-iteration {it}: end time us  min {e.min():8.1f}  median {np.median(e):8.1f}  max {e.max():8.1f}  (n={m.sum()})")
# per-SM clock estimate
b = 0
cyc = t[b, 0, 10] - t[b, 0, 0]
print("example block 0, pair 0: cycles", cyc)
