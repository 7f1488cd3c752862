#!/usr/bin/env python
"""Stage timeline of line_kernel (debug aid).  usage: python tools/trace_line.py [n] [dtype f32|f16] [channels]"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
import tfc_gan_b200 as tfc
lib = tfc._lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dt = sys.argv[2] if len(sys.argv) > 2 else "f32"
ch = sys.argv[3] if len(sys.argv) > 3 else "luma"
g = torch.Generator(device="cuda").manual_seed(0)
fake = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
real = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
if dt == "f16":
    fake, real = fake.half(), real.half()
cfg = tfc.SpectralConfig(grid=4, channels=ch, weight=0.01, input_scale=255.0)
for _ in range(3):
    tfc.spectral_loss_and_grad(fake, real, config=cfg)
nb = 148 * 6
buf = torch.zeros(nb * 6 * 16, dtype=torch.int64, device="cuda")
lib.tfcfft_debug_trace(ctypes.c_void_p(buf.data_ptr()))
tfc.spectral_loss_and_grad(fake, real, config=cfg)
torch.cuda.synchronize()
lib.tfcfft_debug_trace(None)
t = buf.cpu().numpy().reshape(nb, 6, 16).astype(np.float64)
ok = t[:, :, 15] != 0
names = ["load", "rows_fwd", "cols_fwd", "bins", "cols_inv", "rows_inv", "store"]
d = np.diff(t[:, :, :8], axis=2)
print(dt, ch, "tiles traced", int(ok.sum()))
for it in range(6):
    m = ok[:, it]
    if m.any():
        print(f" it{it}: " + " ".join(f"{nm}={d[:, it, i][m].mean():.0f}" for i, nm in enumerate(names)) + f"  total={(t[:, it, 7]-t[:, it, 0])[m].mean():.0f}")
