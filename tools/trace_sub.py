#!/usr/bin/env python
"""Stage timeline of sub_pair_kernel (debug aid).  usage: python tools/trace_sub.py [n] [grid]"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
import tfc_gan_b200 as tfc
lib = tfc._lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(0)
fake = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
real = torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)
cfg = tfc.SpectralConfig(grid=grid, weight=0.01, input_scale=255.0)
for _ in range(3):
    tfc.spectral_loss_and_grad(fake, real, config=cfg)
buf = torch.zeros(148 * 6 * 16, dtype=torch.int64, device="cuda")
for mode_name in ("forward launch only (mode 1)", "both launches (mode 2 overwrites mode 1)"):
    buf.zero_()
    lib.tfcfft_debug_trace(ctypes.c_void_p(buf.data_ptr()))
    if mode_name.startswith("forward"):
        tfc.spectral_terms_per_image(fake, real, config=cfg)
    else:
        tfc.spectral_loss_and_grad(fake, real, config=cfg)
    torch.cuda.synchronize()
    lib.tfcfft_debug_trace(None)
    t = buf.cpu().numpy().reshape(148, 6, 16).astype(np.float64)
    ok = t[:, :, 15] != 0
    print(mode_name, "units traced", int(ok.sum()))
    for it in range(6):
        m = ok[:, it]
        if not m.any():
            continue
        x = t[:, it][m]
        print(f" it{it}: compute wait_full={np.mean(x[:,1]-x[:,0]):7.0f} compute={np.mean(x[:,2]-x[:,1]):7.0f} | loader wait_done={np.mean(np.where(x[:,5]>0, x[:,5]-x[:,4], 0)):7.0f} writeback={np.mean(np.where(x[:,5]>0, x[:,6]-x[:,5], 0)):7.0f} load={np.mean(x[:,7]-x[:,6]):7.0f}  unit period={np.mean(x[:,2]-t[:, max(it-1,0)][m][:,2]):7.0f}")
