#!/usr/bin/env python
"""Stage timeline (global ns timer) of sub_fwd_kernel / sub_inv_kernel (debug aid).
usage: python tools/trace_sub.py [n] [grid]"""
import ctypes, os, sys
os.environ.setdefault("TFCFFT_NO_CLUSTER", "1")  # the stage marks live in the plain (non-cluster) launches
import numpy as np, torch
sys.path.insert(0, ".")
import tfc_gan_b200 as tfc
lib = tfc._lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(0)
pool = [(torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g),
         torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)) for _ in range(4)]
cfg = tfc.SpectralConfig(grid=grid, weight=0.01, input_scale=255.0)
for i in range(8):
    tfc.spectral_loss_and_grad(*pool[i % 4], config=cfg)
nb = 148 * 6
buf = torch.zeros(nb * 6 * 16, dtype=torch.int64, device="cuda")
stages = {"forward launch (sub_fwd_kernel)": ["load", "rows", "cols+store"],
          "inverse launch (sub_inv_kernel; overwrites the forward trace)": ["cols", "rows", "store"]}
for k, (mode_name, names) in enumerate(stages.items()):
    buf.zero_()
    torch.cuda.synchronize()
    lib.tfcfft_debug_trace(ctypes.c_void_p(buf.data_ptr()))
    if mode_name.startswith("forward"):
        tfc.spectral_terms_per_image(*pool[k], config=cfg)
    else:
        tfc.spectral_loss_and_grad(*pool[k], config=cfg)
    torch.cuda.synchronize()
    lib.tfcfft_debug_trace(None)
    t = buf.cpu().numpy().reshape(nb, 6, 16).astype(np.float64)
    ok = t[:, :, 15] != 0
    t0 = t[:, 0, 0][ok[:, 0]].min()
    print(mode_name, "units traced", int(ok.sum()), " (times in us from the first CTA's start)")
    for it in range(6):
        m = ok[:, it]
        if not m.any():
            continue
        x = (t[:, it][m][:, :4] - t0) / 1e3
        print(f" round {it} ({int(m.sum())} CTAs): start min/mean/max {x[:,0].min():.1f}/{x[:,0].mean():.1f}/{x[:,0].max():.1f} | " +
              " | ".join(f"{nm} ends {x[:, i + 1].min():.1f}/{x[:, i + 1].mean():.1f}/{x[:, i + 1].max():.1f}" for i, nm in enumerate(names)))
