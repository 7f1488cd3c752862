#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys
import torch
sys.path.insert(0, ".")
import tfc_gan_b200 as tfc

g = torch.Generator(device="cuda").manual_seed(0)
def pair(n, side, dtype=torch.float32):
    f = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g).to(dtype)
    r = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g).to(dtype)
    return f, r
cases = [
    (3, 64, dict(grid=1)),                       # pair kernel, odd tile count
    (2, 256, dict(grid=4)),                      # pair kernel, several pairs per CTA? (32 tiles)
    (1, 256, dict(grid=4, channels="rgb")),      # pair kernel rgb
    (2, 256, dict(grid=2)),                      # sub-tile D=2
    (2, 256, dict(grid=1)),                      # sub-tile D=4
    (1, 256, dict(grid=1, force_split=True)),    # split kernels
    (1, 256, dict(grid=2, force_generic=True)),  # resident kernel 128
    (1, 512, dict(grid=1)),                      # split 512
    (2, 64, dict(grid=4)),                       # resident kernel 16
]
for n, side, opt in cases:
    f, r = pair(n, side)
    l, t, gr = tfc.spectral_loss_and_grad(f, r, **opt)
    torch.cuda.synchronize()
    print(n, side, opt, float(l), bool(torch.isfinite(gr).all()))
f, r = pair(2, 128, torch.float16)
a, p = tfc.spectral_components(f)
torch.cuda.synchronize()
print("spectra ok", a.shape)
# many pairs per CTA (persistent loop, both buffers, drain): 600 tiles of 64x64 on a 148-SM part -> >2 per CTA
f, r = pair(40, 256)
l, t, gr = tfc.spectral_loss_and_grad(f, r, grid=4)
torch.cuda.synchronize()
print("long loop ok", float(l))
