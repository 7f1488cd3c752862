#!/usr/bin/env python
"""A few fused patch-triplet steps (batch 256, 16 patches) for profiler captures."""
import sys
import torch
sys.path.insert(0, ".")
import tfc_gan_b200 as tfc
g = torch.Generator(device="cuda").manual_seed(0)
pool = [(torch.empty(256, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g),
         torch.empty(256, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)) for _ in range(2)]
neg = [(5 * i + 3) % 16 for i in range(16)]
for i in range(6):
    out, grad = tfc.patch_triplet_loss_and_grad(*pool[i % 2], neg, grid=4)
torch.cuda.synchronize()
print("loss", float(out[1]))
