#!/usr/bin/env python
"""Generator step of a TFC-GAN-style training loop with the B200 loss path dropped in (harness only).

What the reference does in its generator step (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:540-612``): forward the
generator under ``autocast``, relativistic adversarial loss, a patch triplet loss on the 16 tiles, a temperature
triplet loss, LPIPS, the 16-patch FFT loss, ``scaler.scale(loss_G).backward()``.  Three of those terms leave the GPU
once per SAMPLE upstream (PIL / NumPy) and two of them carry no gradient; here they are the fused kernels of this
repo and all of them train the generator:

    loss_FFT      -> tfc_gan_b200.SpectralLoss(grid=4, weight=1/100)
    triplet patch -> tfc_gan_b200.PatchTripletLoss(grid=4)
    temperature   -> tfc_gan_b200.temperature_triplet_loss(fake, T_B, jittered_real, weight=lambda_t)

The networks below are small stand-ins written for this example (a 3-level U-Net and a 4-layer patch
discriminator), NOT the reference's models; LPIPS is omitted (its weights are not available offline).  Run:

    python examples/generator_step.py --steps 5 --batch 32
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/generator_step.py

Under torchrun each rank takes its own shard of the batch (one process per GPU, DDP on the generator); the loss path
needs no collective of its own -- only the logged (amp, pha) terms are averaged.
"""

from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfc_gan_b200 as tfc  # noqa: E402


def block(cin, cout, down=True):
    conv = nn.Conv2d(cin, cout, 4, 2, 1) if down else nn.ConvTranspose2d(cin, cout, 4, 2, 1)
    return nn.Sequential(conv, nn.InstanceNorm2d(cout), nn.LeakyReLU(0.2) if down else nn.ReLU())


class SmallUNet(nn.Module):
    def __init__(self, width=32):
        super().__init__()
        w = width
        self.d1, self.d2, self.d3 = block(3, w), block(w, 2 * w), block(2 * w, 4 * w)
        self.u1, self.u2 = block(4 * w, 2 * w, False), block(4 * w, w, False)
        self.out = nn.Sequential(nn.ConvTranspose2d(2 * w, 3, 4, 2, 1), nn.Tanh())

    def forward(self, x):
        a = self.d1(x)
        b = self.d2(a)
        c = self.d3(b)
        y = self.u1(c)
        y = self.u2(torch.cat([y, b], 1))
        return self.out(torch.cat([y, a], 1))


class SmallPatchDiscriminator(nn.Module):
    def __init__(self, width=32):
        super().__init__()
        w = width
        self.net = nn.Sequential(block(6, w), block(w, 2 * w), block(2 * w, 4 * w), nn.Conv2d(4 * w, 1, 3, 1, 1))

    def forward(self, img, cond):
        return self.net(torch.cat([img, cond], 1))


def generator_step(G, D, opt_G, scaler, real_A, real_B, T_B, B_tf, losses, use_fft=True):
    """One generator update; returns the logged scalars.  ``losses`` = (fft, triplet) modules."""
    crit_fft, crit_trip = losses
    opt_G.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        fake_B = G(real_A)
        pred_fake = D(fake_B, real_A)
        pred_real = D(real_B, real_A).detach()
        loss_gan = nn.functional.binary_cross_entropy_with_logits(pred_fake - pred_real, torch.ones_like(pred_fake))
        # the three fused terms read fake_B (fp16 under autocast) natively and return fp32 scalars
        loss_trip = crit_trip(fake_B, real_B)
        loss_temp = tfc.temperature_triplet_loss(fake_B * 0.5 + 0.5, T_B, B_tf * 0.5 + 0.5, weight=10.0, input_scale=255.0)
        loss_fft = crit_fft(fake_B, real_B) if use_fft else fake_B.new_zeros(())
        loss_G = 0.5 * loss_gan + loss_trip + 0.5 * loss_temp + loss_fft
    scaler.scale(loss_G).backward()
    scaler.step(opt_G)
    scaler.update()
    return dict(G=loss_G.detach(), gan=loss_gan.detach(), trip=loss_trip.detach(), temp=loss_temp.detach(), fft=loss_fft.detach())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--side", type=int, default=256)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(1234 + rank)
    G, D = SmallUNet().cuda(), SmallPatchDiscriminator().cuda()
    if world > 1:
        G = nn.parallel.DistributedDataParallel(G, device_ids=[local])
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    scaler = torch.amp.GradScaler("cuda")
    losses = (tfc.SpectralLoss(grid=4, weight=1 / 100, input_scale=255.0), tfc.PatchTripletLoss(grid=4))
    n, s = args.batch, args.side
    for step in range(args.steps):
        real_A = torch.empty(n, 3, s, s, device="cuda").uniform_(-1, 1)
        real_B = torch.empty(n, 3, s, s, device="cuda").uniform_(-1, 1)
        B_tf = (real_B + 0.1 * torch.randn_like(real_B)).clamp(-1, 1)  # stands in for ColorJitter(real_B)
        T_B = tfc.vectorize_temps(real_B * 0.5 + 0.5)[:, 0]            # the loader's temperature tensor
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        logs = generator_step(G, D, opt_G, scaler, real_A, real_B, T_B, B_tf, losses)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        terms = tfc.dist.global_mean_terms(losses[0].last_terms, n) if world > 1 else losses[0].last_terms
        if rank == 0:
            print(f"step {step}: {dt * 1e3:7.2f} ms  " + "  ".join(f"{k} {float(v):.4f}" for k, v in logs.items()) +
                  f"  (amp {float(terms[0]):.3f}, pha {float(terms[1]):.3f})", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
