#!/usr/bin/env python
"""BASELINE config 3 harness: the 16-patch FFT loss inside a full TFC-GAN training step (U-Net G + PatchGAN D, AMP,
GradScaler, DDP), batch 256 over 8 x B200 (32 per GPU).  SURVEY.md section 8f-4 -- a HARNESS around the hot path, not a
kernel project: convolutions are cuDNN through torch.

What is restated from the reference (``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py``):
  * ``UNetDown`` / ``UNetUp`` / ``GeneratorUNet`` (``:102-174``): 4x4 stride-1 convolutions followed by an
    anti-aliased stride-2 ``BlurPool``, six levels 256 -> 4, skip connections, ``Upsample`` + ``ZeroPad2d`` + conv + tanh;
  * ``Discriminator1`` (``:182-211``): four spectral-normalised 4x4 convolutions with ``BlurPool`` down-sampling and a
    1-channel 16 x 16 patch output on ``cat(img, condition)``;
  * the training step (``:520-638``): generator step under ``autocast`` (relativistic BCE, patch triplet, temperature
    triplet, FFT loss, ``scaler.scale(loss_G).backward()``), discriminator step, one ``scaler.update()``;
  * ``antialiased_cnns.BlurPool`` is not installed here: ``BlurPool`` below is a local stand-in (reflect-pad (1, 2)
    and a fixed depth-wise [1,3,3,1] x [1,3,3,1] / 64 filter, stride 1 or 2);
  * LPIPS (``criterion_lpips``, ``:70-73``) is OMITTED: ``lpips_pytorch`` and its VGG weights are not available offline;
  * ``nn.DataParallel`` (``:444-445``) becomes one process per GPU with DistributedDataParallel.

Two arms per run:
  ``--losses reference``  the loss block as the reference computes it: per-sample tensor -> uint8 -> PIL-style luma ->
                          ``np.fft.rfft2`` on the HOST for the FFT loss (no gradient), per-sample host table look-up
                          for the temperature vectors (no gradient), 16 ``nn.TripletMarginLoss`` calls on patch views;
  ``--losses fused``      the three terms through this repo's fused kernels (``SpectralLoss`` with the GradScaler's
                          scale folded in, ``PatchTripletLoss``, ``temperature_triplet_loss``), all differentiable.
Prints one JSON line per arm (rank 0): ms per step, images/s over all ranks, and the share of the step spent in the
three loss terms (CUDA events around the loss block).

    python examples/tfcgan_train_step.py --batch 32 --steps 10
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/tfcgan_train_step.py --batch 32
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfc_gan_b200 as tfc  # noqa: E402
from tfc_gan_b200 import compat  # noqa: E402


class BlurPool(nn.Module):
    """Stand-in for ``antialiased_cnns.BlurPool(channels, stride)`` (filter size 4): reflect-pad (1, 2), depth-wise
    binomial filter, stride."""

    def __init__(self, channels: int, stride: int = 2):
        super().__init__()
        a = torch.tensor([1.0, 3.0, 3.0, 1.0])
        k = (a[:, None] * a[None, :]) / 64.0
        self.register_buffer("filt", k[None, None].repeat(channels, 1, 1, 1))
        self.stride, self.channels = stride, channels

    def forward(self, x):
        return F.conv2d(F.pad(x, (1, 2, 1, 2), mode="reflect"), self.filt.to(x.dtype), stride=self.stride, groups=self.channels)


class UNetDown(nn.Module):
    def __init__(self, cin, cout, normalize=True, dropout=0.0):
        super().__init__()
        layers = [nn.Conv2d(cin, cout, 4, 1, 1, bias=False)]
        if normalize:
            layers.append(nn.InstanceNorm2d(cout))
        layers += [nn.LeakyReLU(0.2), BlurPool(cout, stride=2)]
        if dropout:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class UNetUp(nn.Module):
    def __init__(self, cin, cout, dropout=0.0):
        super().__init__()
        layers = [nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=False), BlurPool(cout, stride=1), nn.InstanceNorm2d(cout), nn.ReLU(inplace=True)]
        if dropout:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x, skip):
        return torch.cat((self.model(x), skip), 1)


class GeneratorUNet(nn.Module):
    def __init__(self, channels=3):
        super().__init__()
        self.down1 = UNetDown(channels, 64, normalize=False)
        self.down2 = UNetDown(64, 128)
        self.down3 = UNetDown(128, 256, dropout=0.5)
        self.down4 = UNetDown(256, 512, dropout=0.5)
        self.down5 = UNetDown(512, 512, normalize=False)
        self.down6 = UNetDown(512, 512)
        self.up1 = UNetUp(512, 512)
        self.up2 = UNetUp(1024, 512, dropout=0.5)
        self.up3 = UNetUp(1024, 256, dropout=0.5)
        self.up4 = UNetUp(512, 128)
        self.up5 = UNetUp(256, 64)
        self.final = nn.Sequential(nn.Upsample(scale_factor=2), nn.ZeroPad2d((1, 0, 1, 0)), nn.Conv2d(128, channels, 4, padding=1), nn.Tanh())

    def forward(self, x):
        d1 = self.down1(x)
        d2 = self.down2(d1)
        d3 = self.down3(d2)
        d4 = self.down4(d3)
        d5 = self.down5(d4)
        d6 = self.down6(d5)
        u = self.up1(d6, d5)
        u = self.up2(u, d4)
        u = self.up3(u, d3)
        u = self.up4(u, d2)
        u = self.up5(u, d1)
        return self.final(u).half()  # the reference forces HalfTensor output (:173)


class Discriminator1(nn.Module):
    def __init__(self, channels=3):
        super().__init__()

        def block(cin, cout):
            return [nn.utils.parametrizations.spectral_norm(nn.Conv2d(cin, cout, 4, stride=1, padding=1)), nn.LeakyReLU(0.2, inplace=True),
                    BlurPool(cout, stride=2)]

        self.model = nn.Sequential(*block(channels * 2, 64), *block(64, 128), *block(128, 256), *block(256, 512),
                                   nn.ZeroPad2d((1, 0, 1, 0)), nn.Conv2d(512, 1, 4, padding=1, bias=False))

    def forward(self, img, cond):
        return self.model(torch.cat((img, cond), 1)).half()


# ---- the reference's loss block, as written upstream (host detours, no gradient for FFT / temperature) -------------
T_LUT = np.linspace(24, 38, num=256)


def _luma_u8(x):
    """``transforms.ToPILImage()(x).convert("L")`` on the host: (x * 255) -> uint8 with wrap, integer ITU-R 601 luma."""
    u8 = (x.detach().float().cpu().numpy() * 255).astype(np.uint8).astype(np.int64)
    return ((19595 * u8[0] + 38470 * u8[1] + 7471 * u8[2] + 0x8000) >> 16).astype(np.uint8)


def reference_fft_components(t):
    amp, pha = [], []
    for i in range(t.shape[0]):  # one device -> host round trip per sample (:298-302)
        f = np.fft.fftshift(np.fft.rfft2(_luma_u8(t[i])))
        amp.append(torch.from_numpy(np.abs(f).astype(np.float32)))
        pha.append(torch.from_numpy(np.arctan2(f.imag, f.real).astype(np.float32)))
    return torch.stack(amp).cuda(non_blocking=True), torch.stack(pha).cuda(non_blocking=True)


def reference_loss_block(fake_B, real_B, T_B, B_tf, negatives):
    l1 = nn.L1Loss()
    trip = nn.TripletMarginLoss(margin=1.0, p=2)
    fp, rp = compat.make_16_patches(fake_B), compat.make_16_patches(real_B)
    loss_triplet = sum(trip(fp[i], rp[i], rp[negatives[i]]) for i in range(16)) / 16
    la = lp = 0.0
    for i in range(16):  # calculate_ffts (:323-375): 32 x fft_components
        af, pf = reference_fft_components(fp[i])
        ar, pr = reference_fft_components(rp[i])
        la, lp = la + l1(af, ar), lp + l1(pf, pr)
    loss_fft = 0.5 * (la / 16 + lp / 16)

    def temps(x):  # vectorize_temps (:260-268)
        out = [torch.from_numpy(T_LUT[(x[i, 0].detach().float().cpu().numpy() * 255).astype(np.uint8)].astype(np.float32)) for i in range(x.shape[0])]
        return torch.stack(out).unsqueeze(1).cuda(non_blocking=True)

    loss_temp = nn.TripletMarginLoss(margin=1.0, p=2)(temps(fake_B), T_B.float(), temps(B_tf)) * 10
    return loss_triplet, loss_temp, loss_fft


class FusedLossBlock:
    def __init__(self, scaler):
        # the script adds 1/100 * loss_FFT (:607) and backpropagates scaler.scale(loss_G): both folded into the launch
        self.fft = tfc.SpectralLoss(grid=4, input_scale=255.0, grad_scaler=scaler, loss_multiplier=0.01)
        self.triplet = tfc.PatchTripletLoss(grid=4, margin=1.0)

    def __call__(self, fake_B, real_B, T_B, B_tf, negatives):
        loss_triplet = self.triplet(fake_B, real_B, negatives)
        loss_temp = tfc.temperature_triplet_loss(fake_B, T_B, B_tf, weight=10.0, input_scale=255.0, positive_is_temperatures=True)
        return loss_triplet, loss_temp, self.fft(fake_B, real_B)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (config 3: 256 over 8 GPUs = 32)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--losses", choices=["fused", "reference", "both"], default="both")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel as DDP

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42 + rank)
    np.random.seed(42)
    G, D = GeneratorUNet().to(dev), Discriminator1().to(dev)
    if world > 1:
        G, D = DDP(G, device_ids=[local]), DDP(D, device_ids=[local])
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    bce = nn.BCEWithLogitsLoss()
    n = args.batch
    real_A = torch.empty(n, 3, 256, 256, device=dev).uniform_(-1, 1).half()
    real_B = torch.empty(n, 3, 256, 256, device=dev).uniform_(-1, 1).half()
    B_tf = (real_B.float() * 0.8 + 0.1 * torch.randn_like(real_B.float())).clamp(-1, 1).half()  # stands in for ColorJitter (:590-591)
    T_B = tfc.vectorize_temps(real_B)  # the loader's temperature vector (datasets_temp.py:14-35)
    valid = torch.full((n, 1, 16, 16), 0.9, device=dev, dtype=torch.half)
    fake_lbl = torch.zeros((n, 1, 16, 16), device=dev, dtype=torch.half)

    for arm in (["fused", "reference"] if args.losses == "both" else [args.losses]):
        scaler = torch.amp.GradScaler("cuda")
        block = FusedLossBlock(scaler) if arm == "fused" else reference_loss_block
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        loss_ms = step_ms = 0.0
        for it in range(args.warmup + args.steps):
            negatives = compat.draw_negatives(16)
            torch.cuda.synchronize()
            ev[0].record()
            opt_G.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.float16):
                fake_B = G(real_A)
                pred_fake, real_pred = D(fake_B, real_A), D(real_B, real_A)
                loss_gan = bce(pred_fake - real_pred.detach(), valid)
                ev[1].record()
                loss_triplet, loss_temp, loss_fft = block(fake_B, real_B, T_B, B_tf, negatives)
                ev[2].record()
                # fused arm: the 1/100 already sits inside loss_fft's gradient (loss_multiplier); reference arm: as upstream
                loss_G = 0.5 * loss_gan + loss_triplet + 0.5 * loss_temp + 0.01 * loss_fft
            scaler.scale(loss_G).backward()
            scaler.step(opt_G)
            opt_D.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.float16):
                pr, pf = D(real_B, real_A), D(fake_B.detach(), real_A)
                loss_D = 0.5 * (bce(pr - pf, valid) + bce(pf - pr, fake_lbl))
            scaler.scale(loss_D).backward()
            scaler.step(opt_D)
            scaler.update()
            ev[3].record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                loss_ms += ev[1].elapsed_time(ev[2])
                step_ms += ev[0].elapsed_time(ev[3])
        vals = torch.tensor([step_ms / args.steps, loss_ms / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({
                "harness": "tfcgan_train_step", "losses": arm, "n_gpus": world, "per_gpu_batch": n, "global_batch": n * world,
                "ms_per_step": float(vals[0]), "images_per_s": n * world / float(vals[0]) * 1e3,
                "loss_block_ms": float(vals[1]), "loss_block_share": float(vals[1] / vals[0]),
                "loss_G": float(loss_G), "loss_fft": float(loss_fft), "lpips": "omitted (weights unavailable offline)",
                "models": "GeneratorUNet / Discriminator1 restated from TFCGAN_multigpu_patchFFT_16P.py:102-211, local BlurPool",
            }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
