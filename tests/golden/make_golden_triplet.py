#!/usr/bin/env python
"""Golden vectors for the patch triplet loss, produced by the reference's own lines.

Run HERE (the build container), where ``/root/reference`` exists:

    python tests/golden/make_golden_triplet.py

Lifts ``make_16_patches`` and the inline triplet block of the generator step
(``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:227-253, 558-583``; 4-patch block ``TFCGAN_multigpu_patchFFT.py:
468-471, 474-481``) and executes them UNMODIFIED with torch on the CPU, in a namespace that supplies what the
script's module level would have (``triplet_loss = nn.TripletMarginLoss(margin=1.0, p=2)``, ``:75``).  NumPy is
seeded before the block so the negatives it draws can be replayed (``np.random.randint`` streams are frozen).
The gradient w.r.t. ``fake_B`` comes from ``loss_triplet_patch.backward()``, as in the reference's training step.
Only seeds, the drawn negatives and the numbers produced are stored (``golden_triplet.json`` / ``.npz``).
"""

from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from inputs import make_pair  # noqa: E402
from make_golden import F4P, F16P, REF, lift, lines  # noqa: E402


def main():
    cases, arrays = [], {}
    src16 = lift(F16P, ["make_16_patches"])
    blk16 = lines(F16P, 558, 583, "loss_triplet_patch = 1/16*(")
    blk4a = lines(F4P, 468, 471, "fake_B1 = fake_B[")
    blk4 = lines(F4P, 474, 481, "loss_triplet_patch = 0.25*(")
    for grid, kind, seed, npseed, n, dtype in [
        (4, "uniform", 41, 1001, 2, "float32"),
        (4, "tanh", 42, 1002, 3, "float32"),
        (4, "lowpass", 43, 1003, 2, "float64"),
        (2, "uniform", 44, 1004, 2, "float32"),
        (2, "tanh", 45, 1005, 2, "float64"),
    ]:
        f, r = make_pair(kind, seed, (n, 3, 256, 256), dtype)
        fake = torch.from_numpy(f).requires_grad_(True)
        real = torch.from_numpy(r)
        ns = dict(np=np, torch=torch, nn=nn, triplet_loss=nn.TripletMarginLoss(margin=1.0, p=2), fake_B=fake, real_B=real,
                  opt=types.SimpleNamespace(img_width=256, img_height=256))
        np.random.seed(npseed)
        if grid == 4:
            exec(src16, ns)
            names = [f"B{i}" for i in range(1, 17)]
            ns.update(dict(zip(names, ns["make_16_patches"](real))))
            exec(blk16, ns)
            ref = f"{os.path.relpath(F16P, REF)}:227-253,558-583"
        else:
            p = 128  # the loader's four real quadrants (datasets_temp.py:76-118): TL, TR, BL, BR
            ns.update(B1=real[:, :, 0:p, 0:p], B2=real[:, :, 0:p, p:], B3=real[:, :, p:, 0:p], B4=real[:, :, p:, p:])
            exec(blk4a, ns)
            exec(blk4, ns)
            ref = f"{os.path.relpath(F4P, REF)}:468-471,474-481"
        loss = ns["loss_triplet_patch"]
        loss.backward()
        np.random.seed(npseed)  # replay the draws the block made
        negatives = [np.random.randint(grid * grid, size=1).item() for _ in range(grid * grid)]
        g = fake.grad.double().numpy()
        name = f"triplet_g{grid}_{kind}_{seed}_{dtype}"
        cases.append(dict(name=name, ref=ref, grid=grid, kind=kind, seed=seed, numpy_seed=npseed, n=n, dtype=dtype,
                          negatives=negatives, loss=float(loss), grad_l2=float(np.sqrt((g * g).sum())),
                          grad_abs_sum=float(np.abs(g).sum())))
        arrays[name + "_grad_n0_c1_rows60_70"] = g[0, 1, 60:70, :].astype(np.float32)
    # ---- temperature triplet: vectorize_temps + TempVector_PyTorch + the loss lines (patchFFT_16P.py:254-268, 585-595)
    from torchvision import transforms
    FDS = f"{REF}/TFC-GAN-FFT/datasets_temp.py"
    src_t = lift(FDS, ["TempVector_PyTorch"]) + "\n\n" + lines(F16P, 257, 258, "T = np.linspace(24, 38, num=256)") + "\n\n" + lift(F16P, ["vectorize_temps"])
    l_fb = lines(F16P, 587, 587, "TFB_ = vectorize_temps(fake_B)")
    l_tf = lines(F16P, 592, 593, "TBTF = vectorize_temps(B_tf)")
    l_loss = lines(F16P, 595, 595, "loss_temp_g = criterion_temp(TFB_, TB, TBTF)*lambda_t")
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU here; identity on values
    for kind, seed, n, dtype in [("uniform", 51, 2, "float32"), ("tanh", 52, 2, "float16"), ("unit", 53, 3, "float32")]:
        f, r = make_pair(kind, seed, (n, 3, 256, 256), dtype)
        jit, _ = make_pair(kind, seed + 100, (n, 3, 256, 256), dtype)  # stands in for ColorJitter(real_B): an input of the op
        ns = dict(np=np, torch=torch, nn=nn, transforms=transforms, opt=types.SimpleNamespace(batch_size=n, img_height=256, img_width=256),
                  criterion_temp=nn.TripletMarginLoss(margin=1.0, p=2), lambda_t=10,
                  fake_B=torch.from_numpy(f), real_B=torch.from_numpy(r), B_tf=torch.from_numpy(jit))
        exec(src_t, ns)
        # the loader's T_B (datasets_temp.py:65-67): the same table applied to the real image
        ns["TB"] = torch.stack([torch.Tensor(ns["TempVector_PyTorch"](transforms.ToPILImage()(ns["real_B"][t]).convert("RGB"), ns["d"]).make_pixel_vectors())
                                for t in range(n)])
        exec(l_fb, ns)
        exec(l_tf, ns)
        exec(l_loss, ns)
        name = f"temperature_{kind}_{seed}_{dtype}"
        cases.append(dict(name=name, op="temperature", ref=f"{os.path.relpath(F16P, REF)}:257-268,587-595", kind=kind, seed=seed, n=n,
                          dtype=dtype, lambda_t=10, loss=float(ns["loss_temp_g"]), temps_sum=float(ns["TFB_"].double().sum()),
                          tb_sum=float(ns["TB"].double().sum())))
        arrays[name + "_TFB_n0_rows100_104"] = ns["TFB_"][0, 0, 100:104].numpy()
    # ---- regional FFT loss: FFT_Components + regional_fft_loss (patchFFT_withregion_FFT.py:246-265, 353-402)
    FREG = f"{REF}/TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_withregion_FFT.py"
    src_r = lift(FREG, ["FFT_Components", "regional_fft_loss"])
    for kind, seed, n, dtype in [("uniform", 61, 2, "float32"), ("tanh", 62, 2, "float16"), ("lowpass", 63, 3, "float32")]:
        f, r = make_pair(kind, seed, (n, 3, 256, 256), dtype)
        ns = dict(np=np, torch=torch, nn=nn, transforms=transforms, opt=types.SimpleNamespace(batch_size=n, img_height=256, img_width=256),
                  criterion_amp=nn.L1Loss(), criterion_phase=nn.L1Loss())
        exec(src_r, ns)
        loss = ns["regional_fft_loss"](torch.from_numpy(f), torch.from_numpy(r))
        cases.append(dict(name=f"regional_{kind}_{seed}_{dtype}", op="regional", ref=f"{os.path.relpath(FREG, REF)}:246-265,353-402",
                          kind=kind, seed=seed, n=n, dtype=dtype, loss=float(loss)))
    json.dump(dict(torch=torch.__version__, numpy=np.__version__, cases=cases), open(os.path.join(HERE, "golden_triplet.json"), "w"),
              indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_triplet.npz"), **arrays)
    for c in cases:
        print(c["name"], c["loss"], c.get("negatives", ""))


if __name__ == "__main__":
    main()
