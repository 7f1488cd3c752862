#!/usr/bin/env python
"""Generate the golden vectors that pin the oracle to the reference.

Run HERE (the build container), where ``/root/reference`` exists:

    python tests/golden/make_golden.py

The reference scripts cannot be imported (module-level argparse, ``os.makedirs`` of
the author's home, ``torch.cuda.set_device``, missing ``lpips_pytorch`` /
``antialiased_cnns`` -- SURVEY.md §8c), so this script lifts the reference's own
function / class definitions and inline loss blocks out of the files with ``ast`` /
line ranges and executes them UNMODIFIED in a namespace that supplies only what the
scripts' module level would have supplied (``opt``, ``np``, ``torch``,
``transforms``, the two ``nn.L1Loss`` criteria).  ``Tensor.cuda`` is patched to the
identity because this container has no GPU; it does not touch values.

Nothing from the reference is copied into the repo: only seeded-input descriptors
and the numbers the reference code produced are written to ``tests/golden/``.
"""

from __future__ import annotations

import ast
import json
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from inputs import make_gray_pairs, make_pair  # noqa: E402

REF = "/root/reference"
F16P = f"{REF}/TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py"
FGLOB = f"{REF}/TFC-GAN-FFT/TFCGAN_multigpu_globalFFT.py"
F4P = f"{REF}/TFC-GAN-FFT/TFCGAN_multigpu_patchFFT.py"
FEXP = f"{REF}/TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_experiment.py"
FSTN = f"{REF}/TFC-STN/TFCGAN_STN21_Original_NewModel3_B2A.py"
FMSE = f"{REF}/TFC-GAN-FFT/Devcom_MagMSE.py"
FMAE = f"{REF}/TFC-GAN-FFT/eval/Eurecom/Eurecom_MagOther.py"


def lift(path, names):
    """Source text of the named top-level defs/classes of ``path`` (last definition wins)."""
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            found[node.name] = ast.get_source_segment(src, node)
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"{path}: missing {missing}")
    return "\n\n".join(found[n] for n in names)


def lines(path, first, last, must_contain):
    """Dedented source lines [first, last] (1-based) of ``path``; sanity-checked."""
    ls = open(path).read().split("\n")[first - 1 : last]
    block = textwrap.dedent("\n".join(ls))
    if must_contain not in block:
        raise RuntimeError(f"{path}:{first}-{last} does not contain {must_contain!r}")
    return block


def namespace(batch_size, img=256, patch=64):
    from torchvision import transforms

    opt = types.SimpleNamespace(
        batch_size=batch_size, img_height=img, img_width=img, patch_height=patch, patch_width=patch, channels=3
    )
    return dict(
        np=np,
        torch=torch,
        nn=nn,
        transforms=transforms,
        opt=opt,
        criterion_amp=nn.L1Loss(),
        criterion_phase=nn.L1Loss(),
    )


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU here; identity on values
    cases = []
    arrays = {}

    def tensors(kind, seed, n, dtype):
        f, r = make_pair(kind, seed, (n, 3, 256, 256), dtype)
        return torch.from_numpy(f), torch.from_numpy(r)

    # ---- 16-patch: make_16_patches + calculate_ffts (patchFFT_16P.py:227-253, 271-375)
    src16 = lift(F16P, ["make_16_patches", "FFT_Components", "fft_components", "calculate_ffts"])
    for kind, seed, n, dtype in [
        ("uniform", 11, 2, "float32"),
        ("uniform", 12, 2, "float16"),
        ("tanh", 13, 2, "float16"),
        ("lowpass", 14, 3, "float32"),
        ("unit", 15, 2, "float32"),
    ]:
        ns = namespace(n, 256, 64)
        exec(src16, ns)
        fake, real = tensors(kind, seed, n, dtype)
        fp = ns["make_16_patches"](fake)
        rp = ns["make_16_patches"](real)
        loss = ns["calculate_ffts"](*fp, *rp)
        case = dict(name=f"p16_{kind}_{seed}_{dtype}", ref=f"{os.path.relpath(F16P, REF)}:227-375", grid=4,
                    patch_reduce="mean", kind=kind, seed=seed, n=n, dtype=dtype, loss=float(loss))
        if seed == 11:  # keep one patch's spectra (patch B6 = row 1, col 1) for layout / bit parity
            amp, pha = ns["fft_components"](fp[5])
            arrays[case["name"] + "_amp_B6"] = amp.numpy()
            arrays[case["name"] + "_pha_B6"] = pha.numpy()
            case["spectra_patch"] = 5
        cases.append(case)

    # ---- global: fft_components + inline block (globalFFT.py:244-284, 494-499)
    srcg = lift(FGLOB, ["FFT_Components", "fft_components"])
    blkg = lines(FGLOB, 494, 499, "Amp_f, Pha_f = fft_components(fake_B)")
    for kind, seed, n, dtype in [("uniform", 21, 2, "float32"), ("tanh", 22, 2, "float16")]:
        ns = namespace(n, 256, 64)
        exec(srcg, ns)
        ns["fake_B"], ns["real_B"] = tensors(kind, seed, n, dtype)
        exec(blkg, ns)
        cases.append(dict(name=f"global_{kind}_{seed}_{dtype}", ref=f"{os.path.relpath(FGLOB, REF)}:244-284,494-499",
                          grid=1, patch_reduce="mean", kind=kind, seed=seed, n=n, dtype=dtype,
                          loss=float(ns["loss_FFT"]), amp=float(ns["loss_Amp"]), pha=float(ns["loss_Pha"])))

    # ---- 4-patch mean: fft_components + inline blocks (patchFFT.py:241-289, 468-471, 498-511)
    src4 = lift(F4P, ["FFT_Components", "fft_components"])
    blk4a = lines(F4P, 468, 471, "fake_B1 = fake_B[")
    blk4b = lines(F4P, 498, 511, "A1f, P1f = fft_components(fake_B1)")
    for kind, seed, n, dtype in [("uniform", 31, 2, "float32"), ("tanh", 32, 2, "float16")]:
        ns = namespace(n, 256, 128)
        exec(src4, ns)
        fake, real = tensors(kind, seed, n, dtype)
        ns["fake_B"] = fake
        # the loader's quadrant crops (datasets_temp.py:76-118) == contiguous slices of real_B
        ns["B1"] = real[:, :, 0:128, 0:128].contiguous()
        ns["B2"] = real[:, :, 0:128, 128:256].contiguous()
        ns["B3"] = real[:, :, 128:256, 0:128].contiguous()
        ns["B4"] = real[:, :, 128:256, 128:256].contiguous()
        exec(blk4a, ns)
        exec(blk4b, ns)
        cases.append(dict(name=f"p4_{kind}_{seed}_{dtype}", ref=f"{os.path.relpath(F4P, REF)}:241-289,468-471,498-511",
                          grid=2, patch_reduce="mean", kind=kind, seed=seed, n=n, dtype=dtype,
                          loss=float(ns["loss_FFT"]), amp=float(ns["loss_Amp"]), pha=float(ns["loss_Pha"])))

    # ---- 4-patch sum: fft_loss (experiment.py:317-339)
    srce = lift(FEXP, ["FFT_Components", "fft_components", "fft_loss"])
    for kind, seed, n, dtype in [("uniform", 41, 2, "float32")]:
        ns = namespace(n, 256, 128)
        exec(srce, ns)
        fake, real = tensors(kind, seed, n, dtype)
        quads = [real[:, :, y : y + 128, x : x + 128].contiguous() for y in (0, 128) for x in (0, 128)]
        loss = ns["fft_loss"](fake, *quads)
        cases.append(dict(name=f"p4sum_{kind}_{seed}_{dtype}", ref=f"{os.path.relpath(FEXP, REF)}:317-339",
                          grid=2, patch_reduce="sum", kind=kind, seed=seed, n=n, dtype=dtype, loss=float(loss)))

    # ---- STN global_fourier_loss (B2A.py:461-467).  It calls ``fft_components(x, patch=False)``,
    # but that file's own fft_components takes no ``patch`` argument and the 16P/4P versions
    # reshape to the patch shape before looking at ``patch`` (dead code upstream: it cannot run
    # against any fft_components the reference ships).  It is executed here against the STN
    # file's own global fft_components through a one-line adaptor that swallows ``patch``.
    srcs = lift(FSTN, ["FFT_Components", "fft_components", "global_fourier_loss"])
    for kind, seed, n, dtype in [("uniform", 51, 2, "float32")]:
        ns = namespace(n, 256, 64)
        exec(srcs, ns)
        _inner = ns["fft_components"]
        ns["fft_components"] = lambda x, patch=True, _f=_inner: _f(x)
        fake, real = tensors(kind, seed, n, dtype)
        loss = ns["global_fourier_loss"](real, fake)
        cases.append(dict(name=f"stn_global_{kind}_{seed}_{dtype}", ref=f"{os.path.relpath(FSTN, REF)}:461-467,502-542",
                          grid=1, patch_reduce="mean", weight=0.01, kind=kind, seed=seed, n=n, dtype=dtype,
                          loss=float(loss)))

    # ---- make_spectra (patchFFT_16P.py:284-289) on one luma image
    ns = namespace(1)
    exec(lift(F16P, ["FFT_Components"]), ns)
    from torchvision import transforms

    fake, _ = tensors("unit", 61, 1, "float32")
    pil = transforms.ToPILImage()(fake[0]).convert("L")
    arrays["make_spectra_unit_61"] = ns["FFT_Components"](pil).make_spectra().astype(np.float32)
    arrays["luma_unit_61"] = np.array(pil)
    cases.append(dict(name="make_spectra_unit_61", ref=f"{os.path.relpath(F16P, REF)}:284-289,300",
                      kind="unit", seed=61, n=1, dtype="float32"))

    # ---- offline metric: mse_spec (Devcom_MagMSE.py:91-118) and the MAE twin
    import pandas as pd
    from scipy.fft import fft2, fftshift
    from sklearn.metrics import mean_absolute_error, mean_squared_error

    for path, fn, metric in [(FMSE, "mse_spec", "mse"), (FMAE, None, "mae")]:
        src = open(path).read()
        tree = ast.parse(src)
        node = [n for n in tree.body if isinstance(n, ast.FunctionDef) and "_spec" in n.name][0]
        ns = dict(np=np, pd=pd, fft2=fft2, fftshift=fftshift, mean_squared_error=mean_squared_error,
                  mean_absolute_error=mean_absolute_error)
        exec(ast.get_source_segment(src, node), ns)
        reals, fakes = make_gray_pairs(71, 4, 256)
        master = [[i, f"f{i}", fakes[i], f"r{i}", reals[i]] for i in range(4)]
        values, _ = ns[node.name](master)
        cases.append(dict(name=f"mag_{metric}_71", ref=f"{os.path.relpath(path, REF)}:{node.lineno}-{node.end_lineno}",
                          metric=metric, seed=71, n=4, side=256, values=[float(v) for v in values]))

    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    with open(os.path.join(HERE, "golden_cases.json"), "w") as fh:
        json.dump(dict(
            generator="tests/golden/make_golden.py",
            versions=dict(torch=torch.__version__, numpy=np.__version__,
                          torchvision=__import__("torchvision").__version__, PIL=__import__("PIL").__version__),
            cases=cases), fh, indent=1)
    for c in cases:
        print(c["name"], c.get("loss", c.get("values", "")))


if __name__ == "__main__":
    main()
