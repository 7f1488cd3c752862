#!/usr/bin/env python
"""Golden values of the reference's own loss functions on the 20 real face tiles (run HERE, where
``/root/reference`` exists):

    python tests/golden/make_golden_real_tiles.py

Executes, unmodified, the functions lifted by ``make_golden.py`` (``make_16_patches`` + ``calculate_ffts``, the inline
global block, the inline 4-patch block, ``fft_loss``) on the tile pairs of ``real_tiles.load_pairs`` in fp32 and fp16
and commits ONLY the resulting numbers (``golden_real_tiles.json``); no pixel leaves the reference checkout.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from real_tiles import PAIRS, STRIPS, load_pairs  # noqa: E402


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = dict(generator="tests/golden/make_golden_real_tiles.py", strips=list(STRIPS), pairs=[list(p) for p in PAIRS], cases=[])
    for dtype in ("float32", "float16"):
        fake, real = load_pairs(dtype)
        n = fake.shape[0]
        f, r = torch.from_numpy(fake), torch.from_numpy(real)
        # 16-patch (patchFFT_16P.py:227-375)
        ns = mg.namespace(n, 256, 64)
        exec(mg.lift(mg.F16P, ["make_16_patches", "FFT_Components", "fft_components", "calculate_ffts"]), ns)
        l16 = float(ns["calculate_ffts"](*ns["make_16_patches"](f), *ns["make_16_patches"](r)))
        # global (globalFFT.py:244-284, 494-499)
        ns = mg.namespace(n, 256, 64)
        exec(mg.lift(mg.FGLOB, ["FFT_Components", "fft_components"]), ns)
        ns["fake_B"], ns["real_B"] = f, r
        exec(mg.lines(mg.FGLOB, 494, 499, "Amp_f, Pha_f = fft_components(fake_B)"), ns)
        lg, ag, pg = float(ns["loss_FFT"]), float(ns["loss_Amp"]), float(ns["loss_Pha"])
        # 4-patch mean (patchFFT.py:241-289, 468-471, 498-511)
        ns = mg.namespace(n, 256, 128)
        exec(mg.lift(mg.F4P, ["FFT_Components", "fft_components"]), ns)
        ns["fake_B"] = f
        for k, (y, x) in enumerate(((0, 0), (0, 128), (128, 0), (128, 128))):
            ns[f"B{k + 1}"] = r[:, :, y:y + 128, x:x + 128].contiguous()
        exec(mg.lines(mg.F4P, 468, 471, "fake_B1 = fake_B["), ns)
        exec(mg.lines(mg.F4P, 498, 511, "A1f, P1f = fft_components(fake_B1)"), ns)
        l4 = float(ns["loss_FFT"])
        # 4-patch sum: fft_loss (experiment.py:317-339)
        ns = mg.namespace(n, 256, 128)
        exec(mg.lift(mg.FEXP, ["FFT_Components", "fft_components", "fft_loss"]), ns)
        quads = [r[:, :, y:y + 128, x:x + 128].contiguous() for y in (0, 128) for x in (0, 128)]
        l4s = float(ns["fft_loss"](f, *quads))
        out["cases"].append(dict(dtype=dtype, n=n, p16=l16, glob=lg, glob_amp=ag, glob_pha=pg, p4=l4, p4sum=l4s,
                                 checksum=float(np.abs(fake.astype(np.float64)).sum() + np.abs(real.astype(np.float64)).sum())))
        print(dtype, l16, lg, l4, l4s)
    with open(os.path.join(HERE, "golden_real_tiles.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
