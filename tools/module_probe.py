#!/usr/bin/env python
"""Where does the time of the drop-in module path go?  Device / host time of forward and backward, separately."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfc_gan_b200 as tfc  # noqa: E402


def main():
    for grid, n in ((4, 256), (1, 64)):
        g = torch.Generator(device="cuda").manual_seed(1)
        pool = [(torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g),
                 torch.empty(n, 3, 256, 256, device="cuda").uniform_(-1, 1, generator=g)) for _ in range(3)]
        leaves = [f.detach().requires_grad_(True) for f, _ in pool]
        scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
        mod = tfc.SpectralLoss(grid=grid, weight=0.01, input_scale=255.0, grad_scaler=scaler)
        cfg = tfc.SpectralConfig(grid=grid, weight=0.01, input_scale=255.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = [0.0, 0.0, 0.0]
        host = [0.0, 0.0, 0.0]
        steps = 60
        for i in range(steps + 10):
            fk = leaves[i % 3]
            fk.grad = None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev[0].record()
            loss = mod(fk, pool[i % 3][1])
            ev[1].record()
            t1 = time.perf_counter()
            sl = scaler.scale(loss)
            ev[2].record()
            t2 = time.perf_counter()
            sl.backward()
            ev[3].record()
            t3 = time.perf_counter()
            torch.cuda.synchronize()
            if i >= 10:
                for k in range(3):
                    acc[k] += ev[k].elapsed_time(ev[k + 1])
                host[0] += t1 - t0
                host[1] += t2 - t1
                host[2] += t3 - t2
        # the fused call without autograd, same protocol
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fa = 0.0
        for i in range(steps + 10):
            torch.cuda.synchronize()
            e0.record()
            tfc.spectral_loss_and_grad(pool[i % 3][0], pool[i % 3][1], config=cfg)
            e1.record()
            torch.cuda.synchronize()
            if i >= 10:
                fa += e0.elapsed_time(e1)
        print(f"grid {grid} n {n}: device ms  fwd {acc[0] / steps:.3f}  scale {acc[1] / steps:.3f}  bwd {acc[2] / steps:.3f} | "
              f"host ms  fwd {1e3 * host[0] / steps:.3f}  scale {1e3 * host[1] / steps:.3f}  bwd {1e3 * host[2] / steps:.3f} | "
              f"fused call alone {fa / steps:.3f} ms | grad is the kernel's buffer: {leaves[0].grad.data_ptr()}", flush=True)


if __name__ == "__main__":
    main()
