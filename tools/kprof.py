#!/usr/bin/env python
"""Per-kernel device durations and inter-kernel gaps of one workload in its real (warm, back-to-back) sequence,
via torch.profiler (CUPTI activity records; no replay, no cache flush).
usage: python tools/kprof.py [n] [grid] [side] [channels] [dtype]"""
import sys
import torch
sys.path.insert(0, ".")
import tfc_gan_b200 as tfc
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 1
side = int(sys.argv[3]) if len(sys.argv) > 3 else 256
ch = sys.argv[4] if len(sys.argv) > 4 else "luma"
dt = sys.argv[5] if len(sys.argv) > 5 else "f32"
g = torch.Generator(device="cuda").manual_seed(0)
pool = []
for _ in range(4):
    f = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g)
    r = torch.empty(n, 3, side, side, device="cuda").uniform_(-1, 1, generator=g)
    pool.append((f.half(), r.half()) if dt == "f16" else (f, r))
import os
cfg = tfc.SpectralConfig(grid=grid, channels=ch, weight=0.01, input_scale=255.0, use_line=bool(os.environ.get("KPROF_USE_LINE")),
                         use_pair=bool(os.environ.get("KPROF_USE_PAIR")))
for i in range(10):
    tfc.spectral_loss_and_grad(*pool[i % 4], config=cfg)
torch.cuda.synchronize()
STEPS = 40
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(STEPS):
        tfc.spectral_loss_and_grad(*pool[i % 4], config=cfg)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "tfcfft" in e.name]
ev.sort(key=lambda e: e.time_range.start)
per = {}
order = []
for e in ev:
    k = e.name.split("(")[0].replace("void tfcfft::", "")
    if k not in per:
        per[k] = []
        order.append(k)
    per[k].append(e.time_range.end - e.time_range.start)
gaps = [ev[i + 1].time_range.start - ev[i].time_range.end for i in range(len(ev) - 1)]
nk = len(order)
print(f"n={n} grid={grid} side={side} {ch} {dt}: {len(ev)} kernel records, {nk} kernels per call")
tot = 0.0
for k in order:
    v = sorted(per[k])
    med = v[len(v) // 2]
    tot += med
    print(f"  {k:40s} median {med:8.1f} us   min {v[0]:8.1f}   max {v[-1]:8.1f}   n={len(v)}")
if gaps:
    inner = [gp for i, gp in enumerate(gaps) if (i + 1) % nk != 0]
    outer = [gp for i, gp in enumerate(gaps) if (i + 1) % nk == 0]
    med = lambda a: sorted(a)[len(a) // 2] if a else float("nan")
    print(f"  gaps: inside a call median {med(inner):.1f} us, between calls median {med(outer):.1f} us;  sum of kernel medians {tot:.1f} us")
