#!/usr/bin/env python
"""bench.py -- FFT-loss forward+backward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (loss + d loss / d fake) over one batch of synthetic
fake / real images.  Default workload = BASELINE.json configs[1]: global-FFT loss, 256x256,
batch 64 per GPU, fp32 (luma spectrum = the reference's channel handling).
Under torchrun (N > 1) the batch dimension is sharded: every rank processes its own 64 images,
no data-path collective (weak scaling); rank 0 prints ONE JSON line.

value       images/s, inputs resident in HBM, fused tfcfft_loss launch(es) only; inputs rotate over
            a pool of batches larger than L2 so no step sees L2-hot data.  Each step is one replay of a CUDA graph
            that holds one library call (--no-graph: eager calls through Python; reported as `eager` either way).
module_path the same steps through the drop-in nn.Module + autograd backward (expected grad_output folded).
e2e         same metric through the public nn.Module + .backward(), inputs copied from pinned
            host memory every step and the loss read back to the host every step.
roofline    algorithmic bytes (3*C*H*W*4 per image) / device time of the
            hot-path launches, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  the oracle's R1 (torch.fft, fp32, fwd+bwd) on this box's host cores, bounded sample.
--impl reference  the same CPU arm as a stand-alone line (rank 0 only), at the workload's own batch.
--workload patch4-sweep     BASELINE config 4: patch-FFT-4, GLOBAL batch 32..1024 sharded over the N ranks.
--workload combined-512-b32 BASELINE config 5: patch-16 + global loss on the same 512x512 tensors, one summed gradient.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fft_loss_fwd_bwd_images_per_sec"
UNIT = "images/s"

WORKLOADS = {
    # name: (grid, side, per-GPU batch, channels)
    "global-fft-256-b64": dict(grid=1, side=256, batch=64, channels="luma"),      # BASELINE configs[1]
    "patch16-fft-256-b256": dict(grid=4, side=256, batch=256, channels="luma"),   # north_star target shape
    "patch16-fft-256-b32": dict(grid=4, side=256, batch=32, channels="luma"),     # configs[2] per-GPU shard
    "patch4-fft-256-b256": dict(grid=2, side=256, batch=256, channels="luma"),    # configs[3]
    "global-fft-256-b64-rgb": dict(grid=1, side=256, batch=64, channels="rgb"),
    "patch16-fft-256-b256-rgb": dict(grid=4, side=256, batch=256, channels="rgb"),
    "patch16-fft-512-b64": dict(grid=4, side=512, batch=64, channels="luma"),     # configs[4]
    "global-fft-512-b32": dict(grid=1, side=512, batch=32, channels="luma"),      # configs[4]
    "patch16-fft-256-b256-f16": dict(grid=4, side=256, batch=256, channels="luma", dtype="f16"),  # HalfTensor I/O
    "combined-512-b32": dict(grid=4, side=512, batch=32, channels="luma", combined=True),         # configs[4]: patch-16 + global
    "patch4-sweep": dict(grid=2, side=256, batch=256, channels="luma", sweep=(32, 64, 128, 256, 512, 1024)),  # configs[3]
}
DEFAULT_WORKLOAD = "global-fft-256-b64"
VARIANTS = ["patch16-fft-256-b256", "patch4-fft-256-b256", "patch16-fft-256-b256-rgb", "global-fft-256-b64-rgb",
            "patch16-fft-512-b64", "global-fft-512-b32", "patch16-fft-256-b256-f16", "combined-512-b32"]
# algorithmic MFLOP per image (packed forward + one inverse + ~60 flop per bin), for the fp32-issue fraction
MFLOP_PER_IMAGE = {(256, 4, "luma"): 9.9, (256, 4, "rgb"): 29.6, (256, 1, "luma"): 12.5, (256, 1, "rgb"): 37.4,
                   (256, 2, "luma"): 11.2, (512, 4, "luma"): 44.8, (512, 1, "luma"): 55.0}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # CUDA-core FMA peak at the maximum SM clock: 74.4 TFLOP/s
L2_BYTES = 126 << 20


def bytes_per_image(side: int, dtype: str = "f32") -> int:
    return 3 * 3 * side * side * (2 if dtype == "f16" else 4)  # read fake + read real + write grad, RGB


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(workload: str):
    """DRAM bytes per step from the committed ncu --set full capture, if any (profiles/traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks: NVML polling thread during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's R1 (torch.fft) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(wl, budget_s: float, steps: int | None = None, warmup: int = 1, sample_batch: int | None = None):
    """R1 fp32 fwd+bwd on one batch of the workload (its own batch size), all host threads.  Returns a dict."""
    import torch

    import oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    side = wl["side"]
    if sample_batch is None:
        sample_batch = wl["batch"]
    fake = torch.empty(sample_batch, 3, side, side).uniform_(-1, 1, generator=g)
    real = torch.empty(sample_batch, 3, side, side).uniform_(-1, 1, generator=g)

    def step():
        fk = fake.clone().requires_grad_(True)
        loss, _, _ = oracle.spectral_loss_r1(fk, real, grid=wl["grid"], channels=wl["channels"], dtype=torch.float32)
        if wl.get("combined"):  # config 5: patch-16 + global on the same tensors
            loss = loss + oracle.spectral_loss_r1(fk, real, grid=1, channels=wl["channels"], dtype=torch.float32)[0]
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    times = []
    t_end = time.perf_counter() + budget_s
    while True:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if time.perf_counter() > t_end:  # bounded: the CPU arm never runs longer than its budget
            break
    total = sum(times)
    res = {
        "value": sample_batch * len(times) / total,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{len(times)} steps x {sample_batch} images of {side}x{side}, oracle R1 (torch.fft fp32 fwd+bwd, "
                  f"channels={wl['channels']}, grid={wl['grid']}), torch {torch.__version__}, {cores} threads",
        "ms_per_step": 1e3 * total / len(times),
        "steps": len(times),
    }
    # the reference as shipped (R0: uint8 + PIL-style luma + numpy rfft2), forward only, one thread
    try:
        t0 = time.perf_counter()
        oracle.spectral_loss_r0(fake[:4].numpy(), real[:4].numpy(), wl["grid"])
        res["as_shipped_r0_fwd_only_images_per_s"] = 4 / (time.perf_counter() - t0)
    except Exception:
        pass
    return res


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 60.0 if args.steps_given else 20.0
    cb = cpu_arm(wl, budget, steps=args.steps if args.steps_given else None, warmup=max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps"],
        "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **{k: wl[k] for k in ("grid", "side", "channels")},
                   "per_gpu_batch": wl["batch"], "note": "CPU arm: each step is one full batch of the workload on the host cores"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "as_shipped_r0_fwd_only_images_per_s" in cb:
        line["as_shipped_r0_fwd_only_images_per_s"] = cb["as_shipped_r0_fwd_only_images_per_s"]
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_pool(torch, wl, batch=None):
    """Batches of synthetic fake / real images resident in HBM; together larger than 3x L2, so that rotating over
    them never serves a step from L2."""
    dev = torch.device("cuda", torch.cuda.current_device())
    batch = batch or wl["batch"]
    per_batch = 2 * batch * 3 * wl["side"] ** 2 * (2 if wl.get("dtype") == "f16" else 4)
    pool_n = max(2, -(-3 * L2_BYTES // per_batch))
    g = torch.Generator(device=dev).manual_seed(1234 + int(os.environ.get("RANK", "0")))
    pool = []
    for _ in range(pool_n):
        f = torch.empty(batch, 3, wl["side"], wl["side"], device=dev).uniform_(-1, 1, generator=g)
        r = torch.empty_like(f).uniform_(-1, 1, generator=g)
        if wl.get("dtype") == "f16":
            f, r = f.half(), r.half()
        pool.append((f, r))
    return pool


def hot_call(tfc, wl):
    """One step of the hot path on a resident batch: fused loss + gradient (config 5: two fused calls that share one
    gradient buffer -- the second one adds into it)."""
    cfg = tfc.SpectralConfig(grid=wl["grid"], channels=wl["channels"], weight=0.01, input_scale=255.0)
    if not wl.get("combined"):
        return lambda f, r: tfc.spectral_loss_and_grad(f, r, config=cfg)
    # config 5: both grids on the whole batch, the second call adds into the first one's gradient (no add pass).
    # Walking the batch in L2-sized chunks saves HBM bytes but runs every launch at a fraction of a wave: measured
    # 57 k / 69 k / 84 k images/s at chunk 8 / 16 / 32 (profiles/r02_d8_ab.txt), so the whole batch is the default.
    chunk = int(os.environ.get("TFCFFT_COMBINED_CHUNK", "0"))
    return lambda f, r: tfc.multi_grid_loss_and_grad(f, r, grids=(wl["grid"], 1), chunk=chunk, channels=wl["channels"],
                                                     weight=0.01, input_scale=255.0)


def time_kernel_path(tfc, torch, wl, steps, warmup, barrier, graph=True, batch=None):
    """Device-timed hot-path steps on resident inputs.  Returns dict(secs, launches, pool_n, host_us, graph)."""
    pool = make_pool(torch, wl, batch)
    pool_n = len(pool)
    call = hot_call(tfc, wl)
    sink = None
    for i in range(max(warmup, pool_n)):  # also fills the per-device launch caches before any capture
        sink = call(*pool[i % pool_n])
    torch.cuda.synchronize()
    graphs = None
    if graph:
        try:
            cs = torch.cuda.Stream()
            graphs = []
            for f, r in pool:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cs):
                    out = call(f, r)
                graphs.append((g, out))
            for g, _ in graphs:  # one untimed replay each
                g.replay()
            torch.cuda.synchronize()
        except Exception as e:  # capture is an optimisation of the launch path, not a requirement
            sys.stderr.write(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager calls\n")
            graphs = None
            torch.cuda.synchronize()
    barrier()
    tfc.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    if graphs is not None:
        for i in range(steps):
            graphs[(warmup + i) % pool_n][0].replay()
        sink = graphs[(warmup + steps - 1) % pool_n][1]
    else:
        for i in range(steps):
            sink = call(*pool[(warmup + i) % pool_n])
    e1.record()
    host_us = 1e6 * (time.perf_counter() - t0) / steps  # host time to ENQUEUE one step
    torch.cuda.synchronize()
    barrier()
    secs = e0.elapsed_time(e1) / 1e3
    launches = tfc.launch_count()
    if graphs is not None:  # replays do not pass through the library's counter: count one captured call instead
        tfc.reset_launch_count()
        call(*pool[0])
        torch.cuda.synchronize()
        launches = tfc.launch_count() * steps
    assert torch.isfinite(sink[0]).item(), "non-finite loss in the timed region"
    return dict(secs=secs, launches=launches, pool_n=pool_n, host_us=host_us, graph=graphs is not None)


def time_module_path(tfc, torch, wl, steps, warmup):
    """The number a training script sees: ``SpectralLoss(...)(fake, real)`` + ``scaler.scale(loss).backward()`` on
    HBM-resident inputs (autograd in the loop, GradScaler's scale folded into the producing launch)."""
    pool = make_pool(torch, wl)
    pool_n = len(pool)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
    mod = tfc.SpectralLoss(grid=wl["grid"], channels=wl["channels"], weight=0.01, input_scale=255.0, grad_scaler=scaler)
    leaves = [f.detach().requires_grad_(True) for f, _ in pool]

    def step(i):
        fk = leaves[i % pool_n]
        fk.grad = None
        scaler.scale(mod(fk, pool[i % pool_n][1])).backward()
        return fk.grad

    for i in range(max(warmup, pool_n)):
        step(i)
    torch.cuda.synchronize()
    # Python's autograd bookkeeping costs ~0.25 ms of host time per step -- more than the GPU work of these small
    # steps, and hidden behind the networks' kernels in a real training step -- so the step (forward + backward) is
    # captured once per pool batch and replayed: the timing then shows the GPU side of the module path.
    graphs = None
    try:
        cs = torch.cuda.Stream()
        graphs = []
        for i in range(pool_n):
            leaves[i].grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cs):
                scaler.scale(mod(leaves[i], pool[i][1])).backward()
            graphs.append(g)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()
    except Exception as e:
        sys.stderr.write(f"[bench] module-path graph capture failed ({type(e).__name__}: {e}); timing eager autograd\n")
        graphs = None
        torch.cuda.synchronize()
    tfc.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graphs is not None:
        for i in range(steps):
            graphs[(warmup + i) % pool_n].replay()
        g = leaves[(warmup + steps - 1) % pool_n].grad
    else:
        for i in range(steps):
            g = step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    assert g is not None and torch.isfinite(g).all().item()
    if graphs is not None:
        tfc.reset_launch_count()
        step(0)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3, tfc.launch_count() * steps
    return e0.elapsed_time(e1) / 1e3, tfc.launch_count()


def time_triplet(tfc, torch, batch, side, grid, steps, warmup):
    """Device-timed fused patch-triplet loss + gradient (the first 'next' row of the scope table) on resident inputs
    rotating over a pool larger than L2.  Algorithmic bytes: fake read + real read + gradient write (the negative
    patch is another tile of the same real image: an L2 hit by construction, confirmed by ncu --
    profiles/r01_ncu_patch_triplet_summary.txt)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    per_batch = 2 * batch * 3 * side * side * 4
    pool_n = max(2, -(-3 * L2_BYTES // per_batch))
    g = torch.Generator(device=dev).manual_seed(4321)
    pool = [(torch.empty(batch, 3, side, side, device=dev).uniform_(-1, 1, generator=g),
             torch.empty(batch, 3, side, side, device=dev).uniform_(-1, 1, generator=g)) for _ in range(pool_n)]
    neg = [(5 * i + 3) % (grid * grid) for i in range(grid * grid)]
    sink = None
    for i in range(warmup):
        sink = tfc.patch_triplet_loss_and_grad(*pool[i % pool_n], neg, grid=grid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        sink = tfc.patch_triplet_loss_and_grad(*pool[(warmup + i) % pool_n], neg, grid=grid)
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(sink[0][0]).item()
    secs = e0.elapsed_time(e1) / 1e3
    ips = batch * steps / secs
    return ips, ips * 3 * 3 * side * side * 4 / 1e9


def time_temperature(tfc, torch, batch, side, steps, warmup):
    """Device-timed fused temperature triplet loss + gradient (second 'next' row).  Only the red channel takes part:
    algorithmic bytes = fake.R + positive.R + negative.R reads + grad.R write = 4 planes per image."""
    dev = torch.device("cuda", torch.cuda.current_device())
    pool_n = max(2, -(-3 * L2_BYTES // (3 * batch * 3 * side * side * 4)))
    g = torch.Generator(device=dev).manual_seed(777)
    pool = [tuple(torch.empty(batch, 3, side, side, device=dev).uniform_(0, 1, generator=g) for _ in range(3)) for _ in range(pool_n)]
    acc = torch.zeros(batch, 3, side, side, device=dev)
    sink = None
    for i in range(warmup):
        sink = tfc.temperature_triplet_loss_and_grad(*pool[i % pool_n], weight=10.0, accumulate_into=acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        sink = tfc.temperature_triplet_loss_and_grad(*pool[(warmup + i) % pool_n], weight=10.0, accumulate_into=acc)
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(sink[0][0]).item()
    ips = batch * steps / (e0.elapsed_time(e1) / 1e3)
    return ips, ips * 5 * side * side * 4 / 1e9  # accumulate mode also reads the old gradient plane


def time_e2e(tfc, torch, wl, steps, warmup, barrier, dist):
    """Public API (nn.Module + backward) with pinned-host inputs copied in and the loss read back
    every step; H2D of step i+1 overlaps the kernels of step i on a second stream."""
    dev = torch.device("cuda", torch.cuda.current_device())
    shape = (wl["batch"], 3, wl["side"], wl["side"])
    g = torch.Generator().manual_seed(99)
    hd = torch.float16 if wl.get("dtype") == "f16" else torch.float32
    host = [(torch.empty(shape).uniform_(-1, 1, generator=g).to(hd).pin_memory(),
             torch.empty(shape).uniform_(-1, 1, generator=g).to(hd).pin_memory()) for _ in range(2)]
    dbuf = [(torch.empty(shape, device=dev, dtype=hd), torch.empty(shape, device=dev, dtype=hd)) for _ in range(2)]
    if wl.get("combined"):
        class _Multi:
            last_terms = None

            def __call__(self, f, r):
                loss, terms = tfc.multi_grid_loss(f, r, grids=(wl["grid"], 1), channels=wl["channels"], weight=0.01,
                                                  input_scale=255.0, return_terms=True)
                self.last_terms = terms[0]
                return loss
        mod = _Multi()
    else:
        mod = tfc.SpectralLoss(grid=wl["grid"], channels=wl["channels"], weight=0.01, input_scale=255.0)
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(dev)

    def issue_copy(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            dbuf[b][0].copy_(host[b][0], non_blocking=True)
            dbuf[b][1].copy_(host[b][1], non_blocking=True)
            ready[b].record(copy_stream)

    def compute(i):
        b = i % 2
        main.wait_event(ready[b])
        fk = dbuf[b][0].detach().requires_grad_(True)
        loss = mod(fk, dbuf[b][1])
        loss.backward()
        freed[b].record(main)
        terms = mod.last_terms
        if dist is not None:
            terms = tfc.dist.global_mean_terms(terms, wl["batch"])  # the logged value, 2 floats over NCCL
        return loss, terms, fk.grad

    for b in range(2):
        freed[b].record(main)
    total = warmup + steps
    issue_copy(0)
    t0 = None
    last = None
    for i in range(total):
        if i == warmup:
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
        if i + 1 < total:
            issue_copy(i + 1)
        loss, terms, grad = compute(i)
        last = loss.item()  # device -> host read of the step's result, every step
    torch.cuda.synchronize()
    barrier()
    secs = time.perf_counter() - t0
    assert last == last, "nan loss in e2e"
    h2d = 2 * shape[0] * shape[1] * shape[2] * shape[3] * (2 if hd == torch.float16 else 4)
    return secs, h2d, 4


def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs next to its GPU (NVML's ideal-CPU mask) BEFORE the pinned host buffers are
    allocated: 8 ranks x 100 MB of pinned H2D per step otherwise all come out of NUMA node 0."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def run_sweep(args, wl, tfc, torch, world, rank, barrier, maxr, peak):
    """BASELINE config 4: patch-FFT-4 at GLOBAL batch 32..1024 sharded over the ranks (strong scaling in N)."""
    pts = []
    for gb in wl["sweep"]:
        per = gb // world
        if per < 1:
            continue
        steps = max(20, min(args.steps, 400))
        r = time_kernel_path(tfc, torch, wl, steps, 5, barrier, graph=not args.no_graph, batch=per)
        secs = maxr(r["secs"])
        ips = world * per * steps / secs
        mb = per * bytes_per_image(wl["side"]) / 1e6
        pts.append({"global_batch": gb, "per_gpu_batch": per, "value": ips, "unit": UNIT, "ms_per_step": 1e3 * secs / steps,
                    "roofline_frac_per_gpu": ips / world * bytes_per_image(wl["side"]) / 1e9 / peak,
                    "per_gpu_megabytes": mb,
                    # a shard that moves less than ~8 us of HBM traffic is bounded by launch latency / one partial wave
                    "latency_bound": bool(mb * 1e6 / (peak * 1e9) < 8e-6)})
    return pts


def run_ours(args, wl):
    import torch

    import tfc_gan_b200 as tfc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    cpus = bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own log (NCCL_DEBUG as the caller set it) stays on stderr / fd 1 -> stderr; the result line goes out
        # on the private descriptor claimed in main()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(x):
        if dist is None:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    peak, peak_src = peaks()
    sampler = ClockSampler(local)
    sampler.start()
    # launch path of the timed region: eager calls keep the programmatic-dependent-launch overlap BETWEEN calls (the
    # next call's first kernel is scheduled while the previous call drains) and win while the host can enqueue a step
    # faster than the GPU runs it; graph replay wins when it cannot (small shards, many ranks per host).  A short probe
    # decides; all ranks take the same branch.
    use_graph = args.graph and not args.no_graph
    if not args.graph and not args.no_graph:
        ps = max(10, min(args.steps, 60))
        pe = maxr(time_kernel_path(tfc, torch, wl, ps, 5, barrier, graph=False)["secs"])
        pg = maxr(time_kernel_path(tfc, torch, wl, ps, 5, barrier, graph=True)["secs"])
        use_graph = pg < pe
    args.no_graph = not use_graph
    res = time_kernel_path(tfc, torch, wl, args.steps, args.warmup, barrier, graph=use_graph)
    clocks = sampler.stop()
    rank_secs = gather(res["secs"])
    secs = max(rank_secs)
    launches, pool_n = res["launches"], res["pool_n"]
    images = world * wl["batch"] * args.steps
    value = images / secs
    # the same steps as eager calls through Python (what a script without graph capture sees) + host cost per call
    osteps = max(10, min(args.steps, 300))
    other = time_kernel_path(tfc, torch, wl, osteps, 5, barrier, graph=not res["graph"])
    eager = other if res["graph"] else res
    eager_secs = maxr(eager["secs"]) / (osteps if res["graph"] else args.steps)
    graph_secs = (secs / args.steps) if res["graph"] else (maxr(other["secs"]) / osteps if other["graph"] else None)
    host_us = maxr(eager["host_us"])

    mod_secs, mod_launches = (None, None)
    if not wl.get("combined"):
        msteps = max(10, min(args.steps, 300))
        mod_secs, mod_launches = time_module_path(tfc, torch, wl, msteps, 5)
        mod_secs = maxr(mod_secs) / msteps

    e2e_steps = max(3, min(args.steps, 200))
    e_secs, h2d, d2h = time_e2e(tfc, torch, wl, e2e_steps, max(3, min(args.warmup, 10)), barrier, dist)
    e_secs = maxr(e_secs)
    e2e_value = world * wl["batch"] * e2e_steps / e_secs

    bpi = bytes_per_image(wl["side"], wl.get("dtype", "f32"))
    achieved = (wl["batch"] * args.steps * bpi / secs) / 1e9  # per GPU
    tr = traffic_for(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl.get("dtype", "f32"), "data": "synthetic",
        "config": {
            "workload": args.workload, "grid": wl["grid"], "side": wl["side"], "channels": wl["channels"],
            "per_gpu_batch": wl["batch"], "global_batch": world * wl["batch"], "parallelism": f"dp{world}",
            "l2": f"inputs rotate over a pool of {pool_n} batches ({pool_n * 2 * wl['batch'] * 3 * wl['side']**2 * 4 >> 20} MiB > 126 MiB L2)",
            "launch": "one CUDA-graph replay per step (a graph holds one tfcfft_loss call)" if res["graph"] else "eager library calls",
            "channels_note": "luma is the reference's channel handling (.convert('L')); the per-channel reading of "
                             "BASELINE config 2 is variants['global-fft-256-b64-rgb']",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": tr.get("bytes_per_step") if isinstance(tr, dict) else tr,
            "traffic_source": tr.get("source") if isinstance(tr, dict) else "profiles/traffic.json (ncu --set full, dram__bytes_read+write per step)",
            "peak_source": peak_src,
            "kernel": "all launches of one tfcfft_loss call (per-GPU)", "algorithmic_bytes_per_image": bpi,
            "launches_per_step": launches / args.steps,
        },
        "eager": {"value": world * wl["batch"] / eager_secs, "ms_per_step": 1e3 * eager_secs, "host_us_per_call": host_us,
                  "note": "steps enqueued call by call from Python; host_us_per_call = host time to enqueue one step"},
        "graph": ({"value": world * wl["batch"] / graph_secs, "ms_per_step": 1e3 * graph_secs,
                   "note": "one CUDA-graph replay per step"} if graph_secs else None),
        "rank_secs": rank_secs,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "SpectralLoss(fake, real).backward(); loss.item()",
                "note": "PCIe-bound: every step copies its fake / real batch from pinned host memory"},
        "clocks": clocks,
        "gpu_launches": launches,
        "cpu_affinity": f"{len(cpus)} cpus near GPU {local}" if cpus else None,
    }
    if mod_secs is not None:
        mv = world * wl["batch"] / mod_secs
        line["module_path"] = {"value": mv, "unit": UNIT, "ms_per_step": 1e3 * mod_secs,
                               "roofline_frac": mv / world * bpi / 1e9 / peak, "launches_per_step": mod_launches / max(10, min(args.steps, 300)),
                               "api": "SpectralLoss(grad_scaler=scaler)(fake, real); scaler.scale(loss).backward()  (HBM-resident inputs)"}
    if wl.get("sweep"):
        line["sweep"] = run_sweep(args, wl, tfc, torch, world, rank, barrier, maxr, peak)
        line["config"]["sweep"] = "global batch 32..1024 sharded over the ranks (strong scaling); value = the 256-per-GPU point"
    if rank == 0 and world == 1 and not args.no_variants:
        var = {}
        vs = max(10, args.steps // 4)
        for name in VARIANTS:
            w = WORKLOADS[name]
            r = time_kernel_path(tfc, torch, w, vs, 3, barrier, graph=not args.no_graph)
            ips = w["batch"] * vs / r["secs"]
            var[name] = {"value": ips, "unit": UNIT,
                         "roofline_frac": ips * bytes_per_image(w["side"], w.get("dtype", "f32")) / 1e9 / peak}
            mf = MFLOP_PER_IMAGE.get((w["side"], w["grid"], w["channels"]))
            if mf and w["channels"] == "rgb":  # the per-channel modes sit at the fp32 ridge: report the issue-side fraction too
                var[name]["fp32"] = {"achieved_tflops": ips * mf / 1e6, "peak_tflops": FP32_PEAK_TFLOPS,
                                     "frac": ips * mf / 1e6 / FP32_PEAK_TFLOPS, "algorithmic_mflop_per_image": mf}
            if w.get("combined"):
                two = time_kernel_path(tfc, torch, dict(w, combined=False), vs, 3, barrier, graph=not args.no_graph)
                glob = time_kernel_path(tfc, torch, dict(w, combined=False, grid=1), vs, 3, barrier, graph=not args.no_graph)
                var[name]["two_separate_calls_value"] = w["batch"] * vs / (two["secs"] + glob["secs"])
                var[name]["note"] = "patch-16 + global on the same 512x512 tensors, one summed gradient; counted as 3 tensor passes"
        ips, gbs = time_triplet(tfc, torch, 256, 256, 4, vs, 3)
        var["patch16-triplet-256-b256"] = {"value": ips, "unit": UNIT, "roofline_frac": gbs / peak,
                                           "note": "fused TripletMarginLoss fwd+bwd on 16 patches; 3 tensor passes per image"}
        ips, gbs = time_temperature(tfc, torch, 256, 256, vs, 3)
        var["temperature-triplet-256-b256"] = {"value": ips, "unit": UNIT, "roofline_frac": gbs / peak,
                                               "note": "fused temperature-LUT triplet fwd+bwd (red channel only, accumulating into an "
                                                       "existing gradient): 5 single-channel planes per image"}
        # e2e with fp16 host staging, the reference's own input type (HalfTensor, patchFFT_16P.py:524-530): half the PCIe bytes
        if not wl.get("combined"):
            hs, hb, _ = time_e2e(tfc, torch, dict(wl, dtype="f16"), max(3, min(args.steps, 100)), 3, barrier, None)
            var["e2e-f16-host-staging"] = {"value": wl["batch"] * max(3, min(args.steps, 100)) / hs, "unit": UNIT, "h2d_bytes_per_step": hb}
        line["variants"] = var
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_arm(wl, 12.0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if "as_shipped_r0_fwd_only_images_per_s" in cb:
            line["cpu_baseline"]["as_shipped_r0_fwd_only_images_per_s"] = cb["as_shipped_r0_fwd_only_images_per_s"]
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_OUT = None


def claim_stdout():
    """Rank 0 must print exactly ONE line on stdout, but native libraries write there too (NCCL prints its version
    banner with printf when the first communicator comes up).  Keep a private handle to the real stdout for the
    result line and point file descriptor 1 at stderr for everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager library calls instead of CUDA-graph replays")
    ap.add_argument("--graph", action="store_true", help="always time CUDA-graph replays (default: whichever of the two "
                                                         "launch paths a short probe finds faster on this box)")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 2000
    if args.warmup is None:
        args.warmup = 50
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
