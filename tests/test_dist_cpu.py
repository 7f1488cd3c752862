"""Multi-rank host logic on CPU: world_size 2 over gloo.  The per-rank loss is stood in by the oracle
(this is a test of the sharding / normalisation / logged-term all-reduce in ``dist.py``, not of the
kernels)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import tfc_gan_b200 as tfc
from inputs import make_pair


def test_shard_bounds_cover_batch():
    for n in (1, 7, 8, 64, 255):
        for world in (1, 2, 3, 8):
            spans = [tfc.dist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fake, real = make_pair("uniform", 77, (n, 3, 64, 64), "float64")
        fk = tfc.dist.shard_batch(torch.from_numpy(fake)).clone().requires_grad_(True)
        rl = tfc.dist.shard_batch(torch.from_numpy(real))
        local_n = fk.shape[0]
        loss, amp, pha = oracle.spectral_loss_r1(fk, rl, grid=4)       # locally normalised
        scale = tfc.dist.ddp_loss_scale(local_n, n, world)
        (loss * scale).backward()
        # what DDP does to parameter gradients: average over ranks.  Here the "parameter" is a scalar
        # multiplier on fake, d loss / d s = sum(fake * grad).
        pg = (fk.detach() * fk.grad).sum().reshape(1)
        dist.all_reduce(pg)
        pg /= world
        terms = tfc.dist.global_mean_terms(torch.stack([amp.detach(), pha.detach()]), local_n)
        if rank == 0:
            q.put((float(pg), terms.numpy().tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 5])  # equal and ragged shards
def test_two_ranks_reproduce_single_process(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    pg, terms = q.get()
    fake, real = make_pair("uniform", 77, (n, 3, 64, 64), "float64")
    fk = torch.from_numpy(fake).clone().requires_grad_(True)
    loss, amp, pha = oracle.spectral_loss_r1(fk, torch.from_numpy(real), grid=4)
    loss.backward()
    ref_pg = float((fk.detach() * fk.grad).sum())
    assert pg == pytest.approx(ref_pg, rel=1e-10)
    assert terms[0] == pytest.approx(float(amp), rel=1e-10)
    assert terms[1] == pytest.approx(float(pha), rel=1e-10)
