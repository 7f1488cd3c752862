"""Two ranks on real shards reproduce the single-GPU result (``-m gpu``).

One process per rank, each running the CUDA kernels on ITS shard of the batch (``dist.shard_batch``): the logged
``global_mean_terms`` and the DDP-averaged gradient of a parameter upstream of ``fake`` must equal the 1-GPU values
on the whole batch -- for equal and for ragged shards.  With two or more GPUs the ranks own one GPU
each and talk NCCL; on a one-GPU box both ranks share the device and the two tiny collectives go over gloo (two
NCCL ranks cannot share a device) -- the kernels under test are the same.
"""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _data(n):
    g = torch.Generator().manual_seed(2026)
    fake = torch.empty(n, 3, 256, 256).uniform_(-1, 1, generator=g)
    real = torch.empty(n, 3, 256, 256).uniform_(-1, 1, generator=g)
    return fake, real


def _worker(rank, world, port, n, grid, backend, q):
    import tfc_gan_b200 as tfc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    kw = dict(device_id=dev) if backend == "nccl" else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    try:
        fake, real = _data(n)
        fk = tfc.dist.shard_batch(fake).to(dev)
        rl = tfc.dist.shard_batch(real).to(dev)
        local_n = fk.shape[0]
        # a "generator parameter": fake = s * pixels, so d loss / d s = sum(pixels * d loss / d fake)
        s = torch.ones((), device=dev, requires_grad=True)
        mod = tfc.SpectralLoss(grid=grid, weight=0.01, input_scale=255.0)
        loss = mod(fk * s, rl)                                        # locally normalised, like every loss term
        (loss * tfc.dist.ddp_loss_scale(local_n, n, world)).backward()
        pg = s.grad.detach().reshape(1).clone()
        coll = pg if backend == "nccl" else pg.cpu()
        dist.all_reduce(coll)                                         # what DDP does to parameter gradients ...
        coll /= world                                                 # ... and its averaging
        terms = mod.last_terms if backend == "nccl" else mod.last_terms.cpu()
        terms = tfc.dist.global_mean_terms(terms, local_n)
        if rank == 0:
            q.put((float(coll), [float(v) for v in terms], tfc.launch_count()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,grid", [(8, 4), (7, 4), (8, 1), (6, 2)])  # equal and ragged shards
def test_two_ranks_on_cuda_reproduce_one_gpu(n, grid):
    import tfc_gan_b200 as tfc

    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, grid, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    pg, terms, launches = q.get()
    assert launches > 0  # the CUDA library did the work in the rank process
    fake, real = _data(n)
    s = torch.ones((), device="cuda", requires_grad=True)
    mod = tfc.SpectralLoss(grid=grid, weight=0.01, input_scale=255.0)
    loss = mod(fake.cuda() * s, real.cuda())
    loss.backward()
    assert pg == pytest.approx(s.grad.item(), rel=2e-5)
    assert terms[0] == pytest.approx(mod.last_terms[0].item(), rel=1e-6)
    assert terms[1] == pytest.approx(mod.last_terms[1].item(), rel=1e-6)
