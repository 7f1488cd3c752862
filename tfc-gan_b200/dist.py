"""Data-parallel use of the FFT loss: one process per GPU, batch sharded across ranks.

Every (image, patch) is independent, so there is no data-path collective: each rank
computes the loss and ``d loss / d fake`` of its own shard, normalised by its LOCAL batch like every
other loss term of the training script; DDP's gradient averaging of the generator parameters then
reproduces the global-mean loss exactly when shards are equal-sized.  The only exchange is the
2-float all-reduce of the LOGGED terms (``loss_FFT.item()``,
``TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:664``) over NCCL (gloo in the CPU tests).
The reference itself has no parallelism on this path (``nn.DataParallel`` gathers onto one device,
``...patchFFT_16P.py:444-445``).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced shard ``[lo, hi)`` of a batch of ``n`` (first ``n % world`` ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def global_mean_terms(local_terms: torch.Tensor, local_n: int, group=None) -> torch.Tensor:
    """All-reduce of the logged ``(amp, pha)`` terms: the batch-size-weighted mean over ranks, equal to
    the single-process value on the concatenated batch (also for ragged shards)."""
    buf = torch.cat([local_terms.detach().double() * local_n,
                     torch.tensor([float(local_n)], dtype=torch.float64, device=local_terms.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[:-1] / buf[-1]).to(local_terms.dtype)


def ddp_loss_scale(local_n: int, global_n: int, world: int) -> float:
    """Factor that makes ``mean over ranks`` of locally-normalised losses equal the global mean for
    ragged shards (1.0 when shards are equal)."""
    return (local_n * world) / float(global_n)
