"""Document sharding across the GPUs of one box and the single end-of-run collective.

The reference farms files out through Parsl and merges results on the filesystem
(distllm/distributed_embedding.py:160-161, distllm/cli.py:195-245): no communication.  Here each
rank (one process per GPU, ``torchrun``) owns a contiguous range of documents -- so a document's
sentence buffers never straddle ranks and batch composition is a function of
``(world_size, batch_size)`` only -- embeds them locally, and the pooled ``[N_r, H]`` matrices meet
in ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_info() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (
        int(os.environ.get('RANK', '0')),
        int(os.environ.get('WORLD_SIZE', '1')),
        int(os.environ.get('LOCAL_RANK', '0')),
    )


def shard_range(n_units: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced ``[lo, hi)`` slice of ``n_units`` for ``rank`` (sizes differ by <= 1)."""
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(n_units, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_list(items: list, world_size: int, rank: int) -> list:
    lo, hi = shard_range(len(items), world_size, rank)
    return items[lo:hi]


def all_gather_rows(local: torch.Tensor, group: dist.ProcessGroup | None = None) -> torch.Tensor:
    """Concatenate per-rank row blocks ``[n_r, H]`` in rank order on every rank.

    Row counts differ per rank, so counts are exchanged first (``world`` int64s), blocks are padded
    to the largest count, gathered with one ``all_gather_into_tensor`` and the padding is dropped.
    Without an initialised process group this is the identity.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, count, group=group)
    counts_host = counts.tolist()
    n_max = max(counts_host)
    width = local.shape[1]
    padded = local
    if local.shape[0] != n_max:
        padded = torch.zeros((n_max, width), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    gathered = torch.empty((world * n_max, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(c == n_max for c in counts_host):
        return gathered
    blocks = gathered.view(world, n_max, width)
    return torch.cat([blocks[r, : counts_host[r]] for r in range(world)], dim=0)
