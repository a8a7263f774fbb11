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


def partition_host_threads() -> int | None:
    """Give this rank its share of the host's cores: ``RAYON_NUM_THREADS`` (the Rust tokenizers' pool) and
    torch's intra-op threads become ``cores // LOCAL_WORLD_SIZE`` unless the user set them.  Eight ranks each
    starting a pool as wide as the whole host oversubscribe it eightfold exactly where the pipeline is
    host-bound (tokenisation).  Call before the first tokenizer use; returns the thread count or None (single rank)."""
    local_world = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')))
    if local_world <= 1:
        return None
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:   # not Linux
        cores = os.cpu_count() or 1
    share = max(1, cores // local_world)
    os.environ.setdefault('RAYON_NUM_THREADS', str(share))
    if 'OMP_NUM_THREADS' not in os.environ:
        torch.set_num_threads(share)
    return share


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


def all_gather_rows(local: torch.Tensor | None, group: dist.ProcessGroup | None = None,
                    device: torch.device | None = None) -> torch.Tensor:
    """Concatenate per-rank row blocks ``[n_r, H]`` in rank order on every rank.

    Row counts differ per rank, so (count, width) pairs are exchanged first, blocks are padded to the
    largest count, gathered with ONE ``all_gather_into_tensor`` and the padding is dropped.  A rank
    with nothing to contribute passes ``None`` or a zero-row tensor (it need not know the width: a rank
    that received no document never built an encoder) -- every rank still enters both collectives, so
    no rank can leave the others hanging.  Without an initialised process group this is the identity.
    """
    if not (dist.is_available() and dist.is_initialized()):
        if local is None:
            raise ValueError('all_gather_rows(None) needs an initialised process group')
        return local
    world = dist.get_world_size(group)
    if world == 1:
        if local is None:
            raise ValueError('all_gather_rows(None) with a world of one')
        return local
    if local is None:
        if device is None:
            raise ValueError('all_gather_rows(None) needs the device the collective runs on')
        local = torch.empty((0, 0), dtype=torch.float32, device=device)
    if local.is_cuda and dist.get_backend(group) == 'gloo':
        # gloo moves host memory: CUDA rows are staged through the host (CPU tests, and two ranks sharing
        # one GPU in the single-GPU driver test -- NCCL refuses two ranks on one device)
        return all_gather_rows(local.cpu(), group=group).to(local.device)
    shape = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=local.device)
    shapes = torch.empty(world * 2, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(shapes, shape, group=group)
    shapes_host = shapes.view(world, 2).tolist()
    counts_host = [int(n) for n, _ in shapes_host]
    widths = {int(h) for n, h in shapes_host if n > 0}
    if len(widths) > 1:
        raise ValueError(f'ranks disagree on the embedding width: {sorted(widths)}')
    width = widths.pop() if widths else int(local.shape[1])
    n_max = max(counts_host)
    if n_max == 0:
        return torch.empty((0, width), dtype=local.dtype, device=local.device)
    padded = local
    if local.shape[0] != n_max or local.shape[1] != width:
        padded = torch.zeros((n_max, width), dtype=local.dtype, device=local.device)
        if local.shape[0]:
            padded[: local.shape[0]] = local
    gathered = torch.empty((world * n_max, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(c == n_max for c in counts_host):
        return gathered
    blocks = gathered.view(world, n_max, width)
    return torch.cat([blocks[r, : counts_host[r]] for r in range(world)], dim=0)


def all_ranks_ok(ok: bool, device: torch.device, group: dist.ProcessGroup | None = None) -> bool:
    """Collective AND of a per-rank success flag: every rank learns whether ANY rank failed, so a job
    aborts on all ranks together instead of one rank raising while the others wait in a collective."""
    if not (dist.is_available() and dist.is_initialized()):
        return ok
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return int(flag.item()) == 0


# ----------------------------------------------------------------------------------- documents
# the unit of sharding is the DOCUMENT, not the file.  A document is one jsonl line, one
# FASTA record, one line of a sequence-per-line file or one row of a saved HF dataset; a document's
# sentence buffers never straddle ranks, and with one big input file all GPUs still get work.


def count_documents(path, dataset_name: str, header_lines: int = 1) -> int:
    """Documents in ``path`` as the reader of ``dataset_name`` would see them."""
    from pathlib import Path

    path = Path(path)
    if dataset_name in ('jsonl', 'jsonl_chunk'):
        # the readers split ``read_text().strip()`` on newlines (embed/datasets/jsonl.py)
        text = path.read_text().strip()
        return len(text.split('\n')) if text else 0
    if dataset_name == 'fasta':
        return sum(1 for line in path.read_text().split('\n') if line.startswith('>'))
    if dataset_name == 'sequence_per_line':
        return max(0, len(path.read_text().splitlines()) - header_lines)
    if dataset_name == 'huggingface':
        from datasets import Dataset

        return len(Dataset.load_from_disk(str(path)))
    raise ValueError(f'no document counter for dataset {dataset_name!r}')


def plan_document_shards(doc_counts: list[int], world_size: int, rank: int) -> list[tuple[int, int, int]]:
    """This rank's contiguous range of the global document sequence (files in order, documents in file
    order), as ``(file_index, lo, hi)`` pieces with ``[lo, hi)`` local to that file."""
    lo, hi = shard_range(sum(doc_counts), world_size, rank)
    pieces = []
    base = 0
    for i, n in enumerate(doc_counts):
        a, b = max(lo, base), min(hi, base + n)
        if a < b:
            pieces.append((i, a - base, b - base))
        base += n
    return pieces


def materialize_piece(path, lo: int, hi: int, n_docs: int, dataset_name: str, scratch_dir,
                      header_lines: int = 1):
    """A file holding documents ``[lo, hi)`` of ``path`` in the same format.  The whole file is returned
    as is; a partial piece is written under ``scratch_dir`` (readers take a path, distllm's Dataset
    protocol has no row-range argument: embed/datasets/base.py:14-40)."""
    from pathlib import Path

    path = Path(path)
    if lo == 0 and hi == n_docs:
        return path
    scratch_dir = Path(scratch_dir)
    scratch_dir.mkdir(parents=True, exist_ok=True)
    out = scratch_dir / f'{path.stem}.docs{lo}-{hi}{path.suffix}'
    if dataset_name in ('jsonl', 'jsonl_chunk'):
        out.write_text('\n'.join(path.read_text().strip().split('\n')[lo:hi]) + '\n')
    elif dataset_name == 'fasta':
        blocks = ('\n' + path.read_text()).split('\n>')[1:]
        out.write_text(''.join('>' + b.rstrip('\n') + '\n' for b in blocks[lo:hi]))
    elif dataset_name == 'sequence_per_line':
        lines = path.read_text().splitlines()
        out.write_text('\n'.join(lines[:header_lines] + lines[header_lines + lo:header_lines + hi]) + '\n')
    elif dataset_name == 'huggingface':
        from datasets import Dataset

        Dataset.load_from_disk(str(path)).select(range(lo, hi)).save_to_disk(str(out))
    else:
        raise ValueError(f'no document slicer for dataset {dataset_name!r}')
    return out
