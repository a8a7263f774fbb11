"""world_size-2 gloo test of the end-of-run collective and the document sharding."""

from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distllm_b200.sharding import all_gather_rows
from distllm_b200.sharding import shard_range


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_docs: int, tmp: str) -> None:
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_docs, world, rank)
        # "embedding" of document d is a row filled with d: rank order must equal document order
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 8)
        full = all_gather_rows(local)
        assert full.shape == (n_docs, 8)
        assert torch.equal(full[:, 0], torch.arange(n_docs, dtype=torch.float32))
        torch.save(full, os.path.join(tmp, f'full{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_docs', [7, 8])
def test_all_gather_rows_gloo_world2(tmp_path, n_docs):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_docs, str(tmp_path)), nprocs=world, join=True)
    a = torch.load(tmp_path / 'full0.pt')
    b = torch.load(tmp_path / 'full1.pt')
    assert torch.equal(a, b)


def test_all_gather_rows_is_identity_without_process_group():
    x = torch.randn(3, 4)
    assert all_gather_rows(x) is x
