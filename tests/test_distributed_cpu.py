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


def _worker_empty_rank(rank: int, world: int, port: int, tmp: str) -> None:
    from distllm_b200.sharding import all_ranks_ok

    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # a rank without input must still enter the collectives (zero rows, width unknown)
        local = torch.full((3, 8), 5.0) if rank == 0 else None
        full = all_gather_rows(local, device=torch.device('cpu'))
        assert full.shape == (3, 8) and bool((full == 5.0).all())
        nothing = all_gather_rows(None, device=torch.device('cpu'))
        assert nothing.shape[0] == 0
        # collective error flag: one failing rank is seen by every rank
        assert all_ranks_ok(True, torch.device('cpu')) is True
        assert all_ranks_ok(rank != 1, torch.device('cpu')) is False
        torch.save(full, os.path.join(tmp, f'e{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_all_gather_rows_with_an_empty_rank(tmp_path):
    mp.spawn(_worker_empty_rank, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert torch.equal(torch.load(tmp_path / 'e0.pt'), torch.load(tmp_path / 'e1.pt'))


def test_document_shards_cover_every_document_once(tmp_path):
    import json

    from distllm_b200.embed.datasets.fasta import read_fasta
    from distllm_b200.embed.datasets.jsonl import read_jsonl
    from distllm_b200.sharding import count_documents
    from distllm_b200.sharding import materialize_piece
    from distllm_b200.sharding import plan_document_shards

    # planning: contiguous global ranges cut at file borders
    counts = [5, 0, 3, 4]
    for world in (1, 2, 3, 5, 16):
        seen = []
        for rank in range(world):
            for fi, lo, hi in plan_document_shards(counts, world, rank):
                assert 0 <= lo < hi <= counts[fi]
                seen.extend((fi, d) for d in range(lo, hi))
        assert seen == [(fi, d) for fi, n in enumerate(counts) for d in range(n)], world
    assert plan_document_shards([2], 4, 3) == []          # more ranks than documents: an empty rank
    # jsonl: documents are lines
    f = tmp_path / 'a.jsonl'
    f.write_text('\n'.join(json.dumps({'text': f't{i}', 'path': f'p{i}'}) for i in range(7)) + '\n')
    assert count_documents(f, 'jsonl_chunk') == 7
    piece = materialize_piece(f, 2, 5, 7, 'jsonl', tmp_path / 'scratch')
    assert [r['text'] for r in read_jsonl(piece)] == ['t2', 't3', 't4']
    assert materialize_piece(f, 0, 7, 7, 'jsonl', tmp_path / 'scratch') == f
    # fasta: documents are records
    fa = tmp_path / 'x.fasta'
    fa.write_text('>a 1\nMK\nVL\n>b\nAC\n>c\nGG\nTT\n')
    assert count_documents(fa, 'fasta') == 3
    recs = read_fasta(materialize_piece(fa, 1, 3, 3, 'fasta', tmp_path / 'scratch'))
    assert [(r.tag, r.sequence) for r in recs] == [('b', 'AC'), ('c', 'GGTT')]
    # sequence_per_line keeps its header
    sp = tmp_path / 's.txt'
    sp.write_text('header\nAAA\nCCC\nGGG\n')
    assert count_documents(sp, 'sequence_per_line', header_lines=1) == 3
    assert materialize_piece(sp, 1, 2, 3, 'sequence_per_line', tmp_path / 'scratch').read_text() == 'header\nCCC\n'
