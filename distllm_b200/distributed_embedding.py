"""Embedding worker and the one-box multi-GPU driver.

``embedding_worker`` keeps the signature, timers and on-disk result of
distllm/distributed_embedding.py:23-80.  The driver replaces the Parsl pool
(distributed_embedding.py:112-161) with ``torchrun``: one process per GPU, the DOCUMENTS of the input
files sharded contiguously by rank, no traffic between ranks while embedding, and (with ``--gather``)
a single NCCL all-gather of the device-resident pooled embedding matrix at the end.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m distllm_b200.distributed_embedding --config embed.yaml [--gather]
"""

from __future__ import annotations

from argparse import ArgumentParser
from pathlib import Path
from typing import Any

from pydantic import Field
from pydantic import field_validator

from distllm_b200.embed import DatasetConfigs
from distllm_b200.embed import EmbedderConfigs
from distllm_b200.embed import EmbedderResult
from distllm_b200.embed import EncoderConfigs
from distllm_b200.embed import PoolerConfigs
from distllm_b200.embed import WriterConfigs
from distllm_b200.utils import BaseConfig


def embed_file(  # noqa: PLR0913
    input_path: Path,
    output_dir: Path,
    dataset_kwargs: dict[str, Any],
    encoder_kwargs: dict[str, Any],
    pooler_kwargs: dict[str, Any],
    embedder_kwargs: dict[str, Any],
    writer_kwargs: dict[str, Any],
    source_path: Path | None = None,
) -> EmbedderResult:
    """Embed one file, write it under ``output_dir/<uuid4>/`` and hand the result back.

    ``source_path``: when ``input_path`` is a document-range piece of a larger file (the multi-GPU
    driver), the file the piece was cut from; metadata that records the input path (the FASTA reader's
    ``paths``) is pointed back at it before writing."""
    from uuid import uuid4

    from distllm_b200.embed import get_dataset
    from distllm_b200.embed import get_embedder
    from distllm_b200.embed import get_encoder
    from distllm_b200.embed import get_pooler
    from distllm_b200.embed import get_writer
    from distllm_b200.timer import Timer

    total = Timer('finished-embedding', input_path).start()

    with Timer('loaded-encoder', input_path):
        # warm start: the encoder (weights + native workspace) is reused across files
        encoder = get_encoder(encoder_kwargs, register=True)

    dataset = get_dataset(dataset_kwargs)
    pooler = get_pooler(pooler_kwargs)
    embedder = get_embedder(embedder_kwargs)
    writer = get_writer(writer_kwargs)

    with Timer('loaded-dataset', input_path):
        dataloader = dataset.get_dataloader(input_path, encoder)

    with Timer('computed-embeddings', input_path):
        result = embedder.embed(dataloader, encoder, pooler)

    if source_path is not None and result.metadata:
        for meta in result.metadata:
            if meta.get('paths') == str(input_path):
                meta['paths'] = str(source_path)

    dataset_dir = Path(output_dir) / f'{uuid4()}'
    dataset_dir.mkdir(parents=True, exist_ok=True)
    with Timer('wrote-embeddings', input_path):
        writer.write(dataset_dir, result)

    total.stop()
    return result


def embedding_worker(  # noqa: PLR0913
    input_path: Path,
    output_dir: Path,
    dataset_kwargs: dict[str, Any],
    encoder_kwargs: dict[str, Any],
    pooler_kwargs: dict[str, Any],
    embedder_kwargs: dict[str, Any],
    writer_kwargs: dict[str, Any],
) -> None:
    """Embed a single file and save the embeddings (reference-compatible entry point)."""
    embed_file(input_path, output_dir, dataset_kwargs, encoder_kwargs, pooler_kwargs,
               embedder_kwargs, writer_kwargs)


class Config(BaseConfig):
    """YAML schema of a distributed embedding run (distributed_embedding.py:83-109)."""

    # An input directory containing the files to embed.
    input_dir: Path
    # An output directory to save the embeddings.
    output_dir: Path
    # A set of glob patterns to match the input files.
    glob_patterns: list[str] = Field(default=['*'])
    dataset_config: DatasetConfigs
    encoder_config: EncoderConfigs
    pooler_config: PoolerConfigs
    embedder_config: EmbedderConfigs
    writer_config: WriterConfigs
    # Parsed for compatibility with reference YAMLs; the launcher is torchrun, not Parsl.
    compute_config: dict[str, Any] = Field(default_factory=dict)

    @field_validator('input_dir', 'output_dir')
    @classmethod
    def resolve_path(cls, value: Path) -> Path:
        return value.resolve()


def main(argv: list[str] | None = None) -> None:
    """One process per GPU (torchrun).  The global document sequence -- input files in sorted order,
    documents in file order -- is cut into ``world`` contiguous ranges (a document never
    straddles ranks, and one large input file still feeds every GPU); each rank embeds its range with
    ``embed_file`` and, with ``--gather``, the device-resident pooled rows of all ranks meet in ONE
    all-gather (no host round trip before the collective); rank 0 writes ``embeddings_all.npy``."""
    import shutil
    import traceback

    import numpy as np
    import torch
    import torch.distributed as dist

    from distllm_b200.sharding import all_gather_rows
    from distllm_b200.sharding import all_ranks_ok
    from distllm_b200.sharding import count_documents
    from distllm_b200.sharding import materialize_piece
    from distllm_b200.sharding import partition_host_threads
    from distllm_b200.sharding import plan_document_shards
    from distllm_b200.sharding import world_info

    parser = ArgumentParser(description='Embed text (one process per GPU under torchrun)')
    parser.add_argument('--config', type=Path, required=True, help='Path to the .yaml configuration file')
    parser.add_argument('--gather', action='store_true',
                        help='all-gather the pooled embedding matrix; rank 0 writes embeddings_all.npy')
    parser.add_argument('--shard_by', choices=['document', 'file'], default='document',
                        help='unit of the contiguous per-rank ranges (default: document)')
    args = parser.parse_args(argv)

    config = Config.from_yaml(args.config)
    rank, world, local_rank = world_info()
    partition_host_threads()   # before any tokenizer starts its thread pool
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        # one GPU per rank; ranks wrap around only when there are fewer GPUs than ranks (the 2-rank test
        # of this driver on a single-GPU box, which also has to use gloo: B2E_DIST_BACKEND=gloo)
        local_rank %= torch.cuda.device_count()
        torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank) if use_cuda else torch.device('cpu')
    if world > 1 and not dist.is_initialized():
        import os

        dist.init_process_group(os.environ.get('B2E_DIST_BACKEND') or ('nccl' if use_cuda else 'gloo'))

    embedding_dir = config.output_dir / 'embeddings'
    embedding_dir.mkdir(parents=True, exist_ok=True)
    if rank == 0:
        config.write_yaml(config.output_dir / 'config.yaml')

    input_files: list[Path] = []
    for pattern in config.glob_patterns:
        input_files.extend(config.input_dir.glob(pattern))
    input_files = sorted(set(input_files))
    if rank == 0:
        print(f'Found {len(input_files)} input files to embed')

    dataset_kwargs = config.dataset_config.model_dump()
    dataset_name = dataset_kwargs['name']
    header_lines = int(dataset_kwargs.get('header_lines', 1))
    scratch = config.output_dir / '.shards' / f'rank{rank}'
    local_rows: list[torch.Tensor] = []
    error: str | None = None
    try:
        if args.shard_by == 'file':
            counts = [1] * len(input_files)
        else:
            counts = [count_documents(f, dataset_name, header_lines) for f in input_files]
        for file_index, lo, hi in plan_document_shards(counts, world, rank):
            path = input_files[file_index]
            piece = path if args.shard_by == 'file' else materialize_piece(
                path, lo, hi, counts[file_index], dataset_name, scratch, header_lines)
            result = embed_file(
                piece,
                embedding_dir,
                dataset_kwargs=dataset_kwargs,
                encoder_kwargs=config.encoder_config.model_dump(),
                pooler_kwargs=config.pooler_config.model_dump(),
                embedder_kwargs=config.embedder_config.model_dump(),
                writer_kwargs=config.writer_config.model_dump(),
                source_path=None if piece == path else path,
            )
            if args.gather:
                rows = result.device_embeddings
                if rows is None:   # an embedder without the device-resident matrix (third-party plugin)
                    rows = torch.from_numpy(np.ascontiguousarray(result.embeddings)).to(device)
                local_rows.append(rows.to(device=device, dtype=torch.float32))
    except Exception:  # noqa: BLE001  reported collectively below: never raise on a subset of ranks
        error = traceback.format_exc()
    finally:
        shutil.rmtree(scratch, ignore_errors=True)

    # every rank learns whether any rank failed BEFORE the data collective is entered
    if not all_ranks_ok(error is None, device):
        if error is not None:
            print(f'[rank {rank}] embedding failed:\n{error}', flush=True)
        if dist.is_initialized():
            dist.destroy_process_group()
        raise SystemExit(1)

    if args.gather:
        # a rank without documents (world > number of documents) contributes zero rows
        local = torch.cat(local_rows) if local_rows else None
        full = all_gather_rows(local, device=device)
        if rank == 0:
            np.save(config.output_dir / 'embeddings_all.npy', full.cpu().numpy())
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
