"""Embedding worker and the one-box multi-GPU driver.

``embedding_worker`` keeps the signature, timers and on-disk result of
distllm/distributed_embedding.py:23-80.  The driver replaces the Parsl pool
(distributed_embedding.py:112-161) with ``torchrun``: one process per GPU, input files sharded
contiguously by rank, no traffic between ranks while embedding, and (with ``--gather``) a single
NCCL all-gather of the pooled embedding matrix at the end.

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
) -> EmbedderResult:
    """Embed one file, write it under ``output_dir/<uuid4>/`` and hand the result back."""
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
    import numpy as np
    import torch
    import torch.distributed as dist

    from distllm_b200.sharding import all_gather_rows
    from distllm_b200.sharding import shard_list
    from distllm_b200.sharding import world_info

    parser = ArgumentParser(description='Embed text (one process per GPU under torchrun)')
    parser.add_argument('--config', type=Path, required=True, help='Path to the .yaml configuration file')
    parser.add_argument('--gather', action='store_true',
                        help='all-gather the pooled embedding matrix; rank 0 writes embeddings_all.npy')
    args = parser.parse_args(argv)

    config = Config.from_yaml(args.config)
    rank, world, local_rank = world_info()
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')

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

    local_rows = []
    for path in shard_list(input_files, world, rank):
        result = embed_file(
            path,
            embedding_dir,
            dataset_kwargs=config.dataset_config.model_dump(),
            encoder_kwargs=config.encoder_config.model_dump(),
            pooler_kwargs=config.pooler_config.model_dump(),
            embedder_kwargs=config.embedder_config.model_dump(),
            writer_kwargs=config.writer_config.model_dump(),
        )
        local_rows.append(torch.from_numpy(np.ascontiguousarray(result.embeddings)))

    if args.gather:
        if not local_rows:
            raise RuntimeError('rank has no input files; --gather needs at least one file per rank')
        local = torch.cat(local_rows)
        device = torch.device('cuda', local_rank) if torch.cuda.is_available() else torch.device('cpu')
        full = all_gather_rows(local.to(device))
        if rank == 0:
            np.save(config.output_dir / 'embeddings_all.npy', full.cpu().numpy())
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
