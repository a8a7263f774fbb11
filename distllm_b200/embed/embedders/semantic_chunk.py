"""Semantic-chunk embedder: encode sentence buffers, split documents where adjacent buffers
diverge, re-encode the resulting chunks.

Behaviour follows distllm/embed/embedders/semantic_chunk.py:24-294.  What changes is where the
work happens: pass-1 embeddings stay on the GPU, the adjacent cosine distance of *all* documents
is one launch of the fused normalise+dot kernel (``b2e_adjacent_cosine_dist`` with a document-id
vector masking cross-document pairs), and only the ``[N-1]`` distance vector comes back to the host
for the percentile split, which is discrete bookkeeping and stays in numpy.
"""

from __future__ import annotations

from typing import Any
from typing import Literal

import numpy as np
import torch
from pydantic import Field
from torch.utils.data import DataLoader

from distllm_b200 import _native
from distllm_b200.embed.datasets.utils import DataCollator
from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.embed.embedders.full_sequence import compute_embeddings_pair
from distllm_b200.embed.embedders.full_sequence import compute_embeddings_device
from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.embed.poolers.base import Pooler
from distllm_b200.utils import BaseConfig


def adjacent_distances_device(
    buffer_embeds: torch.Tensor,
    doc_id: torch.Tensor | None = None,
) -> torch.Tensor:
    """fp32 ``[N-1]`` cosine distances between consecutive rows (NaN across documents)."""
    return _native.adjacent_cosine_dist(buffer_embeds.contiguous(), doc_id)


def calculate_distances_between_buffer(buffer_embeds: np.ndarray | torch.Tensor) -> np.ndarray:
    """``1 - cos(e_i, e_{i+1})`` for consecutive rows, float64 array of fp32 values (:24-55)."""
    if isinstance(buffer_embeds, np.ndarray):
        if not torch.cuda.is_available():
            raise _native.NativeError('calculate_distances_between_buffer needs a CUDA device')
        buffer_embeds = torch.from_numpy(np.ascontiguousarray(buffer_embeds)).cuda()
    if len(buffer_embeds) < 2:
        return np.zeros(0)
    return adjacent_distances_device(buffer_embeds).cpu().numpy().astype(np.float64)


def build_chunks(
    distances: np.ndarray,
    breakpoint_percentile_threshold: int,
) -> list[tuple[int, int]]:
    """Half-open ``(start, end)`` row groups, split after every distance above the percentile.

    Same rules as semantic_chunk.py:58-102: no distances -> the single (empty) group ``(0, 0)``;
    threshold = ``np.percentile`` (linear interpolation); strictly-greater comparison; the last
    group runs to ``len(distances) + 1``.
    """
    n = len(distances)
    if n == 0:
        return [(0, 0)]
    threshold = np.percentile(distances, breakpoint_percentile_threshold)
    cuts = np.flatnonzero(np.asarray(distances) > threshold) + 1
    starts = np.concatenate(([0], cuts))
    ends = np.concatenate((cuts, [n + 1]))
    return [(int(s), int(e)) for s, e in zip(starts, ends)]


def document_ranges(metadata: list[dict[str, Any]]) -> list[tuple[int, int]]:
    """Runs of consecutive rows sharing ``metadata['path']`` (semantic_chunk.py:150-158)."""
    ranges = []
    start = 0
    for i in range(1, len(metadata)):
        if metadata[i]['path'] != metadata[start]['path']:
            ranges.append((start, i))
            start = i
    ranges.append((start, len(metadata)))
    return ranges


def compute_semantic_chunks(
    dataloader: DataLoader,
    encoder: Encoder,
    pooler: Pooler,
    breakpoint_percentile_threshold: int,
    min_chunk_length: int,
) -> InMemoryDataset:
    """Pass 1 + split: returns the dataset of semantically grouped chunk texts (:105-207)."""
    dataset = dataloader.dataset
    if dataset.metadata is None:
        raise ValueError('Metadata is required for semantic chunking.')
    if dataset.metadata[0].get('path') is None:
        raise ValueError('Metadata path is required for semantic chunking.')

    doc_ranges = document_ranges(dataset.metadata)

    # pass 1: one embedding per sentence buffer, kept on the device
    buffer_embeds = compute_embeddings_device(dataloader, encoder, pooler)
    if encoder.dtype != torch.float32:
        # the reference stores pass-1 results in encoder.dtype before measuring distances
        buffer_embeds = buffer_embeds.to(encoder.dtype).to(torch.float32)

    doc_id = torch.empty(len(dataset), dtype=torch.int32)
    for k, (lo, hi) in enumerate(doc_ranges):
        doc_id[lo:hi] = k
    distances_all = (
        adjacent_distances_device(buffer_embeds, doc_id.to(buffer_embeds.device))
        .cpu()
        .numpy()
        .astype(np.float64)
    )

    row_groups: list[tuple[int, int]] = []
    for lo, hi in doc_ranges:
        groups = build_chunks(distances_all[lo : hi - 1], breakpoint_percentile_threshold)
        row_groups.extend((lo + s, lo + e) for s, e in groups)

    texts = [''.join(m['sentence'] for m in dataset.metadata[s:e]) for s, e in row_groups]
    metas = [dataset.metadata[s] for s, _ in row_groups]

    keep = [i for i, text in enumerate(texts) if len(text) > min_chunk_length]
    texts = [texts[i] for i in keep]
    metas = [metas[i] for i in keep]
    for meta in metas:
        meta.pop('sentence')
    # a chunk is the concatenation of the sentences its rows are centred on: with a sentence token cache on the
    # pass-1 dataset (jsonl_chunk) pass 2 assembles its ids from the same cached pieces
    cache = getattr(dataset, 'token_cache', None)
    index = getattr(dataset, 'sentence_index', None)
    if cache is not None and index is not None:
        parts = [tuple(index[r] for r in range(*row_groups[i])) for i in keep]
        return InMemoryDataset(texts, metas, parts=parts, token_cache=cache)
    return InMemoryDataset(texts, metas)


class SemanticChunkEmbedderConfig(BaseConfig):
    """Configuration for the semantic chunk embedder."""

    name: Literal['semantic_chunk'] = 'semantic_chunk'  # type: ignore[assignment]
    breakpoint_percentile_threshold: int = Field(
        90,
        description='The percentile of cosine dissimilarity that must be '
        'exceeded between a group of sentences and the next to form a chunk. '
        'The smaller this number is, the more chunks will be generated.',
    )
    chunk_batch_size: int = Field(8, description='The batch size for the chunked text.')
    min_chunk_length: int = Field(
        750,
        description='The minimum length of a chunk (number of characters) to '
        'filter out any small chunks.',
    )
    normalize_embeddings: bool = Field(
        False,
        description='Whether to return normalized the embeddings.',
    )


class SemanticChunkEmbedder:
    """Embeds semantically coherent chunks instead of raw buffers."""

    def __init__(self, config: SemanticChunkEmbedderConfig) -> None:
        self.config = config

    def embed(self, dataloader: DataLoader, encoder: Encoder, pooler: Pooler) -> EmbedderResult:
        cfg = self.config
        chunks = compute_semantic_chunks(
            dataloader=dataloader,
            encoder=encoder,
            pooler=pooler,
            breakpoint_percentile_threshold=cfg.breakpoint_percentile_threshold,
            min_chunk_length=cfg.min_chunk_length,
        )
        chunk_loader = DataLoader(
            pin_memory=dataloader.pin_memory,
            batch_size=cfg.chunk_batch_size,
            num_workers=dataloader.num_workers,
            dataset=chunks,
            collate_fn=DataCollator(encoder.tokenizer, cache=getattr(chunks, 'token_cache', None)),
        )
        device_rows, chunk_embeds = compute_embeddings_pair(
            dataloader=chunk_loader,
            encoder=encoder,
            pooler=pooler,
            normalize=cfg.normalize_embeddings,
        )
        return EmbedderResult(embeddings=chunk_embeds, text=chunks.data, metadata=chunks.metadata,
                              device_embeddings=device_rows)
