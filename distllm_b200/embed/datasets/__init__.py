"""Dataset family (readers feeding the hot path; CPU side, same names as the reference)."""

from __future__ import annotations

from typing import Any
from typing import Union

from distllm_b200.embed._factory import build_from_strategies
from distllm_b200.embed.datasets.base import Dataset
from distllm_b200.embed.datasets.fasta import FastaDataset
from distllm_b200.embed.datasets.fasta import FastaDatasetConfig
from distllm_b200.embed.datasets.huggingface import HuggingFaceDataset
from distllm_b200.embed.datasets.huggingface import HuggingFaceDatasetConfig
from distllm_b200.embed.datasets.jsonl import JsonlDataset
from distllm_b200.embed.datasets.jsonl import JsonlDatasetConfig
from distllm_b200.embed.datasets.jsonl_chunk import JsonlChunkDataset
from distllm_b200.embed.datasets.jsonl_chunk import JsonlChunkDatasetConfig
from distllm_b200.embed.datasets.single_line import SequencePerLineDataset
from distllm_b200.embed.datasets.single_line import SequencePerLineDatasetConfig
from distllm_b200.utils import BaseConfig

DatasetConfigs = Union[
    FastaDatasetConfig,
    JsonlDatasetConfig,
    JsonlChunkDatasetConfig,
    SequencePerLineDatasetConfig,
    HuggingFaceDatasetConfig,
]

STRATEGIES: dict[str, tuple[type[BaseConfig], type[Dataset]]] = {
    'fasta': (FastaDatasetConfig, FastaDataset),
    'jsonl': (JsonlDatasetConfig, JsonlDataset),
    'jsonl_chunk': (JsonlChunkDatasetConfig, JsonlChunkDataset),
    'sequence_per_line': (SequencePerLineDatasetConfig, SequencePerLineDataset),
    'huggingface': (HuggingFaceDatasetConfig, HuggingFaceDataset),
}


def get_dataset(kwargs: dict[str, Any]) -> Dataset:
    """Build the dataset named by ``kwargs['name']``; ``ValueError`` on unknown names."""
    return build_from_strategies('dataset', STRATEGIES, kwargs)
