"""Plugin families of the embedding path (same import surface as ``distllm.embed``)."""

from __future__ import annotations

from distllm_b200.embed.datasets import Dataset
from distllm_b200.embed.datasets import DatasetConfigs
from distllm_b200.embed.datasets import get_dataset
from distllm_b200.embed.embedders import Embedder
from distllm_b200.embed.embedders import EmbedderConfigs
from distllm_b200.embed.embedders import EmbedderResult
from distllm_b200.embed.embedders import get_embedder
from distllm_b200.embed.encoders import Encoder
from distllm_b200.embed.encoders import EncoderConfigs
from distllm_b200.embed.encoders import get_encoder
from distllm_b200.embed.poolers import Pooler
from distllm_b200.embed.poolers import PoolerConfigs
from distllm_b200.embed.poolers import get_pooler
from distllm_b200.embed.writers import Writer
from distllm_b200.embed.writers import WriterConfigs
from distllm_b200.embed.writers import get_writer

__all__ = [
    'Dataset', 'DatasetConfigs', 'get_dataset',
    'Embedder', 'EmbedderConfigs', 'EmbedderResult', 'get_embedder',
    'Encoder', 'EncoderConfigs', 'get_encoder',
    'Pooler', 'PoolerConfigs', 'get_pooler',
    'Writer', 'WriterConfigs', 'get_writer',
]
