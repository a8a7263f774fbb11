"""Embedder family: ``full_sequence`` and ``semantic_chunk``."""

from __future__ import annotations

from typing import Any
from typing import Union

from distllm_b200.embed._factory import build_from_strategies
from distllm_b200.embed.embedders.base import Embedder
from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.embed.embedders.full_sequence import FullSequenceEmbedder
from distllm_b200.embed.embedders.full_sequence import FullSequenceEmbedderConfig
from distllm_b200.embed.embedders.semantic_chunk import SemanticChunkEmbedder
from distllm_b200.embed.embedders.semantic_chunk import SemanticChunkEmbedderConfig
from distllm_b200.utils import BaseConfig

EmbedderConfigs = Union[FullSequenceEmbedderConfig, SemanticChunkEmbedderConfig]

STRATEGIES: dict[str, tuple[type[BaseConfig], type[Embedder]]] = {
    'full_sequence': (FullSequenceEmbedderConfig, FullSequenceEmbedder),
    'semantic_chunk': (SemanticChunkEmbedderConfig, SemanticChunkEmbedder),
}


def get_embedder(kwargs: dict[str, Any]) -> Embedder:
    """Build the embedder named by ``kwargs['name']``; ``ValueError`` on unknown names."""
    return build_from_strategies('embedder', STRATEGIES, kwargs)
