"""Embedder protocol and result record (distllm/embed/embedders/base.py:17-58)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any
from typing import Protocol

import numpy as np
from torch.utils.data import DataLoader

from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.embed.poolers.base import Pooler
from distllm_b200.utils import BaseConfig


@dataclass
class EmbedderResult:
    """What a writer consumes: ``embeddings [N,H]``, the ``text`` per row, optional ``metadata``."""

    embeddings: np.ndarray
    text: list[str]
    metadata: list[dict[str, Any]] | None = None


class Embedder(Protocol):
    def __init__(self, config: BaseConfig) -> None: ...

    def embed(self, dataloader: DataLoader, encoder: Encoder, pooler: Pooler) -> EmbedderResult: ...
