"""The five structural interfaces of the embedding path in one place.

distllm keeps one ``base.py`` per plugin family (encoders/base.py:14-55, poolers/base.py:12-42,
embedders/base.py:17-58, datasets/base.py:14-40, writers/base.py:12-41); the interfaces are duck-typed
``typing.Protocol`` classes, so anything with these members plugs in on either side.  The per-family
``base`` modules of this package re-export from here, which keeps the reference's import paths working.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any
from typing import Protocol

import numpy as np
import torch
from torch.utils.data import DataLoader
from transformers import BatchEncoding
from transformers import PreTrainedTokenizer

from distllm_b200.utils import BaseConfig


@dataclass
class EmbedderResult:
    """What an embedder hands to a writer: one embedding row, one text and (optionally) one metadata
    dict per output item."""

    embeddings: np.ndarray                       # [N, H]
    text: list[str]                              # N strings
    metadata: list[dict[str, Any]] | None = None
    # not in the reference (embedders/base.py:17-26): the same rows as a device-resident fp32 tensor, kept
    # by the native embedders so that the multi-GPU driver can all-gather them without a host round trip;
    # writers ignore it
    device_embeddings: torch.Tensor | None = None


class Encoder(Protocol):
    """Token batch in, last hidden state ``[B, S, H]`` out."""

    def __init__(self, config: BaseConfig) -> None: ...
    def encode(self, batch_encoding: BatchEncoding) -> torch.Tensor: ...
    @property
    def tokenizer(self) -> PreTrainedTokenizer: ...
    @property
    def embedding_size(self) -> int: ...
    @property
    def device(self) -> torch.device: ...
    @property
    def dtype(self) -> torch.dtype: ...


class Pooler(Protocol):
    """Hidden states ``[B, S, H]`` + mask ``[B, S]`` in, one vector per sequence ``[B, H]`` out; allowed
    to edit the mask in place (the reference's mean pooler does)."""

    def __init__(self, config: BaseConfig) -> None: ...
    def pool(self, embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor: ...


class Dataset(Protocol):
    """A file in, a DataLoader of tokenised batches out."""

    def __init__(self, config: BaseConfig) -> None: ...
    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader: ...


class Embedder(Protocol):
    """Drives encoder and pooler over a DataLoader."""

    def __init__(self, config: BaseConfig) -> None: ...
    def embed(self, dataloader: DataLoader, encoder: Encoder, pooler: Pooler) -> EmbedderResult: ...


class Writer(Protocol):
    """Persists an ``EmbedderResult``; ``merge`` concatenates the directories several workers wrote."""

    def __init__(self, config: BaseConfig) -> None: ...
    def write(self, output_dir: Path, result: EmbedderResult) -> None: ...
    def merge(self, dataset_dirs: list[Path], output_dir: Path) -> None: ...
