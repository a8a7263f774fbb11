"""Encoder protocol (distllm/embed/encoders/base.py:14-55)."""

from __future__ import annotations

from typing import Protocol

import torch
from transformers import BatchEncoding
from transformers import PreTrainedTokenizer

from distllm_b200.utils import BaseConfig


class Encoder(Protocol):
    """Token ids -> hidden states ``[B, S, H]``."""

    def __init__(self, config: BaseConfig) -> None: ...

    @property
    def dtype(self) -> torch.dtype: ...

    @property
    def device(self) -> torch.device: ...

    @property
    def embedding_size(self) -> int: ...

    @property
    def tokenizer(self) -> PreTrainedTokenizer: ...

    def encode(self, batch_encoding: BatchEncoding) -> torch.Tensor: ...
