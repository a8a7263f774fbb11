"""Pooler protocol (structural, as in distllm/embed/poolers/base.py:12-42)."""

from __future__ import annotations

from typing import Protocol

import torch

from distllm_b200.utils import BaseConfig


class Pooler(Protocol):
    """Reduces hidden states ``[B, S, H]`` to one vector per sequence ``[B, H]``."""

    def __init__(self, config: BaseConfig) -> None: ...

    def pool(self, embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor: ...
