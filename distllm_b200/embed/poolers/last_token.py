"""Last-token pooler (distllm/embed/poolers/last_token.py:12-39) on the native gather kernel.

The reference decides between "left padded / full length -> column S-1" and
"row b -> column len_b - 1" with a host-side ``if`` on a device reduction (a stream sync per
batch); here the decision is taken on the device.
"""

from __future__ import annotations

from typing import Literal

import torch

from distllm_b200 import _native
from distllm_b200.utils import BaseConfig


def last_token_pool(last_hidden_states: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Hidden state of each sequence's final attended token, in the dtype of the input."""
    mask = attention_mask.to(torch.int64).contiguous()
    out = _native.pool_last_token(last_hidden_states.contiguous(), mask)
    return out.to(last_hidden_states.dtype)


class LastTokenPoolerConfig(BaseConfig):
    """Configuration for the LastTokenPooler."""

    name: Literal['last_token'] = 'last_token'  # type: ignore[assignment]


class LastTokenPooler:
    """Uses the final attended token's hidden state as the sequence embedding."""

    native_pool_kind = _native.POOL_LAST_TOKEN

    def __init__(self, config: LastTokenPoolerConfig) -> None:
        self.config = config

    def pool(self, embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        return last_token_pool(embeddings, attention_mask)
