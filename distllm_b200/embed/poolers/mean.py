"""Mean pooler backed by the native masked-mean kernel.

Semantics are those of distllm/embed/poolers/mean.py:13-49, reproduced exactly -- including the
in-place edit of ``attention_mask`` and the cross-row quirk of mean.py:36
(``attention_mask[:, seq_lengths - 1] = 0`` clears column ``len_j - 1`` of *every* row for every
sequence ``j`` of the batch), which makes results depend on batch composition.
"""

from __future__ import annotations

from typing import Literal

import torch

from distllm_b200 import _native
from distllm_b200.utils import BaseConfig


def average_pool(embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Masked mean of ``embeddings [B,S,H]`` -> fp32 ``[B,H]``; edits ``attention_mask`` in place.

    Runs ``b2e_pool_mean`` (one pass over the hidden state).  CUDA tensors only.
    """
    mask = attention_mask
    if mask.dtype != torch.int64 or not mask.is_contiguous():
        # keep the reference's visible side effect on the caller's tensor
        work = mask.to(torch.int64).contiguous()
        out = _native.pool_mean(embeddings.contiguous(), work, _native.POOL_MEAN_REF, True)
        mask.copy_(work.to(mask.dtype))
        return out
    return _native.pool_mean(embeddings.contiguous(), mask, _native.POOL_MEAN_REF, True)


class MeanPoolerConfig(BaseConfig):
    """Configuration for the MeanPooler."""

    name: Literal['mean'] = 'mean'  # type: ignore[assignment]


class MeanPooler:
    """Averages hidden states over the attended tokens, without start/end tokens."""

    #: tells the native embedder which fused epilogue implements this pooler
    native_pool_kind = _native.POOL_MEAN_REF

    def __init__(self, config: BaseConfig) -> None:
        self.config = config

    def pool(self, embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        return average_pool(embeddings, attention_mask)
