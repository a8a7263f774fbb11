"""CPU restatement of the reference poolers and the batch loop.  TEST INFRASTRUCTURE ONLY.

``average_pool``       distllm/embed/poolers/mean.py:13-49
``last_token_pool``    distllm/embed/poolers/last_token.py:12-39
``compute_embeddings`` distllm/embed/embedders/full_sequence.py:20-80 (pre-tokenised batches)

The restatement spells the mean pooler's advanced-indexing quirk out as explicit loops so it cannot
silently inherit the same torch behaviour it is meant to pin.
"""

from __future__ import annotations

from typing import Callable
from typing import Iterable

import numpy as np
import torch


def average_pool(embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """mean.py:32-49.  Edits ``attention_mask`` in place exactly as the reference does."""
    b, s = attention_mask.shape
    seq_lengths = [int(attention_mask[i].sum()) for i in range(b)]      # :32
    for i in range(b):                                                   # :35
        attention_mask[i, 0] = 0
    for length in seq_lengths:                                           # :36 -- column length-1 of
        col = length - 1                                                 # EVERY row, for every
        if col < 0:                                                      # sequence of the batch;
            col += s                                                     # index -1 wraps
        for i in range(b):
            attention_mask[i, col] = 0
    # :39-49: the product/sum stay in the embedding dtype, the divisor is fp32
    weights = attention_mask.to(embeddings.dtype).unsqueeze(-1)
    summed = (embeddings * weights).sum(1)
    counts = attention_mask.sum(1, keepdim=True).to(torch.float32).clamp(min=1e-9)
    return summed / counts


def last_token_pool(last_hidden_states: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """last_token.py:30-39."""
    b, s = attention_mask.shape
    if int(attention_mask[:, -1].sum()) == b:
        return last_hidden_states[:, -1]
    rows = []
    for i in range(b):
        col = int(attention_mask[i].sum()) - 1
        rows.append(last_hidden_states[i, col])
    return torch.stack(rows)


def normalize(pooled: torch.Tensor) -> torch.Tensor:
    """F.normalize(p=2, dim=-1) (full_sequence.py:68-69): x / max(||x||, 1e-12)."""
    norm = pooled.pow(2).sum(-1, keepdim=True).sqrt().clamp(min=1e-12)
    return pooled / norm


def compute_embeddings(
    batches: Iterable[dict[str, torch.Tensor]],
    encode: Callable[[dict[str, torch.Tensor]], torch.Tensor],
    pool: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    do_normalize: bool = False,
) -> np.ndarray:
    """The loop of full_sequence.py:57-78 over already-tokenised batches, in order."""
    out = []
    for batch in batches:
        hidden = encode(batch)
        pooled = pool(hidden, batch['attention_mask'])
        if do_normalize:
            pooled = normalize(pooled)
        out.append(pooled.to(torch.float32))
    return torch.cat(out).numpy()
