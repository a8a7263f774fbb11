"""In-memory dataset, tokenising collator and the DataLoader recipe shared by all readers.

The collator call is the one at distllm/embed/datasets/utils.py:43-50: pad to the longest
sequence of the batch, truncate to ``tokenizer.model_max_length``, return int64 torch tensors.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch
from torch.utils.data import DataLoader
from torch.utils.data import Dataset
from transformers import BatchEncoding
from transformers import PreTrainedTokenizer

from distllm_b200.utils import BaseConfig


class LoaderConfig(BaseConfig):
    """DataLoader knobs every dataset config of the reference carries, with its defaults."""

    num_data_workers: int = 4   # DataLoader worker processes (0: tokenise on the main process)
    batch_size: int = 8         # inference batch size
    pin_memory: bool = True     # page-locked batches for the H2D copy


class InMemoryDataset(Dataset):
    """List of texts with optional per-row metadata."""

    def __init__(self, data: list[str], metadata: list[dict[str, Any]] | None = None) -> None:
        if metadata is not None and len(metadata) != len(data):
            raise AssertionError('metadata and data must have the same length')
        self.data = data
        self.metadata = metadata

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, idx: int) -> str:
        return self.data[idx]


class DataCollator:
    """Tokenises a list of strings into one padded batch.

    Output is, tensor for tensor, what the reference's collator call returns
    (``tokenizer(batch, padding=True, truncation=True, return_tensors='pt')``,
    distllm/embed/datasets/utils.py:43-50).  For fast (Rust-backed) tokenizers the batch is built
    straight from the backend's encodings and padded with numpy: the generic HF path spends ~8x the
    tokenisation time on backend-side padding and on turning every id into a Python object and back into a
    tensor (685 ms vs 80 ms for 512 buffers of 512 tokens on 8 cores), which at >= 10k chunks/s per GPU makes the host feed the end-to-end limit
    (SURVEY 8(f) rank 1).  Slow tokenizers (e.g. ``EsmTokenizer``) take the reference call unchanged.
    """

    def __init__(self, tokenizer: PreTrainedTokenizer, fast: bool = True) -> None:
        self.tokenizer = tokenizer
        self._fast = bool(
            fast
            and getattr(tokenizer, 'is_fast', False)
            and hasattr(tokenizer, '_tokenizer')
            and hasattr(tokenizer, 'set_truncation_and_padding')
            and hasattr(tokenizer, '_get_padding_truncation_strategies'),
        )

    def _collate_fast(self, batch: list[str]) -> BatchEncoding:
        from transformers.utils import PaddingStrategy

        tok = self.tokenizer
        _, truncation, max_length, _ = tok._get_padding_truncation_strategies(
            padding=True, truncation=True)
        # truncate in the backend, pad here: the backend's own padding (it also pads offsets, token
        # strings, word ids ...) costs 3x the tokenisation itself
        tok.set_truncation_and_padding(
            padding_strategy=PaddingStrategy.DO_NOT_PAD, truncation_strategy=truncation,
            max_length=max_length, stride=0, pad_to_multiple_of=None, padding_side=None)
        backend = tok._tokenizer
        encode = getattr(backend, 'encode_batch_fast', backend.encode_batch)  # no offset tracking
        encodings = encode(list(batch), add_special_tokens=True)
        rows = [e.ids for e in encodings]
        lengths = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
        width = int(lengths.max())
        left = tok.padding_side == 'left'
        names = tok.model_input_names
        want_types = 'token_type_ids' in names

        ids = np.full((len(rows), width), tok.pad_token_id, dtype=np.int64)
        types = np.full((len(rows), width), tok.pad_token_type_id, dtype=np.int64) if want_types else None
        for i, (row, n) in enumerate(zip(rows, lengths)):
            span = slice(width - n, width) if left else slice(0, n)
            ids[i, span] = row
            if want_types:
                types[i, span] = encodings[i].type_ids
        cols = np.arange(width)[None, :]
        mask = (cols >= (width - lengths)[:, None]) if left else (cols < lengths[:, None])

        data = {'input_ids': torch.from_numpy(ids)}
        if want_types:
            data['token_type_ids'] = torch.from_numpy(types)
        if 'attention_mask' in names:
            data['attention_mask'] = torch.from_numpy(mask.astype(np.int64))
        return BatchEncoding(data)

    def __call__(self, batch: list[str]) -> BatchEncoding:
        if self._fast and len(batch) > 0 and self.tokenizer.pad_token_id is not None:
            return self._collate_fast(batch)
        return self.tokenizer(batch, padding=True, truncation=True, return_tensors='pt')


def make_dataloader(config: Any, dataset: InMemoryDataset, tokenizer: PreTrainedTokenizer) -> DataLoader:
    """DataLoader with the knobs every dataset config carries (batch_size, workers, pinning)."""
    return DataLoader(
        dataset=dataset,
        batch_size=config.batch_size,
        num_workers=config.num_data_workers,
        pin_memory=config.pin_memory,
        collate_fn=DataCollator(tokenizer),
    )
