"""In-memory dataset, tokenising collator and the DataLoader recipe shared by all readers.

The collator call is the one at distllm/embed/datasets/utils.py:43-50: pad to the longest
sequence of the batch, truncate to ``tokenizer.model_max_length``, return int64 torch tensors.
"""

from __future__ import annotations

from typing import Any

from torch.utils.data import DataLoader
from torch.utils.data import Dataset
from transformers import BatchEncoding
from transformers import PreTrainedTokenizer


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
    """Tokenises a list of strings into one padded batch."""

    def __init__(self, tokenizer: PreTrainedTokenizer) -> None:
        self.tokenizer = tokenizer

    def __call__(self, batch: list[str]) -> BatchEncoding:
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
