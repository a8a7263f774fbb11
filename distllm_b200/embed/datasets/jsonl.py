"""One text per jsonl line (distllm/embed/datasets/jsonl.py:17-73)."""

from __future__ import annotations

import json
from pathlib import Path
from typing import Literal

from torch.utils.data import DataLoader

from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.datasets.utils import make_dataloader
from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.utils import BaseConfig


class JsonlDatasetConfig(BaseConfig):
    name: Literal['jsonl'] = 'jsonl'  # type: ignore[assignment]
    # The name of the text field in the jsonl file
    text_field: str = 'text'
    # Number of data workers for batching.
    num_data_workers: int = 4
    # Inference batch size.
    batch_size: int = 8
    # Whether to pin memory for the dataloader.
    pin_memory: bool = True


def read_jsonl(path: Path) -> list[dict]:
    return [json.loads(line) for line in path.read_text().strip().split('\n')]


class JsonlDataset:
    def __init__(self, config: JsonlDatasetConfig):
        self.config = config

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader:
        texts = [row[self.config.text_field] for row in read_jsonl(data_file)]
        return make_dataloader(self.config, InMemoryDataset(texts), encoder.tokenizer)
