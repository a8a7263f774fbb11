"""One text per jsonl line (distllm/embed/datasets/jsonl.py:17-73)."""

from __future__ import annotations

import json
from pathlib import Path
from typing import Literal

from torch.utils.data import DataLoader

from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.datasets.utils import LoaderConfig
from distllm_b200.embed.datasets.utils import make_dataloader
from distllm_b200.embed.encoders.base import Encoder


class JsonlDatasetConfig(LoaderConfig):
    name: Literal['jsonl'] = 'jsonl'  # type: ignore[assignment]
    text_field: str = 'text'   # which key of each json row holds the text


def read_jsonl(path: Path) -> list[dict]:
    return [json.loads(line) for line in path.read_text().strip().split('\n')]


class JsonlDataset:
    def __init__(self, config: JsonlDatasetConfig):
        self.config = config

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader:
        texts = [row[self.config.text_field] for row in read_jsonl(data_file)]
        return make_dataloader(self.config, InMemoryDataset(texts), encoder.tokenizer)
