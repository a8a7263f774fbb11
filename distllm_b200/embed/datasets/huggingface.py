"""Rows of a HuggingFace ``datasets`` directory saved with ``save_to_disk``
(distllm/embed/datasets/huggingface.py:18-83): one text column, optional metadata columns."""

from __future__ import annotations

from pathlib import Path
from typing import Literal

from pydantic import Field
from torch.utils.data import DataLoader

from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.datasets.utils import LoaderConfig
from distllm_b200.embed.datasets.utils import make_dataloader
from distllm_b200.embed.encoders.base import Encoder


class HuggingFaceDatasetConfig(LoaderConfig):
    name: Literal['huggingface'] = 'huggingface'  # type: ignore[assignment]
    text_field: str = 'text'   # column that holds the text
    metadata_fields: list[str] = Field(default_factory=list)   # columns copied into the metadata


class HuggingFaceDataset:
    def __init__(self, config: HuggingFaceDatasetConfig) -> None:
        self.config = config

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader:
        from datasets import Dataset

        table = Dataset.load_from_disk(str(data_file))
        texts: list[str] = table[self.config.text_field]
        metadata = None
        if self.config.metadata_fields:
            # column-wise read, then one dict per row (the reference iterates rows of the selection)
            cols = {name: table[name] for name in self.config.metadata_fields}
            metadata = [{name: cols[name][i] for name in self.config.metadata_fields}
                        for i in range(len(texts))]
        return make_dataloader(self.config, InMemoryDataset(list(texts), metadata), encoder.tokenizer)
