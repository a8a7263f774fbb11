"""One sequence per line after ``header_lines`` (distllm/embed/datasets/single_line.py:16-68)."""

from __future__ import annotations

from pathlib import Path
from typing import Literal

from torch.utils.data import DataLoader

from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.datasets.utils import LoaderConfig
from distllm_b200.embed.datasets.utils import make_dataloader
from distllm_b200.embed.encoders.base import Encoder


class SequencePerLineDatasetConfig(LoaderConfig):
    name: Literal['sequence_per_line'] = 'sequence_per_line'  # type: ignore[assignment]
    header_lines: int = 1   # lines to skip at the top of the file


class SequencePerLineDataset:
    def __init__(self, config: SequencePerLineDatasetConfig):
        self.config = config

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader:
        lines = data_file.read_text().splitlines()[self.config.header_lines :]
        return make_dataloader(self.config, InMemoryDataset(lines), encoder.tokenizer)
