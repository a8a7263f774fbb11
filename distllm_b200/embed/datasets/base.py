"""Dataset protocol (distllm/embed/datasets/base.py:14-40)."""

from __future__ import annotations

from pathlib import Path
from typing import Protocol

from torch.utils.data import DataLoader

from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.utils import BaseConfig


class Dataset(Protocol):
    def __init__(self, config: BaseConfig) -> None: ...

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader: ...
