"""Writer protocol (distllm/embed/writers/base.py:12-41)."""

from __future__ import annotations

from pathlib import Path
from typing import Protocol

from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.utils import BaseConfig


class Writer(Protocol):
    def __init__(self, config: BaseConfig) -> None: ...

    def write(self, output_dir: Path, result: EmbedderResult) -> None: ...

    def merge(self, dataset_dirs: list[Path], output_dir: Path) -> None: ...
