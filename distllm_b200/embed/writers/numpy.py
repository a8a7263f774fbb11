"""``.npy`` writer: embeddings.npy, text.npy, metadata.npy (pickled dicts), as
distllm/embed/writers/numpy.py:27-69."""

from __future__ import annotations

from pathlib import Path
from typing import Literal

import numpy as np

from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.utils import BaseConfig

_FILES = ('embeddings.npy', 'text.npy', 'metadata.npy')


class NumpyWriterConfig(BaseConfig):
    name: Literal['numpy'] = 'numpy'  # type: ignore[assignment]


class NumpyWriter:
    def __init__(self, config: NumpyWriterConfig) -> None:
        self.config = config

    def write(self, output_dir: Path, result: EmbedderResult) -> None:
        np.save(output_dir / 'embeddings.npy', result.embeddings)
        np.save(output_dir / 'text.npy', result.text)
        if result.metadata is not None:
            np.save(output_dir / 'metadata.npy', result.metadata, allow_pickle=True)

    def merge(self, dataset_dirs: list[Path], output_dir: Path) -> None:
        for fname in _FILES:
            paths = [d / fname for d in dataset_dirs]
            if fname == 'metadata.npy' and not all(p.exists() for p in paths):
                continue
            parts = [np.load(p, allow_pickle=(fname == 'metadata.npy')) for p in paths]
            np.save(output_dir / fname, np.concatenate(parts))
