"""HuggingFace ``datasets`` writer: columns ``text``, ``embeddings`` + one column per metadata key
(the on-disk schema of distllm/embed/writers/huggingface.py:19-92 that the RAG index reads).

The table is assembled column-wise (``Dataset.from_dict``) instead of one Python dict per row.
"""

from __future__ import annotations

from pathlib import Path
from typing import Literal
from typing import Optional

from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.utils import BaseConfig


class HuggingFaceWriterConfig(BaseConfig):
    name: Literal['huggingface'] = 'huggingface'  # type: ignore[assignment]
    # The number of processes to use for writing the dataset
    num_proc: Optional[int] = None  # noqa: UP007


class HuggingFaceWriter:
    def __init__(self, config: HuggingFaceWriterConfig) -> None:
        self.config = config

    def write(self, output_dir: Path, result: EmbedderResult) -> None:
        from datasets import Dataset

        columns: dict[str, list] = {
            'text': list(result.text),
            'embeddings': list(result.embeddings),
        }
        if result.metadata is not None:
            keys: list[str] = []
            for row in result.metadata:
                keys.extend(k for k in row if k not in keys)
            for key in keys:
                columns[key] = [row.get(key) for row in result.metadata]
        Dataset.from_dict(columns).save_to_disk(output_dir)

    def merge(self, dataset_dirs: list[Path], output_dir: Path) -> None:
        from datasets import Dataset
        from datasets import concatenate_datasets

        merged = concatenate_datasets([Dataset.load_from_disk(p) for p in dataset_dirs])
        merged.save_to_disk(output_dir, num_proc=self.config.num_proc)
