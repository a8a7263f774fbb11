"""FASTA reader (distllm/embed/datasets/fasta.py:19-115): upper-cased sequences, tag+path metadata."""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Literal
from typing import Union

from torch.utils.data import DataLoader

from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.datasets.utils import LoaderConfig
from distllm_b200.embed.datasets.utils import make_dataloader
from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.utils import BaseConfig

PathLike = Union[str, Path]


@dataclass
class Sequence:
    sequence: str
    tag: str


def read_fasta(fasta_file: PathLike) -> list[Sequence]:
    """Records start at a line beginning with '>'; the rest of that line is the tag and all
    following lines (newlines removed) are the sequence."""
    text = '\n' + Path(fasta_file).read_text()
    records = []
    for block in text.split('\n>')[1:]:
        tag, _, body = block.partition('\n')
        records.append(Sequence(sequence=body.replace('\n', ''), tag=tag))
    return records


def write_fasta(sequences: Sequence | list[Sequence], fasta_file: PathLike, mode: str = 'w') -> None:
    items = [sequences] if isinstance(sequences, Sequence) else sequences
    with open(fasta_file, mode) as handle:
        handle.writelines(f'>{s.tag}\n{s.sequence}\n' for s in items)


class FastaDatasetConfig(LoaderConfig):
    name: Literal['fasta'] = 'fasta'  # type: ignore[assignment]


class FastaDataset:
    def __init__(self, config: FastaDatasetConfig) -> None:
        self.config = config

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader:
        records = read_fasta(data_file)
        data = [r.sequence.upper() for r in records]
        metadata = [{'tags': r.tag, 'paths': str(data_file)} for r in records]
        return make_dataloader(self.config, InMemoryDataset(data, metadata), encoder.tokenizer)
