"""Config base class and small helpers.

Mirrors the surface of distllm/utils.py:20-112 (``BaseConfig`` with JSON/YAML round-trips and a
``name`` discriminator, ``batch_data``) so reference YAML files load unchanged.
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Literal
from typing import Sequence
from typing import TypeVar
from typing import Union

import yaml
from pydantic import BaseModel

PathLike = Union[str, Path]
T = TypeVar('T')
C = TypeVar('C', bound='BaseConfig')


class BaseConfig(BaseModel):
    """Pydantic model with file round-trips; subclasses pin ``name`` to a Literal."""

    name: Literal[''] = ''

    # -- JSON
    def write_json(self, path: PathLike) -> None:
        Path(path).write_text(json.dumps(self.model_dump(mode='json'), indent=2))

    @classmethod
    def from_json(cls: type[C], path: PathLike) -> C:
        return cls(**json.loads(Path(path).read_text()))

    # -- YAML (dumped through JSON so Paths/enums become plain scalars)
    def write_yaml(self, path: PathLike) -> None:
        plain = json.loads(self.model_dump_json())
        with open(path, 'w') as handle:
            yaml.dump(plain, handle, indent=4, sort_keys=False)

    @classmethod
    def from_yaml(cls: type[C], path: PathLike) -> C:
        with open(path) as handle:
            return cls(**yaml.safe_load(handle))


def batch_data(data: Sequence[T], chunk_size: int) -> list[list[T]]:
    """Split ``data`` into consecutive lists of ``chunk_size`` (the last may be shorter)."""
    if chunk_size <= 0:
        raise ValueError('chunk_size must be positive')
    return [list(data[i : i + chunk_size]) for i in range(0, len(data), chunk_size)]
