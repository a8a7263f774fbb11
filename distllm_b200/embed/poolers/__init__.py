"""Pooler family: ``mean`` and ``last_token`` (same names as distllm/embed/poolers/__init__.py)."""

from __future__ import annotations

from typing import Any
from typing import Union

from distllm_b200.embed._factory import build_from_strategies
from distllm_b200.embed.poolers.base import Pooler
from distllm_b200.embed.poolers.last_token import LastTokenPooler
from distllm_b200.embed.poolers.last_token import LastTokenPoolerConfig
from distllm_b200.embed.poolers.mean import MeanPooler
from distllm_b200.embed.poolers.mean import MeanPoolerConfig
from distllm_b200.utils import BaseConfig

PoolerConfigs = Union[MeanPoolerConfig, LastTokenPoolerConfig]

STRATEGIES: dict[str, tuple[type[BaseConfig], type[Pooler]]] = {
    'mean': (MeanPoolerConfig, MeanPooler),
    'last_token': (LastTokenPoolerConfig, LastTokenPooler),
}


def get_pooler(kwargs: dict[str, Any]) -> Pooler:
    """Build the pooler named by ``kwargs['name']``; ``ValueError`` on unknown names."""
    return build_from_strategies('pooler', STRATEGIES, kwargs)
