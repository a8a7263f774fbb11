"""Writer family: ``huggingface`` and ``numpy``."""

from __future__ import annotations

from typing import Any
from typing import Union

from distllm_b200.embed._factory import build_from_strategies
from distllm_b200.embed.writers.base import Writer
from distllm_b200.embed.writers.huggingface import HuggingFaceWriter
from distllm_b200.embed.writers.huggingface import HuggingFaceWriterConfig
from distllm_b200.embed.writers.numpy import NumpyWriter
from distllm_b200.embed.writers.numpy import NumpyWriterConfig
from distllm_b200.utils import BaseConfig

WriterConfigs = Union[HuggingFaceWriterConfig, NumpyWriterConfig]

STRATEGIES: dict[str, tuple[type[BaseConfig], type[Writer]]] = {
    'huggingface': (HuggingFaceWriterConfig, HuggingFaceWriter),
    'numpy': (NumpyWriterConfig, NumpyWriter),
}


def get_writer(kwargs: dict[str, Any]) -> Writer:
    """Build the writer named by ``kwargs['name']``; ``ValueError`` on unknown names."""
    return build_from_strategies('writer', STRATEGIES, kwargs)
