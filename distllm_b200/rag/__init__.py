"""Retrieval query path on the native kernels."""

from distllm_b200.rag.search import ExactIndex
from distllm_b200.rag.search import ExactIndexConfig
from distllm_b200.rag.search import Retriever
from distllm_b200.rag.search import RetrieverConfig

__all__ = ['ExactIndex', 'ExactIndexConfig', 'Retriever', 'RetrieverConfig']
