"""Re-export: the interfaces live in distllm_b200/embed/protocols.py."""

from distllm_b200.embed.protocols import Embedder
from distllm_b200.embed.protocols import EmbedderResult

__all__ = ['Embedder', 'EmbedderResult']
