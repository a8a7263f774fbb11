"""Re-export: the interfaces live in distllm_b200/embed/protocols.py."""

from distllm_b200.embed.protocols import Dataset

__all__ = ['Dataset']
