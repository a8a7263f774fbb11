"""Re-export: the interfaces live in distllm_b200/embed/protocols.py."""

from distllm_b200.embed.protocols import Encoder

__all__ = ['Encoder']
