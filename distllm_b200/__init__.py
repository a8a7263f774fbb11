"""distllm_b200: B200-native (sm_100a) implementation of distllm's embedding hot path."""

from __future__ import annotations

__version__ = '0.1.0'
