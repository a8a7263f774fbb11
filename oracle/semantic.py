"""CPU restatement of the semantic splitter arithmetic.  TEST INFRASTRUCTURE ONLY.

``calculate_distances_between_buffer`` distllm/embed/embedders/semantic_chunk.py:24-55
``build_chunks``                       distllm/embed/embedders/semantic_chunk.py:58-102
``split_rows``                         the per-document driver, semantic_chunk.py:150-182
"""

from __future__ import annotations

import numpy as np


def calculate_distances_between_buffer(buffer_embeds: np.ndarray) -> np.ndarray:
    emb = buffer_embeds.astype(np.float32)                       # :41
    out = np.zeros(max(len(emb) - 1, 0))                          # :44 float64 container
    for i in range(len(emb) - 1):                                 # :45-53, fp32 arithmetic
        a, b = emb[i], emb[i + 1]
        dot = np.float32(0)
        na = np.float32(0)
        nb = np.float32(0)
        dot = np.dot(a, b)
        na = np.sqrt(np.dot(a, a))
        nb = np.sqrt(np.dot(b, b))
        out[i] = 1 - dot / (na * nb)
    return out


def build_chunks(distances: np.ndarray, breakpoint_percentile_threshold: int) -> list[tuple[int, int]]:
    if len(distances) == 0:                                       # :80-81
        return [(0, 0)]
    threshold = np.percentile(distances, breakpoint_percentile_threshold)   # :83-86
    groups = []
    start = 0
    for i, x in enumerate(distances):                             # :88-97, strict >
        if x > threshold:
            groups.append((start, i + 1))
            start = i + 1
    groups.append((start, len(distances) + 1))                    # :100
    return groups


def split_rows(embeddings: np.ndarray, doc_ranges: list[tuple[int, int]], percentile: int) -> list[tuple[int, int]]:
    rows = []
    for lo, hi in doc_ranges:
        d = calculate_distances_between_buffer(embeddings[lo:hi])
        rows.extend((lo + s, lo + e) for s, e in build_chunks(d, percentile))
    return rows
