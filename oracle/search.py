"""CPU restatement of the exact float32 search the reference delegates to faiss.

distllm/rag/search.py:280-336 calls ``sentence_transformers.quantization.semantic_search_faiss`` with
``corpus_precision='float32'``, ``exact=True`` on a ``faiss.IndexFlatIP`` (search.py:230-233): scores are
fp32 inner products, the k largest per query are returned in descending order.  faiss (pin: none in
pyproject.toml; absent from this image) and sentence_transformers are third-party and not vendored, so
this restates IndexFlatIP's published behaviour; ``normalize_l2`` restates ``faiss.normalize_L2``
(search.py:274).  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import numpy as np


def normalize_l2(x: np.ndarray) -> np.ndarray:
    out = x.astype(np.float32).copy()
    norms = np.sqrt((out * out).sum(axis=1, keepdims=True))
    np.divide(out, norms, out=out, where=norms > 0)
    return out


def topk_inner_product(queries: np.ndarray, corpus: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """(scores [Q,k'], indices [Q,k']) with k' = min(k, N), descending scores, ties by ascending index."""
    scores = queries.astype(np.float64) @ corpus.astype(np.float64).T
    k = min(k, corpus.shape[0])
    order = np.lexsort((np.arange(corpus.shape[0])[None, :].repeat(len(queries), 0), -scores), axis=1)[:, :k]
    return np.take_along_axis(scores, order, axis=1).astype(np.float32), order
