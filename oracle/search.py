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


# ------------------------------------------------------------------------------- ubinary + rescore
# distllm/rag/search.py:34-56 (quantize_dataset -> sentence_transformers.quantization.quantize_embeddings with
# precision='ubinary'), :202-260 (faiss.IndexBinaryFlat over the packed bits) and :280-336 (search through
# semantic_search_faiss with rescore=True, rescore_multiplier).  sentence-transformers (pin >=3.3.1,
# pyproject.toml) and faiss are absent from this image and from /root/reference; what follows restates
# their published behaviour -- PARITY UNPINNED for this branch (no way to run either here):
#
#   quantize_embeddings(x, 'ubinary')   np.packbits(x > 0).reshape(len(x), -1): bit j of a row is
#                                       x[j] > 0, eight dimensions per byte, first dimension in the MSB
#   IndexBinaryFlat.search(q, K)        the K rows with the smallest Hamming distance to the packed query,
#                                       ascending distance; a later row replaces the current worst only when
#                                       STRICTLY closer, so among equal distances the smaller ids stay
#   semantic_search_faiss(rescore=True) the float query is packed for the search, K = top_k *
#                                       rescore_multiplier candidates are fetched, each candidate's bits are
#                                       unpacked to {0, 1} and scored  score = sum_j q[j] * bit[j]  with the
#                                       FLOAT query; the top_k by descending score are returned


def quantize_ubinary(embeddings: np.ndarray) -> np.ndarray:
    """[N, H] float -> [N, H/8] uint8 (H % 8 == 0)."""
    return np.packbits(embeddings > 0).reshape(embeddings.shape[0], -1)


def hamming_topk(query_bits: np.ndarray, corpus_bits: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """(distances [Q,k'], indices [Q,k']), k' = min(k, N): ascending distance, ties by ascending id."""
    table = np.array([bin(i).count('1') for i in range(256)], dtype=np.int32)
    out_d, out_i = [], []
    k = min(k, corpus_bits.shape[0])
    for q in query_bits:
        d = table[np.bitwise_xor(corpus_bits, q[None, :])].sum(axis=1)
        order = np.lexsort((np.arange(len(d)), d))[:k]
        out_d.append(d[order])
        out_i.append(order)
    return np.stack(out_d), np.stack(out_i)


def search_ubinary(queries: np.ndarray, corpus_bits: np.ndarray, top_k: int,
                   rescore_multiplier: int = 2) -> tuple[np.ndarray, np.ndarray]:
    """(scores [Q,k'], indices [Q,k']) of the rescored binary search, descending score; ties keep the
    candidate order (ascending Hamming distance, then id)."""
    _, cand = hamming_topk(quantize_ubinary(queries), corpus_bits, top_k * rescore_multiplier)
    scores, indices = [], []
    for q, ids in zip(queries.astype(np.float32), cand):
        bits = np.unpackbits(corpus_bits[ids], axis=-1).astype(np.float32)
        s = (bits * q[None, :]).sum(axis=1, dtype=np.float64)
        order = np.argsort(-s, kind='stable')[:top_k]
        scores.append(s[order].astype(np.float32))
        indices.append(ids[order])
    return np.stack(scores), np.stack(indices)
