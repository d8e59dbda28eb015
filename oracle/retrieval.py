"""Oracle (test infrastructure): all-pairs scores + top-k of the reference.

Restates ``training/coarse.py:100-105,134-148``: the encodings are held in
float64 arrays (``np.zeros`` default dtype) carrying float32 values; per query
``scores = cell_encodings @ text_encodings[q]`` in float64, ``argsort(-scores)``,
first ``max(top_k)``.  numpy's default quicksort leaves the order of exactly
equal scores unspecified; the oracle (and the CUDA path) fix it to
(score descending, index ascending) = a stable sort.

PINNED: on tie-free data this equals the reference loop verbatim
(``tests/golden/make_golden.py`` runs that loop -> ``retrieval_*.npz``).
"""
from typing import Tuple

import numpy as np


def scores_f64(cell_enc: np.ndarray, text_enc: np.ndarray) -> np.ndarray:
    """[N,D], [Q,D] (float32 values) -> scores [Q,N] float64."""
    return np.asarray(text_enc, dtype=np.float64) @ np.asarray(cell_enc, dtype=np.float64).T


def topk(cell_enc: np.ndarray, text_enc: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """-> (idx [Q,k] int64, scores [Q,k] float64), ordered by (score desc, idx asc)."""
    s = scores_f64(cell_enc, text_enc)
    idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int64), np.take_along_axis(s, idx, axis=1)


def reference_loop(cell_enc: np.ndarray, text_enc: np.ndarray, k: int) -> np.ndarray:
    """The reference's per-query loop, verbatim in structure (training/coarse.py:134-140)."""
    cell_encodings = np.zeros(cell_enc.shape)
    cell_encodings[:] = cell_enc
    text_encodings = np.zeros(text_enc.shape)
    text_encodings[:] = text_enc
    out = np.zeros((len(text_encodings), k), dtype=np.int64)
    for query_idx in range(len(text_encodings)):
        scores = cell_encodings[:] @ text_encodings[query_idx]
        sorted_indices = np.argsort(-1.0 * scores)
        out[query_idx] = sorted_indices[0:k]
    return out


def merge_shards(shard_idx, shard_scores, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard top-k lists (global indices) into a global top-k, same ordering rule.

    shard_idx / shard_scores: sequences of [Q, k_r] arrays.
    """
    idx = np.concatenate(shard_idx, axis=1)
    sc = np.concatenate(shard_scores, axis=1).astype(np.float64)
    order = np.lexsort((idx, -sc), axis=1)[:, :k]
    return np.take_along_axis(idx, order, axis=1), np.take_along_axis(sc, order, axis=1)
