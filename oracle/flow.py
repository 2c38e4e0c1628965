"""Oracle (test infrastructure): the two-stage quantised retrieval control flow.

Follows reference radiant/storage/redis_store.py:757-861 (canonical) and
radiant/storage/chroma_store.py:563-691 (the one backend whose stage 1 is a
binary search):

    candidate_k = int(top_k * rescore_multiplier) if use_rescoring else top_k
    stage 1     : top candidate_k by Hamming distance (min_similarity ignored,
                  language/doc_level filter applied)           [R4, R12]
    no rescoring: candidates[:top_k] with placeholder score 1.0 (chroma_store.py:624-631)
    stage 2     : int8 rows preferred, float32 fallback; rescore_candidates [R3];
                  rescored[:top_k] kept where score >= min_similarity
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .hamming import hamming_topk
from .quantize import quantize_ubinary
from .rescore import rescore_f32


def two_stage_search(
    queries: np.ndarray,
    codes: np.ndarray,
    rescore_rows: np.ndarray,
    top_k: int,
    rescore_multiplier: float = 4.0,
    use_rescoring: bool = True,
    min_similarity: float = 0.0,
    valid: Optional[np.ndarray] = None,
    exact: bool = False,
) -> List[Tuple[np.ndarray, np.ndarray]]:
    """queries f32 [Q, D]; codes u8 [N, D/8]; rescore_rows i8|f32 [N, D].

    Returns one (rows int64, scores f32) pair per query.
    """
    queries = np.asarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries[None, :]
    candidate_k = int(top_k * rescore_multiplier) if use_rescoring else top_k
    qcodes = quantize_ubinary(queries)
    _dist, cand = hamming_topk(codes, qcodes, candidate_k, valid=valid)
    out = []
    for qi in range(queries.shape[0]):
        ids = cand[qi]
        ids = ids[ids >= 0]
        if not use_rescoring:
            ids = ids[:top_k]
            out.append((ids.astype(np.int64), np.ones(ids.size, dtype=np.float32)))
            continue
        if ids.size == 0:
            out.append((np.empty(0, np.int64), np.empty(0, np.float32)))
            continue
        out.append(
            rescore_f32(queries[qi], rescore_rows[ids], ids, top_k=top_k,
                        min_similarity=min_similarity, exact=exact)
        )
    return out
