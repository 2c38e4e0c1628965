"""Oracle (test infrastructure): Reciprocal Rank Fusion, row R10 of SURVEY.md 8(a).

Follows reference radiant/agents/fusion.py:61-102: for each run, for each 1-based
rank, ``score[id] = score.get(id, 0.0) + 1.0 / (rrf_k + rank)`` in float64, ids kept
in first-insertion order, Python's stable ``sort(reverse=True)`` (equal scores keep
first-insertion order), cut ``[:top_k]``.  A doc appearing twice in one run is
counted twice.  Pinned by tests/golden/rrf_cases.json (reference RRFAgent output).
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def rrf_fuse(runs: Sequence[Sequence[int]], top_k: int, rrf_k: float = 60) -> Tuple[np.ndarray, np.ndarray]:
    """runs: lists of integer doc ids (best first).  -> (ids int64 [M], scores f64 [M])."""
    score = {}
    for run in runs:
        for rank, doc in enumerate(run, start=1):
            doc = int(doc)
            score[doc] = score.get(doc, 0.0) + (1.0 / (rrf_k + rank))
    fused: List[Tuple[int, float]] = list(score.items())  # first-insertion order
    fused.sort(key=lambda x: x[1], reverse=True)  # stable
    fused = fused[:top_k]
    return (
        np.asarray([d for d, _ in fused], dtype=np.int64),
        np.asarray([s for _, s in fused], dtype=np.float64),
    )
