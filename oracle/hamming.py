"""Oracle (test infrastructure): exact Hamming top-k, row R4 of SURVEY.md 8(a).

PARITY UNPINNED: no reference code computes a Hamming distance (SURVEY.md
section 0.2).  The stage is *documented* in the reference at
docs/BINARY_QUANTIZATION_README.md:84-100 ("Quantize query to binary, search
with Hamming distance, retrieve 4x candidate documents") and sits at the call
sites radiant/storage/redis_store.py:799-809, radiant/storage/chroma_store.py:588-619,
radiant/storage/pgvector_store.py:794-802.  Restated as

    dist[n] = popcount(code[n] XOR qcode)         (integer, exact)
    order   = (dist ascending, row id ascending)  (canonical tie rule)
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def _as_u64_view(codes: np.ndarray) -> Tuple[np.ndarray, int]:
    """View u8 [N, B] as u64 words when B is a multiple of 8 (faster popcount)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    b = codes.shape[-1]
    if b % 8 == 0:
        return codes.view(np.uint64), 8
    if b % 4 == 0:
        return codes.view(np.uint32), 4
    return codes, 1


def hamming_distances(codes: np.ndarray, qcode: np.ndarray) -> np.ndarray:
    """u8 [N, B], u8 [B] -> int32 [N] Hamming distances."""
    cw, _ = _as_u64_view(codes)
    qw, _ = _as_u64_view(np.asarray(qcode, dtype=np.uint8)[None, :])
    return np.bitwise_count(cw ^ qw).sum(axis=1, dtype=np.int64).astype(np.int32)


def hamming_topk(
    codes: np.ndarray,
    qcodes: np.ndarray,
    k: int,
    valid: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k by (dist asc, row asc) for each query.

    codes  u8 [N, B]; qcodes u8 [Q, B]; valid optional bool [N] row predicate
    (the doc_level / language filter of reference redis_store.py:669-687).
    Returns (dist int32 [Q, k], row int64 [Q, k]); missing slots are
    (2**31 - 1, -1).
    """
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    qcodes = np.ascontiguousarray(qcodes, dtype=np.uint8)
    if qcodes.ndim == 1:
        qcodes = qcodes[None, :]
    n = codes.shape[0]
    nq = qcodes.shape[0]
    out_d = np.full((nq, k), np.iinfo(np.int32).max, dtype=np.int32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    rows_all = np.arange(n, dtype=np.int64)
    for qi in range(nq):
        d = hamming_distances(codes, qcodes[qi]).astype(np.int64)
        rows = rows_all
        if valid is not None:
            rows = rows_all[valid]
            d = d[valid]
        if rows.size == 0:
            continue
        key = (d << 32) | rows  # unique composite key
        kk = min(k, key.size)
        if kk < key.size:
            part = np.partition(key, kk - 1)[:kk]
        else:
            part = key
        part = np.sort(part)
        out_d[qi, :kk] = (part >> 32).astype(np.int32)
        out_i[qi, :kk] = part & 0xFFFFFFFF
    return out_d, out_i
