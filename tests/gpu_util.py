"""Helpers shared by the -m gpu parity tests (all compute goes through the C ABI)."""

import numpy as np
import pytest
import torch

import oracle
from radiant_rag_b200 import synthetic
from radiant_rag_b200.index import DenseIndex

REL = 1e-5       # north-star tolerance, float32 rescoring / BM25
ABS_FLOOR = 1e-6


def require_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


def ulp_diff_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in float32 ulps (same-sign finite values)."""
    ai = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-(2 ** 31)) - ai, ai)
    bi = np.where(bi < 0, np.int64(-(2 ** 31)) - bi, bi)
    return np.abs(ai - bi)


def build_index(corpus, int8=True, f32=True, tags=None, row_base=0):
    ranges = oracle.calculate_int8_ranges(corpus) if int8 and len(corpus) else None
    idx = DenseIndex(corpus.shape[1], device=0, store_int8=int8 and ranges is not None, store_f32=f32,
                     int8_ranges=ranges, row_base=row_base)
    if len(corpus):
        idx.add(corpus, tags)
    return idx, ranges


def assert_lists_match_tie_aware(got_ids, got_s, want_ids, want_s, rel=REL, floor=ABS_FLOOR, ctx=""):
    """Same length; scores within tolerance position by position; ids equal except where
    the wanted scores of the swapped ids are within the tolerance of each other."""
    assert len(got_ids) == len(want_ids), (ctx, len(got_ids), len(want_ids))
    want = dict(zip(want_ids, want_s))
    for p, (g, w) in enumerate(zip(got_ids, want_ids)):
        assert abs(got_s[p] - want_s[p]) <= rel * abs(want_s[p]) + floor, (ctx, p, got_s[p], want_s[p])
        if g != w:
            assert g in want, (ctx, p, g, "not among wanted ids")
            assert abs(want[g] - want[w]) <= 2 * (rel * abs(want[w]) + floor), (ctx, p, g, w)
