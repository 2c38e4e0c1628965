"""Deterministic synthetic corpora for parity tests and bench.py.

Two families:

* ``normal_unit_rows`` - ``numpy.random.default_rng(seed)`` N(0,1) rows,
  L2-normalised (reference embeddings are normalised, radiant/llm/local_models.py:164).
  Used for BASELINE config 1, the one config the reference's CPU path runs in full.

* counter-based rows (``hash_rows_f32`` / ``hash_query_rows_f32`` / Zipf token
  streams): every element is a pure function of (seed, row, column) built from
  the splitmix64 finaliser and integer arithmetic only, so the CUDA generator in
  ``csrc/synth.cu`` and this NumPy restatement are BIT-IDENTICAL and any sub-range
  of a 100M-row corpus can be regenerated on the CPU for a parity check without
  materialising the whole thing.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

_U64 = np.uint64
_GOLDEN = _U64(0x9E3779B97F4A7C15)
_M1 = _U64(0xBF58476D1CE4E5B9)
_M2 = _U64(0x94D049BB133111EB)

SEED_QUERY_SALT = 0x51ED270B
SEED_MIX_SALT = 0x2545F491
SEED_LEN_SALT = 0x0D15EA5E
SEED_TOK_SALT = 0x7F4A7C15


def splitmix64(counter: np.ndarray, seed: int) -> np.ndarray:
    """z = mix(counter + seed * golden) - identical to ``rr_splitmix64`` in csrc/synth.cu."""
    with np.errstate(over="ignore"):
        z = counter.astype(_U64) + _U64(seed & 0xFFFFFFFFFFFFFFFF) * _GOLDEN
        z = (z ^ (z >> _U64(30))) * _M1
        z = (z ^ (z >> _U64(27))) * _M2
        z = z ^ (z >> _U64(31))
    return z


def _irwin_hall4(z: np.ndarray) -> np.ndarray:
    """Sum of the four 16-bit fields of z, centred: int64 in [-131070, 131070]."""
    m = _U64(0xFFFF)
    s = (z & m) + ((z >> _U64(16)) & m) + ((z >> _U64(32)) & m) + ((z >> _U64(48)) & m)
    return s.astype(np.int64) - 131070


def value_shift(dim: int) -> int:
    """Power-of-two scale so that rows have roughly unit L2 norm: value = s * 2^-shift."""
    # std(s) = 65536/sqrt(3) ~= 37837; want std(value) ~= 1/sqrt(dim)
    return int(round(np.log2(37837.0 * np.sqrt(dim))))


def hash_rows_f32(row_start: int, n_rows: int, dim: int, seed: int) -> np.ndarray:
    """Corpus rows [row_start, row_start + n_rows) as f32 [n_rows, dim]."""
    rows = np.arange(row_start, row_start + n_rows, dtype=np.uint64)[:, None]
    cols = np.arange(dim, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        ctr = rows * _U64(dim) + cols
    s = _irwin_hall4(splitmix64(ctr, seed))
    return (s.astype(np.float32) * np.float32(2.0 ** -value_shift(dim))).astype(np.float32)


def query_source_row(q_index: np.ndarray, n_corpus: int, seed: int) -> np.ndarray:
    """Corpus row each odd-numbered query is derived from."""
    z = splitmix64(np.asarray(q_index, dtype=np.uint64), seed ^ SEED_MIX_SALT)
    return (z % _U64(max(n_corpus, 1))).astype(np.int64)


def hash_query_rows_f32(q_start: int, n_q: int, dim: int, seed: int, n_corpus: int) -> np.ndarray:
    """Query rows: even queries are fresh noise; odd queries copy 3/4 of the
    dimensions of a corpus row (so small Hamming distances and real neighbours
    occur) and redraw the rest."""
    qi = np.arange(q_start, q_start + n_q, dtype=np.uint64)
    cols = np.arange(dim, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        ctr = qi[:, None] * _U64(dim) + cols
    zq = splitmix64(ctr, seed ^ SEED_QUERY_SALT)
    fresh = _irwin_hall4(zq)
    src = query_source_row(qi, n_corpus, seed).astype(np.uint64)
    with np.errstate(over="ignore"):
        cctr = src[:, None] * _U64(dim) + cols
    copied = _irwin_hall4(splitmix64(cctr, seed))
    zm = splitmix64(ctr, seed ^ SEED_MIX_SALT)
    take_copy = ((qi[:, None] & _U64(1)) == _U64(1)) & ((zm & _U64(3)) != _U64(0))
    s = np.where(take_copy, copied, fresh)
    return (s.astype(np.float32) * np.float32(2.0 ** -value_shift(dim))).astype(np.float32)


def normal_unit_rows(n: int, dim: int, seed: int) -> np.ndarray:
    """i.i.d. N(0,1) rows, L2-normalised per row (BASELINE config 1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


# ---- BM25 token streams (BASELINE config 3) ------------------------------------

def zipf_cdf_u32(n_terms: int) -> np.ndarray:
    """Cumulative Zipf(s=1) distribution over term ranks quantised to uint32
    thresholds: term = first index with u <= table[index].  Computed on the host
    in both the CPU and the GPU path and uploaded, so both sides search the SAME
    integer table."""
    p = 1.0 / np.arange(1, n_terms + 1, dtype=np.float64)
    cdf = np.cumsum(p) / p.sum()
    t = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint64)
    t[-1] = 4294967295
    return t.astype(np.uint32)


def doc_lengths(row_start: int, n_rows: int, seed: int, mean_len: int = 200) -> np.ndarray:
    """Poisson-like document lengths, integer arithmetic only: mean + (s * 49 >> 17), >= 1."""
    rows = np.arange(row_start, row_start + n_rows, dtype=np.uint64)
    s = _irwin_hall4(splitmix64(rows, seed ^ SEED_LEN_SALT))
    return np.maximum(1, mean_len + ((s * 49) >> 17)).astype(np.int64)


def zipf_tokens(global_pos_start: int, n_tokens: int, seed: int, cdf: np.ndarray) -> np.ndarray:
    """Token ids for global token positions [start, start + n): Zipf over len(cdf) terms."""
    pos = np.arange(global_pos_start, global_pos_start + n_tokens, dtype=np.uint64)
    u = (splitmix64(pos, seed ^ SEED_TOK_SALT) >> _U64(32)).astype(np.uint32)
    return np.searchsorted(cdf, u, side="left").astype(np.int32)


def zipf_corpus(n_docs: int, n_terms: int, seed: int, mean_len: int = 200) -> Tuple[np.ndarray, np.ndarray]:
    """(doc_ptr int64 [n_docs+1], tokens int32 [total]) - document d owns global token
    positions [doc_ptr[d], doc_ptr[d+1])."""
    lens = doc_lengths(0, n_docs, seed, mean_len)
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    toks = zipf_tokens(0, int(ptr[-1]), seed, zipf_cdf_u32(n_terms))
    return ptr, toks


def zipf_queries(n_q: int, q_len: int, n_terms: int, seed: int) -> np.ndarray:
    """int32 [n_q, q_len] query term ids from the same Zipf law (repeats allowed)."""
    cdf = zipf_cdf_u32(n_terms)
    pos = np.arange(n_q * q_len, dtype=np.uint64)
    u = (splitmix64(pos, seed ^ SEED_QUERY_SALT) >> _U64(32)).astype(np.uint32)
    return np.searchsorted(cdf, u, side="left").astype(np.int32).reshape(n_q, q_len)
