"""Oracle (test infrastructure): embedding quantisation, rows R1/R2 of SURVEY.md 8(a).

PARITY UNPINNED.  The reference forwards to the third-party
``sentence_transformers.quantization.quantize_embeddings`` (reference
radiant/storage/quantization.py:21-26, 104-108; call sites
radiant/storage/redis_store.py:328, 340-344, radiant/storage/chroma_store.py:258,
269-273, 591, radiant/storage/pgvector_store.py:429, 443-447).  That package is
not installed here and not pinned by the reference (requirements.txt:47
``sentence-transformers>=3.2.0``), so the functions below restate the published
algorithm of sentence-transformers 3.x:

    ubinary : np.packbits(emb > 0).reshape(N, -1)
    int8    : starts = ranges[0]; steps = (ranges[1] - ranges[0]) / 255
              ((emb - starts) / steps - 128).astype(np.int8)

The only checks the reference holds are shape/dtype ones
(tools/validate_quantization.py:142 ``get_binary_dimension(384) == 48``,
:159-160 ubinary -> uint8 of width 48, :169-170 int8 dtype/shape).
"""

from __future__ import annotations

import numpy as np


def quantize_ubinary(emb: np.ndarray) -> np.ndarray:
    """f32 [N, D] -> u8 [N, ceil(D/8)]; bit = (x > 0); dim 8b is the MSB of byte b.

    Zeros and NaN map to bit 0 (strict ``>``).  D not a multiple of 8 is padded
    with zero bits in the low bits of the last byte (np.packbits behaviour).
    """
    emb = np.asarray(emb, dtype=np.float32)
    if emb.ndim == 1:
        emb = emb[None, :]
    n = emb.shape[0]
    return np.packbits(emb > 0, axis=-1).reshape(n, -1)


def calculate_int8_ranges(emb: np.ndarray) -> np.ndarray:
    """[N, D] -> [2, D] per-dimension min / max.

    Follows reference radiant/storage/quantization.py:159-182.
    """
    emb = np.asarray(emb, dtype=np.float32)
    return np.vstack([np.min(emb, axis=0), np.max(emb, axis=0)])


def quantize_int8(emb: np.ndarray, ranges: np.ndarray) -> np.ndarray:
    """f32 [N, D] + f32 ranges [2, D] -> i8 [N, D].

    All arithmetic is IEEE float32 in the order ``(x - lo) / step - 128`` with
    ``step = (hi - lo) / 255`` and the cast truncates toward zero.  The reference
    never clips; out-of-range inputs are undefined behaviour of the C cast there.
    This oracle (and the CUDA kernel) SATURATE to [-128, 127] instead, which is
    identical whenever the input lies inside the calibrated range (the only
    defined case).  NaN maps to 0.
    """
    emb = np.asarray(emb, dtype=np.float32)
    if emb.ndim == 1:
        emb = emb[None, :]
    ranges = np.asarray(ranges, dtype=np.float32)
    starts = ranges[0, :]
    steps = (ranges[1, :] - ranges[0, :]) / np.float32(255)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = (emb - starts) / steps - np.float32(128)
    v = np.where(np.isnan(v), np.float32(0), v)
    v = np.clip(v, np.float32(-128), np.float32(127))
    return np.trunc(v).astype(np.int8)


def quantize_int8_symmetric_query(q: np.ndarray, ranges: np.ndarray) -> np.ndarray:
    """Query-side int8 codes for the symmetric int8 x int8 extension mode.

    The reference never quantises queries to int8; the
    north-star's bit-exact int8 dot product needs one, so queries go through the
    same affine map as documents, saturated.
    """
    return quantize_int8(q, ranges)
