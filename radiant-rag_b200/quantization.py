"""GPU-backed mirror of reference radiant/storage/quantization.py (same function names,
argument meaning and return types; NumPy in, NumPy out).

    quantize_embeddings   :74-108   -> rr_quantize_ubinary / rr_quantize_int8
    embedding_to_bytes    :111-121
    bytes_to_embedding    :124-136
    get_binary_dimension  :139-156
    calculate_int8_ranges :159-182
    rescore_candidates    :185-222  -> rr_rescore_f32

Unlike the reference this does not need sentence-transformers; it needs the CUDA
extension and raises if it is missing (no CPU fallback).
"""

from __future__ import annotations

import logging
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .index import _stream, to_device

logger = logging.getLogger(__name__)

QUANTIZATION_AVAILABLE = True  # reference flag name; here it means "CUDA extension present"


def _device(device: Optional[int]) -> torch.device:
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    _lib.init(dev.index)
    return dev


def quantize_embeddings(
    embeddings: np.ndarray,
    precision: str = "binary",
    ranges: Optional[np.ndarray] = None,
    device: Optional[int] = None,
) -> np.ndarray:
    """Quantise float32 embeddings [N, D] to "ubinary"/"binary" (packed sign bits) or
    "int8"/"uint8" (affine, calibrated by ``ranges`` [2, D]; batch min/max when None, as
    sentence-transformers does)."""
    if not isinstance(embeddings, np.ndarray):
        embeddings = np.array(embeddings, dtype=np.float32)
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    if emb.ndim == 1:
        emb = emb[None, :]
    n, d = emb.shape
    dev = _device(device)
    e = to_device(emb, dev, torch.float32)
    if precision in ("ubinary", "binary"):
        nbytes = (d + 7) // 8
        stride = (nbytes + 3) // 4 * 4
        out = torch.empty((n, stride), dtype=torch.uint8, device=dev)
        _lib.call("rr_quantize_ubinary", e.data_ptr(), n, d, out.data_ptr(), stride, _stream())
        codes = out[:, :nbytes].cpu().numpy()
        if precision == "binary":  # signed variant of sentence-transformers: packbits - 128 as int8
            return (codes.astype(np.int16) - 128).astype(np.int8)
        return codes
    if precision in ("int8", "uint8"):
        if ranges is None:
            ranges = calculate_int8_ranges(emb)
        r = to_device(np.asarray(ranges, dtype=np.float32), dev, torch.float32)
        out = torch.empty((n, d), dtype=torch.int8, device=dev)
        _lib.call("rr_quantize_int8", e.data_ptr(), n, d, r.data_ptr(), out.data_ptr(), _stream())
        q = out.cpu().numpy()
        if precision == "uint8":
            return (q.astype(np.int16) + 128).astype(np.uint8)
        return q
    if precision == "float32":
        return emb
    raise ValueError(f"Precision {precision!r} is not supported")


def embedding_to_bytes(embedding: np.ndarray) -> bytes:
    return embedding.tobytes()


def bytes_to_embedding(data: bytes, dtype: np.dtype, shape: tuple) -> np.ndarray:
    return np.frombuffer(data, dtype=dtype).reshape(shape)


def get_binary_dimension(float_dimension: int) -> int:
    if float_dimension % 8 != 0:
        logger.warning(
            f"Float dimension {float_dimension} is not divisible by 8. Binary embedding will be padded."
        )
    return (float_dimension + 7) // 8


def calculate_int8_ranges(embeddings: np.ndarray) -> np.ndarray:
    """[N, D] -> [2, D] min / max per dimension (host-side, as in the reference)."""
    if not isinstance(embeddings, np.ndarray):
        embeddings = np.array(embeddings, dtype=np.float32)
    return np.vstack([np.min(embeddings, axis=0), np.max(embeddings, axis=0)])


def rescore_candidates(
    query_embedding: np.ndarray,
    candidate_embeddings: Sequence[np.ndarray],
    candidate_ids: List[str],
    device: Optional[int] = None,
) -> List[tuple]:
    """Rescore candidates with higher-precision rows: [(doc_id, score)] sorted by score
    descending, ties in candidate order (the reference's stable sort)."""
    if len(candidate_embeddings) == 0:
        return []
    q = np.asarray(query_embedding)
    if q.dtype != np.float32:
        q = q.astype(np.float32)
    rows = np.stack([np.asarray(e) for e in candidate_embeddings])
    c = rows.shape[0]
    if c > _lib.RR_MAX_K:
        raise ValueError(f"at most {_lib.RR_MAX_K} candidates per call")
    dev = _device(device)
    if rows.dtype == np.int8:
        rt, dt = to_device(rows, dev, torch.int8), _lib.RR_I8
    else:
        rt, dt = to_device(rows.astype(np.float32, copy=False), dev, torch.float32), _lib.RR_F32
    qt = to_device(q[None, :], dev, torch.float32)
    cand = torch.arange(c, dtype=torch.int64, device=dev)[None, :].contiguous()
    score = torch.empty((1, c), dtype=torch.float32, device=dev)
    idx = torch.empty((1, c), dtype=torch.int64, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    _lib.call("rr_rescore_f32", qt.data_ptr(), 1, q.shape[0], rt.data_ptr(), dt, c, 0, cand.data_ptr(),
              c, c, float("-inf"), score.data_ptr(), idx.data_ptr(), count.data_ptr(), _stream())
    m = int(count.item())
    order = idx[0, :m].cpu().tolist()
    vals = score[0, :m].cpu().tolist()
    return [(candidate_ids[i], float(s)) for i, s in zip(order, vals)]
