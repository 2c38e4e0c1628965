"""Device-resident quantised index of ONE shard and the batched search entry points.

This is the "quantized index layout in radiant/storage" the north-star names.  HBM
layout (row = position in upsert order, all arrays row-major, rows contiguous):

    codes  uint8 [cap, words*4]   packed sign bits, np.packbits byte order, row padded
                                  with zero bits to a multiple of 16 bytes
    int8   int8  [cap, dim]       optional (precision "int8"/"both")
    f32    f32   [cap, dim]       optional (float rescoring / exact scan)
    tags   uint8 [cap]            bits 0-1 doc_level (1 child, 2 parent), bits 2-7 language id

torch is used for allocation, streams and host<->device copies only; every
computation goes through the C ABI (``_lib.call``) into the sm_100a kernels.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .synthetic import value_shift

ArrayLike = Union[np.ndarray, torch.Tensor]

LEVEL_CHILD = 1
LEVEL_PARENT = 2
_LEVEL_BITS = {"child": LEVEL_CHILD, "parent": LEVEL_PARENT}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def code_words(dim: int) -> int:
    """u32 words per packed row: ceil(dim/32) rounded up to a multiple of 4 (16 bytes)."""
    w = (dim + 31) // 32
    return (w + 3) // 4 * 4


def to_device(x: ArrayLike, device: torch.device, dtype: torch.dtype) -> torch.Tensor:
    """Host array (pinned on the way) or tensor -> contiguous device tensor of dtype."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        if t.device.type == "cpu" and not t.is_pinned():
            t = t.pin_memory()
        t = t.to(device, non_blocking=True)
    return t.contiguous()


class LanguageTable:
    """language code <-> 6-bit id stored in the tag byte (0 = none)."""

    def __init__(self) -> None:
        self._ids: Dict[str, int] = {}

    def id_for(self, code: Optional[str], create: bool = False) -> int:
        if not code:
            return 0
        if code in self._ids:
            return self._ids[code]
        if not create:
            return 63  # never assigned to a row unless the table is full: matches nothing new
        if len(self._ids) >= 62:
            raise ValueError("more than 62 distinct language codes in one index")
        self._ids[code] = len(self._ids) + 1
        return self._ids[code]


def make_tag(doc_level: Optional[str], lang_id: int) -> int:
    """Level bits: 1 = "child", 2 = "parent", 0 = any other stored value - the reference's filters
    match the stored string exactly (redis_store.py:684, 915-916), so such a row passes neither."""
    return _LEVEL_BITS.get("child" if doc_level is None else doc_level, 0) | (lang_id << 2)


def tag_predicate(level_value: Optional[str], lang_id: int) -> Tuple[int, int]:
    """(mask, value) such that a row passes when (tag & mask) == value."""
    mask = 0
    value = 0
    if level_value in _LEVEL_BITS:
        mask |= 0x03
        value |= _LEVEL_BITS[level_value]
    if lang_id:
        mask |= 0xFC
        value |= lang_id << 2
    return mask, value


class DenseIndex:
    """Packed binary codes + int8 / float32 rows of one shard, resident in HBM."""

    def __init__(
        self,
        dim: int,
        device: Union[int, str, torch.device] = 0,
        store_int8: bool = True,
        store_f32: bool = True,
        int8_ranges: Optional[ArrayLike] = None,
        row_base: int = 0,
        capacity: int = 0,
    ) -> None:
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RadiantB200Error("DenseIndex needs a CUDA device; there is no CPU fallback")
        _lib.init(self.device.index or 0)
        self.dim = int(dim)
        self.words = code_words(self.dim)
        if self.words > _lib.RR_MAX_WORDS:
            raise ValueError(f"dim {dim} > {_lib.RR_MAX_WORDS * 32} unsupported by the Hamming scan")
        self.store_int8 = bool(store_int8)
        self.store_f32 = bool(store_f32)
        self.row_base = int(row_base)
        self.n = 0
        self._cap = 0
        self.codes: Optional[torch.Tensor] = None
        self.int8: Optional[torch.Tensor] = None
        self.f32: Optional[torch.Tensor] = None
        self.tags: Optional[torch.Tensor] = None
        self.tags_uniform = True  # every row has the default tag: the scan can skip the tag read
        self.use_tensor_cores = True   # batched stage 1 on tcgen05 (rr_hamming_topk_tc)
        self.tc_min_queries = 16       # below this the POPC scan (memory-bound) is the faster path
        self._tc_overflow: Optional[torch.Tensor] = None
        self.last_tc_redone = 0        # queries the last checked tensor-core call redid on the exact path
        self.tc_exact_min_queries = 8  # exact float32 scan: batches from here on take the TF32 filter + float64 refine
        self._inv_norm: Optional[torch.Tensor] = None  # 1/|row| of the float32 rows (steers the TF32 filter only)
        self._inv_norm_n = 0
        self.ranges: Optional[torch.Tensor] = None
        if int8_ranges is not None:
            self.set_int8_ranges(int8_ranges)
        if capacity:
            self._reserve(capacity)

    def _activate(self) -> None:
        """CUDA's current device is per host thread; the reference drives retrieval from
        worker threads (radiant/orchestrator.py:994-998), so select ours on entry."""
        if torch.cuda.current_device() != (self.device.index or 0):
            torch.cuda.set_device(self.device)

    # ---- storage ---------------------------------------------------------------
    def _reserve(self, cap: int) -> None:
        if cap <= self._cap:
            return
        cap = max(cap, int(self._cap * 1.5), 1024)

        def grow(old: Optional[torch.Tensor], shape, dtype) -> torch.Tensor:
            new = torch.empty(shape, dtype=dtype, device=self.device)
            if old is not None and self.n:
                new[: self.n].copy_(old[: self.n])
            return new

        self.codes = grow(self.codes, (cap, self.words * 4), torch.uint8)
        self.tags = grow(self.tags, (cap,), torch.uint8)
        if self.store_int8:
            self.int8 = grow(self.int8, (cap, self.dim), torch.int8)
        if self.store_f32:
            self.f32 = grow(self.f32, (cap, self.dim), torch.float32)
        self._cap = cap

    def set_int8_ranges(self, ranges: ArrayLike) -> None:
        """[2, dim] min/max calibration (reference calculate_int8_ranges,
        radiant/storage/quantization.py:159-182, or the .npy of tools/calibrate_int8_ranges.py)."""
        r = to_device(ranges, self.device, torch.float32)
        if tuple(r.shape) != (2, self.dim):
            raise ValueError(f"int8 ranges must have shape (2, {self.dim}), got {tuple(r.shape)}")
        self.ranges = r

    def add(self, emb: ArrayLike, tags: Optional[ArrayLike] = None) -> Tuple[int, int]:
        """Append float32 rows [m, dim]: quantises to ubinary (+ int8) on device.
        Returns the local row range [lo, hi)."""
        self._activate()
        e = to_device(emb, self.device, torch.float32)
        if e.ndim == 1:
            e = e[None, :]
        if e.shape[1] != self.dim:
            raise ValueError(f"embedding dim {e.shape[1]} != index dim {self.dim}")
        m = e.shape[0]
        lo, hi = self.n, self.n + m
        self._reserve(hi)
        _lib.call("rr_quantize_ubinary", e.data_ptr(), m, self.dim, self.codes[lo:hi].data_ptr(),
                  self.words * 4, _stream())
        if self.store_int8:
            if self.ranges is None:
                raise ValueError("int8 storage needs calibration ranges (set_int8_ranges)")
            _lib.call("rr_quantize_int8", e.data_ptr(), m, self.dim, self.ranges.data_ptr(),
                      self.int8[lo:hi].data_ptr(), _stream())
        if self.store_f32:
            self.f32[lo:hi].copy_(e)
        if tags is None:
            self.tags[lo:hi].fill_(LEVEL_CHILD)
        else:
            self.tags[lo:hi].copy_(to_device(tags, self.device, torch.uint8))
            self.tags_uniform = False
        self.n = hi
        return lo, hi

    def add_quantized(self, codes: ArrayLike, int8_rows: Optional[ArrayLike] = None,
                      f32_rows: Optional[ArrayLike] = None, tags: Optional[ArrayLike] = None) -> Tuple[int, int]:
        """Append rows that are ALREADY quantised, byte for byte as the reference stores them
        (wire formats): codes u8 [m, ceil(dim/8)] = ``np.packbits`` rows, the payload
        of the ``...:doc_binary:{id}`` side keys (redis_store.py:329-338); int8_rows i8 [m, dim], the
        ``...:doc_int8:{id}`` payload (:339-349); f32_rows the float32 embeddings if kept.
        Nothing is recomputed.  Returns the local row range [lo, hi)."""
        self._activate()
        c = to_device(codes, self.device, torch.uint8)
        if c.ndim == 1:
            c = c[None, :]
        nbytes = (self.dim + 7) // 8
        if c.shape[1] != nbytes:
            raise ValueError(f"packed code rows must have {nbytes} bytes for dim {self.dim}, got {c.shape[1]}")
        m = c.shape[0]
        if self.store_int8 and int8_rows is None:
            raise ValueError("this index stores int8 rows: int8_rows is required")
        if self.store_f32 and f32_rows is None:
            raise ValueError("this index stores float32 rows: f32_rows is required")
        lo, hi = self.n, self.n + m
        self._reserve(hi)
        self.codes[lo:hi].zero_()  # row padding up to 16 bytes must be 0
        self.codes[lo:hi, :nbytes].copy_(c)
        if self.store_int8:
            r = to_device(int8_rows, self.device, torch.int8).reshape(m, self.dim)
            self.int8[lo:hi].copy_(r)
        if self.store_f32:
            self.f32[lo:hi].copy_(to_device(f32_rows, self.device, torch.float32).reshape(m, self.dim))
        if tags is None:
            self.tags[lo:hi].fill_(LEVEL_CHILD)
        else:
            self.tags[lo:hi].copy_(to_device(tags, self.device, torch.uint8))
            self.tags_uniform = False
        self.n = hi
        return lo, hi

    def save(self, path: str) -> None:
        """Persist the shard: one uncompressed .npz with the packed codes, the
        int8 / float32 rows that are kept, tags, calibration ranges and the scalar settings."""
        self._activate()
        torch.cuda.current_stream(self.device).synchronize()
        n = self.n
        arrays = {
            "dim": np.int64(self.dim), "row_base": np.int64(self.row_base), "n": np.int64(n),
            "codes": self.codes[:n, : (self.dim + 7) // 8].cpu().numpy() if n else np.zeros((0, (self.dim + 7) // 8), np.uint8),
            "tags": self.tags[:n].cpu().numpy() if n else np.zeros((0,), np.uint8),
        }
        if self.store_int8:
            arrays["int8"] = self.int8[:n].cpu().numpy() if n else np.zeros((0, self.dim), np.int8)
        if self.store_f32:
            arrays["f32"] = self.f32[:n].cpu().numpy() if n else np.zeros((0, self.dim), np.float32)
        if self.ranges is not None:
            arrays["ranges"] = self.ranges.cpu().numpy()
        with open(path, "wb") as fh:
            np.savez(fh, **arrays)

    @classmethod
    def load(cls, path: str, device: Union[int, str, torch.device] = 0) -> "DenseIndex":
        """Rebuild a shard saved by ``save`` on `device` (rows are copied, not re-quantised)."""
        with np.load(path) as z:
            dim, n = int(z["dim"]), int(z["n"])
            idx = cls(dim, device=device, store_int8="int8" in z.files, store_f32="f32" in z.files,
                      int8_ranges=z["ranges"] if "ranges" in z.files else None, row_base=int(z["row_base"]),
                      capacity=n)
            if n:
                idx.add_quantized(z["codes"], z["int8"] if "int8" in z.files else None,
                                  z["f32"] if "f32" in z.files else None, z["tags"])
                idx.tags_uniform = bool((z["tags"] == LEVEL_CHILD).all())
        return idx

    def set_row(self, row: int, emb: ArrayLike, tag: int) -> None:
        """Overwrite one existing row (upsert of a known doc_id)."""
        self._activate()
        e = to_device(emb, self.device, torch.float32).reshape(1, self.dim)
        _lib.call("rr_quantize_ubinary", e.data_ptr(), 1, self.dim, self.codes[row:row + 1].data_ptr(),
                  self.words * 4, _stream())
        if self.store_int8:
            if self.ranges is None:
                raise ValueError("int8 storage needs calibration ranges (set_int8_ranges)")
            _lib.call("rr_quantize_int8", e.data_ptr(), 1, self.dim, self.ranges.data_ptr(),
                      self.int8[row:row + 1].data_ptr(), _stream())
        if self.store_f32:
            self.f32[row].copy_(e[0])
            if row < self._inv_norm_n:
                _lib.call("rr_row_inv_norms_f32", self.f32[row:row + 1].data_ptr(), 1, self.dim,
                          self._inv_norm[row:row + 1].data_ptr(), _stream())
        self.tags[row] = tag
        if tag != LEVEL_CHILD:
            self.tags_uniform = False

    def delete_row(self, row: int) -> Optional[int]:
        """Remove one row by moving the LAST row into its slot (the arrays stay dense).  Returns the old
        index of the row that moved (``n - 1`` before the call), or None when the last row itself was
        removed.  Keeps the cached row norms of the exact tensor-core scan consistent."""
        self._activate()
        if not 0 <= row < self.n:
            raise IndexError(f"row {row} outside [0, {self.n})")
        last = self.n - 1
        moved = None
        if row != last:
            self.codes[row].copy_(self.codes[last])
            self.tags[row] = self.tags[last]
            if self.int8 is not None:
                self.int8[row].copy_(self.int8[last])
            if self.f32 is not None:
                self.f32[row].copy_(self.f32[last])
            if self._inv_norm is not None and row < self._inv_norm_n:
                if last < self._inv_norm_n:
                    self._inv_norm[row] = self._inv_norm[last]
                else:  # the moved row's norm was not cached yet
                    _lib.call("rr_row_inv_norms_f32", self.f32[row:row + 1].data_ptr(), 1, self.dim,
                              self._inv_norm[row:row + 1].data_ptr(), _stream())
            moved = last
        self.n = last
        self._inv_norm_n = min(self._inv_norm_n, self.n)
        return moved

    def set_tag(self, row: int, tag: int) -> None:
        self.tags[row] = tag
        if tag != LEVEL_CHILD:
            self.tags_uniform = False

    def clear(self) -> None:
        self.n = 0
        self.tags_uniform = True
        self._inv_norm_n = 0

    def invalidate_norms(self) -> None:
        """Call after writing ``self.f32`` directly (bulk loaders that bypass add / set_row)."""
        self._inv_norm_n = 0

    def _row_inv_norms(self) -> torch.Tensor:
        """1 / |row| for rows [0, n), computed once per row (rr_row_inv_norms_f32) and kept."""
        if self._inv_norm is None or self._inv_norm.shape[0] < self.n:
            new = torch.empty((max(self._cap, self.n),), dtype=torch.float32, device=self.device)
            if self._inv_norm is not None and self._inv_norm_n:
                new[: self._inv_norm_n].copy_(self._inv_norm[: self._inv_norm_n])
            self._inv_norm = new
        if self._inv_norm_n < self.n:
            lo = self._inv_norm_n
            _lib.call("rr_row_inv_norms_f32", self.f32[lo:self.n].data_ptr(), self.n - lo, self.dim,
                      self._inv_norm[lo:self.n].data_ptr(), _stream())
            self._inv_norm_n = self.n
        return self._inv_norm

    # ---- kernels -----------------------------------------------------------------
    def quantize_queries(self, queries: ArrayLike) -> Tuple[torch.Tensor, torch.Tensor]:
        """f32 [q, dim] -> (device f32 queries, device packed codes uint8 [q, words*4])."""
        self._activate()
        qf = to_device(queries, self.device, torch.float32)
        if qf.ndim == 1:
            qf = qf[None, :]
        qc = torch.empty((qf.shape[0], self.words * 4), dtype=torch.uint8, device=self.device)
        _lib.call("rr_quantize_ubinary", qf.data_ptr(), qf.shape[0], self.dim, qc.data_ptr(),
                  self.words * 4, _stream())
        return qf, qc

    def _tag_args(self, tag_mask: int, tag_value: int):
        if tag_mask == 0 or self.tags is None:
            return None, 0, 0
        if self.tags_uniform:
            # every row carries the default tag: decide the predicate once on the host
            if (LEVEL_CHILD & tag_mask) == tag_value:
                return None, 0, 0
        return self.tags.data_ptr(), tag_mask, tag_value

    def _hamming_topk_popc(self, qcodes: torch.Tensor, k: int, tag_mask: int, tag_value: int
                           ) -> Tuple[torch.Tensor, torch.Tensor]:
        q = qcodes.shape[0]
        dist = torch.empty((q, k), dtype=torch.int32, device=self.device)
        idx = torch.empty((q, k), dtype=torch.int64, device=self.device)
        lib = _lib.load()
        ws_bytes = lib.rr_hamming_topk_workspace_bytes(self.n, self.words, q, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        tptr, tm, tv = self._tag_args(tag_mask, tag_value)
        _lib.call("rr_hamming_topk", self.codes.data_ptr() if self.codes is not None else None, self.n,
                  self.words, tptr, tm, tv, qcodes.data_ptr(), q, k, self.row_base, dist.data_ptr(),
                  idx.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        return dist, idx

    def _hamming_topk_tc(self, qcodes: torch.Tensor, k: int, tag_mask: int, tag_value: int,
                         ovf: Optional[torch.Tensor] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Tensor-core path -> (dist, idx, ovf, flags): the device overflow counter of this call
        (`ovf`: an existing counter the kernel adds to, instead of a fresh zeroed one) and one flag
        per query whose bounded candidate list overflowed."""
        q = qcodes.shape[0]
        dpad = self.words * 32  # padded width: padding bits are 0 (-1) in rows and queries alike
        q_pm1 = torch.empty((q, dpad), dtype=torch.int8, device=self.device)
        _lib.call("rr_unpack_codes_pm1", qcodes.data_ptr(), q, self.words * 4, dpad, q_pm1.data_ptr(),
                  _stream())
        dist = torch.empty((q, k), dtype=torch.int32, device=self.device)
        idx = torch.empty((q, k), dtype=torch.int64, device=self.device)
        flags = torch.empty((q,), dtype=torch.uint8, device=self.device)
        if ovf is None:
            ovf = torch.zeros(1, dtype=torch.int32, device=self.device)
        lib = _lib.load()
        ws_bytes = lib.rr_tc_search_workspace_bytes(self.n, q, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        tptr, tm, tv = self._tag_args(tag_mask, tag_value)
        _lib.call("rr_hamming_topk_tc", self.codes.data_ptr(), self.n, self.words, tptr, tm, tv,
                  q_pm1.data_ptr(), q, k, self.row_base, dist.data_ptr(), idx.data_ptr(), ovf.data_ptr(),
                  flags.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        return dist, idx, ovf, flags

    def hamming_topk(self, qcodes: torch.Tensor, k: int, tag_mask: int = 0, tag_value: int = 0,
                     use_tc: Optional[bool] = None, check_overflow: bool = True
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Exact top-k by (dist asc, row asc).  -> (dist int32 [q,k], idx int64 [q,k]).

        Batches of >= tc_min_queries run on the tensor cores (packed codes expanded on chip
        to a 0/255 u8 operand in tensor memory); both paths return identical results.  The
        tensor-core path keeps, per query, the rows that beat a sampled bound in a bounded list:
        the queries whose list overflows (clustered rows, selective filters) are redone on the
        POPC path - only those.  check_overflow=False skips that host-side check (one device
        sync) and accumulates the counter in ``tc_overflow_total()`` for the caller to verify."""
        dist, idx, ovf, flags = self._hamming_topk_auto(qcodes, k, tag_mask, tag_value, use_tc,
                                                        accumulate=not check_overflow)
        if ovf is not None and check_overflow and int(ovf.item()) != 0:
            bad = torch.nonzero(flags).flatten()
            self.last_tc_redone = int(bad.numel())
            d2, i2 = self._hamming_topk_popc(qcodes[bad].contiguous(), k, tag_mask, tag_value)
            dist[bad] = d2
            idx[bad] = i2
        return dist, idx

    def _hamming_topk_auto(self, qcodes: torch.Tensor, k: int, tag_mask: int, tag_value: int,
                           use_tc: Optional[bool] = None, accumulate: bool = False):
        """-> (dist, idx, ovf, flags): ovf is the device overflow counter of a tensor-core call, None
        when the POPC path ran (always exact).  accumulate: the kernel adds straight into the
        index's running counter (``tc_overflow_total``) - no per-call counter, no extra launches."""
        self._activate()
        q = qcodes.shape[0]
        if use_tc is None:
            use_tc = self.use_tensor_cores and q >= self.tc_min_queries and self.n >= 4096
        if not use_tc or self.n == 0:
            dist, idx = self._hamming_topk_popc(qcodes, k, tag_mask, tag_value)
            return dist, idx, None, None
        if accumulate:
            if self._tc_overflow is None:
                self._tc_overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
            return self._hamming_topk_tc(qcodes, k, tag_mask, tag_value, ovf=self._tc_overflow)
        return self._hamming_topk_tc(qcodes, k, tag_mask, tag_value)

    def tc_overflow_total(self) -> int:
        """Overflow events of tensor-core calls made with check_overflow=False (0 = all exact)."""
        return 0 if self._tc_overflow is None else int(self._tc_overflow.item())

    def tc_overflow_reset(self) -> None:
        if self._tc_overflow is not None:
            self._tc_overflow.zero_()

    def rescore_source(self, prefer_int8: bool = True) -> Tuple[torch.Tensor, int]:
        """int8 rows preferred, float32 fallback (reference redis_store.py:820-840)."""
        if prefer_int8 and self.int8 is not None:
            return self.int8, _lib.RR_I8
        if self.f32 is not None:
            return self.f32, _lib.RR_F32
        if self.int8 is not None:
            return self.int8, _lib.RR_I8
        raise _lib.RadiantB200Error("index stores neither int8 nor float32 rows: cannot rescore")

    def rescore(self, queries: torch.Tensor, cand_idx: torch.Tensor, top_k: int,
                min_similarity: float = 0.0, prefer_int8: bool = True
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> (score f32 [q,top_k], idx int64 [q,top_k] (-1 padded), count int32 [q])."""
        q, c = cand_idx.shape
        rows, dt = self.rescore_source(prefer_int8)
        score = torch.empty((q, top_k), dtype=torch.float32, device=self.device)
        idx = torch.empty((q, top_k), dtype=torch.int64, device=self.device)
        count = torch.empty((q,), dtype=torch.int32, device=self.device)
        if q >= 8:
            # batches: the scoring kernel spreads a query's candidates over several CTAs (more
            # rows in flight per SM), then one small ranking kernel; same arithmetic, same order
            s = self.score_candidates(queries, cand_idx, prefer_int8)
            _lib.call("rr_rank_scored_f32", s.data_ptr(), cand_idx.data_ptr(), q, c, top_k,
                      float(min_similarity), score.data_ptr(), idx.data_ptr(), count.data_ptr(), _stream())
            return score, idx, count
        _lib.call("rr_rescore_f32", queries.data_ptr(), q, self.dim, rows.data_ptr(), dt, self.n,
                  self.row_base, cand_idx.data_ptr(), c, top_k, float(min_similarity), score.data_ptr(),
                  idx.data_ptr(), count.data_ptr(), _stream())
        return score, idx, count

    def score_candidates(self, queries: torch.Tensor, cand_idx: torch.Tensor,
                         prefer_int8: bool = True) -> torch.Tensor:
        """Scores of the candidates THIS shard owns, -inf elsewhere.  f32 [q, c]."""
        q, c = cand_idx.shape
        rows, dt = self.rescore_source(prefer_int8)
        out = torch.empty((q, c), dtype=torch.float32, device=self.device)
        _lib.call("rr_score_candidates_f32", queries.data_ptr(), q, self.dim, rows.data_ptr(), dt,
                  self.n, self.row_base, cand_idx.data_ptr(), c, out.data_ptr(), _stream())
        return out

    def rescore_int8_symmetric(self, queries_i8: torch.Tensor, cand_idx: torch.Tensor, top_k: int
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Extension mode: exact int8 x int8 -> int32 rescoring."""
        if self.int8 is None:
            raise _lib.RadiantB200Error("index has no int8 rows")
        q, c = cand_idx.shape
        score = torch.empty((q, top_k), dtype=torch.int32, device=self.device)
        idx = torch.empty((q, top_k), dtype=torch.int64, device=self.device)
        count = torch.empty((q,), dtype=torch.int32, device=self.device)
        _lib.call("rr_rescore_i8", queries_i8.data_ptr(), q, self.dim, self.int8.data_ptr(), self.n,
                  self.row_base, cand_idx.data_ptr(), c, top_k, score.data_ptr(), idx.data_ptr(),
                  count.data_ptr(), _stream())
        return score, idx, count

    def search_quantized(
        self,
        queries: ArrayLike,
        top_k: int,
        rescore_multiplier: float = 4.0,
        use_rescoring: bool = True,
        min_similarity: float = 0.0,
        tag_mask: int = 0,
        tag_value: int = 0,
        prefer_int8: bool = True,
        check_overflow: bool = True,
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Batched two-stage retrieval (reference redis_store.py:757-861 for one query).

        -> (idx int64 [q, top_k] (-1 padded), score f32 [q, top_k], count int32 [q]).
        Without rescoring the score is the placeholder 1.0 (reference chroma_store.py:624-631).
        """
        qf, qc = self.quantize_queries(queries)
        if not 1 <= int(top_k) <= _lib.RR_MAX_K:
            raise ValueError(f"top_k={top_k} outside [1, {_lib.RR_MAX_K}]")
        candidate_k = int(top_k * rescore_multiplier) if use_rescoring else top_k
        candidate_k = max(int(top_k), min(candidate_k, _lib.RR_MAX_K))  # documented limit: INTEGRATION.md
        # Stage 2 is queued behind stage 1 before the tensor-core overflow counter is read, so
        # the (single) host sync of a checked call sits at its end and the GPU never idles
        # between the stages; an overflow (adversarial data) redoes the call on the POPC path.
        _dist, cand, ovf, flags = self._hamming_topk_auto(qc, candidate_k, tag_mask, tag_value,
                                                          accumulate=not check_overflow)

        def stage2(q_rows, cand_rows):
            if not use_rescoring:
                idx = cand_rows[:, :top_k].contiguous()
                score = torch.ones(idx.shape, dtype=torch.float32, device=self.device)
                count = (idx >= 0).sum(dim=1).to(torch.int32)
                return idx, score, count
            score, idx, count = self.rescore(q_rows, cand_rows, top_k, min_similarity, prefer_int8)
            return idx, score, count

        out = stage2(qf, cand)
        if ovf is not None and check_overflow and int(ovf.item()) != 0:
            # only the queries whose candidate list overflowed go through the POPC path again
            bad = torch.nonzero(flags).flatten()
            self.last_tc_redone = int(bad.numel())
            _dist, cand_b = self._hamming_topk_popc(qc[bad].contiguous(), candidate_k, tag_mask, tag_value)
            out_b = stage2(qf[bad].contiguous(), cand_b)
            for t, tb in zip(out, out_b):
                t[bad] = tb
        return out

    def search_exact(self, queries: ArrayLike, top_k: int, min_similarity: float = 0.0,
                     tag_mask: int = 0, tag_value: int = 0, use_tc: Optional[bool] = None,
                     check_overflow: bool = True
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Exact float32 cosine scan (reference redis_store.py:863-952), batched.

        Batches of >= ``tc_exact_min_queries`` run the filter on the tensor cores
        (rr_exact_search_f32_tc: TF32 products, exact float64 refine of the survivors) with
        results bit-identical to the CUDA-core scan; a checked call redoes the queries whose
        candidate list overflowed with rr_exact_search_f32, an unchecked call (graph capture)
        accumulates the counter (``tc_overflow_total``)."""
        self._activate()
        if self.f32 is None:
            raise _lib.RadiantB200Error("index has no float32 rows: exact search unavailable")
        if not 1 <= int(top_k) <= _lib.RR_MAX_K:
            raise ValueError(f"top_k={top_k} outside [1, {_lib.RR_MAX_K}]")
        qf = to_device(queries, self.device, torch.float32)
        if qf.ndim == 1:
            qf = qf[None, :]
        q = qf.shape[0]
        lib = _lib.load()
        tptr, tm, tv = self._tag_args(tag_mask, tag_value)

        def cuda_cores(q_rows):
            qn = q_rows.shape[0]
            score = torch.empty((qn, top_k), dtype=torch.float32, device=self.device)
            idx = torch.empty((qn, top_k), dtype=torch.int64, device=self.device)
            count = torch.empty((qn,), dtype=torch.int32, device=self.device)
            ws_bytes = lib.rr_exact_search_f32_workspace_bytes(self.n, qn, top_k)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            _lib.call("rr_exact_search_f32", self.f32.data_ptr(), self.n, self.dim, tptr, tm, tv,
                      q_rows.data_ptr(), qn, top_k, float(min_similarity), self.row_base, score.data_ptr(),
                      idx.data_ptr(), count.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
            return idx, score, count

        if use_tc is None:
            use_tc = self.use_tensor_cores and q >= self.tc_exact_min_queries
        if use_tc:
            use_tc = bool(lib.rr_exact_search_f32_tc_supported(self.n, self.dim, q, top_k)) \
                and qf.data_ptr() % 16 == 0
        if not use_tc:
            return cuda_cores(qf.contiguous())
        qf = qf.contiguous()
        score = torch.empty((q, top_k), dtype=torch.float32, device=self.device)
        idx = torch.empty((q, top_k), dtype=torch.int64, device=self.device)
        count = torch.empty((q,), dtype=torch.int32, device=self.device)
        flags = torch.empty((q,), dtype=torch.uint8, device=self.device)
        if check_overflow:
            ovf = torch.zeros(1, dtype=torch.int32, device=self.device)
        else:
            if self._tc_overflow is None:
                self._tc_overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
            ovf = self._tc_overflow
        inv = self._row_inv_norms()
        ws_bytes = lib.rr_exact_search_f32_tc_workspace_bytes(self.n, q, top_k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        _lib.call("rr_exact_search_f32_tc", self.f32.data_ptr(), inv.data_ptr(), self.n, self.dim, tptr, tm, tv,
                  qf.data_ptr(), q, top_k, float(min_similarity), self.row_base, score.data_ptr(),
                  idx.data_ptr(), count.data_ptr(), ovf.data_ptr(), flags.data_ptr(), ws.data_ptr(), ws_bytes,
                  _stream())
        if check_overflow and int(ovf.item()) != 0:
            bad = torch.nonzero(flags).flatten()
            self.last_tc_redone = int(bad.numel())
            i2, s2, c2 = cuda_cores(qf[bad].contiguous())
            idx[bad] = i2
            score[bad] = s2
            count[bad] = c2
        elif check_overflow:
            self.last_tc_redone = 0
        return idx, score, count

    def search_int8_exact(self, queries_i8: ArrayLike, top_k: int, tag_mask: int = 0,
                          tag_value: int = 0, use_tc: Optional[bool] = None
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Exact int8 x int8 -> int32 search (BASELINE config 4).  -> (idx, score int32)."""
        self._activate()
        if self.int8 is None:
            raise _lib.RadiantB200Error("index has no int8 rows")
        if self.row_base + self.n >= (1 << 32):
            raise ValueError("global rows do not fit the 32-bit row field of the int32-score merge")
        qi = to_device(queries_i8, self.device, torch.int8)
        if qi.ndim == 1:
            qi = qi[None, :]
        q = qi.shape[0]
        score = torch.empty((q, top_k), dtype=torch.int32, device=self.device)
        idx = torch.empty((q, top_k), dtype=torch.int64, device=self.device)
        lib = _lib.load()
        tptr, tm, tv = self._tag_args(tag_mask, tag_value)
        use_tc = (self.dim % 128 == 0 and 128 <= self.dim <= 1024 and q >= 8 and self.n >= 4096
                  if use_tc is None else use_tc)
        def dp4a(q_rows):
            qn = q_rows.shape[0]
            s2 = torch.empty((qn, top_k), dtype=torch.int32, device=self.device)
            i2 = torch.empty((qn, top_k), dtype=torch.int64, device=self.device)
            ws_bytes = lib.rr_int8_search_topk_workspace_bytes(self.n, qn, top_k)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            _lib.call("rr_int8_search_topk", self.int8.data_ptr(), self.n, self.dim, tptr, tm, tv,
                      q_rows.data_ptr(), qn, top_k, self.row_base, s2.data_ptr(), i2.data_ptr(),
                      ws.data_ptr(), ws_bytes, _stream())
            return i2, s2

        if not use_tc:
            return dp4a(qi)
        ovf = torch.zeros(1, dtype=torch.int32, device=self.device)
        flags = torch.empty((q,), dtype=torch.uint8, device=self.device)
        ws_bytes = lib.rr_tc_search_workspace_bytes(self.n, q, top_k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        _lib.call("rr_int8_search_topk_tc", self.int8.data_ptr(), self.n, self.dim, tptr, tm, tv,
                  qi.data_ptr(), q, top_k, self.row_base, score.data_ptr(), idx.data_ptr(),
                  ovf.data_ptr(), flags.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        if int(ovf.item()) != 0:  # redo only the queries whose list overflowed, on the CUDA-core path
            bad = torch.nonzero(flags).flatten()
            self.last_tc_redone = int(bad.numel())
            i2, s2 = dp4a(qi[bad].contiguous())
            idx[bad] = i2
            score[bad] = s2
        return idx, score

    def quantize_int8_queries(self, queries: ArrayLike) -> torch.Tensor:
        """Query-side int8 codes for the symmetric extension mode (same affine map)."""
        self._activate()
        qf = to_device(queries, self.device, torch.float32)
        if qf.ndim == 1:
            qf = qf[None, :]
        out = torch.empty(qf.shape, dtype=torch.int8, device=self.device)
        _lib.call("rr_quantize_int8", qf.data_ptr(), qf.shape[0], self.dim, self.ranges.data_ptr(),
                  out.data_ptr(), _stream())
        return out


# ---- synthetic rows generated on device (bench / tests) -----------------------------

def synth_rows_device(row_start: int, n_rows: int, dim: int, seed: int,
                      device: Union[int, torch.device] = 0) -> torch.Tensor:
    dev = torch.device("cuda", device) if isinstance(device, int) else device
    _lib.init(dev.index or 0)
    out = torch.empty((n_rows, dim), dtype=torch.float32, device=dev)
    _lib.call("rr_synth_rows_f32", out.data_ptr(), row_start, n_rows, dim, seed, value_shift(dim), _stream())
    return out


def synth_query_rows_device(q_start: int, n_q: int, dim: int, seed: int, n_corpus: int,
                            device: Union[int, torch.device] = 0) -> torch.Tensor:
    dev = torch.device("cuda", device) if isinstance(device, int) else device
    _lib.init(dev.index or 0)
    out = torch.empty((n_q, dim), dtype=torch.float32, device=dev)
    _lib.call("rr_synth_query_rows_f32", out.data_ptr(), q_start, n_q, dim, seed, n_corpus,
              value_shift(dim), _stream())
    return out
