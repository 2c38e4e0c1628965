"""BM25 index served from the GPU, behind the reference's ``BM25Index`` /
``PersistentBM25Index`` interface (reference radiant/storage/bm25_index.py).

Host side (this file, Python like the reference): tokenizer :50-58, the document /
df / idf bookkeeping of ``_rebuild_index`` :100-137, ``add_document`` :139-180 (with its
stale-idf behaviour: an incremental add only refreshes the idf of the new document's
own terms), ``remove_document`` :182-216, JSON round trip :275-327 and the persistent,
thread-safe wrapper :330-709.

Device side (``Bm25DeviceIndex``): a tile-sharded inverted CSR with one float64 impact
per posting; ``search`` of :218-270 becomes ``rr_bm25_topk``.  The device index is a
derived cache: it is rebuilt lazily from the host bookkeeping after any mutation
(idf/avgdl are COPIED from the host tables, never recomputed on device - SURVEY.md R7).
"""

from __future__ import annotations

import gzip
import json
import logging
import os
import threading
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

from . import _lib
from .base import StoredDoc
from .index import _stream, to_device

logger = logging.getLogger(__name__)

_INDEX_EXTENSIONS = (".json.gz", ".pickle", ".json", ".pkl")


def _normalize_index_path(path: Path) -> Path:
    """Strip a known index extension so paths with and without one name the same files."""
    name = path.name
    for ext in _INDEX_EXTENSIONS:  # longest first
        if name.endswith(ext):
            return path.parent / name[: -len(ext)]
    return path


def _tokenize(text: str) -> List[str]:
    """lower-case, every non-alphanumeric character becomes a separator, tokens of
    length <= 1 are dropped (Unicode-aware ``str.isalnum``)."""
    out: List[str] = []
    cur: List[str] = []
    for ch in text.lower():
        if ch.isalnum():
            cur.append(ch)
        elif cur:
            if len(cur) > 1:
                out.append("".join(cur))
            cur = []
    if len(cur) > 1:
        out.append("".join(cur))
    return out


class Bm25DeviceIndex:
    """Tile-sharded inverted CSR + per-posting float64 impacts in HBM (one shard).

    tile t owns rows [t*tile_docs, (t+1)*tile_docs); postings are grouped by
    (tile, term); inside a segment rows ascend (or, with bank_interleave=True, cycle over
    row % 16 - the exact kernel's conflict-free order, which the batched path cannot use):
        tile_term_ptr int64 [n_tiles, n_terms+1]
        post_row      int32 [P]   (read as uint32 by the kernels)
        post_impact   f64   [P]   idf_t * (tf*(k1+1)) / (tf + k1*((1-b) + (b*len)/avgdl))
    For the batched filter-and-refine path (csrc/bm25_fast.cu) the <= 32 terms with the largest
    document frequency ("head" terms: most of the posting mass under Zipf) are ALSO kept as
    dense float64 columns per tile:
        head_slot     int32 [n_terms]                 slot of a head term, -1 otherwise
        head_imp      f64   [n_tiles, n_head, tile_docs]  impact, 0.0 where the term is absent
        post_pack     int64 [P]   row-in-tile << 32 | round(impact * 2^fx_shift): what the filter
                                  pass reads (one 8-byte load per posting, integer accumulation)
        head_max      f32   [n_head] upper bound of the head term's float32 impact over all rows
    """

    FAST_MIN_DOCS = 65_536   # below this the exact kernel alone is as fast (few tiles)
    FAST_MAX_TILE = 1024
    FAST_MAX_QLEN = 64       # tokens per query the fixed-point filter sums have headroom for
    MAX_HEAD = 64
    MAX_TABLE_BYTES = 8 << 30  # [n_tiles, n_terms+1] int64 offset table

    def __init__(self, device: torch.device, n_docs: int, n_terms: int, tile_docs: int,
                 tile_term_ptr: torch.Tensor, post_row: torch.Tensor, post_impact: torch.Tensor,
                 row_base: int = 0, head_slot: Optional[torch.Tensor] = None,
                 head_imp: Optional[torch.Tensor] = None, fast_ok: bool = False,
                 post_pack: Optional[torch.Tensor] = None, head_max: Optional[torch.Tensor] = None,
                 fx_shift: int = 0) -> None:
        if row_base + n_docs >= (1 << 32):
            # the (score, row) merge keys carry rows as 32-bit values (csrc/merge.cuh)
            raise ValueError(f"global rows up to {row_base + n_docs} do not fit the 32-bit row field of the BM25 merge")
        self.device = device
        self.n_docs = n_docs
        self.n_terms = n_terms
        self.tile_docs = tile_docs
        self.n_tiles = int(tile_term_ptr.shape[0]) if n_docs else 0
        self.tile_term_ptr = tile_term_ptr
        self.post_row = post_row
        self.post_impact = post_impact
        self.row_base = row_base
        self.head_slot = head_slot
        self.head_imp = head_imp
        self.head_imp_h = (head_imp.to(torch.float16).contiguous()
                           if head_imp is not None and head_imp.numel() else None)
        self.post_pack = post_pack
        self.head_max = head_max
        self.fx_shift = int(fx_shift)
        self.n_head = int(head_imp.shape[1]) if head_imp is not None else 0
        self.fast_ok = bool(fast_ok)
        self.post_term: Optional[torch.Tensor] = None  # int32 [P], kept with keep_structure=True
        self.post_tf: Optional[torch.Tensor] = None    # int32 [P]
        self.doc_len: Optional[torch.Tensor] = None    # int32 [n_docs]
        self.fast_min_docs = self.FAST_MIN_DOCS
        self._inexact: Optional[torch.Tensor] = None  # device u32: flagged queries of unchecked calls
        self.last_flagged = 0                          # queries the last checked call had to redo

    @property
    def n_postings(self) -> int:
        return int(self.post_row.numel())

    @classmethod
    def build(
        cls,
        doc_ptr,
        doc_terms,
        n_terms: int,
        idf,
        avgdl: Optional[float],
        k1: float,
        b: float,
        device=0,
        tile_docs: int = 1024,
        row_base: int = 0,
        doc_len=None,
        bank_interleave: bool = False,
        head_terms: int = 64,
        sharded: bool = False,
        group: Any = None,
        chunk_docs: int = 2_000_000,
        keep_structure: bool = False,
    ) -> "Bm25DeviceIndex":
        """doc_ptr int64 [N+1] / doc_terms int32 [T]: the documents of THIS shard as CSR of
        term ids; idf f64 [n_terms] (0 for unknown terms), avgdl, k1, b: global tables
        copied from the host index (idf=None: document frequencies are counted on device and
        idf is evaluated ON THE HOST with the reference's scalar expression,
        np.log((n - df + 0.5) / (df + 0.5) + 1.0), bm25_index.py:131-135).  doc_len int32 [N] overrides diff(doc_ptr) (the
        reference keeps ``doc_lengths`` separately).  Sorting / counting uses torch on the
        device (index build is not the hot path); impacts come from rr_bm25_impacts.
        bank_interleave: order each (tile, term) segment round-robin over row % 16 (see
        _bank_interleave) instead of ascending rows; disables the batched path.
        head_terms: how many of the most frequent terms also get dense per-tile columns.
        sharded=True (one process per GPU, torch.distributed initialised): these documents are one
        row block of a larger corpus - with idf=None / avgdl=None the document frequencies, the
        document count and the token count are summed over the ranks of `group` first, so every
        shard bakes the GLOBAL idf / avgdl into its impacts."""
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        _lib.init(dev.index or 0)
        ptr = to_device(doc_ptr, dev, torch.int64)
        terms_all = to_device(doc_terms, dev, torch.int32)
        n = int(ptr.numel() - 1)
        v = int(n_terms)
        if n <= 0 or v <= 0 or terms_all.numel() == 0:
            z64 = torch.zeros((1, max(v, 0) + 1), dtype=torch.int64, device=dev)
            return cls(dev, 0, v, tile_docs, z64, torch.zeros(0, dtype=torch.int32, device=dev),
                       torch.zeros(0, dtype=torch.float64, device=dev), row_base)
        n_tiles = (n + tile_docs - 1) // tile_docs
        if n_tiles * (v + 1) * 8 > cls.MAX_TABLE_BYTES:
            # the dense [tile, term] offset table is what limits small tiles on large vocabularies:
            # widen the tiles (the exact kernel takes up to 16384 documents per tile)
            while tile_docs < 16384 and ((n + tile_docs - 1) // tile_docs) * (v + 1) * 8 > cls.MAX_TABLE_BYTES:
                tile_docs *= 2
            n_tiles = (n + tile_docs - 1) // tile_docs
            if n_tiles * (v + 1) * 8 > cls.MAX_TABLE_BYTES:
                raise ValueError(f"BM25 offset table of {n_tiles} tiles x {v} terms exceeds {cls.MAX_TABLE_BYTES >> 30} GiB; "
                                 "shard the documents over more GPUs or prune the vocabulary")
        lens = ptr[1:] - ptr[:-1]
        dlen = lens.to(torch.int32) if doc_len is None else to_device(doc_len, dev, torch.int32)
        # ---- pass 1: (tile, term, row) order.  Tiles are independent, so the documents are sorted
        # in chunks of whole tiles (bounded temporaries: a 12.5M-document shard holds 2.5G tokens)
        chunk = max(tile_docs, (int(chunk_docs) // tile_docs) * tile_docs)
        tile_term_ptr = torch.empty((n_tiles, v + 1), dtype=torch.int64, device=dev)
        df_dev = torch.zeros(v, dtype=torch.int64, device=dev)
        rows_c, term_c, tf_c = [], [], []
        n_post = 0
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            t0, t1 = c0 // tile_docs, (c1 + tile_docs - 1) // tile_docs
            a, b_ = int(ptr[c0].item()), int(ptr[c1].item())
            rows = torch.repeat_interleave(torch.arange(c0, c1, dtype=torch.int64, device=dev), lens[c0:c1])
            # key = (tile * V + term) * tile_docs + row_in_tile  -> sorted by (tile, term, row)
            key = ((rows // tile_docs) * v + terms_all[a:b_].to(torch.int64)) * tile_docs + (rows % tile_docs)
            del rows
            key, _ = torch.sort(key)
            ukey, tf = torch.unique_consecutive(key, return_counts=True)
            del key
            if bank_interleave and ukey.numel() > 1:
                ukey, tf = cls._bank_interleave(ukey, tf, tile_docs)
            tt = ukey // tile_docs  # tile * V + term, non-decreasing
            # offsets of every (tile, term) segment of these tiles: ONE searchsorted over the sorted
            # segment ids; row t of the table is flat[(t - t0)*V : (t - t0)*V + V + 1] (a strided view,
            # no dense bincount / cumsum / gather temporaries)
            flat = torch.searchsorted(tt, torch.arange(t0 * v, t1 * v + 1, dtype=torch.int64, device=dev))
            tile_term_ptr[t0:t1] = flat.as_strided((t1 - t0, v + 1), (v, 1)) + n_post
            del flat
            term = tt % v
            rows_c.append(((tt // v) * tile_docs + (ukey % tile_docs)).to(torch.int32))
            del ukey, tt
            df_dev += torch.bincount(term, minlength=v)
            term_c.append(term.to(torch.int32))
            tf_c.append(tf.to(torch.int32))
            n_post += int(term.numel())
            del term, tf
        del terms_all
        post_row = torch.cat(rows_c) if len(rows_c) > 1 else rows_c[0]
        del rows_c
        if idf is None or avgdl is None:
            df_glob = df_dev.clone()
            stats = torch.tensor([n, int(lens.sum().item()) if doc_len is None else int(dlen.sum(dtype=torch.int64).item())],
                                 dtype=torch.int64, device=dev)
            if sharded:
                import torch.distributed as dist
                if dist.is_initialized() and dist.get_world_size(group) > 1:
                    dist.all_reduce(df_glob, group=group)
                    dist.all_reduce(stats, group=group)
            n_glob, tok_glob = (int(x) for x in stats.cpu().tolist())
            if avgdl is None:
                avgdl = tok_glob / n_glob  # Python ints -> one double divide (bm25_index.py:117)
            if idf is None:
                df = df_glob.cpu().numpy()
                idf = np.zeros(v, dtype=np.float64)
                for t in np.nonzero(df)[0]:
                    d = int(df[t])
                    idf[t] = np.log((n_glob - d + 0.5) / (d + 0.5) + 1.0)
            del df_glob
        post_term = torch.cat(term_c) if len(term_c) > 1 else term_c[0]
        post_tf = torch.cat(tf_c) if len(tf_c) > 1 else tf_c[0]
        del term_c, tf_c
        inst = cls(dev, n, v, tile_docs, tile_term_ptr, post_row, torch.empty(n_post, dtype=torch.float64, device=dev),
                   row_base)
        inst.post_term, inst.post_tf, inst.doc_len = post_term, post_tf, dlen
        inst._fast_layout = bool(not bank_interleave and tile_docs % 128 == 0 and tile_docs <= cls.FAST_MAX_TILE
                                 and n < (1 << 32))
        inst._compute_impacts(idf, float(avgdl), float(k1), float(b))
        # head terms: the most frequent terms that are dense enough to pay for a column
        if inst._fast_layout:
            n_head = min(int(head_terms), cls.MAX_HEAD, v, max(0, _lib.load().rr_bm25_fast_max_head(tile_docs)))
            order = torch.argsort(df_dev, descending=True, stable=True)[:n_head]
            order = order[df_dev[order] * 8 >= n]
            inst.head_slot = torch.full((v,), -1, dtype=torch.int32, device=dev)
            inst.head_slot[order] = torch.arange(int(order.numel()), dtype=torch.int32, device=dev)
            inst.n_head = int(order.numel())
        inst._derive_fast()
        if not keep_structure:  # per-posting term / tf are only needed to refresh the impacts later
            inst.post_term = inst.post_tf = None
        del df_dev
        torch.cuda.current_stream().synchronize()  # temporaries die here
        return inst

    def _compute_impacts(self, idf, avgdl: float, k1: float, b: float) -> None:
        """post_impact for every posting in the reference's operation order (rr_bm25_impacts), from
        the per-posting tf / term / document length and the GIVEN idf table and avgdl."""
        dev = self.device
        idf_t = to_device(idf, dev, torch.float64)
        if idf_t.numel() < self.n_terms:  # tables of a host index that has grown no new terms for these rows
            idf_t = torch.cat([idf_t, torch.zeros(self.n_terms - idf_t.numel(), dtype=torch.float64, device=dev)])
        n_post = int(self.post_row.numel())
        step = 128_000_000
        for o in range(0, n_post, step):
            e = min(n_post, o + step)
            post_idf = idf_t[self.post_term[o:e].to(torch.int64)].contiguous()
            post_len = self.doc_len[self.post_row[o:e].to(torch.int64)].contiguous()
            tf = self.post_tf[o:e].contiguous()
            _lib.call("rr_bm25_impacts", tf.data_ptr(), post_len.data_ptr(), post_idf.data_ptr(), e - o, k1, b, avgdl,
                      self.post_impact[o:e].data_ptr(), _stream())
            del post_idf, post_len, tf

    def _derive_fast(self) -> None:
        """The arrays the batched path reads, derived from post_impact: fixed-point packed postings,
        dense head columns and their per-term maxima; decides whether the path is usable."""
        cls = type(self)
        dev, n_post, tile_docs = self.device, int(self.post_row.numel()), self.tile_docs
        self.fast_ok = False
        self.post_pack = self.head_imp = self.head_max = self.head_imp_h = None
        self.fx_shift = 0
        if not getattr(self, "_fast_layout", False) or n_post == 0:
            return
        lo, hi = torch.aminmax(self.post_impact)
        finite = bool(torch.isfinite(self.post_impact).all().item())
        # the filter needs positive impacts well inside the float32 range ...
        if not (finite and float(lo) >= 2.0 ** -100 and float(hi) <= 2.0 ** 100):
            return
        # ... and a fixed-point scale under which 64 tokens of the largest impact stay below 2^30
        fx_shift = 30 - int(np.ceil(np.log2(float(hi) * cls.FAST_MAX_QLEN)))
        if not 8 <= fx_shift <= 60:
            return
        post_pack = torch.empty(n_post, dtype=torch.int64, device=dev)
        n_head = self.n_head
        head_imp = torch.zeros((self.n_tiles, max(n_head, 1), tile_docs), dtype=torch.float64, device=dev)
        step = 64_000_000
        for o in range(0, n_post, step):
            e = min(n_post, o + step)
            r64 = self.post_row[o:e].to(torch.int64)
            q_imp = torch.round(self.post_impact[o:e] * float(2.0 ** fx_shift)).to(torch.int64)
            post_pack[o:e] = ((r64 % tile_docs) << 32) | q_imp
            del q_imp
            if n_head:
                slot = self.head_slot[self.post_term[o:e].to(torch.int64)].to(torch.int64)
                sel = torch.nonzero(slot >= 0).flatten()
                rs = r64[sel]
                pos = ((rs // tile_docs) * n_head + slot[sel]) * tile_docs + rs % tile_docs
                head_imp.view(-1)[pos] = self.post_impact[o:e][sel]
                del slot, sel, rs, pos
            del r64
        if n_head:
            # the filter reads the head impacts as float16 (half the shared memory per column, so
            # twice the head terms); the refine reads the float64 ones
            if float(head_imp.max()) >= 60000.0:
                return
            head_imp_h = head_imp.to(torch.float16).contiguous()
            head_max = head_imp_h.amax(dim=(0, 2)).to(torch.float32).contiguous()  # bounds what the filter adds
        else:
            head_imp = head_imp[:, :0, :].contiguous()
            head_imp_h = torch.zeros(8, dtype=torch.float16, device=dev)
            head_max = torch.zeros(1, dtype=torch.float32, device=dev)
        self.head_imp_h = head_imp_h
        self.post_pack, self.head_imp, self.head_max, self.fx_shift, self.fast_ok = post_pack, head_imp, head_max, fx_shift, True

    def refresh(self, idf, avgdl: float, k1: float, b: float) -> None:
        """Re-evaluate every impact from NEW idf / avgdl tables without re-sorting the postings
        (an ``add_document`` changes avgdl, i.e. every impact, but not the postings of the documents
        already indexed): O(postings) element-wise work instead of a rebuild.  Needs
        ``keep_structure=True`` at build time."""
        if self.post_term is None:
            raise _lib.RadiantB200Error("this index was built without keep_structure=True: rebuild it instead")
        if torch.cuda.current_device() != (self.device.index or 0):
            torch.cuda.set_device(self.device)
        self._compute_impacts(idf, float(avgdl), float(k1), float(b))
        self._derive_fast()

    SMEM_BANKS64 = 16  # 8-byte accumulators: 16 bank pairs per half-warp

    @classmethod
    def _bank_interleave(cls, ukey: torch.Tensor, tf: torch.Tensor, tile_docs: int):
        """Reorder the postings of every (tile, term) segment so that consecutive postings
        cycle through row % 16: the exact kernel's shared-memory read-modify-write of the
        float64 accumulators (16 lanes of a half-warp = 16 consecutive postings) then hits
        16 different bank pairs instead of a random multiset (~3-way conflicts).
        ukey = (tile*V + term) * tile_docs + row_in_tile, sorted."""
        nb = cls.SMEM_BANKS64
        dev = ukey.device
        seg = ukey // tile_docs
        bank = (ukey % tile_docs) % nb
        grp, perm1 = torch.sort(seg * nb + bank, stable=True)  # rows stay ascending inside a group
        del seg, bank
        _, gcount = torch.unique_consecutive(grp, return_counts=True)
        gstart = torch.cumsum(gcount, 0) - gcount
        rank = torch.arange(grp.numel(), dtype=torch.int64, device=dev) - torch.repeat_interleave(gstart, gcount)
        del gcount, gstart
        per_bank = (tile_docs + nb - 1) // nb + 1
        key3 = ((grp // nb) * per_bank + rank) * nb + (grp % nb)  # (segment, rank, bank)
        del grp, rank
        _, perm2 = torch.sort(key3)
        del key3
        order = perm1[perm2]
        return ukey[order].contiguous(), tf[order].contiguous()

    # ---- search ----------------------------------------------------------------------
    def _out(self, q: int, k: int, out_words: Optional[torch.Tensor]):
        """Result tensors; with out_words (int64 [2, q, k]) scores and rows are the two planes of ONE
        buffer, which is what the sharded exchange gathers."""
        if out_words is None:
            return (torch.empty((q, k), dtype=torch.float64, device=self.device),
                    torch.empty((q, k), dtype=torch.int64, device=self.device))
        return out_words[0].view(torch.float64), out_words[1]

    def _search_exact(self, qt: torch.Tensor, k: int, out_words: Optional[torch.Tensor] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        q, ql = qt.shape
        score, idx = self._out(q, k, out_words)
        count = torch.empty((q,), dtype=torch.int32, device=self.device)
        lib = _lib.load()
        ws_bytes = lib.rr_bm25_topk_workspace_bytes(self.n_tiles, q, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        _lib.call("rr_bm25_topk", self.tile_term_ptr.data_ptr(), self.post_row.data_ptr(),
                  self.post_impact.data_ptr(), self.n_tiles, self.tile_docs, self.n_terms, self.n_docs,
                  qt.data_ptr(), q, ql, k, self.row_base, score.data_ptr(), idx.data_ptr(),
                  count.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        return idx, score, count

    def _search_fast(self, qt: torch.Tensor, k: int, counter: torch.Tensor,
                     out_words: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        q, ql = qt.shape
        score, idx = self._out(q, k, out_words)
        count = torch.empty((q,), dtype=torch.int32, device=self.device)
        flags = torch.empty((q,), dtype=torch.uint8, device=self.device)
        lib = _lib.load()
        ws_bytes = lib.rr_bm25_fast_workspace_bytes(self.n_tiles, self.tile_docs, self.n_docs, q, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        _lib.call("rr_bm25_topk_fast", self.tile_term_ptr.data_ptr(), self.post_row.data_ptr(),
                  self.post_impact.data_ptr(), self.post_pack.data_ptr(), self.fx_shift,
                  self.head_slot.data_ptr(), self.head_imp.data_ptr() if self.n_head else None,
                  self.head_imp_h.data_ptr() if self.n_head else None,
                  self.head_max.data_ptr(), self.n_head, self.n_tiles,
                  self.tile_docs, self.n_terms, self.n_docs, qt.data_ptr(), q, ql, k, self.row_base,
                  score.data_ptr(), idx.data_ptr(), count.data_ptr(), flags.data_ptr(), counter.data_ptr(),
                  ws.data_ptr(), ws_bytes, _stream())
        return idx, score, count, flags

    def uses_fast_path(self, q: int, k: int, q_len: int = 8) -> bool:
        return bool(self.fast_ok and self.n_docs >= self.fast_min_docs and q >= 1 and k <= 1000
                    and q_len <= self.FAST_MAX_QLEN)

    def search_batch_into(self, q_terms, k: int, check: bool = True) -> torch.Tensor:
        """``search_batch`` with scores and rows written into one int64 [2, Q, k] buffer (plane 0 =
        float64 score bits, plane 1 = global rows): the form the sharded exchange moves in ONE collective."""
        qt = to_device(q_terms, self.device, torch.int32)
        if qt.ndim == 1:
            qt = qt[None, :]
        words = torch.empty((2, qt.shape[0], k), dtype=torch.int64, device=self.device)
        self.search_batch(qt, k, check=check, out_words=words)
        return words

    def search_batch(self, q_terms, k: int, check: bool = True, exact: bool = False,
                     out_words: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """q_terms int32 [Q, L] term ids in query-token order, -1 = unknown / padding.
        -> (idx int64 [Q,k] (-1 padded), score f64 [Q,k], count int32 [Q]);
        order (score desc, row asc), score > 0 only; float64 scores equal the reference's bits.

        Large indexes run the batched filter-and-refine path (csrc/bm25_fast.cu), which proves
        the exactness of every query it answers and flags the ones it cannot (degenerate ties at
        the bound); a checked call (one device sync at its end) redoes those with the exact
        kernel.  check=False skips the sync - for CUDA-graph capture - and accumulates the number
        of flagged queries in ``inexact_total()`` for the caller to verify.  exact=True forces the
        exact kernel for the whole batch."""
        if torch.cuda.current_device() != (self.device.index or 0):
            torch.cuda.set_device(self.device)
        qt = to_device(q_terms, self.device, torch.int32)
        if qt.ndim == 1:
            qt = qt[None, :]
        q, ql = qt.shape
        if exact or ql == 0 or self.n_docs == 0 or not self.uses_fast_path(q, k, ql):
            return self._search_exact(qt, k, out_words)
        if not check:
            if self._inexact is None:
                self._inexact = torch.zeros(1, dtype=torch.int32, device=self.device)
            idx, score, count, _flags = self._search_fast(qt, k, self._inexact, out_words)
            return idx, score, count
        counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        idx, score, count, flags = self._search_fast(qt, k, counter, out_words)
        self.last_flagged = int(counter.item())
        if self.last_flagged:
            bad = torch.nonzero(flags).flatten()
            r_idx, r_score, r_count = self._search_exact(qt[bad].contiguous(), k)
            idx[bad] = r_idx
            score[bad] = r_score
            count[bad] = r_count
        return idx, score, count

    def inexact_total(self) -> int:
        """Queries flagged by ``check=False`` calls since the last reset (0 = all exact)."""
        return 0 if self._inexact is None else int(self._inexact.item())

    def inexact_reset(self) -> None:
        if self._inexact is not None:
            self._inexact.zero_()


def synth_zipf_corpus_device(n_docs: int, n_terms: int, seed: int, mean_len: int = 200,
                             device=0, row_start: int = 0, chunk_docs: int = 4_000_000
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic Zipf documents [row_start, row_start + n_docs) generated ON DEVICE, bit-identical
    to ``synthetic.zipf_corpus`` restricted to those rows (bench / parity only).
    Document d owns the GLOBAL token positions [P(d), P(d+1)), P = prefix sum of all document
    lengths from document 0, so a shard's tokens do not depend on how the corpus is sharded.
    -> (doc_ptr int64 [n_docs+1] local offsets, tokens int32 [total])."""
    from . import synthetic

    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    _lib.init(dev.index or 0)
    pos0 = 0
    for lo in range(0, row_start, chunk_docs):  # global position of this shard's first token
        m = min(chunk_docs, row_start - lo)
        tmp = torch.empty(m, dtype=torch.int32, device=dev)
        _lib.call("rr_synth_doc_lengths", tmp.data_ptr(), lo, m, seed, mean_len, _stream())
        pos0 += int(tmp.sum(dtype=torch.int64).item())
    lens = torch.empty(n_docs, dtype=torch.int32, device=dev)
    _lib.call("rr_synth_doc_lengths", lens.data_ptr(), row_start, n_docs, seed, mean_len, _stream())
    ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens.to(torch.int64), 0, out=ptr[1:])
    total = int(ptr[-1].item())
    cdf_d = torch.from_numpy(synthetic.zipf_cdf_u32(n_terms).view(np.int32)).to(dev)
    toks = torch.empty(total, dtype=torch.int32, device=dev)
    _lib.call("rr_synth_zipf_tokens", toks.data_ptr(), pos0, total, seed, cdf_d.data_ptr(), n_terms, _stream())
    return ptr, toks


class BM25Index:
    """Host bookkeeping with the reference's attribute names; ``search`` runs on the GPU.

    Attributes kept for callers that reach in (reference orchestrator / persistent
    wrapper): doc_ids, doc_tokens, doc_id_set, doc_id_to_idx, k1, b, avgdl, doc_lengths,
    idf, term_doc_freqs, dirty, needs_rebuild.
    """

    def __init__(
        self,
        doc_ids: Optional[List[str]] = None,
        doc_tokens: Optional[List[List[str]]] = None,
        k1: float = 1.5,
        b: float = 0.75,
        needs_rebuild: bool = True,
        device: int = 0,
        tile_docs: int = 1024,
    ) -> None:
        self.doc_ids: List[str] = list(doc_ids or [])
        self.doc_tokens: List[List[str]] = list(doc_tokens or [])
        self.doc_id_set: Set[str] = set(self.doc_ids)
        self.doc_id_to_idx: Dict[str, int] = {d: i for i, d in enumerate(self.doc_ids)}
        self.k1 = k1
        self.b = b
        self.avgdl = 0.0
        self.doc_lengths: List[int] = []
        self.idf: Dict[str, float] = {}
        self.term_doc_freqs: Dict[str, int] = {}
        self.dirty = False
        self.needs_rebuild = needs_rebuild
        self._device = device
        self._tile_docs = tile_docs
        self._vocab: Dict[str, int] = {}
        self._doc_term_ids: List[np.ndarray] = []
        self._gpu: Optional[Bm25DeviceIndex] = None
        self._gpu_params: Tuple[float, float] = (k1, b)
        self._gpu_docs = 0              # documents covered by self._gpu
        self._gpu_version = -1          # host-table version its impacts were computed from
        self._delta: Optional[Bm25DeviceIndex] = None   # documents added since (rows _gpu_docs ..)
        self._delta_version = -1
        self._tables_version = 0
        self.delta_rebuild_fraction = 0.125   # a delta larger than this share of the main index triggers a rebuild
        for toks in self.doc_tokens:
            self._doc_term_ids.append(self._term_ids(toks, create=True))
        if self.doc_ids and self.needs_rebuild:
            self._rebuild_index()

    # ---- vocabulary -----------------------------------------------------------------
    def _term_ids(self, tokens: Sequence[str], create: bool) -> np.ndarray:
        vocab = self._vocab
        if create:
            return np.fromiter((vocab.setdefault(t, len(vocab)) for t in tokens), dtype=np.int32,
                               count=len(tokens))
        return np.fromiter((vocab.get(t, -1) for t in tokens), dtype=np.int32, count=len(tokens))

    # ---- bookkeeping (same arithmetic as the reference) --------------------------------
    def _rebuild_index(self) -> None:
        self._gpu = None
        if not self.doc_ids:
            self.avgdl = 0.0
            self.doc_lengths = []
            self.idf = {}
            self.term_doc_freqs = {}
            self.needs_rebuild = False
            return
        self.doc_id_set = set(self.doc_ids)
        self.doc_id_to_idx = {d: i for i, d in enumerate(self.doc_ids)}
        self.doc_lengths = [len(t) for t in self.doc_tokens]
        self.avgdl = sum(self.doc_lengths) / len(self.doc_lengths) if self.doc_lengths else 0.0
        df: Dict[str, int] = {}
        for toks in self.doc_tokens:
            for t in set(toks):
                df[t] = df.get(t, 0) + 1
        self.term_doc_freqs = df
        n = len(self.doc_ids)
        self.idf = {t: np.log((n - d + 0.5) / (d + 0.5) + 1.0) for t, d in df.items()}
        self.needs_rebuild = False

    def add_document(self, doc_id: str, tokens: List[str]) -> bool:
        if doc_id in self.doc_id_set:
            return False
        self.doc_ids.append(doc_id)
        self.doc_tokens.append(tokens)
        self._doc_term_ids.append(self._term_ids(tokens, create=True))
        self.doc_id_set.add(doc_id)
        self.doc_id_to_idx[doc_id] = len(self.doc_ids) - 1
        self.doc_lengths.append(len(tokens))
        n = len(self.doc_ids)
        self.avgdl = (self.avgdl * (n - 1) + len(tokens)) / n
        # only the new document's own terms get a fresh idf (with the current n)
        for t in set(tokens):
            d = self.term_doc_freqs.get(t, 0) + 1
            self.term_doc_freqs[t] = d
            self.idf[t] = np.log((n - d + 0.5) / (d + 0.5) + 1.0)
        self.dirty = True
        # device side: the postings of the documents already indexed do not change, only their
        # impacts (avgdl moved, the idf of this document's terms moved): refreshed lazily, the new
        # documents go to a small delta index - no O(N) rebuild per add
        self._tables_version += 1
        return True

    def remove_document(self, doc_id: str) -> bool:
        if doc_id not in self.doc_id_set:
            return False
        i = self.doc_id_to_idx.pop(doc_id)
        del self.doc_ids[i]
        del self.doc_tokens[i]
        del self._doc_term_ids[i]
        self.doc_id_set.discard(doc_id)
        for other, j in self.doc_id_to_idx.items():
            if j > i:
                self.doc_id_to_idx[other] = j - 1
        self.needs_rebuild = True
        self.dirty = True
        self._gpu = None
        return True

    # ---- device index ---------------------------------------------------------------
    def _csr(self, lo: int, hi: int):
        docs = self._doc_term_ids[lo:hi]
        lens = np.fromiter((a.size for a in docs), dtype=np.int64, count=len(docs))
        ptr = np.zeros(lens.size + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        terms = np.concatenate(docs) if docs else np.zeros(0, dtype=np.int32)
        return ptr, terms

    def _idf_table(self, v: int) -> np.ndarray:
        idf = np.zeros(v, dtype=np.float64)
        for t, val in self.idf.items():
            tid = self._vocab.get(t)
            if tid is not None and tid < v:
                idf[tid] = float(val)
        return idf

    def device_indexes(self) -> Tuple[Bm25DeviceIndex, Optional[Bm25DeviceIndex]]:
        """(main, delta): the GPU index over rows [0, n_main) and, after incremental adds, a small
        one over rows [n_main, n) - both with impacts from the CURRENT host tables (idf / avgdl are
        copied, never recomputed: the reference's stale-idf state is what gets scored).

        A full build sorts every posting (O(N)); an ``add_document`` does not need one: the main
        index re-evaluates its impacts in place (``refresh``, element-wise) and the new documents are
        indexed on their own.  Removals, parameter changes and a delta beyond
        ``delta_rebuild_fraction`` of the main index rebuild."""
        if self.needs_rebuild:
            self._rebuild_index()
        n = len(self.doc_ids)
        grown = n - self._gpu_docs
        full = (self._gpu is None or self._gpu_params != (self.k1, self.b) or grown < 0
                or grown > max(1024, int(self._gpu_docs * self.delta_rebuild_fraction))
                or self._gpu.post_term is None)
        if full:
            ptr, terms = self._csr(0, n)
            v = len(self._vocab)
            self._gpu = Bm25DeviceIndex.build(
                ptr, terms, v, self._idf_table(v), float(self.avgdl), float(self.k1), float(self.b),
                device=self._device, tile_docs=self._tile_docs,
                doc_len=np.asarray(self.doc_lengths, dtype=np.int32), keep_structure=True)
            self._gpu_params = (self.k1, self.b)
            self._gpu_docs, self._gpu_version = n, self._tables_version
            self._delta, self._delta_version = None, -1
            return self._gpu, None
        if self._gpu_version != self._tables_version:
            self._gpu.refresh(self._idf_table(self._gpu.n_terms), float(self.avgdl), float(self.k1), float(self.b))
            self._gpu_version = self._tables_version
        if grown == 0:
            self._delta = None
            return self._gpu, None
        if self._delta is None or self._delta_version != self._tables_version:
            ptr, terms = self._csr(self._gpu_docs, n)
            v = len(self._vocab)
            self._delta = Bm25DeviceIndex.build(
                ptr, terms, v, self._idf_table(v), float(self.avgdl), float(self.k1), float(self.b),
                device=self._device, tile_docs=self._tile_docs, row_base=self._gpu_docs,
                doc_len=np.asarray(self.doc_lengths[self._gpu_docs:], dtype=np.int32))
            self._delta_version = self._tables_version
        return self._gpu, self._delta

    def device_index(self) -> Bm25DeviceIndex:
        """The GPU index over ALL current documents (forces a full build if a delta is pending)."""
        main, delta = self.device_indexes()
        if delta is not None:
            self._gpu = None
            main, _ = self.device_indexes()
        return main

    def _device_search(self, qt: np.ndarray, k: int):
        main, delta = self.device_indexes()
        idx, score, count = main.search_batch(qt, k)
        if delta is None or delta.n_docs == 0:
            return idx, score, count
        i2, s2, _c2 = delta.search_batch(qt, k)
        q = idx.shape[0]
        out_s = torch.empty((q, k), dtype=torch.float64, device=idx.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=idx.device)
        out_c = torch.empty((q,), dtype=torch.int32, device=idx.device)
        s_all = torch.cat([score, s2], dim=1).contiguous()
        i_all = torch.cat([idx, i2], dim=1).contiguous()
        _lib.call("rr_merge_scores_f64", s_all.data_ptr(), i_all.data_ptr(), q, 2 * k, k, out_s.data_ptr(),
                  out_i.data_ptr(), out_c.data_ptr(), _stream())
        return out_i, out_s, out_c

    def _query_term_ids(self, query_tokens: Sequence[str]) -> np.ndarray:
        ids = self._term_ids(query_tokens, create=False)
        # "if term not in self.idf: continue" - a vocabulary term whose idf was dropped by a
        # rebuild (all its documents removed) is unknown too
        for i, t in enumerate(query_tokens):
            if ids[i] >= 0 and t not in self.idf:
                ids[i] = -1
        return ids

    def search_batch(self, queries_tokens: Sequence[Sequence[str]], top_k: int
                     ) -> List[List[Tuple[str, float]]]:
        if self.needs_rebuild:
            self._rebuild_index()
        nq = len(queries_tokens)
        if not self.doc_ids or nq == 0:
            return [[] for _ in range(nq)]
        width = max(1, max(len(q) for q in queries_tokens))
        qt = np.full((nq, width), -1, dtype=np.int32)
        for i, toks in enumerate(queries_tokens):
            if toks:
                qt[i, : len(toks)] = self._query_term_ids(toks)
        k = max(1, min(int(top_k), _lib.RR_MAX_K))
        if k != int(top_k):
            logger.warning(f"BM25 top_k={top_k} outside [1, {_lib.RR_MAX_K}]: {k} results per query are returned")
        idx, score, count = self._device_search(qt, k)
        idx_h, score_h, count_h = idx.cpu().tolist(), score.cpu().tolist(), count.cpu().tolist()
        out: List[List[Tuple[str, float]]] = []
        for i in range(nq):
            m = min(count_h[i], top_k)
            out.append([(self.doc_ids[r], float(s)) for r, s in zip(idx_h[i][:m], score_h[i][:m])])
        return out

    def search(self, query_tokens: List[str], top_k: int) -> List[Tuple[str, float]]:
        """[(doc_id, score)] by (score desc, row asc); scores are the reference's float64
        values bit for bit."""
        if not query_tokens or top_k <= 0:
            return []
        return self.search_batch([query_tokens], top_k)[0]

    def __len__(self) -> int:
        return len(self.doc_ids)

    # ---- serialisation (same JSON schema as the reference, version 2) -----------------------
    def to_dict(self) -> Dict[str, Any]:
        return {"version": 2, "doc_ids": self.doc_ids, "doc_tokens": self.doc_tokens,
                "k1": self.k1, "b": self.b}

    @classmethod
    def from_dict(cls, data: Dict[str, Any], device: int = 0) -> "BM25Index":
        index = cls(doc_ids=data.get("doc_ids", []), doc_tokens=data.get("doc_tokens", []),
                    k1=data.get("k1", 1.5), b=data.get("b", 0.75), needs_rebuild=True, device=device)
        if index.doc_ids and index.needs_rebuild:
            index._rebuild_index()
        return index

    @classmethod
    def from_reference(cls, ref_index, device: int = 0, tile_docs: int = 1024) -> "BM25Index":
        """Adopt a live reference ``BM25Index`` INCLUDING its current idf / avgdl / df tables
        (they may be stale after incremental adds; they are copied, not recomputed)."""
        inst = cls(doc_ids=list(ref_index.doc_ids), doc_tokens=list(ref_index.doc_tokens),
                   k1=ref_index.k1, b=ref_index.b, needs_rebuild=False, device=device,
                   tile_docs=tile_docs)
        inst.avgdl = float(ref_index.avgdl)
        inst.doc_lengths = list(ref_index.doc_lengths)
        inst.idf = dict(ref_index.idf)
        inst.term_doc_freqs = dict(ref_index.term_doc_freqs)
        inst.needs_rebuild = bool(ref_index.needs_rebuild)
        return inst


class PersistentBM25Index:
    """Thread-safe, persistent wrapper with the reference's method set
    (reference bm25_index.py:330-709); ``search`` hydrates ``StoredDoc``s from the store."""

    def __init__(self, config: Any, store: Any, device: int = 0) -> None:
        self._config = config
        self._store = store
        self._device = device
        self._lock = threading.RLock()
        self._index: Optional[BM25Index] = None
        self._unsaved_count = 0
        Path(config.index_path).parent.mkdir(parents=True, exist_ok=True)

    def _paths(self) -> Tuple[Path, Path]:
        base = _normalize_index_path(Path(self._config.index_path))
        return base.with_suffix(".json.gz"), base.with_suffix(".tmp.gz")

    def _load_or_create_index(self) -> BM25Index:
        json_path, _ = self._paths()
        if json_path.exists():
            try:
                with gzip.open(json_path, "rt", encoding="utf-8") as f:
                    data = json.load(f)
                index = BM25Index.from_dict(data, device=self._device)
                index.k1 = self._config.k1
                index.b = self._config.b
                logger.info(f"Loaded BM25 index (JSON) with {len(index)} documents")
                return index
            except Exception as e:
                logger.warning(f"Failed to load JSON BM25 index: {e}")
        # legacy pickle files of the reference hold reference-class instances; they are
        # migrated by the reference itself, not unpickled here (out of scope, security)
        return BM25Index(k1=self._config.k1, b=self._config.b, device=self._device)

    @property
    def index(self) -> BM25Index:
        if self._index is None:
            with self._lock:
                if self._index is None:
                    self._index = self._load_or_create_index()
        return self._index

    def save(self) -> bool:
        with self._lock:
            if self._index is None or not self._index.dirty:
                return True
            json_path, tmp_path = self._paths()
            try:
                with gzip.open(tmp_path, "wt", encoding="utf-8") as f:
                    json.dump(self._index.to_dict(), f, separators=(",", ":"))
                os.replace(tmp_path, json_path)
                self._index.dirty = False
                self._unsaved_count = 0
                return True
            except Exception as e:
                logger.error(f"Failed to save BM25 index: {e}")
                try:
                    if tmp_path.exists():
                        tmp_path.unlink()
                except Exception:
                    pass
                return False

    def _maybe_auto_save(self) -> None:
        if self._unsaved_count >= self._config.auto_save_threshold:
            self.save()

    def add_document(self, doc_id: str, content: str) -> bool:
        tokens = _tokenize(content)
        if not tokens:
            return False
        with self._lock:
            added = self.index.add_document(doc_id, tokens)
            if added:
                self._unsaved_count += 1
                self._maybe_auto_save()
            return added

    def add_documents_batch(self, documents: List[Tuple[str, str]]) -> int:
        added = 0
        with self._lock:
            for doc_id, content in documents:
                tokens = _tokenize(content)
                if tokens and self.index.add_document(doc_id, tokens):
                    added += 1
            if added:
                self._unsaved_count += added
                self._maybe_auto_save()
        return added

    def remove_document(self, doc_id: str) -> bool:
        with self._lock:
            removed = self.index.remove_document(doc_id)
            if removed:
                self._unsaved_count += 1
                self._maybe_auto_save()
            return removed

    def _hydrate(self, results: List[Tuple[str, float]]) -> List[Tuple[StoredDoc, float]]:
        out: List[Tuple[StoredDoc, float]] = []
        for doc_id, score in results:
            doc = self._store.get_doc(doc_id)
            if doc is not None:  # documents missing from the store are dropped silently
                out.append((doc, score))
        return out

    def search(self, query: str, top_k: int) -> List[Tuple[StoredDoc, float]]:
        tokens = _tokenize(query)
        if not tokens:
            return []
        with self._lock:
            results = self.index.search(tokens, top_k)
        return self._hydrate(results)

    def search_batch(self, queries: List[str], top_k: int) -> List[List[Tuple[StoredDoc, float]]]:
        """New batched surface: one kernel launch for all queries."""
        toks = [_tokenize(q) for q in queries]
        with self._lock:
            results = self.index.search_batch(toks, top_k)
        return [self._hydrate(r) for r in results]

    def build_from_store(self, limit: int = 0) -> int:
        max_docs = limit or self._config.max_documents
        doc_ids = self._store.list_doc_ids_with_embeddings(limit=max_docs)
        documents: List[Tuple[str, str]] = []
        for doc_id in doc_ids:
            doc = self._store.get_doc(doc_id)
            if doc and doc.content:
                documents.append((doc_id, doc.content))
        with self._lock:
            self._index = BM25Index(k1=self._config.k1, b=self._config.b, device=self._device)
            for doc_id, content in documents:
                tokens = _tokenize(content)
                if tokens:
                    self._index.add_document(doc_id, tokens)  # incremental path, as the reference
            self._index.dirty = True
            self.save()
        return len(self._index)

    def sync_with_store(self) -> Tuple[int, int]:
        store_ids = set(self._store.list_doc_ids_with_embeddings(limit=self._config.max_documents))
        with self._lock:
            index_ids = set(self.index.doc_ids)
            removed = sum(1 for d in index_ids - store_ids if self.index.remove_document(d))
            added = 0
            for doc_id in store_ids - index_ids:
                doc = self._store.get_doc(doc_id)
                if doc and doc.content:
                    tokens = _tokenize(doc.content)
                    if tokens and self.index.add_document(doc_id, tokens):
                        added += 1
            if added or removed:
                self._index.dirty = True
                self.save()
        return added, removed

    def clear(self) -> None:
        with self._lock:
            self._index = BM25Index(k1=self._config.k1, b=self._config.b, device=self._device)
            self._unsaved_count = 0
            json_path, _ = self._paths()
            for p in (json_path, json_path.with_name(json_path.name.replace(".json.gz", ".pkl"))):
                if p.exists():
                    try:
                        p.unlink()
                    except Exception as e:
                        logger.warning(f"Failed to delete index file {p}: {e}")

    def __len__(self) -> int:
        return len(self.index)

    def get_stats(self) -> Dict[str, Any]:
        with self._lock:
            idx = self.index
            json_path, _ = self._paths()
            return {
                "document_count": len(idx),
                "unique_terms": len(idx.idf),
                "avg_doc_length": idx.avgdl,
                "k1": idx.k1,
                "b": idx.b,
                "dirty": idx.dirty,
                "needs_rebuild": idx.needs_rebuild,
                "index_path": str(json_path),
                "storage_format": "json.gz" if json_path.exists() else "json.gz (pending)",
            }
