"""B200VectorStore: the reference's ``BaseVectorStore`` served from a GPU index.

Drop-in for RedisVectorStore / ChromaVectorStore / PgVectorStore behind
``DenseRetrievalAgent`` and the orchestrator (constructor injection):

* ``retrieve_by_embedding``           -> exact float32 cosine scan on device, semantics
  of reference radiant/storage/redis_store.py:863-952 (the exact linear scan);
* ``retrieve_by_embedding_quantized`` -> the two-stage flow of reference
  redis_store.py:757-861 with stage 1 done as the documented exact Hamming search;
* ``retrieve_batch`` / ``retrieve_batch_quantized`` are the new batched surface.

Document text/metadata I/O is NOT accelerated: documents live in a host dict or in
an ``inner`` store (any reference backend) that this class delegates to; the GPU
holds only codes / int8 / float32 rows and a tag byte per row.
"""

from __future__ import annotations

import logging
import threading
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .base import BaseVectorStore, StoredDoc, normalize_doc_level
from .config import QuantizationConfig
from .index import DenseIndex, LanguageTable, make_tag, tag_predicate

logger = logging.getLogger(__name__)


class B200VectorStore(BaseVectorStore):
    def __init__(
        self,
        embedding_dim: Optional[int] = None,
        device: int = 0,
        quantization: Optional[Any] = None,
        int8_ranges: Optional[np.ndarray] = None,
        inner: Optional[Any] = None,
        max_content_chars: int = 200_000,
    ) -> None:
        """quantization: a QuantizationConfig (this package's or the reference's);
        int8_ranges: [2, D] calibration, or loaded from ``quantization.int8_ranges_file``
        (reference redis_store.py:174-181); inner: optional reference store that keeps
        the documents themselves."""
        self._device = device
        self._quant_config = quantization or QuantizationConfig()
        self._inner = inner
        self._max_chars = max_content_chars
        self._docs: Dict[str, StoredDoc] = {}
        self._row_of: Dict[str, int] = {}
        self._id_of: List[str] = []
        self._langs = LanguageTable()
        self._lock = threading.RLock()
        self._index: Optional[DenseIndex] = None
        self._embedding_dim = embedding_dim
        self._int8_ranges = None
        if int8_ranges is not None:
            self._int8_ranges = np.asarray(int8_ranges, dtype=np.float32)
        elif getattr(self._quant_config, "int8_ranges_file", None):
            try:
                self._int8_ranges = np.load(self._quant_config.int8_ranges_file).astype(np.float32)
            except Exception as e:  # same tolerance as the reference: warn and go on
                logger.warning(f"Failed to load int8 ranges: {e}")
        if embedding_dim:
            self._ensure_index(embedding_dim)

    # ---- index management ----------------------------------------------------------
    def _ensure_index(self, dim: int) -> None:
        """Create the GPU index on first use (name kept from the reference, which
        ``RadiantRAG.clear_index`` calls: radiant/app.py:1320)."""
        with self._lock:
            if self._index is not None:
                if dim != self._index.dim:
                    raise ValueError(f"embedding dim {dim} != index dim {self._index.dim}")
                return
            precision = getattr(self._quant_config, "precision", "both")
            want_int8 = precision in ("int8", "both") and self._int8_ranges is not None
            if precision in ("int8", "both") and self._int8_ranges is None and \
                    getattr(self._quant_config, "enabled", False):
                logger.warning("int8 precision requested without calibration ranges; "
                               "rescoring falls back to float32 rows")
            self._embedding_dim = dim
            self._index = DenseIndex(dim, device=self._device, store_int8=want_int8, store_f32=True,
                                     int8_ranges=self._int8_ranges if want_int8 else None)

    @property
    def index(self) -> Optional[DenseIndex]:
        return self._index

    def ping(self) -> bool:
        try:
            _lib.init(self._device)
            return True if self._inner is None else bool(self._inner.ping())
        except Exception:
            return False

    def make_doc_id(self, content: str, meta: Optional[Dict[str, Any]] = None) -> str:
        return self._default_make_doc_id(content, meta)

    # ---- writes ----------------------------------------------------------------------
    def _prep_doc(self, content: str, meta: Optional[Dict[str, Any]], default_level: str):
        meta = dict(meta or {})
        if len(content) > self._max_chars:
            content = content[: self._max_chars]
            meta["truncated"] = True
        level = str(meta.get("doc_level", default_level))
        lang = str(meta.get("language_code", "en"))
        return content, meta, level, lang

    def _store_doc(self, doc_id: str, content: str, meta: Dict[str, Any]) -> None:
        self._docs[doc_id] = StoredDoc(doc_id=doc_id, content=content, meta=meta)

    def upsert(self, doc_id: str, content: str, embedding: List[float],
               meta: Optional[Dict[str, Any]] = None) -> None:
        self.upsert_batch([{"doc_id": doc_id, "content": content, "embedding": embedding, "meta": meta}])

    def upsert_doc_only(self, doc_id: str, content: str, meta: Optional[Dict[str, Any]] = None) -> None:
        content, meta, _level, _lang = self._prep_doc(content, meta, "parent")
        with self._lock:
            if self._inner is not None:
                self._inner.upsert_doc_only(doc_id, content, meta)
            else:
                self._store_doc(doc_id, content, meta)

    def upsert_batch(self, documents: List[Dict[str, Any]]) -> int:
        """Batch insert/update; unlike the reference's batch path
        (redis_store.py:476-532) the quantised rows ARE written here.
        All-or-nothing: the whole batch is validated and staged first, the GPU append runs, and
        only then are the id <-> row maps and the document store updated - a bad document in the
        middle of a batch leaves the store exactly as it was."""
        if not documents:
            return 0
        with self._lock:
            staged = []           # (doc_id, content, meta, emb, tag)
            dim = self._index.dim if self._index is not None else self._embedding_dim
            for d in documents:
                emb = np.asarray(d["embedding"], dtype=np.float32)
                if emb.ndim != 1 or emb.size == 0:
                    raise ValueError(f"{d.get('doc_id')}: embedding must be a non-empty 1-D vector")
                if dim is None:
                    dim = int(emb.shape[0])
                if emb.shape[0] != dim:
                    raise ValueError(f"{d.get('doc_id')}: embedding dim {emb.shape[0]} != index dim {dim}")
                content, meta, level, lang = self._prep_doc(d["content"], d.get("meta"), "child")
                staged.append((d["doc_id"], content, meta, emb, level, lang))
            self._ensure_index(dim)
            idx = self._index
            if idx.store_int8 and idx.ranges is None:
                raise ValueError("int8 storage needs calibration ranges")
            # rows: existing ids are overwritten in place, new ids appended (last one of a repeated id wins)
            new_pos: Dict[str, int] = {}
            new_rows: List[np.ndarray] = []
            new_tags: List[int] = []
            new_ids: List[str] = []
            updates: List[Tuple[int, np.ndarray, int]] = []
            langs_before = dict(self._langs._ids)
            try:
                for doc_id, _content, _meta, emb, level, lang in staged:
                    tag = make_tag(level, self._langs.id_for(lang, create=True))
                    if doc_id in self._row_of:
                        updates.append((self._row_of[doc_id], emb, tag))
                    elif doc_id in new_pos:
                        new_rows[new_pos[doc_id]] = emb
                        new_tags[new_pos[doc_id]] = tag
                    else:
                        new_pos[doc_id] = len(new_rows)
                        new_ids.append(doc_id)
                        new_rows.append(emb)
                        new_tags.append(tag)
                base = idx.n
                if new_rows:
                    idx.add(np.stack(new_rows), np.asarray(new_tags, dtype=np.uint8))
            except Exception:
                self._langs._ids = langs_before
                raise
            # ---- commit (nothing below can fail half way on the GPU side)
            for row, emb, tag in updates:
                idx.set_row(row, emb, tag)
            for j, doc_id in enumerate(new_ids):
                self._row_of[doc_id] = base + j
                self._id_of.append(doc_id)
            for doc_id, content, meta, emb, _level, _lang in staged:
                if self._inner is not None:
                    self._inner.upsert(doc_id, content, emb.tolist(), meta)
                else:
                    self._store_doc(doc_id, content, meta)
            return len(documents)

    def import_quantized_batch(self, documents: List[Dict[str, Any]]) -> int:
        """Bulk-load documents whose quantised rows already exist in the reference's wire format
        (SURVEY.md 8f): each dict has ``doc_id``, ``content``, ``meta`` and ``binary`` = the raw
        bytes of the ``...:doc_binary:{id}`` key (np.packbits of the embedding, D/8 bytes,
        redis_store.py:329-338) - or of the pgvector ``*_binary.embedding`` BYTEA column
        (pgvector_store.py:322-355, same bytes) - optionally ``int8`` = the raw bytes of
        ``...:doc_int8:{id}`` / the ``*_int8.embedding`` BYTEA column (D bytes, :339-349) and
        ``embedding`` (float32, needed when the index keeps float rows).
        The payloads are stored as they are; new doc_ids only; all-or-nothing like upsert_batch."""
        if not documents:
            return 0
        with self._lock:
            first = documents[0]
            dim = self._embedding_dim or (len(first["embedding"]) if first.get("embedding") is not None
                                          else len(first["binary"]) * 8)
            self._ensure_index(dim)
            idx = self._index
            nbytes = (idx.dim + 7) // 8
            codes, i8, f32, tags, staged = [], [], [], [], []
            seen = set()
            langs_before = dict(self._langs._ids)
            try:
                for d in documents:
                    doc_id = d["doc_id"]
                    if doc_id in self._row_of or doc_id in seen:
                        raise ValueError(f"import_quantized_batch: {doc_id} is already indexed (use upsert)")
                    seen.add(doc_id)
                    content, meta, level, lang = self._prep_doc(d["content"], d.get("meta"), "child")
                    c = np.frombuffer(bytes(d["binary"]), dtype=np.uint8)
                    if c.size != nbytes:
                        raise ValueError(f"{doc_id}: binary payload has {c.size} bytes, expected {nbytes}")
                    codes.append(c)
                    if idx.store_int8:
                        if d.get("int8") is None:
                            raise ValueError(f"{doc_id}: the index keeps int8 rows but no int8 payload was given")
                        r = np.frombuffer(bytes(d["int8"]), dtype=np.int8)
                        if r.size != idx.dim:
                            raise ValueError(f"{doc_id}: int8 payload has {r.size} bytes, expected {idx.dim}")
                        i8.append(r)
                    if idx.store_f32:
                        if d.get("embedding") is None:
                            raise ValueError(f"{doc_id}: the index keeps float32 rows but no embedding was given")
                        e = np.asarray(d["embedding"], dtype=np.float32)
                        if e.shape != (idx.dim,):
                            raise ValueError(f"{doc_id}: embedding shape {e.shape} != ({idx.dim},)")
                        f32.append(e)
                    tags.append(make_tag(level, self._langs.id_for(lang, create=True)))
                    staged.append((doc_id, content, meta))
                base = idx.n
                idx.add_quantized(np.stack(codes), np.stack(i8) if i8 else None, np.stack(f32) if f32 else None,
                                  np.asarray(tags, dtype=np.uint8))
            except Exception:
                self._langs._ids = langs_before
                raise
            for j, (doc_id, content, meta) in enumerate(staged):
                self._row_of[doc_id] = base + j
                self._id_of.append(doc_id)
                if self._inner is not None:
                    self._inner.upsert_doc_only(doc_id, content, meta)
                else:
                    self._store_doc(doc_id, content, meta)
            return len(documents)

    def import_pgvector_rows(self, doc_rows, binary_rows, int8_rows=None) -> int:
        """Bulk-load from the reference's pgvector tables (pgvector_store.py:322-355, 421-458):
        ``doc_rows`` = rows of the leaf table as (doc_id, content, meta dict, embedding list or
        None); ``binary_rows`` / ``int8_rows`` = rows of ``<table>_binary`` / ``<table>_int8`` as
        (doc_id, embedding BYTEA) - ``bytes`` or ``memoryview`` as psycopg2 returns them.  Documents
        without a binary row are skipped (they were stored through the reference's batch path,
        which writes no quantised rows - SURVEY.md 0.5).  SQL I/O itself stays in the reference."""
        b = {doc_id: bytes(payload) for doc_id, payload in binary_rows}
        i8 = {doc_id: bytes(payload) for doc_id, payload in (int8_rows or [])}
        docs = []
        for doc_id, content, meta, embedding in doc_rows:
            if doc_id not in b:
                continue
            docs.append({"doc_id": doc_id, "content": content, "meta": meta, "binary": b[doc_id],
                         "int8": i8.get(doc_id), "embedding": embedding})
        return self.import_quantized_batch(docs)

    def upsert_doc_only_batch(self, documents: List[Dict[str, Any]]) -> int:
        for d in documents:
            self.upsert_doc_only(d["doc_id"], d["content"], d.get("meta"))
        return len(documents)

    def delete_doc(self, doc_id: str) -> bool:
        with self._lock:
            found = doc_id in self._docs or doc_id in self._row_of
            if self._inner is not None:
                found = bool(self._inner.delete_doc(doc_id)) or found
            self._docs.pop(doc_id, None)
            row = self._row_of.pop(doc_id, None)
            if row is not None:
                moved_from = self._index.delete_row(row)  # swap-with-last keeps the arrays dense
                if moved_from is not None:
                    moved = self._id_of[moved_from]
                    self._id_of[row] = moved
                    self._row_of[moved] = row
                self._id_of.pop()
            return found

    # ---- reads -------------------------------------------------------------------------
    def get_doc(self, doc_id: str) -> Optional[StoredDoc]:
        if self._inner is not None:
            return self._inner.get_doc(doc_id)
        return self._docs.get(doc_id)

    def has_embedding(self, doc_id: str) -> bool:
        return doc_id in self._row_of

    def list_doc_ids(self, pattern: str = "*", limit: int = 10_000) -> List[str]:
        if self._inner is not None:
            return self._inner.list_doc_ids(pattern, limit)
        return list(self._docs.keys())[:limit]

    def list_doc_ids_with_embeddings(self, limit: int = 10_000) -> List[str]:
        return self._id_of[:limit]

    def get_index_info(self) -> Dict[str, Any]:
        idx = self._index
        return {
            "backend": "b200",
            "num_docs": len(self._id_of),
            "embedding_dim": self._embedding_dim,
            "quantization_enabled": bool(getattr(self._quant_config, "enabled", False)),
            "has_int8": bool(idx is not None and idx.int8 is not None),
            "device": self._device,
        }

    def drop_index(self, delete_documents: bool = False) -> bool:
        with self._lock:
            if self._index is not None:
                self._index.clear()
            self._row_of.clear()
            self._id_of.clear()
            if delete_documents:
                self._docs.clear()
                if self._inner is not None:
                    self._inner.drop_index(delete_documents=True)
            return True

    def count_documents(self) -> int:
        if self._inner is not None:
            return self._inner.count_documents()
        return len(self._docs)

    # ---- retrieval ---------------------------------------------------------------------
    def _predicate(self, language_filter: Optional[str], doc_level_filter: Optional[str]) -> Tuple[int, int]:
        return tag_predicate(normalize_doc_level(doc_level_filter), self._langs.id_for(language_filter))

    @staticmethod
    def _check_k(top_k: int) -> int:
        """The kernels select at most RR_MAX_K entries per query: say so instead of clamping."""
        if int(top_k) < 1:
            raise ValueError(f"top_k must be >= 1, got {top_k}")
        if int(top_k) > _lib.RR_MAX_K:
            raise ValueError(f"top_k={top_k} exceeds the limit of {_lib.RR_MAX_K} results per query")
        return int(top_k)

    def _hydrate(self, idx_row, score_row, count: int) -> List[Tuple[StoredDoc, float]]:
        out: List[Tuple[StoredDoc, float]] = []
        for r, s in zip(idx_row[:count], score_row[:count]):
            if r < 0:
                continue
            doc = self.get_doc(self._id_of[r])
            if doc is not None:
                out.append((doc, float(s)))
        return out

    def retrieve_batch(self, queries, top_k: int, min_similarity: float = 0.0,
                       language_filter: Optional[str] = None, doc_level_filter: Optional[str] = None
                       ) -> List[List[Tuple[StoredDoc, float]]]:
        """Exact float32 cosine retrieval for a batch of queries [Q, D]."""
        top_k = self._check_k(top_k)
        # the lock is held from the launch to the row -> doc_id translation: delete_doc's
        # swap-with-last and a growing upsert must not move rows under a search in flight
        with self._lock:
            if self._index is None or self._index.n == 0:
                return [[] for _ in range(len(queries))]
            mask, value = self._predicate(language_filter, doc_level_filter)
            idx, score, count = self._index.search_exact(queries, top_k, min_similarity, mask, value)
            idx_h, score_h, count_h = idx.cpu().tolist(), score.cpu().tolist(), count.cpu().tolist()
            return [self._hydrate(idx_h[i], score_h[i], count_h[i]) for i in range(len(idx_h))]

    def retrieve_batch_quantized(self, queries, top_k: int, min_similarity: float = 0.0,
                                 rescore_multiplier: Optional[float] = None,
                                 use_rescoring: Optional[bool] = None,
                                 language_filter: Optional[str] = None,
                                 doc_level_filter: Optional[str] = None
                                 ) -> List[List[Tuple[StoredDoc, float]]]:
        """Two-stage quantised retrieval for a batch of queries [Q, D]."""
        if not getattr(self._quant_config, "enabled", False):
            return self.retrieve_batch(queries, top_k, min_similarity, language_filter, doc_level_filter)
        top_k = self._check_k(top_k)
        mult = rescore_multiplier if rescore_multiplier is not None else self._quant_config.rescore_multiplier
        use = use_rescoring if use_rescoring is not None else self._quant_config.use_rescoring
        if use and int(top_k * mult) > _lib.RR_MAX_K:
            logger.warning(f"candidate_k = int({top_k} * {mult}) exceeds the kernel limit {_lib.RR_MAX_K}; "
                           f"{_lib.RR_MAX_K} candidates are rescored instead")
        with self._lock:
            if self._index is None or self._index.n == 0:
                return [[] for _ in range(len(queries))]
            mask, value = self._predicate(language_filter, doc_level_filter)
            idx, score, count = self._index.search_quantized(
                queries, top_k, rescore_multiplier=mult, use_rescoring=use, min_similarity=min_similarity,
                tag_mask=mask, tag_value=value)
            idx_h, score_h, count_h = idx.cpu().tolist(), score.cpu().tolist(), count.cpu().tolist()
            return [self._hydrate(idx_h[i], score_h[i], count_h[i]) for i in range(len(idx_h))]

    def retrieve_by_embedding(
        self,
        query_embedding: List[float],
        top_k: int,
        min_similarity: float = 0.0,
        ef_runtime: Optional[int] = None,
        language_filter: Optional[str] = None,
        doc_level_filter: Optional[str] = None,
    ) -> List[Tuple[StoredDoc, float]]:
        q = np.asarray(query_embedding, dtype=np.float32)[None, :]
        if self._index is not None and q.shape[1] != self._index.dim:
            raise ValueError(f"query dim {q.shape[1]} != index dim {self._index.dim}")
        return self.retrieve_batch(q, top_k, min_similarity, language_filter, doc_level_filter)[0]

    def retrieve_by_embedding_quantized(
        self,
        query_embedding: List[float],
        top_k: int,
        min_similarity: float = 0.0,
        rescore_multiplier: Optional[float] = None,
        use_rescoring: Optional[bool] = None,
        language_filter: Optional[str] = None,
        doc_level_filter: Optional[str] = None,
    ) -> List[Tuple[StoredDoc, float]]:
        q = np.asarray(query_embedding, dtype=np.float32)[None, :]
        if self._index is not None and q.shape[1] != self._index.dim:
            raise ValueError(f"query dim {q.shape[1]} != index dim {self._index.dim}")
        return self.retrieve_batch_quantized(q, top_k, min_similarity, rescore_multiplier, use_rescoring,
                                             language_filter, doc_level_filter)[0]
