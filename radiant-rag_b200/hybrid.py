"""One product call for the batched hybrid retrieval step: dense -> BM25 -> RRF on device.

This is the path BASELINE.json's metric names ("hybrid retrieval queries/sec @ top-10").  The
reference runs it one query at a time through three agents:

    RadiantRAG.search(query, mode="hybrid", top_k)          radiant/app.py:1178-1249
    RAGOrchestrator._run_retrieval / _fuse_results          radiant/orchestrator.py:918-1151, 1153-1196

``HybridSearch.search_batch`` is the device-level form (tensors in, tensors out, replayable from
a CUDA graph); ``BatchedRetrieval`` is the host-level form with the reference's objects (query
strings in, ``[(StoredDoc, score)]`` out) that replaces the orchestrator's per-sub-query loop
and two-thread pool with ONE batched call per phase.

Row ids: the dense index and the BM25 index must number documents identically (row = position
in upsert order); both return GLOBAL rows (``row_base`` + local row) so the fused
list is the same on one GPU and on a row-sharded corpus.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .agents import rrf_fuse_runs_device
from .index import DenseIndex, to_device
from .sharded import GpuShardOps, ShardedBM25Search, ShardedDenseSearch


@dataclass
class HybridResult:
    """Static-shape outputs of one hybrid step (device tensors).

    idx int64 [Q, top_k] (-1 padded) / score f64 [Q, top_k] / count int32 [Q]: the RRF-fused list,
    (score desc, first-insertion asc) exactly as RRFAgent._execute orders it.  The two input runs
    are returned as well: the reference's ``search`` falls back to the non-empty run when the
    other one is empty (app.py:1234-1239) and the host wrapper needs their own scores for that."""
    idx: torch.Tensor
    score: torch.Tensor
    count: torch.Tensor
    dense_idx: torch.Tensor
    dense_score: torch.Tensor
    dense_count: torch.Tensor
    bm25_idx: torch.Tensor
    bm25_score: torch.Tensor
    bm25_count: torch.Tensor

    def tensors(self) -> Tuple[torch.Tensor, ...]:
        return (self.idx, self.score, self.count, self.dense_idx, self.dense_score, self.dense_count,
                self.bm25_idx, self.bm25_score, self.bm25_count)


class HybridSearch:
    """dense top-``dense_top_k`` (two-stage quantised, or the exact scan) + BM25
    top-``bm25_top_k`` fused by RRF into the top-``top_k``, for a batch of queries, entirely on
    the device: three C-ABI calls chains and no host decision in between.

    dense: a ``DenseIndex`` (this rank's shard); bm25: a ``Bm25DeviceIndex`` over the same rows.
    With ``torch.distributed`` initialised (one process per GPU) both halves exchange their
    per-shard candidates (sharded.py) and every rank returns the same fused lists."""

    def __init__(self, dense: DenseIndex, bm25: Any, group: Optional[Any] = None,
                 rescore_multiplier: float = 4.0, prefer_int8: bool = True,
                 dense_mode: str = "quantized", comm: Optional[Any] = None, overlap: bool = True,
                 comm_sparse: Optional[Any] = None) -> None:
        """comm: an ``nccl.NcclComm`` over the same ranks - the candidate exchanges are then issued on
        the current stream, which lets ``GraphedHybridSearch`` capture the WHOLE sharded step (kernels
        and collectives) into one CUDA graph.
        overlap: unchecked steps run the BM25 half on a second stream next to the dense half (the two
        are independent until RRF); in a captured step they become parallel branches of the graph.
        When sharded this needs ``comm_sparse``, a SECOND communicator for the BM25 exchange: one
        communicator must see its collectives in one order on every rank, and two branches of a
        graph have none."""
        if dense_mode not in ("quantized", "exact"):
            raise ValueError("dense_mode must be 'quantized' or 'exact'")
        self.dense_index = dense
        self.bm25_index = bm25
        self.device = dense.device
        self.group = group
        self.rescore_multiplier = float(rescore_multiplier)
        self.prefer_int8 = bool(prefer_int8)
        self.dense_mode = dense_mode
        self.comm = comm
        self.overlap = bool(overlap)
        self._side_stream: Optional[torch.cuda.Stream] = None
        self.ops = GpuShardOps(dense)
        self.dense = ShardedDenseSearch(self.ops, group, comm)
        self.comm_sparse = comm_sparse
        self.sparse = ShardedBM25Search(bm25, self.ops, group, comm_sparse if comm_sparse is not None else comm)

    def search_batch(self, queries, q_terms, top_k: int = 10, dense_top_k: int = 100,
                     bm25_top_k: int = 100, rrf_k: float = 60, min_similarity: float = 0.0,
                     tag_mask: int = 0, tag_value: int = 0, check: bool = True) -> HybridResult:
        """queries f32 [Q, D] (host or device); q_terms int32 [Q, L] BM25 term ids in query-token
        order (-1 = unknown / padding), identical on every rank.

        check=False leaves out every host synchronisation (the step can then be captured into a
        CUDA graph); the tensor-core overflow counter and the BM25 inexact counter are accumulated
        on the device instead and must be read by the caller (``unchecked_events``)."""
        for name, kk in (("top_k", top_k), ("dense_top_k", dense_top_k), ("bm25_top_k", bm25_top_k)):
            if not 1 <= int(kk) <= _lib.RR_MAX_K:
                raise ValueError(f"{name}={kk} outside [1, {_lib.RR_MAX_K}]")
        side = None
        if self.overlap and not check and (self.dense.world() == 1 or self.comm_sparse is not None):
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            side = self._side_stream
            main = torch.cuda.current_stream(self.device)
            queries = to_device(queries, self.device, torch.float32)
            q_terms = to_device(q_terms, self.device, torch.int32)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                b_idx, b_score, b_count = self.sparse.search_batch(q_terms, bm25_top_k, check=False)
            for t in (b_idx, b_score, b_count):
                t.record_stream(main)
        if self.dense_mode == "quantized":
            d_idx, d_score, d_count = self.dense.search_quantized(
                queries, dense_top_k, rescore_multiplier=self.rescore_multiplier,
                min_similarity=min_similarity, tag_mask=tag_mask, tag_value=tag_value,
                prefer_int8=self.prefer_int8, check_overflow=check)
        else:
            if self.dense.world() > 1:
                raise ValueError("dense_mode='exact' is a single-GPU path")
            d_idx, d_score, d_count = self.dense_index.search_exact(queries, dense_top_k, min_similarity,
                                                                    tag_mask, tag_value, check_overflow=check)
        if side is not None:
            torch.cuda.current_stream(self.device).wait_stream(side)
        else:
            b_idx, b_score, b_count = self.sparse.search_batch(q_terms, bm25_top_k, check=check)
        f_idx, f_score, f_count = rrf_fuse_runs_device([d_idx, b_idx], top_k, rrf_k)
        return HybridResult(f_idx, f_score, f_count, d_idx, d_score, d_count, b_idx, b_score, b_count)

    def unchecked_events(self) -> int:
        """Tensor-core list overflows + BM25 inexact flags accumulated by ``check=False`` calls on
        this rank since the last ``reset_unchecked_events`` (0 = every result was exact)."""
        n = self.dense_index.tc_overflow_total()
        fn = getattr(self.bm25_index, "inexact_total", None)
        return n + (int(fn()) if fn else 0)

    def reset_unchecked_events(self) -> None:
        self.dense_index.tc_overflow_reset()
        fn = getattr(self.bm25_index, "inexact_reset", None)
        if fn:
            fn()


class GraphedHybridSearch:
    """CUDA-graph replay of one fixed-shape hybrid step: on ONE GPU, or on a row-sharded corpus when
    the ``HybridSearch`` was built with an ``NcclComm`` (the exchanges are then part of the graph;
    every rank must replay in lock-step, as with any collective).

        g = GraphedHybridSearch(hybrid, n_queries, dim, q_len, top_k=10, ...)
        res = g(queries_host_pinned, q_terms_host_pinned)      # HybridResult of static tensors

    Host inputs go through two staging buffers on a copy stream (the H2D of step i+1 overlaps
    step i).  Results are the graph's static outputs: copy them out before the next replay."""

    def __init__(self, hybrid: HybridSearch, n_queries: int, dim: int, q_len: int, warmup: int = 3,
                 **search_kwargs: Any) -> None:
        self.hybrid = hybrid
        self.device = hybrid.device
        dev = self.device
        if hybrid.dense.world() > 1 and hybrid.comm is None:
            raise ValueError("capturing a sharded step needs HybridSearch(..., comm=NcclComm(...))")
        self.static_q = torch.zeros((n_queries, dim), dtype=torch.float32, device=dev)
        self.static_t = torch.full((n_queries, q_len), -1, dtype=torch.int32, device=dev)
        kwargs = dict(search_kwargs, check=False)

        def step() -> HybridResult:
            return hybrid.search_batch(self.static_q, self.static_t, **kwargs)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        launches0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.static_out = step()
        self.kernels_per_replay = _lib.launch_count - launches0
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._stage = [(torch.empty_like(self.static_q), torch.empty_like(self.static_t)) for _ in range(2)]
        self._filled = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._turn = 0

    def load(self, queries: torch.Tensor, q_terms: torch.Tensor) -> None:
        if queries.is_cuda and q_terms.is_cuda:
            self.static_q.copy_(queries, non_blocking=True)
            self.static_t.copy_(q_terms, non_blocking=True)
            return
        b = self._turn
        self._turn ^= 1
        main = torch.cuda.current_stream(self.device)
        cs = self._copy_stream
        cs.wait_event(self._consumed[b])
        with torch.cuda.stream(cs):
            self._stage[b][0].copy_(queries, non_blocking=True)
            self._stage[b][1].copy_(q_terms, non_blocking=True)
            self._filled[b].record(cs)
        main.wait_event(self._filled[b])
        self.static_q.copy_(self._stage[b][0], non_blocking=True)
        self.static_t.copy_(self._stage[b][1], non_blocking=True)
        self._consumed[b].record(main)

    def replay(self) -> HybridResult:
        self.graph.replay()
        _lib.launch_count += self.kernels_per_replay
        return self.static_out

    def __call__(self, queries: torch.Tensor, q_terms: torch.Tensor) -> HybridResult:
        self.load(queries, q_terms)
        return self.replay()


class BatchedRetrieval:
    """Host-level batched replacement of the reference's retrieval phase.

    The reference embeds the sub-queries, then loops ``store.retrieve_by_embedding`` and
    ``index.search`` per sub-query in two threads, de-duplicates by ``doc_id`` (first occurrence
    wins, orchestrator.py:953-965, 980-987) and fuses the two merged runs with RRFAgent
    (:1153-1196); ``RadiantRAG.search`` does the same for one query (app.py:1178-1249).  Here the
    whole batch is one embedding call, one dense search, one BM25 search and one RRF launch.

    store: ``B200VectorStore``; bm25_index: this package's ``PersistentBM25Index`` built over the
    same store; local_models: anything with ``embed(list[str])`` or ``embed_single(str)``;
    config: a ``RetrievalConfig`` (dense_top_k, bm25_top_k, fused_top_k, rrf_k, min_similarity,
    search_scope)."""

    def __init__(self, store: Any, bm25_index: Any, local_models: Any, config: Any,
                 use_quantized: Optional[bool] = None) -> None:
        self._store = store
        self._index = bm25_index
        self._local_models = local_models
        self._config = config
        if use_quantized is None:
            use_quantized = bool(getattr(getattr(store, "_quant_config", None), "enabled", False))
        self._use_quantized = use_quantized

    def _doc_level_filter(self, search_scope: Optional[str]) -> Optional[str]:
        scope = search_scope or getattr(self._config, "search_scope", "leaves")
        if scope == "parents":
            return "parent"
        if scope == "all":
            return None
        return "child"

    def _embed(self, queries: Sequence[str]) -> np.ndarray:
        embed = getattr(self._local_models, "embed", None)
        vecs = embed(list(queries)) if embed else [self._local_models.embed_single(q) for q in queries]
        return np.asarray(vecs, dtype=np.float32)

    def dense_batch(self, queries: Sequence[str], top_k: Optional[int] = None,
                    search_scope: Optional[str] = None) -> List[List[Tuple[Any, float]]]:
        k = top_k or self._config.dense_top_k
        level = self._doc_level_filter(search_scope)
        vecs = self._embed(queries)
        if self._use_quantized:
            return self._store.retrieve_batch_quantized(vecs, k, self._config.min_similarity,
                                                        doc_level_filter=level)
        return self._store.retrieve_batch(vecs, k, self._config.min_similarity, doc_level_filter=level)

    def bm25_batch(self, queries: Sequence[str], top_k: Optional[int] = None) -> List[List[Tuple[Any, float]]]:
        return self._index.search_batch(list(queries), top_k or self._config.bm25_top_k)

    def search_batch(self, queries: Sequence[str], mode: str = "hybrid", top_k: int = 10,
                     search_scope: Optional[str] = None) -> List[List[Tuple[Any, float]]]:
        """Batched ``RadiantRAG.search`` (app.py:1178-1249): every query is searched with
        ``top_k`` in both retrievers, fused with ``top_k``; if one of a query's runs is empty
        the other run is returned as it is."""
        from .agents import RRFAgent

        if mode not in ("hybrid", "dense", "bm25"):
            raise ValueError("mode must be 'hybrid', 'dense' or 'bm25'")
        nq = len(queries)
        dense = self.dense_batch(queries, top_k, search_scope) if mode in ("hybrid", "dense") else [[] for _ in range(nq)]
        if mode == "dense":
            return dense
        sparse = self.bm25_batch(queries, top_k) if mode in ("hybrid", "bm25") else [[] for _ in range(nq)]
        if mode == "bm25":
            return sparse
        both = [i for i in range(nq) if dense[i] and sparse[i]]
        fused = RRFAgent(self._config, device=getattr(self._store, "_device", 0)).fuse_batch(
            [[dense[i], sparse[i]] for i in both], top_k=top_k) if both else []
        out: List[List[Tuple[Any, float]]] = [dense[i] if dense[i] else sparse[i] for i in range(nq)]
        for j, i in enumerate(both):
            out[i] = fused[j][:top_k]
        return out

    def run_retrieval(self, sub_queries: Sequence[str], retrieval_mode: str = "hybrid",
                      search_scope: Optional[str] = None
                      ) -> Tuple[List[Tuple[Any, float]], List[Tuple[Any, float]], List[Tuple[Any, float]]]:
        """Batched ``_run_retrieval`` + ``_fuse_results`` for ONE user query expanded into
        ``sub_queries``: -> (dense_retrieved, bm25_retrieved, fused).  Each retriever's per-sub-query
        lists are concatenated in sub-query order keeping the first occurrence of a doc_id
        (orchestrator.py:953-965, 980-987); the merged runs are fused with the config's
        ``fused_top_k`` / ``rrf_k`` when both are non-empty, otherwise the non-empty one is the
        result (:1195-1196)."""
        from .agents import RRFAgent

        def merged(per_query: List[List[Tuple[Any, float]]]) -> List[Tuple[Any, float]]:
            seen, out = set(), []
            for results in per_query:
                for doc, score in results:
                    doc_id = getattr(doc, "doc_id", id(doc))
                    if doc_id not in seen:
                        seen.add(doc_id)
                        out.append((doc, score))
            return out

        dense = merged(self.dense_batch(sub_queries, None, search_scope)) if retrieval_mode in ("hybrid", "dense") else []
        sparse = merged(self.bm25_batch(sub_queries, None)) if retrieval_mode in ("hybrid", "bm25") else []
        lists = [r for r in (dense, sparse) if r]
        if len(lists) > 1:
            fused = RRFAgent(self._config, device=getattr(self._store, "_device", 0)).fuse_batch([lists])[0]
        elif lists:
            fused = lists[0]
        else:
            fused = []
        return dense, sparse, fused


# ---- the step after the path: batched cross-encoder rerank and auto-merge ---------

def rerank_batch(local_models: Any, queries: Sequence[str], docs_per_query: Sequence[List[Tuple[Any, float]]],
                 config: Any, top_k: Optional[int] = None) -> List[List[Tuple[Any, float]]]:
    """``CrossEncoderRerankingAgent._execute`` (radiant/agents/rerank.py:64-117) for a batch of
    queries with ONE cross-encoder call: the reference scores each query's candidates with its own
    ``cross_encoder.predict`` (radiant/llm/local_models.py:251-280); the (query, document) pairs
    of all queries are independent, so they are flattened into a single ``predict``.  Per query
    the result is the reference's: candidates = first max(k * candidate_multiplier,
    min_candidates) documents, texts cut to ``max_doc_chars``, stable sort by score descending,
    first k.  Falls back to per-query ``local_models.rerank`` when no ``cross_encoder`` is exposed."""
    k = top_k or config.top_k
    num_candidates = max(k * config.candidate_multiplier, config.min_candidates)
    cands = [list(docs[:num_candidates]) for docs in docs_per_query]
    texts = [[d.content[: config.max_doc_chars] for d, _ in c] for c in cands]
    encoder = getattr(local_models, "cross_encoder", None)
    out: List[List[Tuple[Any, float]]] = []
    if encoder is None:
        for q, c, t in zip(queries, cands, texts):
            out.append([(c[i][0], s) for i, s in local_models.rerank(q, t, top_k=k)] if c else [])
        return out
    pairs = [(q, t) for q, ts in zip(queries, texts) for t in ts]
    scores = list(encoder.predict(pairs, show_progress_bar=False)) if pairs else []
    pos = 0
    for c in cands:
        s = [(i, float(scores[pos + i])) for i in range(len(c))]
        pos += len(c)
        s.sort(key=lambda x: x[1], reverse=True)  # stable, as the reference
        out.append([(c[i][0], sc) for i, sc in s[:k]])
    return out


def automerge_batch(store: Any, docs_per_query: Sequence[List[Tuple[Any, float]]], min_children: int,
                    max_parent_chars: int, top_k: Optional[int] = None) -> List[List[Tuple[Any, float]]]:
    """The parent look-ups of ``HierarchicalAutoMergingAgent`` (radiant/agents/automerge.py:87-135) for
    a batch of queries: every distinct parent id of the batch is fetched ONCE (the reference calls
    ``store.get_doc`` per parent per query), the merge rule itself is unchanged - children of a parent
    with >= min_children hits are replaced by the parent (best child score) if it exists and is
    short enough; per doc_id the best score wins; sorted by score descending."""
    wanted = set()
    grouped = []
    for docs in docs_per_query:
        by_parent, passthrough = {}, []
        for doc, score in docs:
            meta = getattr(doc, "meta", {}) or {}
            parent_id = str(meta.get("parent_id", "")).strip()
            if str(meta.get("doc_level", "child")) == "child" and parent_id:
                by_parent.setdefault(parent_id, []).append((doc, score))
            else:
                passthrough.append((doc, score))
        grouped.append((by_parent, passthrough))
        wanted.update(p for p, ch in by_parent.items() if len(ch) >= min_children)
    parents = {p: store.get_doc(p) for p in wanted}
    out = []
    for by_parent, passthrough in grouped:
        merged: List[Tuple[Any, float]] = []
        for parent_id, children in by_parent.items():
            parent = parents.get(parent_id) if len(children) >= min_children else None
            if parent is not None and len(parent.content) <= max_parent_chars:
                merged.append((parent, max(s for _, s in children)))
            else:
                merged.extend(children)
        best = {}
        for doc, score in passthrough + merged:
            if doc.doc_id not in best or score > best[doc.doc_id][1]:
                best[doc.doc_id] = (doc, score)
        result = list(best.values())
        result.sort(key=lambda x: x[1], reverse=True)
        out.append(result[: (top_k or len(passthrough) + sum(len(c) for c in by_parent.values()))])
    return out
