"""Retrieval agents with the reference's names, kwargs and error behaviour.

    DenseRetrievalAgent  reference radiant/agents/dense.py:26-141
    BM25RetrievalAgent   reference radiant/agents/bm25.py:25-101
    RRFAgent             reference radiant/agents/fusion.py:24-114
    AgentResult / run()  reference radiant/agents/base_agent.py:145-184, 468-584

Only what the hot path needs of ``BaseAgent`` is mirrored (run / execute / _execute /
_on_error, enabled flag, PARTIAL status on recovered errors); metrics, correlation ids
and structured logging stay in the reference.  The attributes the reference's
orchestrator reaches into are kept: ``_store``, ``_config``, ``_index``,
``_local_models``, ``_get_doc_level_filter`` (radiant/orchestrator.py:939-951, 977-978).
"""

from __future__ import annotations

import logging
import time
import uuid
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .index import _stream, to_device

logger = logging.getLogger(__name__)

RRF_MAX_LEN = 4096
RRF_MAX_RUNS = 16


try:  # inside a Radiant RAG installation the agents ARE reference agents (same lifecycle, metrics,
    # structured logging): subclass the reference's BaseAgent and return its AgentResult
    from radiant.agents.base_agent import (  # type: ignore
        AgentCategory, AgentMetrics, AgentResult, AgentStatus, BaseAgent as _ReferenceBaseAgent,
    )
    HAVE_REFERENCE_AGENTS = True
except Exception:  # stand-alone: mirror what the orchestrator touches (base_agent.py:43-184, 468-584)
    HAVE_REFERENCE_AGENTS = False
    _ReferenceBaseAgent = None

    class AgentCategory(Enum):
        RETRIEVAL = "retrieval"
        POST_RETRIEVAL = "post_retrieval"
        UTILITY = "utility"

    class AgentStatus(Enum):
        SUCCESS = "success"
        PARTIAL = "partial"
        FAILED = "failed"
        SKIPPED = "skipped"
        TIMEOUT = "timeout"

    @dataclass
    class AgentMetrics:
        agent_name: str
        agent_category: str
        run_id: str
        correlation_id: str
        start_time: float = 0.0
        end_time: float = 0.0
        duration_ms: float = 0.0
        status: AgentStatus = AgentStatus.SUCCESS
        error_message: Optional[str] = None
        items_processed: int = 0
        items_returned: int = 0
        llm_calls: int = 0
        retrieval_calls: int = 0
        confidence: float = 0.0
        custom: Dict[str, Any] = field(default_factory=dict)

        def to_dict(self) -> Dict[str, Any]:
            d = dict(self.__dict__)
            d["status"] = self.status.value
            return d

    @dataclass
    class AgentResult:
        data: Any
        success: bool = True
        status: AgentStatus = AgentStatus.SUCCESS
        error: Optional[str] = None
        warnings: List[str] = field(default_factory=list)
        metrics: Optional[AgentMetrics] = None

        def add_warning(self, message: str) -> None:
            self.warnings.append(message)
            if self.status == AgentStatus.SUCCESS:
                self.status = AgentStatus.PARTIAL

        def to_dict(self) -> Dict[str, Any]:
            data_dict = self.data if isinstance(self.data, dict) else None
            if hasattr(self.data, "to_dict"):
                data_dict = self.data.to_dict()
            return {"data": data_dict, "success": self.success, "status": self.status.value,
                    "error": self.error, "warnings": self.warnings,
                    "metrics": self.metrics.to_dict() if self.metrics else None}


if HAVE_REFERENCE_AGENTS:

    class _Agent(_ReferenceBaseAgent):  # type: ignore[misc]
        """The reference's lifecycle (run / execute / hooks / metrics) unchanged; subclasses give
        ``name``, ``category`` and ``_execute`` / ``_on_error``."""

        AGENT_NAME = "Agent"
        AGENT_CATEGORY = "RETRIEVAL"

        def __init__(self, enabled: bool = True, store: Any = None, local_models: Any = None) -> None:
            super().__init__(llm=None, store=store, local_models=local_models, enabled=enabled)

        @property
        def name(self) -> str:
            return self.AGENT_NAME

        @property
        def category(self):
            return getattr(AgentCategory, self.AGENT_CATEGORY)

        @property
        def description(self) -> str:
            return self.__doc__ or self.AGENT_NAME

else:

    class _Agent:
        AGENT_NAME = "Agent"
        AGENT_CATEGORY = "RETRIEVAL"

        def __init__(self, enabled: bool = True, store: Any = None, local_models: Any = None) -> None:
            self._enabled = enabled
            self._store = store
            self._local_models = local_models
            self._total_executions = 0
            self._total_successes = 0
            self._total_failures = 0
            self._total_duration_ms = 0.0

        @property
        def name(self) -> str:
            return self.AGENT_NAME

        @property
        def category(self) -> AgentCategory:
            return getattr(AgentCategory, self.AGENT_CATEGORY)

        def _execute(self, **kwargs: Any) -> Any:  # pragma: no cover - abstract
            raise NotImplementedError

        def _on_error(self, error: Exception, metrics: Any = None, **kwargs: Any) -> Optional[Any]:
            return None

        def run(self, correlation_id: Optional[str] = None, **kwargs: Any) -> AgentResult:
            """Same contract as the reference's BaseAgent.run (base_agent.py:468-584): SKIPPED when
            disabled, metrics on every result, PARTIAL with the ``_on_error`` fallback, FAILED otherwise."""
            if not self._enabled:
                return AgentResult(data=None, success=True, status=AgentStatus.SKIPPED)
            metrics = AgentMetrics(agent_name=self.name, agent_category=self.category.value,
                                   run_id=str(uuid.uuid4()), correlation_id=correlation_id or str(uuid.uuid4())[:8],
                                   start_time=time.time())
            try:
                data = self._execute(**kwargs)
                metrics.end_time = time.time()
                metrics.duration_ms = (metrics.end_time - metrics.start_time) * 1000
                if isinstance(data, (list, tuple)):
                    metrics.items_returned = len(data)
                self._total_executions += 1
                self._total_successes += 1
                self._total_duration_ms += metrics.duration_ms
                return AgentResult(data=data, success=True, status=AgentStatus.SUCCESS, metrics=metrics)
            except Exception as e:  # noqa: BLE001 - same catch-all as the reference lifecycle
                metrics.end_time = time.time()
                metrics.duration_ms = (metrics.end_time - metrics.start_time) * 1000
                metrics.status = AgentStatus.FAILED
                metrics.error_message = str(e)
                self._total_executions += 1
                self._total_failures += 1
                self._total_duration_ms += metrics.duration_ms
                logger.error(f"{self.name} execution failed: {e}")
                fallback = self._on_error(e, metrics, **kwargs)
                if fallback is not None:
                    return AgentResult(data=fallback, success=True, status=AgentStatus.PARTIAL,
                                       warnings=[f"Recovered from error: {e}"], metrics=metrics)
                return AgentResult(data=None, success=False, status=AgentStatus.FAILED, error=str(e), metrics=metrics)

        def execute(self, correlation_id: Optional[str] = None, **kwargs: Any) -> Any:
            result = self.run(correlation_id=correlation_id, **kwargs)
            if not result.success and result.data is None:
                raise RuntimeError(f"{self.name} failed: {result.error or 'Unknown error'}")
            return result.data


class DenseRetrievalAgent(_Agent):
    AGENT_NAME = "DenseRetrievalAgent"
    AGENT_CATEGORY = "RETRIEVAL"

    def __init__(self, store: Any, local: Any, config: Any, enabled: bool = True,
                 use_quantized: Optional[bool] = None) -> None:
        """use_quantized: route through ``retrieve_by_embedding_quantized``;
        default = the store's ``quantization.enabled``."""
        if store is None:
            raise ValueError("DenseRetrievalAgent requires a vector store")
        if local is None:
            raise ValueError("DenseRetrievalAgent requires local models")
        super().__init__(enabled=enabled, store=store, local_models=local)
        self._store = store
        self._local_models = local
        self._config = config
        if use_quantized is None:
            use_quantized = bool(getattr(getattr(store, "_quant_config", None), "enabled", False))
        self._use_quantized = use_quantized

    def _get_doc_level_filter(self, search_scope: Optional[str] = None) -> Optional[str]:
        scope = search_scope or self._config.search_scope
        if scope == "parents":
            return "parent"
        if scope == "all":
            return None
        return "child"  # "leaves" and anything unknown

    def _execute(self, query: str, top_k: Optional[int] = None, search_scope: Optional[str] = None,
                 **kwargs: Any) -> List[Tuple[Any, float]]:
        k = top_k or self._config.dense_top_k
        level = self._get_doc_level_filter(search_scope)
        query_vec = self._local_models.embed_single(query)
        if self._use_quantized:
            return self._store.retrieve_by_embedding_quantized(
                query_embedding=query_vec, top_k=k, min_similarity=self._config.min_similarity,
                doc_level_filter=level)
        return self._store.retrieve_by_embedding(
            query_embedding=query_vec, top_k=k, min_similarity=self._config.min_similarity,
            doc_level_filter=level)

    def execute_batch(self, queries: Sequence[str], top_k: Optional[int] = None,
                      search_scope: Optional[str] = None) -> List[List[Tuple[Any, float]]]:
        """New batched surface: one embedding call + one GPU search for all queries."""
        k = top_k or self._config.dense_top_k
        level = self._get_doc_level_filter(search_scope)
        embed = getattr(self._local_models, "embed", None)
        vecs = embed(list(queries)) if embed else [self._local_models.embed_single(q) for q in queries]
        vecs = np.asarray(vecs, dtype=np.float32)
        if self._use_quantized:
            return self._store.retrieve_batch_quantized(vecs, k, self._config.min_similarity,
                                                        doc_level_filter=level)
        return self._store.retrieve_batch(vecs, k, self._config.min_similarity, doc_level_filter=level)

    def _on_error(self, error: Exception, metrics: Any = None, **kwargs: Any) -> Optional[List[Tuple[Any, float]]]:
        logger.warning(f"Dense retrieval failed: {error}")
        return []


class BM25RetrievalAgent(_Agent):
    AGENT_NAME = "BM25RetrievalAgent"
    AGENT_CATEGORY = "RETRIEVAL"

    def __init__(self, bm25_index: Any, config: Any, enabled: bool = True) -> None:
        super().__init__(enabled=enabled)
        self._index = bm25_index
        self._config = config

    def _execute(self, query: str, top_k: Optional[int] = None, **kwargs: Any) -> List[Tuple[Any, float]]:
        k = top_k or self._config.bm25_top_k
        return self._index.search(query, top_k=k)

    def execute_batch(self, queries: Sequence[str], top_k: Optional[int] = None
                      ) -> List[List[Tuple[Any, float]]]:
        k = top_k or self._config.bm25_top_k
        return self._index.search_batch(list(queries), top_k=k)

    def _on_error(self, error: Exception, metrics: Any = None, **kwargs: Any) -> Optional[List[Tuple[Any, float]]]:
        logger.warning(f"BM25 retrieval failed: {error}")
        return []


def rrf_fuse_device(run_idx: torch.Tensor, run_off: Sequence[int], k: int, rrf_k: float = 60
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Batched RRF on device.  run_idx int64 [Q, L]: the runs of each query concatenated
    (segment bounds ``run_off``, -1 pads a short run at its tail; ids < 2^32).
    -> (idx int64 [Q,k] (-1 padded), score f64 [Q,k], count int32 [Q])."""
    import ctypes as C

    dev = run_idx.device
    q, length = run_idx.shape
    n_runs = len(run_off) - 1
    if length != run_off[-1]:
        raise ValueError("run_off[-1] must equal run_idx.shape[1]")
    idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    score = torch.empty((q, k), dtype=torch.float64, device=dev)
    count = torch.empty((q,), dtype=torch.int32, device=dev)
    off = (C.c_int32 * (n_runs + 1))(*[int(x) for x in run_off])
    _lib.call("rr_rrf_fuse", run_idx.contiguous().data_ptr(), off, n_runs, q, float(rrf_k), k,
              idx.data_ptr(), score.data_ptr(), count.data_ptr(), _stream())
    return idx, score, count


def rrf_fuse_runs_device(runs: Sequence[torch.Tensor], k: int, rrf_k: float = 60
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Batched RRF over runs that stay where their kernels wrote them: ``runs[r]`` is a
    contiguous int64 [Q, L_r] device tensor (-1 pads a short run at its tail; ids < 2^32).
    Same arithmetic and order as ``rrf_fuse_device`` without the concatenation copy.
    -> (idx int64 [Q,k] (-1 padded), score f64 [Q,k], count int32 [Q])."""
    import ctypes as C

    n_runs = len(runs)
    if n_runs == 0:
        raise ValueError("at least one run is needed")
    dev = runs[0].device
    q = runs[0].shape[0]
    for r in runs:
        if r.dtype != torch.int64 or r.ndim != 2 or r.shape[0] != q or not r.is_contiguous():
            raise ValueError("runs must be contiguous int64 [Q, L_r] tensors with the same Q")
    idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    score = torch.empty((q, k), dtype=torch.float64, device=dev)
    count = torch.empty((q,), dtype=torch.int32, device=dev)
    ptrs = (C.c_void_p * n_runs)(*[r.data_ptr() for r in runs])
    lens = (C.c_int32 * n_runs)(*[int(r.shape[1]) for r in runs])
    _lib.call("rr_rrf_fuse_runs", ptrs, lens, n_runs, q, float(rrf_k), k, idx.data_ptr(),
              score.data_ptr(), count.data_ptr(), _stream())
    return idx, score, count


class RRFAgent(_Agent):
    AGENT_NAME = "RRFAgent"
    AGENT_CATEGORY = "POST_RETRIEVAL"

    def __init__(self, config: Any, enabled: bool = True, device: int = 0) -> None:
        super().__init__(enabled=enabled)
        self._config = config
        self._device = device

    def _execute(self, runs: List[List[Tuple[Any, float]]], top_k: Optional[int] = None,
                 rrf_k: Optional[int] = None, **kwargs: Any) -> List[Tuple[Any, float]]:
        return self.fuse_batch([runs], top_k=top_k, rrf_k=rrf_k)[0]

    def fuse_batch(self, batch_runs: Sequence[List[List[Tuple[Any, float]]]],
                   top_k: Optional[int] = None, rrf_k: Optional[int] = None
                   ) -> List[List[Tuple[Any, float]]]:
        """Fuse the runs of many queries in one launch.  Every query has the same number of
        runs; run lengths may differ (shorter ones are padded)."""
        k = top_k or self._config.fused_top_k
        c = rrf_k or self._config.rrf_k
        nq = len(batch_runs)
        if nq == 0:
            return []
        n_runs = max(len(r) for r in batch_runs)
        if n_runs == 0:
            return [[] for _ in range(nq)]
        if n_runs > RRF_MAX_RUNS:
            raise ValueError(f"at most {RRF_MAX_RUNS} runs can be fused")
        seg = [max([len(r[j]) if j < len(r) else 0 for r in batch_runs] + [0]) for j in range(n_runs)]
        off = [0]
        for s in seg:
            off.append(off[-1] + s)
        if off[-1] == 0:
            return [[] for _ in range(nq)]
        if off[-1] > RRF_MAX_LEN:
            raise ValueError(f"total run length {off[-1]} > {RRF_MAX_LEN}")
        _lib.init(self._device)
        mat = np.full((nq, off[-1]), -1, dtype=np.int64)
        doc_tables: List[List[Any]] = []
        for qi, runs in enumerate(batch_runs):
            ids: Dict[str, int] = {}
            table: List[Any] = []
            for j, run in enumerate(runs):
                for pos, (doc, _score) in enumerate(run):
                    key = doc.doc_id
                    if key not in ids:
                        ids[key] = len(table)
                        table.append(doc)
                    else:
                        table[ids[key]] = doc  # doc_map keeps the LAST object seen for an id
                    mat[qi, off[j] + pos] = ids[key]
            doc_tables.append(table)
        dev = torch.device("cuda", self._device)
        if torch.cuda.current_device() != self._device:
            torch.cuda.set_device(dev)
        kk = max(1, min(int(k), _lib.RR_MAX_K))
        if kk != int(k):
            logger.warning(f"RRF top_k={k} outside [1, {_lib.RR_MAX_K}]: {kk} fused results are returned")
        idx, score, count = rrf_fuse_device(to_device(mat, dev, torch.int64), off, kk, c)
        idx_h, score_h, count_h = idx.cpu().tolist(), score.cpu().tolist(), count.cpu().tolist()
        out: List[List[Tuple[Any, float]]] = []
        for qi in range(nq):
            m = min(count_h[qi], k)
            out.append([(doc_tables[qi][i], float(s)) for i, s in zip(idx_h[qi][:m], score_h[qi][:m])])
        return out

    def _on_error(self, error: Exception, metrics: Any = None, **kwargs: Any) -> Optional[List[Tuple[Any, float]]]:
        logger.warning(f"RRF fusion failed: {error}")
        return []
