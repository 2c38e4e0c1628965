"""Row-sharded retrieval across the GPUs of one box.

One process per GPU (torchrun), ``torch.distributed`` over NCCL/NVLink.  The corpus is
cut into contiguous row blocks ``[g*N/G, (g+1)*N/G)``; queries are replicated.  Only
candidate lists cross NVLink:

  dense, exact single-index semantics in two small steps
    1. local Hamming top-k'           -> ONE all_gather [G, Q, k'] of packed (dist << 40 | row)
                                         keys -> merge (dist asc, row asc)
    2. each shard scores the global candidates it OWNS (-inf elsewhere)
                                      -> all_reduce(MAX) [Q, k'] -> rank, cut, filter
  BM25     local top-k (global idf/avgdl baked into the shard's impacts)
                                      -> all_gather -> merge (score desc, row asc)
  int8     (config 4) local exact int8 top-k -> all_gather -> merge (score desc, row asc)
  RRF      runs on the merged, replicated lists on every rank.

Payloads are tiny (Q*k'*12 B per rank), so the collectives are latency-bound; they are
issued on the same stream as the kernels (NCCL stream semantics) with no host sync in
between.  The compute ops are injected (``ops``): ``GpuShardOps`` (the product; CUDA
only, no CPU fallback) or a test double that the world_size-2 gloo tests supply.
"""

from __future__ import annotations

from typing import Any, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .index import DenseIndex, _stream


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: [lo, hi)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


class GpuShardOps:
    """Compute ops of one shard, all through the C ABI."""

    def __init__(self, index: DenseIndex) -> None:
        self.index = index
        self.device = index.device

    def quantize_queries(self, queries):
        return self.index.quantize_queries(queries)

    def hamming_topk(self, qcodes, k, tag_mask=0, tag_value=0, check_overflow=True):
        return self.index.hamming_topk(qcodes, k, tag_mask, tag_value, check_overflow=check_overflow)

    def merge_hamming(self, dist_all: torch.Tensor, idx_all: torch.Tensor, k: int):
        q, n_in = idx_all.shape
        out_d = torch.empty((q, k), dtype=torch.int32, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        _lib.call("rr_merge_hamming", dist_all.data_ptr(), idx_all.data_ptr(), q, n_in, k,
                  out_d.data_ptr(), out_i.data_ptr(), _stream())
        return out_d, out_i

    def pack_hamming(self, dist_loc: torch.Tensor, idx_loc: torch.Tensor) -> torch.Tensor:
        """(dist int32, global row int64) [Q, k] -> int64 keys dist << 40 | row (-1 = padding):
        the two lists of a shard travel in ONE collective."""
        keys = torch.empty(idx_loc.shape, dtype=torch.int64, device=self.device)
        _lib.call("rr_pack_hamming", dist_loc.data_ptr(), idx_loc.data_ptr(), idx_loc.numel(),
                  keys.data_ptr(), _stream())
        return keys

    def merge_hamming_gathered(self, keys_all: torch.Tensor, k: int):
        """keys_all int64 [G, Q, k'] exactly as all_gather_into_tensor leaves it (no transpose)."""
        g, q, k_in = keys_all.shape
        out_d = torch.empty((q, k), dtype=torch.int32, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        _lib.call("rr_merge_hamming_gathered", keys_all.data_ptr(), g, q, k_in, k, out_d.data_ptr(),
                  out_i.data_ptr(), _stream())
        return out_d, out_i

    def score_candidates(self, queries_f32, cand_idx, prefer_int8=True):
        return self.index.score_candidates(queries_f32, cand_idx, prefer_int8)

    def search_int8_exact(self, queries_i8, top_k, tag_mask=0, tag_value=0):
        return self.index.search_int8_exact(queries_i8, top_k, tag_mask, tag_value)

    def rescore(self, queries_f32, cand_idx, top_k, min_similarity, prefer_int8=True):
        score, idx, count = self.index.rescore(queries_f32, cand_idx, top_k, min_similarity, prefer_int8)
        return idx, score, count

    def rank_scored(self, scores, cand_idx, top_k, min_similarity):
        q, c = cand_idx.shape
        out_s = torch.empty((q, top_k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((q, top_k), dtype=torch.int64, device=self.device)
        out_c = torch.empty((q,), dtype=torch.int32, device=self.device)
        _lib.call("rr_rank_scored_f32", scores.data_ptr(), cand_idx.data_ptr(), q, c, top_k,
                  float(min_similarity), out_s.data_ptr(), out_i.data_ptr(), out_c.data_ptr(), _stream())
        return out_i, out_s, out_c

    def merge_scores_f64(self, score_all: torch.Tensor, idx_all: torch.Tensor, k: int):
        q, n_in = idx_all.shape
        out_s = torch.empty((q, k), dtype=torch.float64, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        out_c = torch.empty((q,), dtype=torch.int32, device=self.device)
        _lib.call("rr_merge_scores_f64", score_all.data_ptr(), idx_all.data_ptr(), q, n_in, k,
                  out_s.data_ptr(), out_i.data_ptr(), out_c.data_ptr(), _stream())
        return out_i, out_s, out_c

    def merge_scores_f64_gathered(self, words_all: torch.Tensor, k: int):
        """words_all int64 [G, 2, Q, k'] exactly as ONE all_gather of the shards' (score bits, row)
        buffers leaves it (no transpose copies)."""
        g, _two, q, k_in = words_all.shape
        out_s = torch.empty((q, k), dtype=torch.float64, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        out_c = torch.empty((q,), dtype=torch.int32, device=self.device)
        _lib.call("rr_merge_scores_f64_gathered", words_all.data_ptr(), g, q, k_in, k, out_s.data_ptr(),
                  out_i.data_ptr(), out_c.data_ptr(), _stream())
        return out_i, out_s, out_c

    def merge_scores_i32(self, score_all: torch.Tensor, idx_all: torch.Tensor, k: int):
        q, n_in = idx_all.shape
        out_s = torch.empty((q, k), dtype=torch.int32, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        _lib.call("rr_merge_scores_i32", score_all.data_ptr(), idx_all.data_ptr(), q, n_in, k,
                  out_s.data_ptr(), out_i.data_ptr(), _stream())
        return out_i, out_s


def _world(group: Optional[Any]) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _gather_raw(t: torch.Tensor, group: Optional[Any], comm: Optional[Any] = None) -> torch.Tensor:
    """[Q, k] on every rank -> [G, Q, k] on every rank (the collective's native layout).
    comm: an ``nccl.NcclComm`` - the collective is then issued on the CURRENT stream (capturable
    into a CUDA graph) instead of going through torch.distributed's process group."""
    world = _world(group)
    t = t.contiguous()
    buf = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if comm is not None and t.is_cuda:
        comm.all_gather(t, buf)
    elif t.is_cuda:
        dist.all_gather_into_tensor(buf, t, group=group)
    else:  # gloo (CPU tests of the host logic)
        dist.all_gather(list(buf.unbind(0)), t, group=group)
    return buf


def _gather_lists(t: torch.Tensor, group: Optional[Any], comm: Optional[Any] = None) -> torch.Tensor:
    """[Q, k] on every rank -> [Q, G*k] (rank-major within a query) on every rank."""
    world = _world(group)
    if world == 1:
        return t
    buf = _gather_raw(t, group, comm)
    return buf.permute(1, 0, 2).reshape(t.shape[0], world * t.shape[1]).contiguous()


def _all_reduce_max(t: torch.Tensor, group: Optional[Any], comm: Optional[Any] = None) -> None:
    if comm is not None and t.is_cuda:
        comm.all_reduce(t, "max")
    else:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)


class ShardedDenseSearch:
    """Two-stage quantised retrieval over a row-sharded corpus with single-index results."""

    def __init__(self, ops: Any, group: Optional[Any] = None, comm: Optional[Any] = None) -> None:
        self.ops = ops
        self.group = group
        self.comm = comm  # nccl.NcclComm: collectives on the current stream (CUDA-graph capturable)

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def overflow_total(self) -> int:
        """Tensor-core candidate-list overflows of ``check_overflow=False`` calls, summed over ALL
        ranks (an overflow on one shard makes the merged result inexact on every rank, so the
        per-rank counter alone is not enough).  Collective: every rank must call it."""
        index = getattr(self.ops, "index", None)
        local = index.tc_overflow_total() if index is not None else 0
        if self.world() == 1:
            return local
        t = torch.tensor([local], dtype=torch.int64, device=getattr(self.ops, "device", "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def search_quantized(self, queries, top_k: int, rescore_multiplier: float = 4.0,
                         use_rescoring: bool = True, min_similarity: float = 0.0,
                         tag_mask: int = 0, tag_value: int = 0, prefer_int8: bool = True,
                         check_overflow: bool = True
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """queries f32 [Q, D], identical on every rank.  -> (idx [Q,top_k] global rows,
        score f32 [Q,top_k], count int32 [Q]), identical on every rank."""
        ops = self.ops
        qf, qc = ops.quantize_queries(queries)
        candidate_k = int(top_k * rescore_multiplier) if use_rescoring else top_k
        candidate_k = max(1, min(candidate_k, _lib.RR_MAX_K))
        d_loc, i_loc = ops.hamming_topk(qc, candidate_k, tag_mask, tag_value, check_overflow=check_overflow)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1 and hasattr(ops, "pack_hamming"):
            # (dist, row) packed into one int64 key per entry: one collective, merged in place
            keys_all = _gather_raw(ops.pack_hamming(d_loc, i_loc), self.group, self.comm)
            _d, cand = ops.merge_hamming_gathered(keys_all, candidate_k)
        elif world > 1:
            d_all = _gather_lists(d_loc, self.group, self.comm)
            i_all = _gather_lists(i_loc, self.group, self.comm)
            _d, cand = ops.merge_hamming(d_all, i_all, candidate_k)
        else:
            cand = i_loc
        if not use_rescoring:
            idx = cand[:, :top_k].contiguous()
            score = torch.ones(idx.shape, dtype=torch.float32, device=idx.device)
            count = (idx >= 0).sum(dim=1).to(torch.int32)
            return idx, score, count
        # (also at world size 1: the scoring kernel splits a query's candidates over several CTAs,
        #  which beats the fused one-CTA-per-query score+rank kernel: 0.315 vs 0.331 ms per config-2 step)
        s = ops.score_candidates(qf, cand, prefer_int8)
        if world > 1:
            _all_reduce_max(s, self.group, self.comm)
        return ops.rank_scored(s, cand, top_k, min_similarity)


class ShardedBM25Search:
    """BM25 over a row-sharded document space.  ``local`` is a Bm25DeviceIndex (or a test
    double with the same ``search_batch``) built over this rank's documents with the
    GLOBAL idf / avgdl tables and ``row_base`` = first global row of the shard."""

    def __init__(self, local: Any, ops: Any, group: Optional[Any] = None, comm: Optional[Any] = None) -> None:
        self.local = local
        self.ops = ops
        self.group = group
        self.comm = comm

    def search_batch(self, q_terms, k: int, check: bool = True
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """check=False: no host synchronisation (see Bm25DeviceIndex.search_batch)."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1 and hasattr(self.ops, "merge_scores_f64_gathered") and hasattr(self.local, "search_batch_into"):
            # scores and rows of a shard travel in ONE collective and are merged where they land
            words = self.local.search_batch_into(q_terms, k, check=check)  # int64 [2, Q, k]
            return self.ops.merge_scores_f64_gathered(_gather_raw(words, self.group, self.comm), k)
        idx, score, count = self.local.search_batch(q_terms, k, check=check)
        if world == 1:
            return idx, score, count
        s_all = _gather_lists(score, self.group, self.comm)
        i_all = _gather_lists(idx, self.group, self.comm)
        return self.ops.merge_scores_f64(s_all, i_all, k)


class ShardedInt8Search:
    """BASELINE config 4: exact int8 x int8 -> int32 search over a row-sharded corpus.
    Every rank searches its shard (tensor cores for batches), the per-shard top-k
    (score, global row) lists are all_gathered and merged by (score desc, row asc)."""

    def __init__(self, ops: Any, group: Optional[Any] = None, comm: Optional[Any] = None) -> None:
        self.ops = ops
        self.group = group
        self.comm = comm

    def search(self, queries_i8, top_k: int, tag_mask: int = 0, tag_value: int = 0
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries_i8 int8 [Q, D], identical on every rank -> (idx int64 [Q,k], score int32 [Q,k])."""
        idx, score = self.ops.search_int8_exact(queries_i8, top_k, tag_mask, tag_value)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return idx, score
        s_all = _gather_lists(score, self.group, self.comm)
        i_all = _gather_lists(idx, self.group, self.comm)
        return self.ops.merge_scores_i32(s_all, i_all, top_k)
