"""CUDA-graph replay of a fixed-shape retrieval step.

A search step is ~10 short kernels (plus two or three NCCL collectives when the corpus is
sharded); at serving batch sizes the GPU time of a step is a few hundred microseconds, the
same order as the host time needed to issue it from Python.  ``GraphedSearch`` captures one
step (kernels, memsets, collectives) into a CUDA graph with static input / output buffers,
so a step costs one host->device copy of the queries, one graph launch and the read-back.

Only calls that need no host decision can be captured: the tensor-core overflow counter is
accumulated on the device (``check_overflow=False``) and read by the caller when it consumes
results (``DenseIndex.tc_overflow_total``).
"""

from __future__ import annotations

from typing import Any, Callable, Tuple

import torch

from . import _lib


class _StagedInput:
    """Host queries reach the graph's static input through two device staging buffers filled on a
    copy stream, so the host->device copy of step i+1 overlaps the kernels of step i (the
    host runs ahead of the GPU once a step is a graph launch)."""

    def _init_staging(self) -> None:
        dev = self.device
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._staging = [torch.empty_like(self.static_in) for _ in range(2)]
        self._filled = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._turn = 0

    def load(self, queries: torch.Tensor) -> None:
        """Stage this step's queries into the graph's input buffer."""
        if queries.is_cuda:
            self.static_in.copy_(queries, non_blocking=True)
            return
        b = self._turn
        self._turn ^= 1
        main = torch.cuda.current_stream(self.device)
        cs = self._copy_stream
        cs.wait_event(self._consumed[b])  # no-op until the buffer has been used once
        with torch.cuda.stream(cs):
            self._staging[b].copy_(queries, non_blocking=True)  # async for pinned host memory
            self._filled[b].record(cs)
        main.wait_event(self._filled[b])
        self.static_in.copy_(self._staging[b], non_blocking=True)
        self._consumed[b].record(main)


class GraphedSearch(_StagedInput):
    """graph = GraphedSearch(lambda q: search.search_quantized(q, 10, check_overflow=False),
                             n_queries, dim, device)
    idx, score, count = graph(queries)      # queries: host (pinned) or device f32 [n_queries, dim]

    The returned tensors are the graph's static outputs: they are overwritten by the next
    replay, copy them out (``.to('cpu', non_blocking=True)`` or ``copy_``) before replaying."""

    def __init__(self, step: Callable[[torch.Tensor], Any], n_queries: int, dim: int, device,
                 dtype: torch.dtype = torch.float32, warmup: int = 3) -> None:
        self.device = torch.device(device)
        self.static_in = torch.zeros((n_queries, dim), dtype=dtype, device=self.device)
        # warm-up on a side stream (lazy module loads, NCCL channel setup, allocator growth)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step(self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        launches0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.static_out = step(self.static_in)
        self.kernels_per_replay = _lib.launch_count - launches0  # our kernels inside the graph
        self._init_staging()

    def replay(self) -> Any:
        self.graph.replay()
        _lib.launch_count += self.kernels_per_replay
        return self.static_out

    def __call__(self, queries: torch.Tensor) -> Any:
        self.load(queries)
        return self.replay()


def _capture(fn: Callable[[], Any], device: torch.device, warmup: int = 3) -> Tuple[torch.cuda.CUDAGraph, Any, int]:
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn()
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    graph = torch.cuda.CUDAGraph()
    launches0 = _lib.launch_count
    with torch.cuda.graph(graph):
        out = fn()
    return graph, out, _lib.launch_count - launches0


class GraphedShardedSearch(_StagedInput):
    """The row-sharded two-stage step (sharded.ShardedDenseSearch.search_quantized) with its
    three compute segments captured as CUDA graphs and the NCCL exchanges issued eagerly
    between them:

        graph 1  quantise queries, local Hamming top-k', pack (dist, row) into int64 keys
        NCCL     ONE all_gather of the packed lists
        graph 2  merge to the global top-k', score the candidates this shard owns
        NCCL     all_reduce(MAX) of the scores
        graph 3  rank, cut, filter

    Per step the host issues 3 graph launches and 2 collectives instead of ~25 calls, which
    is what bounds a sharded step whose GPU time is ~100 us.  Results are identical to the
    eager path (same kernels, same order)."""

    def __init__(self, search: Any, n_queries: int, dim: int, top_k: int, rescore_multiplier: float = 4.0,
                 min_similarity: float = 0.0, prefer_int8: bool = True, tag_mask: int = 0,
                 tag_value: int = 0) -> None:
        import torch.distributed as dist

        ops = search.ops
        self.group = search.group
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if self.world < 2:
            raise ValueError("GraphedShardedSearch is for world_size >= 2; use GraphedSearch on one GPU")
        self.device = torch.device(ops.device)
        dev = self.device
        nq = n_queries
        cand_k = max(1, min(int(top_k * rescore_multiplier), _lib.RR_MAX_K))
        self.static_in = torch.zeros((nq, dim), dtype=torch.float32, device=dev)

        def seg1():
            qf, qc = ops.quantize_queries(self.static_in)
            d, i = ops.hamming_topk(qc, cand_k, tag_mask, tag_value, check_overflow=False)
            return qf, ops.pack_hamming(d, i)  # (dist, row) as one int64 key per entry

        self.g1, (qf, self.key_loc), k1 = _capture(seg1, dev)
        self.key_buf = torch.full((self.world, nq, cand_k), -1, dtype=torch.int64, device=dev)

        def seg2():
            _d, cand = ops.merge_hamming_gathered(self.key_buf, cand_k)  # gathered layout, no transpose
            return cand, ops.score_candidates(qf, cand, prefer_int8)

        self.g2, (cand, self.scores), k2 = _capture(seg2, dev)

        def seg3():
            return ops.rank_scored(self.scores, cand, top_k, min_similarity)

        self.g3, self.static_out, k3 = _capture(seg3, dev)
        self.kernels_per_replay = k1 + k2 + k3
        self._init_staging()

    def replay(self) -> Any:
        import torch.distributed as dist

        self.g1.replay()
        dist.all_gather_into_tensor(self.key_buf, self.key_loc, group=self.group)
        self.g2.replay()
        dist.all_reduce(self.scores, op=dist.ReduceOp.MAX, group=self.group)
        self.g3.replay()
        _lib.launch_count += self.kernels_per_replay
        return self.static_out

    def __call__(self, queries: torch.Tensor) -> Any:
        self.load(queries)
        return self.replay()
