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
    """The row-sharded two-stage step (sharded.ShardedDenseSearch.search_quantized) as ONE CUDA graph:
    quantise, local Hamming top-k', pack, all_gather, merge, score the owned candidates,
    all_reduce(MAX), rank - kernels and collectives alike.  The collectives are NCCL calls issued on
    the capturing stream (nccl.NcclComm); per step the host launches one graph instead of ~25 calls,
    which is what bounds a sharded step whose GPU time is ~100 us.  Results are identical to the
    eager path (same kernels, same order).  Every rank must replay in lock-step."""

    def __init__(self, search: Any, n_queries: int, dim: int, top_k: int, rescore_multiplier: float = 4.0,
                 min_similarity: float = 0.0, prefer_int8: bool = True, tag_mask: int = 0,
                 tag_value: int = 0) -> None:
        import torch.distributed as dist

        from .nccl import NcclComm

        self.search = search
        self.group = search.group
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if self.world < 2:
            raise ValueError("GraphedShardedSearch is for world_size >= 2; use GraphedSearch on one GPU")
        self.device = torch.device(search.ops.device)
        if search.comm is None:
            search.comm = NcclComm(self.device, self.group)
        self.static_in = torch.zeros((n_queries, dim), dtype=torch.float32, device=self.device)

        def step():
            return search.search_quantized(self.static_in, top_k, rescore_multiplier=rescore_multiplier,
                                           min_similarity=min_similarity, tag_mask=tag_mask, tag_value=tag_value,
                                           prefer_int8=prefer_int8, check_overflow=False)

        self.graph, self.static_out, self.kernels_per_replay = _capture(step, self.device)
        self._init_staging()

    def replay(self) -> Any:
        self.graph.replay()
        _lib.launch_count += self.kernels_per_replay
        return self.static_out

    def __call__(self, queries: torch.Tensor) -> Any:
        self.load(queries)
        return self.replay()
