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


class GraphedSearch:
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

    def load(self, queries: torch.Tensor) -> None:
        """Stage this step's queries into the graph's input buffer (async for pinned host memory)."""
        self.static_in.copy_(queries, non_blocking=True)

    def replay(self) -> Any:
        self.graph.replay()
        _lib.launch_count += self.kernels_per_replay
        return self.static_out

    def __call__(self, queries: torch.Tensor) -> Any:
        self.load(queries)
        return self.replay()
