#!/usr/bin/env python
"""BASELINE config 4: 10M x 1024-dim int8-quantised exact search on tensor cores, batch 4096,
row-sharded over 2 / 4 GPUs (torchrun, one rank per GPU; NCCL all_gather of the per-shard top-k).

    python -m torch.distributed.run --nproc-per-node 4 tools/config4_bench.py [rows_total] [batch]

Shards are generated on device.  Rank 0 prints one JSON line: device-timed queries/s (max over
ranks), the int8 tensor throughput per GPU, and a parity check of a sample of queries against
the CUDA-core DP4A path of the same shard set (bit-exact scores and ids)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200 import synthetic  # noqa: E402
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402
from radiant_rag_b200.sharded import GpuShardOps, ShardedInt8Search, _gather_lists, shard_range  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dim, seed, top_k = 1024, 3, 10
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)  # NCCL's version banner goes to stderr
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lo, hi = shard_range(n_total, rank, world)
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    index = DenseIndex(dim, device=local, store_int8=True, store_f32=False, int8_ranges=ranges, row_base=lo,
                       capacity=hi - lo)
    for a in range(lo, hi, 250_000):
        index.add(synth_rows_device(a, min(250_000, hi - a), dim, seed, dev))
    q8 = index.quantize_int8_queries(synth_query_rows_device(0, nq, dim, seed, n_total, dev))
    torch.cuda.synchronize()
    ops = GpuShardOps(index)
    search = ShardedInt8Search(ops)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        flush.fill_(1)
        idx, score = search.search(q8, top_k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps, total = 3, 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx, score = search.search(q8, top_k)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    t = torch.tensor([total / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # parity on a sample: the DP4A CUDA-core path over the same shards, merged the same way
    sample = 64
    i_loc, s_loc = index.search_int8_exact(q8[:sample].contiguous(), top_k, use_tc=False)
    if world > 1:
        i_ref, s_ref = ops.merge_scores_i32(_gather_lists(s_loc, None), _gather_lists(i_loc, None), top_k)
    else:
        i_ref, s_ref = i_loc, s_loc
    same = bool(torch.equal(i_ref, idx[:sample]) and torch.equal(s_ref, score[:sample]))
    if rank == 0:
        print(json.dumps({
            "workload": f"config4: {n_total} x {dim} int8 exact search (tcgen05 kind::i8), batch {nq}, top-{top_k}, "
                        f"row-sharded x{world} (NCCL all_gather of per-shard top-k + merge)",
            "ms_per_batch": ms, "queries_per_s": nq / (ms * 1e-3), "rows_per_gpu": hi - lo,
            "int8_TOPS_per_gpu": 2.0 * (hi - lo) * dim * nq / (ms * 1e-3) / 1e12,
            "sample_matches_dp4a_path_bit_exact": same, "sample_queries": sample,
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
