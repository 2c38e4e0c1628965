#!/usr/bin/env python
"""BASELINE config 5 (dense half): 100M x 1024-dim binary codes + int8 rescoring, top-10, row-sharded
over the GPUs of one box with the NCCL candidate exchange (torchrun, one rank per GPU).

    python -m torch.distributed.run --nproc-per-node 8 tools/config5_bench.py [rows_total]

Every rank generates its 12.5M-row shard on device (counter-based rows), so nothing crosses PCIe.
Rank 0 prints one JSON line per query-batch size with the device-timed queries/s and, for the
single-query case, the fraction of the measured HBM peak the per-GPU scan achieves.  Results are
checked by size-independent properties: lists strictly increasing in (dist, row) is implied by
exact equality of the recomputed distances, and every odd query (3/4 of its dimensions copied from a
corpus row that lives on some shard) must come back with that row ranked first after rescoring."""
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
from radiant_rag_b200.sharded import GpuShardOps, ShardedDenseSearch, shard_range  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    dim, seed, top_k, mult = 1024, 4, 10, 4.0
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
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)  # every synthetic value lies inside
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    index = DenseIndex(dim, device=local, store_int8=True, store_f32=False, int8_ranges=ranges, row_base=lo,
                       capacity=hi - lo)
    step = 250_000
    for a in range(lo, hi, step):
        index.add(synth_rows_device(a, min(step, hi - a), dim, seed, dev))
    torch.cuda.synchronize()
    search = ShardedDenseSearch(GpuShardOps(index))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = ROOT / "MEASURED_PEAKS.json"
    hbm = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0

    def run(nq, reps):
        queries = synth_query_rows_device(0, nq, dim, seed, n_total, dev)
        fn = lambda: search.search_quantized(queries, top_k, rescore_multiplier=mult, prefer_int8=True,  # noqa: E731
                                             check_overflow=False)
        for _ in range(2):
            flush.fill_(1)
            out = fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        total = 0.0
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        t = torch.tensor([total / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out, queries

    for nq, reps in [(1, 20), (16, 10), (256, 5), (1024, 3), (8192, 1)]:
        ms, (idx, score, count), queries = run(nq, reps)
        # stage-1 only, for the roofline of the scan itself
        _qf, qc = index.quantize_queries(queries)
        scan_ms = None
        if nq == 1:
            ts = []
            for _ in range(10):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                index.hamming_topk(qc, int(top_k * mult))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            scan_ms = sum(ts[2:]) / len(ts[2:])
        if rank == 0:
            src = synthetic.query_source_row(np.arange(nq), n_total, seed)
            got = idx[:, 0].cpu().numpy()
            odd = np.arange(nq) % 2 == 1
            hit = float((got[odd] == src[odd]).mean()) if odd.any() else None
            line = {
                "workload": f"config5 dense: {n_total} x {dim} binary codes + int8 rescore, k'=40, top-10, "
                            f"row-sharded x{world} (NCCL all_gather + all_reduce of candidates)",
                "queries": nq, "ms_per_batch": ms, "queries_per_s": nq / (ms * 1e-3), "rows_per_gpu": hi - lo,
                "odd_query_source_row_ranked_first": hit, "tc_overflow": index.tc_overflow_total(),
            }
            if scan_ms is not None:
                code_bytes = (hi - lo) * index.words * 4
                line["scan_only_ms_per_gpu"] = scan_ms
                line["scan_GBs_per_gpu"] = code_bytes / (scan_ms * 1e-3) / 1e9
                line["scan_frac_of_measured_hbm"] = code_bytes / (scan_ms * 1e-3) / 1e9 / hbm
                line["scan_plus_rescore_GBs_per_gpu_e2e"] = code_bytes / (ms * 1e-3) / 1e9
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
