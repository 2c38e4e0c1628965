#!/usr/bin/env python
"""Micro-benchmark of the Hamming scan (rr_hamming_topk) over several shapes.

Reports, per shape, the CUDA-event time of scan + merge with L2 flushed between
iterations, the algorithmic code bytes / time as a fraction of the measured HBM copy
peak, and the word-popcount rate.  Used to fill DESIGN.md and profiles/."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402


def build(n, dim, seed=4, pm1=True):
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False, capacity=n)
    step = 500_000
    for lo in range(0, n, step):
        idx.add(synth_rows_device(lo, min(step, n - lo), dim, seed))
    torch.cuda.synchronize()
    return idx


def time_call(fn, flush, reps=10, warm=3):
    for _ in range(warm):
        flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return sum(ts) / len(ts), ts[0]


def main():
    peaks = ROOT / "MEASURED_PEAKS.json"
    hbm = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = [  # (rows, dim, [(q, k), ...])
        (1_000_000, 768, [(1, 200), (2, 200), (4, 200), (16, 200), (64, 200), (256, 200)]),
        (12_500_000, 1024, [(1, 40), (2, 40), (4, 40), (8, 40), (64, 40), (256, 40), (1024, 40)]),
    ]
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        shapes = shapes[:1]
    for n, dim, cases in shapes:
        t0 = time.time()
        idx = build(n, dim)
        build_s = time.time() - t0
        for q, k in cases:
            queries = synth_query_rows_device(0, q, dim, 4, n)
            _qf, qc = idx.quantize_queries(queries)
            mean_ms, best_ms = time_call(lambda: idx.hamming_topk(qc, k, use_tc=False), flush)
            code_bytes = n * idx.words * 4
            tc = None
            if q >= 16:
                tc_ms, tc_best = time_call(lambda: idx.hamming_topk(qc, k, use_tc=True, check_overflow=False), flush)
                macs = q * n * dim
                tc = {"ms_mean": round(tc_ms, 4), "ms_best": round(tc_best, 4),
                      "int8_TOPS": round(2 * macs / (tc_ms * 1e-3) / 1e12, 1),
                      
                      "queries_per_s": round(q / (tc_ms * 1e-3), 1), "overflow": idx.tc_overflow_total()}
            out = {
                "rows": n, "dim": dim, "q": q, "k": k, "ms_mean": round(mean_ms, 4), "ms_best": round(best_ms, 4),
                "tensor_core": tc,
                "code_GB": round(code_bytes / 1e9, 3),
                "hbm_GBs": round(code_bytes / (mean_ms * 1e-3) / 1e9, 1),
                "hbm_frac_of_measured": round(code_bytes / (mean_ms * 1e-3) / 1e9 / hbm, 4),
                "Tpopc32_per_s": round(q * n * idx.words / (mean_ms * 1e-3) / 1e12, 3),
                "queries_per_s": round(q / (mean_ms * 1e-3), 1), "build_s": round(build_s, 1),
            }
            print(json.dumps(out), flush=True)
        del idx
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
