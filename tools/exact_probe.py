#!/usr/bin/env python
"""Times the exact float32 cosine scan (R5, rr_exact_search_f32) and int8 exact search (config 4 shape, scaled)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402
from radiant_rag_b200 import synthetic  # noqa: E402


def ev(fn, reps=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


n, dim = 1_000_000, 768
idx = DenseIndex(dim, device=0, store_int8=False, store_f32=True, capacity=n)
for lo in range(0, n, 125_000):
    idx.add(synth_rows_device(lo, 125_000, dim, 1))
for q in (1, 2, 4, 8, 64):
    qs = synth_query_rows_device(0, q, dim, 1, n)
    ms = ev(lambda: idx.search_exact(qs, 10, use_tc=False), reps=5 if q >= 8 else 20)
    print(json.dumps({"op": "exact_f32_cuda_cores", "rows": n, "dim": dim, "q": q, "ms": ms,
                      "GBs": n * dim * 4 * ((q + 7) // 8) / ms / 1e6, "queries_per_s": q / ms * 1e3}), flush=True)
idx.search_exact(synth_query_rows_device(0, 8, dim, 1, n), 10, use_tc=True)  # row norms are computed once, here
for q, k in ((8, 10), (64, 10), (128, 10), (256, 10), (1024, 10), (64, 100), (1024, 100)):
    qs = synth_query_rows_device(0, q, dim, 1, n)
    ms = ev(lambda: idx.search_exact(qs, k, use_tc=True, check_overflow=False))
    print(json.dumps({"op": "exact_f32_tc", "rows": n, "dim": dim, "q": q, "k": k, "ms": ms,
                      "GBs_per_query_block": n * dim * 4 * ((q + 127) // 128) / ms / 1e6,
                      "tf32_TFLOPs": 2.0 * n * dim * q / ms / 1e9, "queries_per_s": q / ms * 1e3,
                      "overflow_events": idx.tc_overflow_total()}), flush=True)
    idx.tc_overflow_reset()
if "--f32-only" in sys.argv:
    sys.exit(0)
del idx
torch.cuda.empty_cache()
# config 4 shape scaled to one GPU: 2.5M x 1024 int8, 4096 queries
n, dim, q = 2_500_000, 1024, 4096
bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
idx = DenseIndex(dim, device=0, store_int8=True, store_f32=False, int8_ranges=ranges, capacity=n)
for lo in range(0, n, 250_000):
    idx.add(synth_rows_device(lo, 250_000, dim, 3))
q8 = idx.quantize_int8_queries(synth_query_rows_device(0, q, dim, 3, n))
ms = ev(lambda: idx.search_int8_exact(q8, 10, use_tc=True), reps=2)
print(json.dumps({"op": "int8_exact_tc", "rows": n, "dim": dim, "q": q, "ms": ms, "int8_TOPS": 2.0 * n * dim * q / ms / 1e9,
                  "queries_per_s": q / ms * 1e3}), flush=True)
