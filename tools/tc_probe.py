#!/usr/bin/env python
"""Runs the tensor-core Hamming search a few times on one shape (for ncu)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402

n, dim, q, k = (int(x) for x in sys.argv[1:5])
idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False, capacity=n)
for lo in range(0, n, 500_000):
    idx.add(synth_rows_device(lo, min(500_000, n - lo), dim, 4))
qc = idx.quantize_queries(synth_query_rows_device(0, q, dim, 4, n))[1]
for _ in range(3):
    idx.hamming_topk(qc, k, use_tc=True, check_overflow=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    idx.hamming_topk(qc, k, use_tc=True, check_overflow=False)
b.record()
torch.cuda.synchronize()
import os
print("debug", os.environ.get("RR_TC_DEBUG", "0"), "ms_per_call", a.elapsed_time(b) / 5)
print("overflow", idx.tc_overflow_total())
