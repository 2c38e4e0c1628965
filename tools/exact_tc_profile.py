#!/usr/bin/env python
"""The tensor-core exact float32 scan (rr_exact_search_f32_tc) on 1M x 768 rows, 64 queries, three times, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'tc_tf32|tx_refine|tc_select_lists' \\
        --launch-skip 8 -c 4 -o out python tools/exact_tc_profile.py
(a call launches: qnorm, sample scan, tau, filter scan, refine, select; the first call also the row norms)"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402

n, dim, q = 1_000_000, 768, int(sys.argv[1]) if len(sys.argv) > 1 else 64
idx = DenseIndex(dim, device=0, store_int8=False, store_f32=True, capacity=n)
for lo in range(0, n, 125_000):
    idx.add(synth_rows_device(lo, 125_000, dim, 1))
qs = synth_query_rows_device(0, q, dim, 1, n)
for _ in range(3):
    out = idx.search_exact(qs, 10, use_tc=True, check_overflow=False)
torch.cuda.synchronize()
print("overflow events", idx.tc_overflow_total(), "first ids", out[0][:2, :3].tolist())
