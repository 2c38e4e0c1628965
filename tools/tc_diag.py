#!/usr/bin/env python
"""Diagnostic: tensor-core Hamming top-k vs the POPC scan on a small shape; prints the first mismatches."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402

n, dim, q, k = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (4096, 128, 16, 8)
idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False, capacity=n)
idx.add(synth_rows_device(0, n, dim, 4))
qc = idx.quantize_queries(synth_query_rows_device(0, q, dim, 4, n))[1]
d_tc, i_tc = idx.hamming_topk(qc, k, use_tc=True, check_overflow=False)[:2]
d_pc, i_pc = idx.hamming_topk(qc, k, use_tc=False)[:2]
torch.cuda.synchronize()
print("overflow", idx.tc_overflow_total())
print("dist equal", bool(torch.equal(d_tc, d_pc)), "idx equal", bool(torch.equal(i_tc, i_pc)))
bad = (d_tc != d_pc).any(dim=1).nonzero().flatten().tolist()
print("queries with wrong distances:", bad[:32], "of", q)
for qi in bad[:4]:
    print("q", qi, "tc", d_tc[qi].tolist(), i_tc[qi].tolist())
    print("q", qi, "pc", d_pc[qi].tolist(), i_pc[qi].tolist())
