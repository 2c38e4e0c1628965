#!/usr/bin/env python
"""Tiny invocation of every kernel family (for `compute-sanitizer --tool memcheck`)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200 import synthetic  # noqa: E402
from radiant_rag_b200.agents import rrf_fuse_device  # noqa: E402
from radiant_rag_b200.bm25_index import Bm25DeviceIndex  # noqa: E402
from radiant_rag_b200.index import DenseIndex  # noqa: E402

n, dim, nq = 5000, 256, 20
corpus = synthetic.hash_rows_f32(0, n, dim, seed=3)
queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=3, n_corpus=n)
ranges = np.stack([corpus.min(axis=0), corpus.max(axis=0)]).astype(np.float32)  # per-dim calibration
idx = DenseIndex(dim, device=0, store_int8=True, store_f32=True, int8_ranges=ranges)
idx.add(corpus, np.where(np.arange(n) % 2 == 0, 1, 2).astype(np.uint8))
qf, qc = idx.quantize_queries(queries)
a = idx.hamming_topk(qc, 40, use_tc=False)
b = idx.hamming_topk(qc, 40, use_tc=True, check_overflow=False)
assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
c = idx.hamming_topk(qc[:1].contiguous(), 40, 3, 1, use_tc=False)
idx.search_quantized(queries, 10)
idx.search_exact(queries[:9], 10)
q8 = idx.quantize_int8_queries(queries)
x = idx.search_int8_exact(q8, 10, use_tc=True)
y = idx.search_int8_exact(q8, 10, use_tc=False)
assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
idx.rescore_int8_symmetric(q8, a[1], 10)
ptr, toks = synthetic.zipf_corpus(3000, 400, seed=2, mean_len=30)
avgdl = float(np.diff(ptr).mean())
bm = Bm25DeviceIndex.build(ptr, toks, 400, None, avgdl, 1.5, 0.75, device=0, tile_docs=1024)  # idf built on device/host
bi, bs, bc = bm.search_batch(synthetic.zipf_queries(nq, 5, 400, seed=2), 20)
runs = torch.cat([a[1][:, :20], bi], dim=1).contiguous()
rrf_fuse_device(runs, [0, 20, 40], 10, 60)
torch.cuda.synchronize()
print("sanitize smoke ok")
