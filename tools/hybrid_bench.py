#!/usr/bin/env python
"""BASELINE config 3 on one GPU: BM25 over 1M synthetic docs (50k vocab, Zipf, ~200 tokens)
fused via RRF with the dense top-100, batch of 1024 queries, fused top-10.

Prints one JSON line with per-stage CUDA-event times and the hybrid queries/s, plus a
float64 bit-exactness check of a sample of BM25 results against the NumPy oracle restricted
to the documents returned (impacts recomputed on the host)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200 import _lib, synthetic  # noqa: E402
from radiant_rag_b200.agents import rrf_fuse_device  # noqa: E402
from radiant_rag_b200.bm25_index import Bm25DeviceIndex  # noqa: E402
from radiant_rag_b200.index import DenseIndex, _stream, synth_query_rows_device, synth_rows_device  # noqa: E402


def ev_time(fn, flush, reps=5, warm=2):
    for _ in range(warm):
        flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def main():
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    v, mean_len, nq, qlen, seed, dim = 50_000, 200, 1024, 8, 2, 768
    dev = torch.device("cuda", 0)
    _lib.init(0)
    t0 = time.time()
    lens = torch.empty(n_docs, dtype=torch.int32, device=dev)
    _lib.call("rr_synth_doc_lengths", lens.data_ptr(), 0, n_docs, seed, mean_len, _stream())
    ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens.to(torch.int64), 0, out=ptr[1:])
    total = int(ptr[-1].item())
    cdf = synthetic.zipf_cdf_u32(v)
    cdf_d = torch.from_numpy(cdf.view(np.int32)).to(dev)
    toks = torch.empty(total, dtype=torch.int32, device=dev)
    _lib.call("rr_synth_zipf_tokens", toks.data_ptr(), 0, total, seed, cdf_d.data_ptr(), v, _stream())
    avgdl = total / n_docs
    bm = Bm25DeviceIndex.build(ptr, toks, v, None, avgdl, 1.5, 0.75, device=0, tile_docs=int(os.environ.get('RR_BM25_TILE', '8192')))
    del toks
    build_s = time.time() - t0
    qt = torch.from_numpy(synthetic.zipf_queries(nq, qlen, v, seed)).to(dev)

    index = DenseIndex(dim, device=0, store_int8=False, store_f32=True, capacity=n_docs)
    for lo in range(0, n_docs, 125_000):
        index.add(synth_rows_device(lo, min(125_000, n_docs - lo), dim, seed))
    queries = synth_query_rows_device(0, nq, dim, seed, n_docs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    out = {}

    def dense():
        out["dense"] = index.search_quantized(queries, 100, rescore_multiplier=4.0, prefer_int8=False,
                                              check_overflow=False)

    def sparse():
        out["bm25"] = bm.search_batch(qt, 100)

    def fuse():
        runs = torch.cat([out["dense"][0], out["bm25"][0]], dim=1).contiguous()
        out["fused"] = rrf_fuse_device(runs, [0, 100, 200], 10, 60)

    def hybrid():
        dense()
        sparse()
        fuse()

    t_dense = ev_time(dense, flush)
    t_bm25 = ev_time(sparse, flush)
    t_rrf = ev_time(fuse, flush)
    t_all = ev_time(hybrid, flush)
    assert index.tc_overflow_total() == 0

    # parity of a sample: BM25 scores of the returned docs recomputed on the host in the
    # reference's float64 operation order from the device postings (exact equality)
    b_idx, b_score, b_count = out["bm25"]
    post_rows = bm.post_row.cpu().numpy().astype(np.int64)
    post_imp = bm.post_impact.cpu().numpy()
    ttp = bm.tile_term_ptr.cpu().numpy()
    qt_h = qt.cpu().numpy()
    checked = 0
    for qi in range(0, nq, 128):
        m = int(b_count[qi])
        rows = b_idx[qi, :m].cpu().numpy()
        want = np.zeros(m, dtype=np.float64)
        for t in qt_h[qi]:
            for j, r in enumerate(rows):
                tile = r // bm.tile_docs
                lo, hi = ttp[tile, t], ttp[tile, t + 1]
                hit = np.nonzero(post_rows[lo:hi] == r)[0]  # segment order is bank-interleaved
                if hit.size:
                    want[j] += post_imp[lo + hit[0]]
        assert np.array_equal(want, b_score[qi, :m].cpu().numpy()), qi
        checked += m
    # CPU port of the same BM25 top-100 (NumPy, one core) on a bounded sample of the batch: scores
    # accumulated per document in query-token order from the same postings, exact top-k by
    # (score desc, row asc); timed, and compared bit for bit with the GPU's lists
    sample = 16
    n_tiles = ttp.shape[0]
    rows_all = np.arange(n_docs)
    agree = 0
    t0 = time.perf_counter()
    for qi in range(sample):
        sc = np.zeros(n_docs, dtype=np.float64)
        for t in qt_h[qi]:
            for tile in range(n_tiles):
                lo, hi = ttp[tile, t], ttp[tile, t + 1]
                if hi > lo:
                    sc[post_rows[lo:hi]] += post_imp[lo:hi]  # a document occurs once per segment
        order = np.lexsort((rows_all, -sc))[:100]
        order = order[sc[order] > 0]
        m = int(b_count[qi])
        if (m == order.size and np.array_equal(b_idx[qi, :m].cpu().numpy(), order)
                and np.array_equal(b_score[qi, :m].cpu().numpy(), sc[order])):
            agree += 1
    cpu_s = time.perf_counter() - t0
    f_idx, f_score, f_count = out["fused"]
    line = {
        "workload": "config3: BM25 over %d docs (50k vocab, Zipf, ~200 tokens) + dense top-100 (1M x 768 two-stage) "
                    "fused by RRF, batch 1024, fused top-10" % n_docs,
        "postings": bm.n_postings, "index_build_s": round(build_s, 1),
        "ms": {"dense_top100": t_dense, "bm25_top100": t_bm25, "rrf_top10": t_rrf, "hybrid_total": t_all},
        "hybrid_queries_per_s": nq / (t_all * 1e-3),
        "bm25_queries_per_s": nq / (t_bm25 * 1e-3),
        "bm25_posting_GBs": float((bm.post_impact.numel() and 12.0) * sum(
            int(ttp[:, t + 1].sum() - ttp[:, t].sum()) for t in qt_h.ravel()) / (t_bm25 * 1e-3) / 1e9),
        "bm25_scores_checked_exact": checked,
        "fused_nonempty": int((f_count > 0).sum()),
        "cpu_baseline": {"value": sample / cpu_s, "unit": "BM25 top-100 queries/s", "cores": 1, "kind": "port",
                         "sample": f"{sample} of the {nq} queries over all {n_docs} docs, NumPy scatter-add over the "
                                   "same postings + lexsort, 1 thread",
                         "gpu_matches_cpu_on_sample_bit_exact": f"{agree}/{sample}"},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
