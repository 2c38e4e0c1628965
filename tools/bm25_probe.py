#!/usr/bin/env python
"""BM25 at BASELINE config 3 (1M docs, 1024 queries x 8 tokens, top-100) on one GPU: per-kernel
CUDA-event times of the batched path, the exact kernel's time on a sub-batch, for profiling runs
(ncu -k regex:bm25_) and tuning.  usage: bm25_probe.py [n_docs] [n_queries] [reps] [tile_docs] [head_terms]"""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200 import _lib, synthetic  # noqa: E402
from radiant_rag_b200.bm25_index import Bm25DeviceIndex, synth_zipf_corpus_device  # noqa: E402


def main():
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    tile = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    head_terms = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    v, seed = 50_000, 2
    ptr, toks = synth_zipf_corpus_device(n_docs, v, seed, 200, device=0)
    bm = Bm25DeviceIndex.build(ptr, toks, v, None, None, 1.5, 0.75, device=0, tile_docs=tile, head_terms=head_terms)
    bm.fast_min_docs = 0
    del ptr, toks
    qt = torch.from_numpy(synthetic.zipf_queries(nq, 8, v, seed)).cuda()
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        bm.search_batch(qt, 100, check=False)
    torch.cuda.synchronize()
    lib.rr_bm25_timing(1)
    parts = [0.0] * 4
    buf = (C.c_float * 4)()
    for _ in range(reps):
        flush.fill_(1)
        bm.search_batch(qt, 100, check=False)
        lib.rr_bm25_last_timing_ms(buf)
        parts = [p + float(b) for p, b in zip(parts, buf)]
    lib.rr_bm25_timing(0)
    print(json.dumps({"lib": str(_lib.LIB_PATH.name), "tile": tile, "n_docs": n_docs, "queries": nq, "n_head": bm.n_head, "postings": bm.n_postings,
                      "flagged": bm.inexact_total(),
                      "ms": dict(zip(["sample", "tau", "filter", "refine"], [p / reps for p in parts]))}))


if __name__ == "__main__":
    main()
