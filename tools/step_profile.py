#!/usr/bin/env python
"""One hybrid step (BASELINE config 3) issued eagerly a few times on real inputs, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'rr::|tc_i8|bm25_|tau_keys|rescore_ring|tc_select' \\
        --launch-skip <build + 2 steps> -c 13 -o out python tools/step_profile.py

usage: step_profile.py [steps]     (prints how many rr:: kernels the build and each step launch)"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200 import _lib, synthetic  # noqa: E402
from radiant_rag_b200.bm25_index import Bm25DeviceIndex, synth_zipf_corpus_device  # noqa: E402
from radiant_rag_b200.hybrid import HybridSearch  # noqa: E402
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n, dim, nq, v, seed = 1_000_000, 768, 1024, 50_000, 2
    index = DenseIndex(dim, device=0, store_int8=False, store_f32=True, capacity=n)
    for a in range(0, n, 125_000):
        index.add(synth_rows_device(a, 125_000, dim, seed))
    ptr, toks = synth_zipf_corpus_device(n, v, seed, 200, device=0)
    bm = Bm25DeviceIndex.build(ptr, toks, v, None, None, 1.5, 0.75, device=0)
    del ptr, toks
    queries = synth_query_rows_device(0, nq, dim, seed, n)
    qt = torch.from_numpy(synthetic.zipf_queries(nq, 8, v, seed)).cuda()
    hybrid = HybridSearch(index, bm, rescore_multiplier=4.0, prefer_int8=False, overlap=False)  # one stream: a stable launch order for --launch-skip
    torch.cuda.synchronize()
    build_launches = _lib.launch_count
    for _ in range(steps):
        res = hybrid.search_batch(queries, qt, top_k=10, dense_top_k=100, bm25_top_k=100, rrf_k=60, check=False)
    torch.cuda.synchronize()
    if len(sys.argv) > 2 and sys.argv[2] == "exchange":
        # the merge kernels of the row-sharded exchange on buffers laid out as 8 ranks' all_gather would
        # leave them (every "rank" contributes this GPU's local lists): for ncu -k regex:merge_|pack_
        from radiant_rag_b200.sharded import GpuShardOps
        ops = GpuShardOps(index)
        qf, qc = index.quantize_queries(queries)
        d, i = index.hamming_topk(qc, 400, check_overflow=False)
        keys = ops.pack_hamming(d, i)
        keys_all = keys.unsqueeze(0).repeat(8, 1, 1).contiguous()
        words = bm.search_batch_into(qt, 100, check=False)
        words_all = words.unsqueeze(0).repeat(8, 1, 1, 1).contiguous()
        for _ in range(3):
            ops.merge_hamming_gathered(keys_all, 400)
            ops.merge_scores_f64_gathered(words_all, 100)
        torch.cuda.synchronize()
    print(json.dumps({"rr_launches_in_build": build_launches, "rr_launches_per_step": (_lib.launch_count - build_launches) // steps,
                      "unchecked_events": hybrid.unchecked_events(), "fused_nonempty": int((res.count > 0).sum())}))


if __name__ == "__main__":
    main()
