"""CPU, world_size 2 over gloo: the row-sharded search returns exactly the
single-index answer.  Compute ops are the oracle-backed test double (tests/cpu_ops.py);
what is under test is the product's sharding / allgather / merge / all-reduce plumbing."""

import os
import socket
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from oracle.bm25 import BM25Oracle
        from radiant_rag_b200 import synthetic
        from radiant_rag_b200.sharded import ShardedBM25Search, ShardedDenseSearch, ShardedInt8Search, shard_range
        from tests.cpu_ops import CpuBm25Shard, CpuShardOps

        n, dim, nq, top_k = 3001, 128, 12, 10  # odd n: ragged last shard
        corpus = synthetic.hash_rows_f32(0, n, dim, seed=21)
        queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=21, n_corpus=n)
        ranges = oracle.calculate_int8_ranges(corpus)
        i8 = oracle.quantize_int8(corpus, ranges)
        lo, hi = shard_range(n, rank, world)
        ops = CpuShardOps(corpus[lo:hi], i8[lo:hi], lo)
        search = ShardedDenseSearch(ops)
        results = {}
        for tag, kwargs in {"rescore": {}, "norescore": {"use_rescoring": False},
                            "minsim": {"min_similarity": 0.3}}.items():
            idx, score, count = search.search_quantized(queries, top_k, rescore_multiplier=4.0, **kwargs)
            results[tag] = (idx.numpy(), score.numpy(), count.numpy())

        # config 4: exact int8 search, per-shard top-k merged by (score desc, row asc)
        q8 = oracle.quantize_int8(queries, ranges)
        iidx, iscore = ShardedInt8Search(ops).search(q8, 7)
        results["int8"] = (iidx.numpy(), iscore.numpy())

        # BM25: shard-local postings, global idf / avgdl
        n_docs, v = 1203, 300
        ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=4, mean_len=25)
        glob = BM25Oracle(ptr, toks, v)
        dlo, dhi = shard_range(n_docs, rank, world)
        sub_ptr = ptr[dlo:dhi + 1] - ptr[dlo]
        sub = BM25Oracle(sub_ptr, toks[ptr[dlo]:ptr[dhi]], v, idf=glob.idf, avgdl=glob.avgdl)
        sub.known = glob.known
        qt = synthetic.zipf_queries(nq, 5, v, seed=4)
        bidx, bscore, bcount = ShardedBM25Search(CpuBm25Shard(sub, dlo), ops).search_batch(qt, 15)
        results["bm25"] = (bidx.numpy(), bscore.numpy(), bcount.numpy())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
                 **{f"{k}_{j}": a for k, t in results.items() for j, a in enumerate(t)})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_equals_single_index():
    import oracle
    from oracle.bm25 import BM25Oracle
    from radiant_rag_b200 import synthetic

    world = 2
    out_dir = tempfile.mkdtemp()
    mp.spawn(_worker, args=(world, _free_port(), out_dir), nprocs=world, join=True)
    r0 = np.load(os.path.join(out_dir, "rank0.npz"))
    r1 = np.load(os.path.join(out_dir, "rank1.npz"))
    for k in r0.files:  # every rank ends with the same, replicated answer
        assert np.array_equal(r0[k], r1[k]), k

    n, dim, nq, top_k = 3001, 128, 12, 10
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=21)
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=21, n_corpus=n)
    ranges = oracle.calculate_int8_ranges(corpus)
    codes, i8 = oracle.quantize_ubinary(corpus), oracle.quantize_int8(corpus, ranges)
    for tag, kwargs in {"rescore": {}, "norescore": {"use_rescoring": False},
                        "minsim": {"min_similarity": 0.3}}.items():
        want = oracle.two_stage_search(queries, codes, i8, top_k, 4.0, exact=True, **kwargs)
        idx, score, count = r0[f"{tag}_0"], r0[f"{tag}_1"], r0[f"{tag}_2"]
        for qi, (w_ids, w_s) in enumerate(want):
            m = int(count[qi])
            assert m == len(w_ids), (tag, qi)
            assert idx[qi, :m].tolist() == w_ids.tolist(), (tag, qi)
            assert np.array_equal(score[qi, :m], w_s), (tag, qi)

    want_r, want_s = oracle.int8_exact_topk(oracle.quantize_int8(queries, ranges), i8, 7)
    assert np.array_equal(r0["int8_0"], want_r) and np.array_equal(r0["int8_1"], want_s)

    n_docs, v = 1203, 300
    ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=4, mean_len=25)
    glob = BM25Oracle(ptr, toks, v)
    qt = synthetic.zipf_queries(nq, 5, v, seed=4)
    for qi in range(nq):
        rows, sc = glob.search(qt[qi].tolist(), 15)
        m = int(r0["bm25_2"][qi])
        assert r0["bm25_0"][qi, :m].tolist() == rows.tolist(), qi
        assert r0["bm25_1"][qi, :m].tolist() == sc.tolist(), qi  # float64 bit-exact across shards
