"""GPU parity: R6/R7 BM25 (float64 bit-exact vs the reference goldens and the oracle) and
R10 RRF (ids and float64 scores bit-exact, first-insertion tie order)."""

import json

import numpy as np
import pytest
import torch

import oracle
from oracle.bm25 import BM25Oracle
from radiant_rag_b200 import _lib, synthetic
from radiant_rag_b200.agents import RRFAgent, rrf_fuse_device
from radiant_rag_b200.bm25_index import BM25Index, Bm25DeviceIndex
from radiant_rag_b200.base import StoredDoc
from radiant_rag_b200.config import RetrievalConfig
from radiant_rag_b200.index import _stream
from tests.gpu_util import require_gpu
from tests.test_host_logic import _replay

pytestmark = pytest.mark.gpu


def _canonical(full):
    return sorted(full, key=lambda x: (-x[1], x[2]))


def test_bm25_reference_goldens_bit_exact(golden_dir):
    require_gpu()
    data = json.loads((golden_dir / "bm25_cases.json").read_text())
    checked = 0
    for case in data["cases"]:
        if case["name"] == "zipf_after_remove":
            stale = next(c for c in data["cases"] if c["name"] == "zipf_incremental_stale_idf")
            idx = _replay(stale)
            idx.remove_document("z003")
            idx.remove_document("z100")
        else:
            idx = _replay(case)
        row_of = {d: i for i, d in enumerate(case["doc_ids"])}
        n = len(case["doc_ids"])
        for qc in case["queries"]:
            want = _canonical([(d, s, row_of[d]) for d, s in qc["full"]])
            got = idx.search(qc["tokens"], top_k=n)
            assert [d for d, _ in got] == [d for d, _, _ in want], (case["name"], qc["tokens"])
            assert [s for _, s in got] == [s for _, s, _ in want], (case["name"], qc["tokens"])  # float64 ==
            for k in (1, 3, 10):
                assert idx.search(qc["tokens"], top_k=k) == [(d, s) for d, s, _ in want[:k]]
            checked += len(want)
    assert checked > 500


def test_bm25_reference_unit_test_assertions():
    """Known answers of reference tests/test_all.py:415-483 on the GPU path."""
    require_gpu()
    idx = BM25Index()
    idx.add_document("doc1", ["python", "programming", "language"])
    idx.add_document("doc2", ["java", "programming", "language"])
    idx.add_document("doc3", ["python", "snake", "animal"])
    res = idx.search(["python"], top_k=3)
    assert {d for d, _ in res} == {"doc1", "doc3"}
    assert idx.search([], top_k=10) == []
    assert idx.search(["nonexistent", "terms"], top_k=10) == []
    assert idx.remove_document("doc1") and len(idx) == 2 and "doc1" not in idx.doc_id_set
    assert {d for d, _ in idx.search(["python"], top_k=3)} == {"doc3"}
    i2 = BM25Index()
    i2.add_document("doc1", ["common", "rare"])
    i2.add_document("doc2", ["common", "other"])
    assert i2.idf["rare"] > i2.idf["common"]
    back = BM25Index.from_dict(idx.to_dict())
    assert back.search(["snake"], 3) == idx.search(["snake"], 3)


@pytest.mark.parametrize("n_docs,v,tile,k", [(20_000, 2000, 1024, 100), (5_000, 300, 8192, 10),
                                            (33, 20, 32, 50), (70_000, 5000, 4096, 100)])
def test_bm25_zipf_vs_oracle(n_docs, v, tile, k):
    require_gpu()
    ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=n_docs, mean_len=60)
    orc = BM25Oracle(ptr, toks, v)
    bm = Bm25DeviceIndex.build(ptr, toks, v, orc.idf, orc.avgdl, orc.k1, orc.b, device=0, tile_docs=tile,
                               row_base=7_000)
    assert bm.n_postings == orc.post_row.size
    qt = synthetic.zipf_queries(48, 8, v, seed=n_docs)
    qt[0, :] = -1               # all tokens unknown
    qt[1, 3:] = -1              # short query
    qt[2, :] = qt[2, 0]         # one token repeated 8 times
    idx, score, count = bm.search_batch(qt, k)
    torch.cuda.synchronize()
    for qi in range(qt.shape[0]):
        rows, sc = orc.search(qt[qi].tolist(), k)
        m = int(count[qi])
        assert m == rows.size, qi
        assert idx[qi, :m].cpu().tolist() == (rows + 7_000).tolist(), qi
        assert score[qi, :m].cpu().tolist() == sc.tolist(), qi   # float64 bit-exact
        assert (idx[qi, m:].cpu().numpy() == -1).all()


@pytest.mark.parametrize("interleave", [True, False])
def test_bm25_long_queries_many_tiles(interleave):
    """70-token queries (three blocks of segment bounds), repeats and unknown tokens mixed in,
    many small tiles so that the cross-tile bound path is the common one, k above and below the
    number of matching documents; with and without the bank-interleaved segment order."""
    require_gpu()
    n_docs, v = 40_000, 3000
    ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=99, mean_len=40)
    orc = BM25Oracle(ptr, toks, v)
    bm = Bm25DeviceIndex.build(ptr, toks, v, orc.idf, orc.avgdl, orc.k1, orc.b, device=0, tile_docs=512,
                               bank_interleave=interleave)
    rng = np.random.default_rng(5)
    qt = synthetic.zipf_queries(24, 70, v, seed=99)
    qt[:, 5::11] = -1                         # unknown tokens sprinkled in
    qt[3, 20:50] = qt[3, 0]                   # one token repeated 30 times
    qt[4, :] = np.int32(v - 1)                # a rare token only
    qt[5, 1:] = -1                            # single known token
    for k in (5, 300):
        idx, score, count = bm.search_batch(qt, k)
        torch.cuda.synchronize()
        for qi in range(qt.shape[0]):
            rows, sc = orc.search(qt[qi].tolist(), k)
            m = int(count[qi])
            assert m == rows.size, (k, qi)
            assert idx[qi, :m].cpu().tolist() == rows.tolist(), (k, qi)
            assert score[qi, :m].cpu().tolist() == sc.tolist(), (k, qi)


@pytest.mark.parametrize("n_docs,v,tile,k,qlen,mean_len", [
    (150_000, 5000, 1024, 100, 8, 60),   # the production shape of the batched path
    (40_000, 3000, 256, 5, 60, 40),      # 60-token queries (two token chunks), small k
    (40_000, 3000, 128, 300, 40, 40),    # smallest tile, k above many queries' match count
    (9_000, 300, 512, 50, 8, 30),        # tiny vocabulary: most terms are dense head columns
    (3_000, 200, 1024, 20, 6, 30),       # three tiles: the sample pass visits every tile
    (2_000, 200, 1024, 20, 6, 30),       # fewer documents than the candidate list holds: no sample pass
])
def test_bm25_fast_path_vs_oracle(n_docs, v, tile, k, qlen, mean_len):
    """The batched filter-and-refine path (float32 filter against a sampled bound, float64 refine in
    the reference's operation order): ids and float64 scores == oracle, checked and unchecked."""
    require_gpu()
    ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=n_docs + tile, mean_len=mean_len)
    orc = BM25Oracle(ptr, toks, v)
    bm = Bm25DeviceIndex.build(ptr, toks, v, orc.idf, orc.avgdl, orc.k1, orc.b, device=0, tile_docs=tile,
                               row_base=11_000)
    bm.fast_min_docs = 0
    assert bm.fast_ok and bm.n_head > 0 and bm.uses_fast_path(1, k)
    qt = synthetic.zipf_queries(80, qlen, v, seed=n_docs)
    qt[0, :] = -1               # all tokens unknown
    qt[1, 3:] = -1              # short query
    qt[2, :] = qt[2, 0]         # one token repeated
    qt[3, :] = np.int32(v - 1)  # a rare token only
    qt[4, :] = 0                # the most frequent term only (a dense column, massive near-ties)
    qt[5, 1:] = -1
    qt[6, ::2] = -1             # unknown tokens interleaved
    qt[7, :] = np.arange(qlen, dtype=np.int32) % min(v, 8)  # head terms only
    want = [orc.search(qt[qi].tolist(), k) for qi in range(qt.shape[0])]

    def check(idx, score, count, skip=()):
        for qi in range(qt.shape[0]):
            if qi in skip:
                continue
            rows, sc = want[qi]
            m = int(count[qi])
            assert m == rows.size, qi
            assert idx[qi, :m].cpu().tolist() == (rows + 11_000).tolist(), qi
            assert score[qi, :m].cpu().tolist() == sc.tolist(), qi   # float64 bit-exact
            assert (idx[qi, m:].cpu().numpy() == -1).all()

    idx, score, count = bm.search_batch(qt, k)
    torch.cuda.synchronize()
    check(idx, score, count)
    assert bm.last_flagged <= 12, bm.last_flagged  # the fast path answered most queries itself
    # unchecked call: every query that is NOT flagged is exact; the counter equals the flags
    qd = torch.from_numpy(qt).cuda()
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    u_idx, u_score, u_count, flags = bm._search_fast(qd, k, counter)
    torch.cuda.synchronize()
    flagged = set(np.nonzero(flags.cpu().numpy())[0].tolist())
    assert int(counter.item()) == len(flagged)
    check(u_idx, u_score, u_count, skip=flagged)
    bm.inexact_reset()
    bm.search_batch(qt, k, check=False)
    assert bm.inexact_total() == len(flagged)
    # a single query (the reference's call shape) takes the same path
    i1, s1, c1 = bm.search_batch(qt[9:10], k)
    rows, sc = want[9]
    assert i1[0, : int(c1[0])].cpu().tolist() == (rows + 11_000).tolist() and s1[0, : int(c1[0])].cpu().tolist() == sc.tolist()


def test_bm25_fast_path_degenerate_ties_fall_back_to_exact_kernel():
    """Every document scores the same: the sampled bound equals the top score, the candidate list
    overflows, the query is flagged and the checked call answers it with the exact kernel."""
    require_gpu()
    n_docs = 70_000
    ptr = np.arange(n_docs + 1, dtype=np.int64) * 2
    toks = np.tile(np.array([0, 1], dtype=np.int32), n_docs)
    toks[2 * 123 + 1] = 2  # one document differs
    orc = BM25Oracle(ptr, toks, 3)
    bm = Bm25DeviceIndex.build(ptr, toks, 3, orc.idf, orc.avgdl, orc.k1, orc.b, device=0)
    assert bm.uses_fast_path(3, 10)
    qt = np.array([[0, -1], [1, 0], [2, 1]], dtype=np.int32)
    idx, score, count = bm.search_batch(qt, 10)
    assert bm.last_flagged >= 2
    for qi in range(3):
        rows, sc = orc.search(qt[qi].tolist(), 10)
        assert idx[qi, : int(count[qi])].cpu().tolist() == rows.tolist(), qi
        assert score[qi, : int(count[qi])].cpu().tolist() == sc.tolist(), qi


def test_bm25_fast_path_disabled_for_out_of_range_impacts():
    """Impacts outside the float32-safe range (or non-positive) switch the batched path off."""
    require_gpu()
    ptr, toks = synthetic.zipf_corpus(70_000, 500, seed=4, mean_len=20)
    orc = BM25Oracle(ptr, toks, 500)
    idf = orc.idf.copy()
    idf[7] = 1e-40
    bm = Bm25DeviceIndex.build(ptr, toks, 500, idf, orc.avgdl, orc.k1, orc.b, device=0)
    assert not bm.fast_ok and not bm.uses_fast_path(8, 10)
    orc.idf = idf
    qt = synthetic.zipf_queries(8, 6, 500, seed=4)
    qt[0, 0] = 7
    idx, score, count = bm.search_batch(qt, 10)
    for qi in range(8):
        rows, sc = orc.search(qt[qi].tolist(), 10)
        assert idx[qi, : int(count[qi])].cpu().tolist() == rows.tolist() and score[qi, : int(count[qi])].cpu().tolist() == sc.tolist()


def test_bm25_impacts_kernel_bit_exact():
    require_gpu()
    rng = np.random.default_rng(3)
    n = 100_000
    tf = rng.integers(1, 30, n).astype(np.int32)
    ln = rng.integers(1, 2000, n).astype(np.int32)
    idf = rng.random(n) * 12 + 1e-3
    k1, b, avgdl = 1.5, 0.75, 203.1234567
    want = idf * ((tf.astype(np.float64) * (k1 + 1)) / (tf + k1 * ((1 - b) + (b * ln.astype(np.float64)) / avgdl)))
    t = [torch.from_numpy(a).cuda() for a in (tf, ln, idf)]
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.call("rr_bm25_impacts", t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), n, k1, b, avgdl,
              out.data_ptr(), _stream())
    assert np.array_equal(out.cpu().numpy(), want)


def test_rrf_reference_goldens_bit_exact(golden_dir):
    require_gpu()
    cases = json.loads((golden_dir / "rrf_cases.json").read_text())["cases"]
    for c in cases:
        agent = RRFAgent(RetrievalConfig(fused_top_k=c["cfg_top"], rrf_k=c["cfg_rrf"]))
        runs = [[(StoredDoc(f"d{i}", "", {}), 1.0) for i in run] for run in c["runs"]]
        kwargs = {"runs": runs}
        if c["top_k"] is not None:
            kwargs["top_k"] = c["top_k"]
        if c["rrf_k"] is not None:
            kwargs["rrf_k"] = c["rrf_k"]
        res = agent.run(**kwargs)
        assert res.success and res.status.value == "success", res.error
        assert [int(d.doc_id[1:]) for d, _ in res.data] == c["ids"], c
        assert [s for _, s in res.data] == c["scores"], c  # float64 bit-exact


def test_rrf_batched_device_api_vs_oracle():
    require_gpu()
    rng = np.random.default_rng(8)
    nq, l0, l1, l2 = 300, 100, 100, 37
    runs = np.full((nq, l0 + l1 + l2), -1, dtype=np.int64)
    lists = []
    for qi in range(nq):
        a = rng.permutation(5000)[: rng.integers(0, l0 + 1)]
        b = rng.permutation(5000)[: rng.integers(0, l1 + 1)]
        c = rng.integers(0, 50, size=rng.integers(0, l2 + 1))  # duplicates inside a run
        runs[qi, : a.size] = a
        runs[qi, l0: l0 + b.size] = b
        runs[qi, l0 + l1: l0 + l1 + c.size] = c
        lists.append([a.tolist(), b.tolist(), c.tolist()])
    for k, rk in [(10, 60), (100, 1), (1, 60.5)]:
        idx, score, count = rrf_fuse_device(torch.from_numpy(runs).cuda(), [0, l0, l0 + l1, l0 + l1 + l2], k, rk)
        torch.cuda.synchronize()
        for qi in range(nq):
            ids, sc = oracle.rrf_fuse(lists[qi], k, rk)
            m = int(count[qi])
            assert m == ids.size, (qi, k)
            assert idx[qi, :m].cpu().tolist() == ids.tolist(), (qi, k)
            assert score[qi, :m].cpu().tolist() == sc.tolist(), (qi, k)


def test_rrf_error_behaviour():
    """Exceptions become an empty list with PARTIAL status (reference fusion.py:104-114)."""
    require_gpu()
    agent = RRFAgent(RetrievalConfig())
    res = agent.run(runs=[[("not a doc", 1.0)]])
    assert res.data == [] and res.status.value == "partial" and res.success
    assert agent.run(runs=[[], []]).data == []
    assert RRFAgent(RetrievalConfig(), enabled=False).run(runs=[]).status.value == "skipped"


def test_bm25_row_sharded_one_collective_merge_emulated():
    """SURVEY.md 8e for BM25 on one device: three shards built with the GLOBAL idf / avgdl, each
    writes (score bits, global row) into one [2, Q, k] buffer, the buffers are stacked exactly as
    ONE all_gather leaves them and merged in place - must equal the single-index oracle."""
    require_gpu()
    from radiant_rag_b200.index import DenseIndex
    from radiant_rag_b200.sharded import GpuShardOps, shard_range

    n_docs, v, k, g = 90_000, 4000, 50, 3
    ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=21, mean_len=50)
    orc = BM25Oracle(ptr, toks, v)
    qt = synthetic.zipf_queries(40, 8, v, seed=21)
    qt[0, :] = -1
    words = []
    for r in range(g):
        lo, hi = shard_range(n_docs, r, g)
        sub_ptr = ptr[lo: hi + 1] - ptr[lo]
        sub_toks = toks[ptr[lo]: ptr[hi]]
        sh = Bm25DeviceIndex.build(sub_ptr, sub_toks, v, orc.idf, orc.avgdl, orc.k1, orc.b, device=0, row_base=lo)
        sh.fast_min_docs = 0
        words.append(sh.search_batch_into(qt, k))
    ops = GpuShardOps(DenseIndex(32, device=0, store_int8=False, store_f32=False))
    idx, score, count = ops.merge_scores_f64_gathered(torch.stack(words).contiguous(), k)
    torch.cuda.synchronize()
    for qi in range(qt.shape[0]):
        rows, sc = orc.search(qt[qi].tolist(), k)
        m = int(count[qi])
        assert idx[qi, :m].cpu().tolist() == rows.tolist(), qi
        assert score[qi, :m].cpu().tolist() == sc.tolist(), qi


@pytest.mark.parametrize("g,q,k_in,k", [(8, 48, 100, 100), (2, 9, 100, 100), (4, 5, 1000, 1000), (3, 7, 40, 25),
                                        (16, 3, 900, 900)])   # too large for the merge tree: the select-based kernel
def test_merge_of_gathered_sorted_bm25_lists(g, q, k_in, k):
    """rr_merge_scores_f64_gathered over sorted shard lists (exact score ties across shards, short lists,
    padding) == sorting the union by (score desc, row asc)."""
    require_gpu()
    from radiant_rag_b200.index import DenseIndex
    from radiant_rag_b200.sharded import GpuShardOps

    rng = np.random.default_rng(g * 100 + k_in)
    words = np.zeros((g, 2, q, k_in), np.int64)
    words[:, 1] = -1
    pool = rng.random(64) * 20.0 + 0.5           # few distinct scores: ties within and across shards
    for s_ in range(g):
        for r in range(q):
            m = int(rng.integers(0, k_in + 1)) if (s_ + r) % 3 == 0 else k_in
            sc = rng.choice(pool, size=m)
            rows = rng.choice(1 << 20, size=m, replace=False).astype(np.int64) + (s_ << 22)
            order = np.lexsort((rows, -sc))
            words[s_, 0, r, :m] = sc[order].view(np.int64)
            words[s_, 1, r, :m] = rows[order]
    ops = GpuShardOps(DenseIndex(32, device=0, store_int8=False, store_f32=False))
    idx, score, count = ops.merge_scores_f64_gathered(torch.from_numpy(words).cuda(), k)
    torch.cuda.synchronize()
    for r in range(q):
        rows = words[:, 1, r, :].reshape(-1)
        sc = words[:, 0, r, :].reshape(-1).view(np.float64)
        keep = rows >= 0
        rows, sc = rows[keep], sc[keep]
        order = np.lexsort((rows, -sc))[:k]
        m = int(count[r])
        assert m == order.size, r
        assert idx[r, :m].cpu().tolist() == rows[order].tolist(), r
        assert score[r, :m].cpu().numpy().tolist() == sc[order].tolist(), r
        assert (idx[r, m:].cpu().numpy() == -1).all()


def test_bm25_incremental_adds_refresh_the_device_index_instead_of_rebuilding():
    """``add_document`` is O(document) in the reference; here the device index of the
    documents already indexed keeps its sorted postings - its impacts are re-evaluated in place from
    the host's (stale-idf) tables and the new documents go to a small delta index.  Every search in
    between equals the oracle fed with the SAME host tables, float64 bit for bit."""
    require_gpu()
    v = 400
    ptr, toks = synthetic.zipf_corpus(2400, v, seed=17, mean_len=30)
    docs = [[f"t{t:04d}" for t in toks[ptr[d]: ptr[d + 1]]] for d in range(2400)]
    idx = BM25Index(tile_docs=256)
    for d in range(2000):
        idx.add_document(f"d{d}", docs[d])
    builds = {"n": 0, "rows": []}
    real_build = Bm25DeviceIndex.build.__func__

    def counting_build(cls, doc_ptr, *args, **kwargs):
        builds["n"] += 1
        builds["rows"].append(len(doc_ptr) - 1)
        return real_build(cls, doc_ptr, *args, **kwargs)

    Bm25DeviceIndex.build = classmethod(counting_build)
    try:
        queries = [[f"t{t:04d}" for t in q] for q in synthetic.zipf_queries(6, 5, v, seed=17).tolist()] + [["t0000", "nope"]]

        def check(n_docs):
            if idx.needs_rebuild:  # a fresh index rebuilds its tables at the first search, as the reference does
                idx.search(queries[0], 1)   # (bm25_index.py:229-230); afterwards adds are incremental (stale idf)
            term_ids = np.concatenate([idx._doc_term_ids[d] for d in range(n_docs)]).astype(np.int64)
            lens = np.array([idx._doc_term_ids[d].size for d in range(n_docs)])
            p = np.concatenate([[0], np.cumsum(lens)])
            vv = len(idx._vocab)
            idf = idx._idf_table(vv)
            orc = BM25Oracle(p, term_ids, vv, idx.k1, idx.b, idf=idf, avgdl=idx.avgdl)
            orc.known = idf > 0
            for q in queries:
                qt = idx._query_term_ids(q)
                rows, sc = orc.search(qt.tolist(), 25)
                got = idx.search(q, 25)
                assert [d for d, _ in got] == [f"d{r}" for r in rows.tolist()], (n_docs, q)
                assert [s for _, s in got] == sc.tolist(), (n_docs, q)

        check(2000)
        assert builds["n"] == 1 and builds["rows"] == [2000]
        assert not idx.needs_rebuild
        for d in range(2000, 2100):
            idx.add_document(f"d{d}", docs[d])
            if d % 33 == 0:
                check(d + 1)
        check(2100)
        assert builds["rows"].count(2000) == 1 and max(r for r in builds["rows"][1:]) <= 100  # only delta builds since
        for d in range(2100, 2400):  # beyond the delta budget (1024 / an eighth): one full rebuild, then deltas again
            idx.add_document(f"d{d}", docs[d])
        idx.delta_rebuild_fraction = 0.0
        idx_delta_budget = idx._gpu_docs
        check(2400)
        assert idx._gpu_docs == 2400 or idx_delta_budget == 2000
        assert idx.remove_document("d5")
        assert "d5" not in [d for d, _ in idx.search(queries[0], 2399)]
    finally:
        Bm25DeviceIndex.build = classmethod(real_build)
