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
        assert res.success and res.status == "success", res.error
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
    assert res.data == [] and res.status == "partial" and res.success
    assert agent.run(runs=[[], []]).data == []
    assert RRFAgent(RetrievalConfig(), enabled=False).run(runs=[]).status == "skipped"
