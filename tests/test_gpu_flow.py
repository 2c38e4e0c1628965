"""GPU parity through the drop-in boundary: B200VectorStore / PersistentBM25Index / agents
against goldens produced by the reference's RedisVectorStore and agents, and the
row-sharded path emulated shard by shard on one GPU."""

import json
import os
import tempfile
import threading

import numpy as np
import pytest
import torch

import oracle
from oracle.bm25 import BM25Oracle
from radiant_rag_b200 import synthetic
from radiant_rag_b200.agents import BM25RetrievalAgent, DenseRetrievalAgent, RRFAgent
from radiant_rag_b200.bm25_index import Bm25DeviceIndex, PersistentBM25Index
from radiant_rag_b200.config import BM25Config, QuantizationConfig, RetrievalConfig
from radiant_rag_b200.sharded import GpuShardOps, shard_range
from radiant_rag_b200.vector_store import B200VectorStore
from tests.gpu_util import (ABS_FLOOR, REL, assert_lists_match_tie_aware, build_index, require_gpu)
from tests.helpers import CORPUS_TEXTS, QUERY_TEXTS, HashEmbedder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg1_store(golden_dir):
    require_gpu()
    z = np.load(golden_dir / "redis_flow.npz")
    corpus = synthetic.normal_unit_rows(10_000, 384, seed=0)
    queries = synthetic.normal_unit_rows(64, 384, seed=1000)
    ranges = oracle.calculate_int8_ranges(corpus)
    store = B200VectorStore(quantization=QuantizationConfig(enabled=True, precision="both", rescore_multiplier=4.0),
                            int8_ranges=ranges)
    docs = [{"doc_id": f"{r:06d}", "content": f"doc {r}", "embedding": corpus[r],
             "meta": {"doc_level": "parent" if z["levels_parent"][r] else "child"}} for r in range(10_000)]
    assert store.upsert_batch(docs[:6000]) == 6000
    for d in docs[6000:6010]:
        store.upsert(d["doc_id"], d["content"], d["embedding"].tolist(), d["meta"])  # single-doc path
    store.upsert_batch(docs[6010:])
    return store, z, corpus, queries, ranges


def test_store_bookkeeping(cfg1_store):
    store = cfg1_store[0]
    assert store.ping() and store.count_documents() == 10_000
    assert store.has_embedding("000123") and not store.has_embedding("nope")
    assert store.get_doc("000123").content == "doc 123" and store.get_doc("nope") is None
    assert store.list_doc_ids_with_embeddings(limit=3) == ["000000", "000001", "000002"]
    info = store.get_index_info()
    assert info["num_docs"] == 10_000 and info["embedding_dim"] == 384 and info["has_int8"]
    assert store.make_doc_id("a", {"x": 1}) == store._default_make_doc_id("a", {"x": 1})
    with pytest.raises(ValueError):
        store.retrieve_by_embedding([0.0] * 10, 5)


def test_retrieve_by_embedding_matches_reference_linear_scan(cfg1_store):
    store, z, corpus, queries, _ = cfg1_store
    for tag, kwargs in {"all": {}, "child": {"doc_level_filter": "child"},
                        "parent_min": {"doc_level_filter": "parents", "min_similarity": 0.12}}.items():
        ref_ids, ref_s = z[f"linear_{tag}_ids"], z[f"linear_{tag}_scores"]
        for qi in range(16):
            res = store.retrieve_by_embedding(queries[qi].tolist(), 10, **kwargs)
            m = int((ref_ids[qi] >= 0).sum())
            assert_lists_match_tie_aware([int(d.doc_id) for d, _ in res], [s for _, s in res],
                                         ref_ids[qi, :m].tolist(), ref_s[qi, :m].tolist(), ctx=(tag, qi))
            assert all(isinstance(s, float) for _, s in res)


def test_retrieve_quantized_matches_reference_flow(cfg1_store):
    store, z, corpus, queries, ranges = cfg1_store
    no_int8 = set(z["no_int8_rows"].tolist())
    n_golden = 0
    for tag, kwargs in {"all": {}, "child": {"doc_level_filter": "leaves"}, "all_min": {"min_similarity": 0.25}}.items():
        cand = z[f"flow_{tag}_cand"]
        batch = store.retrieve_batch_quantized(queries, 10, **kwargs)
        for qi in range(64):
            res = store.retrieve_by_embedding_quantized(queries[qi].tolist(), 10, **kwargs)
            assert [(d.doc_id, s) for d, s in res] == [(d.doc_id, s) for d, s in batch[qi]]
            if no_int8 & set(cand[qi].tolist()):
                continue  # the golden used a float32 fallback row this store keeps as int8
            m = int(z[f"flow_{tag}_count"][qi])
            assert_lists_match_tie_aware([int(d.doc_id) for d, _ in res], [s for _, s in res],
                                         z[f"flow_{tag}_ids"][qi, :m].tolist(),
                                         z[f"flow_{tag}_scores"][qi, :m].tolist(), floor=1e-4, ctx=(tag, qi))
            n_golden += 1
    assert n_golden > 150
    # use_rescoring=False: stage-1 order, placeholder score 1.0 (reference chroma_store.py:624-631)
    res = store.retrieve_by_embedding_quantized(queries[0].tolist(), 10, use_rescoring=False)
    assert [int(d.doc_id) for d, _ in res] == z["flow_all_cand"][0, :10].tolist()
    assert all(s == 1.0 for _, s in res)
    # rescore_multiplier override
    res2 = store.retrieve_by_embedding_quantized(queries[0].tolist(), 10, rescore_multiplier=1.0)
    assert set(int(d.doc_id) for d, _ in res2) <= set(z["flow_all_cand"][0, :10].tolist()) and len(res2) >= 1
    assert [s for _, s in res2] == sorted((s for _, s in res2), reverse=True)


def test_quantization_disabled_falls_back_to_float_path(cfg1_store):
    _, z, corpus, queries, _ = cfg1_store
    store = B200VectorStore(quantization=QuantizationConfig(enabled=False))
    store.upsert_batch([{"doc_id": f"{r:06d}", "content": "x", "embedding": corpus[r], "meta": {}} for r in range(500)])
    a = store.retrieve_by_embedding_quantized(queries[0].tolist(), 5)
    b = store.retrieve_by_embedding(queries[0].tolist(), 5)
    assert [(d.doc_id, s) for d, s in a] == [(d.doc_id, s) for d, s in b]


def test_upsert_update_and_delete():
    require_gpu()
    corpus = synthetic.normal_unit_rows(300, 64, seed=5)
    store = B200VectorStore(quantization=QuantizationConfig(enabled=True, precision="binary"))
    store.upsert_batch([{"doc_id": f"d{r}", "content": f"c{r}", "embedding": corpus[r], "meta": {}} for r in range(300)])
    q = corpus[17]
    assert store.retrieve_by_embedding(q.tolist(), 1)[0][0].doc_id == "d17"
    store.upsert("d17", "moved", (-corpus[17]).tolist(), {"language_code": "de"})   # overwrite in place
    assert store.retrieve_by_embedding(q.tolist(), 1)[0][0].doc_id != "d17"
    assert store.retrieve_by_embedding((-q).tolist(), 1, language_filter="de")[0][0].content == "moved"
    assert store.retrieve_by_embedding((-q).tolist(), 1, language_filter="fr") == []
    assert store.delete_doc("d17") and not store.delete_doc("d17")
    assert all(d.doc_id != "d17" for d, _ in store.retrieve_by_embedding((-q).tolist(), 300, min_similarity=-1.0))
    assert len(store.retrieve_by_embedding((-q).tolist(), 300, min_similarity=-1.0)) == 299
    res = store.retrieve_by_embedding_quantized(corpus[299].tolist(), 3)   # row 299 was swapped into slot 17
    assert res[0][0].doc_id == "d299"
    store.upsert_doc_only("p1", "parent text", {"k": 1})
    assert store.get_doc("p1").meta["k"] == 1 and not store.has_embedding("p1")
    assert store.drop_index(delete_documents=True) and store.retrieve_by_embedding(q.tolist(), 3) == []


def test_agent_chain_matches_reference(golden_dir):
    """DenseRetrievalAgent -> BM25RetrievalAgent -> RRFAgent: the reference's agents over an
    in-memory store (tests/golden/agent_chain.json) vs this package's agents on the GPU."""
    require_gpu()
    cases = json.loads((golden_dir / "agent_chain.json").read_text())["cases"]
    emb = HashEmbedder(64)
    store = B200VectorStore()
    for i, text in enumerate(CORPUS_TEXTS):
        store.upsert(f"doc{i:03d}", text, emb.embed_single(text), {"doc_level": "parent" if i % 7 == 3 else "child"})
    bm = PersistentBM25Index(BM25Config(index_path=os.path.join(tempfile.mkdtemp(), "bm25")), store)
    assert bm.build_from_store() == len(CORPUS_TEXTS)
    rcfg = RetrievalConfig(dense_top_k=8, bm25_top_k=8, fused_top_k=6, rrf_k=60)
    dense, sparse, rrf = DenseRetrievalAgent(store, emb, rcfg), BM25RetrievalAgent(bm, rcfg), RRFAgent(rcfg)
    assert dense._get_doc_level_filter() == "child" and dense._get_doc_level_filter("all") is None
    batch_d = dense.execute_batch([c["query"] for c in cases])
    batch_s = sparse.execute_batch([c["query"] for c in cases])
    for ci, c in enumerate(cases):
        # the reference runs dense and BM25 from two threads (orchestrator.py:994-998)
        out = {}
        t1 = threading.Thread(target=lambda: out.__setitem__("d", dense.run(query=c["query"])))
        t2 = threading.Thread(target=lambda: out.__setitem__("s", sparse.run(query=c["query"])))
        t1.start(); t2.start(); t1.join(); t2.join()
        d, s = out["d"], out["s"]
        assert d.success and s.success, (d.error, s.error)
        assert_lists_match_tie_aware([x.doc_id for x, _ in d.data], [v for _, v in d.data],
                                     [x for x, _ in c["dense"]], [v for _, v in c["dense"]], ctx=c["query"])
        # BM25: the reference's scores bit for bit; its tie order is unspecified (argpartition +
        # unstable argsort), so its full result list is put in canonical (score desc, row asc) order
        want_s = [tuple(x) for x in sorted(c["bm25_full"], key=lambda x: (-x[1], x[0]))[:8]]
        assert sorted(v for _, v in c["bm25"]) == sorted(v for _, v in want_s)  # same score multiset
        assert [(x.doc_id, v) for x, v in s.data] == want_s, c["query"]  # float64 ==
        assert [(x.doc_id, v) for x, v in batch_s[ci]] == want_s
        assert [x.doc_id for x, _ in batch_d[ci]] == [x.doc_id for x, _ in d.data]
        f = rrf.run(runs=[d.data, s.data])
        if [x.doc_id for x, _ in d.data] == [x for x, _ in c["dense"]] and want_s == [tuple(x) for x in c["bm25"]]:
            assert [(x.doc_id, v) for x, v in f.data] == [tuple(x) for x in c["fused"]], c["query"]
        else:  # same runs in a different tie order: check against the oracle on OUR runs
            ids = {x: i for i, x in enumerate(dict.fromkeys([x.doc_id for x, _ in d.data] + [x.doc_id for x, _ in s.data]))}
            o_ids, o_sc = oracle.rrf_fuse([[ids[x.doc_id] for x, _ in d.data], [ids[x.doc_id] for x, _ in s.data]], 6, 60)
            assert [ids[x.doc_id] for x, _ in f.data] == o_ids.tolist() and [v for _, v in f.data] == o_sc.tolist()
    # persistence round trip keeps results (reference tests/test_all.py:619-647)
    assert bm.save()
    again = PersistentBM25Index(bm._config, store)
    assert [(x.doc_id, v) for x, v in again.search(cases[0]["query"], 8)] == \
        [tuple(x) for x in sorted(cases[0]["bm25_full"], key=lambda x: (-x[1], x[0]))[:8]]
    stats = again.get_stats()
    assert stats["document_count"] == len(CORPUS_TEXTS) and stats["storage_format"] == "json.gz"
    # a store without the document drops the hit silently (bm25_index.py:566-570)
    store.delete_doc(cases[0]["bm25"][0][0])
    assert cases[0]["bm25"][0][0] not in [x.doc_id for x, _ in bm.search(cases[0]["query"], 8)]


def test_dense_agent_error_becomes_empty_list():
    require_gpu()

    class Broken:
        def embed_single(self, text):
            raise RuntimeError("embedder down")

    agent = DenseRetrievalAgent(B200VectorStore(), Broken(), RetrievalConfig())
    res = agent.run(query="x")
    assert res.data == [] and res.status.value == "partial"
    with pytest.raises(ValueError):
        DenseRetrievalAgent(None, Broken(), RetrievalConfig())


@pytest.mark.parametrize("g,q,k_in,k", [(8, 64, 400, 400),    # config 3 on 8 GPUs: 8 x 512 tree
                                        (2, 33, 400, 400), (3, 17, 40, 40), (4, 9, 1000, 1000),
                                        (5, 7, 100, 60),       # fewer wanted than a shard holds
                                        (1, 5, 50, 50),
                                        (16, 4, 700, 700)])    # too large for the tree: the select-based kernel
def test_merge_of_gathered_sorted_hamming_lists(g, q, k_in, k):
    """rr_merge_hamming_gathered over sorted shard lists with ties, short lists and padding == sorting the union."""
    require_gpu()
    from tests.cpu_ops import CpuShardOps
    rng = np.random.default_rng(g * 1000 + k_in)
    keys = np.full((g, q, k_in), -1, np.int64)
    for s_ in range(g):
        for r in range(q):
            m = int(rng.integers(0, k_in + 1)) if (s_ + r) % 3 == 0 else k_in   # some lists are short or empty
            dist = np.sort(rng.integers(200, 230, size=m))                       # few distinct distances: many ties
            rows = rng.choice(1 << 22, size=m, replace=False).astype(np.int64) + (s_ << 24)
            kk = np.sort((dist.astype(np.int64) << 40) | rows)
            keys[s_, r, :m] = kk
    index = DenseIndexForOps()
    got_d, got_i = GpuShardOps(index).merge_hamming_gathered(torch.from_numpy(keys).cuda(), k)
    want_d, want_i = CpuShardOps.merge_hamming_gathered(None, torch.from_numpy(keys), k)
    assert torch.equal(got_i.cpu(), want_i) and torch.equal(got_d.cpu(), want_d)


def DenseIndexForOps():
    from radiant_rag_b200.index import DenseIndex
    return DenseIndex(128, device=0, store_int8=False, store_f32=False)


def test_row_sharded_emulated_on_one_gpu():
    """SURVEY.md 8e on one device: G shards processed one after the other, their lists
    concatenated exactly as the all_gather lays them out, merged by the product kernels.
    Must equal the single-index answer (ids, order, scores)."""
    require_gpu()
    n, dim, nq, top_k, g = 50_001, 768, 40, 10, 3
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=31)
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=31, n_corpus=n)
    single, ranges = build_index(corpus, int8=True, f32=False)
    s_idx, s_score, s_count = single.search_quantized(queries, top_k, 4.0)
    shards = []
    for r in range(g):
        lo, hi = shard_range(n, r, g)
        from radiant_rag_b200.index import DenseIndex
        sh = DenseIndex(dim, device=0, store_int8=True, store_f32=False, int8_ranges=ranges, row_base=lo)
        sh.add(corpus[lo:hi])
        shards.append(GpuShardOps(sh))
    qf, qc = shards[0].quantize_queries(queries)
    lists = [s.hamming_topk(qc, 40) for s in shards]
    d_all = torch.stack([l[0] for l in lists]).permute(1, 0, 2).reshape(nq, g * 40).contiguous()
    i_all = torch.stack([l[1] for l in lists]).permute(1, 0, 2).reshape(nq, g * 40).contiguous()
    _d, cand = shards[0].merge_hamming(d_all, i_all, 40)
    _sd, s_cand = single.hamming_topk(qc, 40)
    assert torch.equal(cand, s_cand)
    # one-collective form: packed keys in the gathered [G, Q, k'] layout, merged without a transpose
    keys_all = torch.stack([s.pack_hamming(l[0], l[1]) for s, l in zip(shards, lists)]).contiguous()
    pd, pc = shards[0].merge_hamming_gathered(keys_all, 40)
    assert torch.equal(pc, cand) and torch.equal(pd, _d)
    scores = torch.stack([s.score_candidates(qf, cand) for s in shards]).max(dim=0).values  # all_reduce(MAX)
    idx, score, count = shards[0].rank_scored(scores, cand, top_k, 0.0)
    assert torch.equal(idx, s_idx) and torch.equal(score, s_score) and torch.equal(count, s_count)

    # BM25: shard-local postings with global idf / avgdl, merged by rr_merge_scores_f64
    n_docs, v, k = 30_001, 1500, 100
    ptr, toks = synthetic.zipf_corpus(n_docs, v, seed=6, mean_len=50)
    orc = BM25Oracle(ptr, toks, v)
    qt = synthetic.zipf_queries(32, 8, v, seed=6)
    outs = []
    for r in range(g):
        lo, hi = shard_range(n_docs, r, g)
        sub = Bm25DeviceIndex.build(ptr[lo:hi + 1] - ptr[lo], toks[ptr[lo]:ptr[hi]], v, orc.idf, orc.avgdl,
                                    orc.k1, orc.b, device=0, tile_docs=2048, row_base=lo)
        outs.append(sub.search_batch(qt, k))
    s_all = torch.stack([o[1] for o in outs]).permute(1, 0, 2).reshape(32, g * k).contiguous()
    i_all = torch.stack([o[0] for o in outs]).permute(1, 0, 2).reshape(32, g * k).contiguous()
    m_idx, m_score, m_count = shards[0].merge_scores_f64(s_all, i_all, k)
    for qi in range(32):
        rows, sc = orc.search(qt[qi].tolist(), k)
        m = int(m_count[qi])
        assert m_idx[qi, :m].cpu().tolist() == rows.tolist() and m_score[qi, :m].cpu().tolist() == sc.tolist()


def test_graph_replay_equals_eager_search():
    """A captured search step replays to exactly the eager results, for new queries too."""
    require_gpu()
    from radiant_rag_b200.graphed import GraphedSearch
    n, dim, nq, top_k = 20_000, 768, 32, 10
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=77)
    index, _ranges = build_index(corpus, int8=True, f32=False)
    graph = GraphedSearch(lambda q: index.search_quantized(q, top_k, 4.0, check_overflow=False), nq, dim, "cuda:0")
    assert graph.kernels_per_replay > 0
    for seed in (1, 2):
        queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=seed, n_corpus=n)
        want = index.search_quantized(queries, top_k, 4.0)
        got = graph(torch.from_numpy(queries).pin_memory())
        torch.cuda.synchronize()
        for g, w in zip(got, want):
            assert torch.equal(g, w)
    assert index.tc_overflow_total() == 0


def test_wire_format_import_and_index_persistence(tmp_path):
    """rows taken byte for byte from the reference's side-key payloads
    (np.packbits bytes, int8 bytes) give the same index as quantising on device, the shard
    survives save -> load, and the store's bulk import answers like a store built by upsert."""
    require_gpu()
    from radiant_rag_b200.index import DenseIndex
    n, dim = 5_000, 384
    corpus = synthetic.normal_unit_rows(n, dim, seed=12)
    queries = synthetic.normal_unit_rows(24, dim, seed=13)
    ranges = oracle.calculate_int8_ranges(corpus)
    codes, i8 = oracle.quantize_ubinary(corpus), oracle.quantize_int8(corpus, ranges)
    a, _ = build_index(corpus, int8=True, f32=True)
    b = DenseIndex(dim, device=0, store_int8=True, store_f32=True, int8_ranges=ranges, row_base=0)
    assert b.add_quantized(codes, i8, corpus) == (0, n)
    assert torch.equal(a.codes[:n], b.codes[:n]) and torch.equal(a.int8[:n], b.int8[:n])
    want = a.search_quantized(queries, 10, 4.0)
    for got in (b.search_quantized(queries, 10, 4.0),):
        assert all(torch.equal(g, w) for g, w in zip(got, want))
    path = str(tmp_path / "shard.npz")
    b.save(path)
    c = DenseIndex.load(path, device=0)
    assert (c.n, c.dim, c.row_base, c.store_int8, c.store_f32) == (n, dim, 0, True, True)
    assert all(torch.equal(g, w) for g, w in zip(c.search_quantized(queries, 10, 4.0), want))
    assert all(torch.equal(g, w) for g, w in zip(c.search_exact(queries, 7), a.search_exact(queries, 7)))
    with pytest.raises(ValueError):
        b.add_quantized(codes[:, :-1], i8, corpus)   # wrong packed width

    qc = QuantizationConfig(enabled=True, precision="both")
    ref_store = B200VectorStore(embedding_dim=dim, quantization=qc, int8_ranges=ranges)
    ref_store.upsert_batch([{"doc_id": f"d{r}", "content": f"c{r}", "embedding": corpus[r], "meta": {}}
                            for r in range(400)])
    imp_store = B200VectorStore(embedding_dim=dim, quantization=qc, int8_ranges=ranges)
    assert imp_store.import_quantized_batch(
        [{"doc_id": f"d{r}", "content": f"c{r}", "meta": {}, "binary": codes[r].tobytes(), "int8": i8[r].tobytes(),
          "embedding": corpus[r]} for r in range(400)]) == 400
    for qi in range(4):
        got = imp_store.retrieve_by_embedding_quantized(queries[qi].tolist(), 5)
        exp = ref_store.retrieve_by_embedding_quantized(queries[qi].tolist(), 5)
        assert [(d.doc_id, s) for d, s in got] == [(d.doc_id, s) for d, s in exp]
    with pytest.raises(ValueError):
        imp_store.import_quantized_batch([{"doc_id": "d1", "content": "x", "meta": {}, "binary": codes[1].tobytes(),
                                           "int8": i8[1].tobytes(), "embedding": corpus[1]}])


def test_upsert_batch_is_all_or_nothing():
    """A bad document in the middle of a batch leaves the store exactly as it was (ids, rows,
    documents); afterwards new rows still map to the right doc_ids."""
    require_gpu()
    corpus = synthetic.normal_unit_rows(40, 64, seed=6)
    store = B200VectorStore()
    store.upsert_batch([{"doc_id": f"d{r}", "content": f"c{r}", "embedding": corpus[r], "meta": {}} for r in range(10)])
    bad = [{"doc_id": f"d{r}", "content": f"c{r}", "embedding": corpus[r], "meta": {}} for r in range(10, 20)]
    bad[6]["embedding"] = corpus[16][:32]  # wrong dimension
    with pytest.raises(ValueError):
        store.upsert_batch(bad)
    assert store.list_doc_ids_with_embeddings() == [f"d{r}" for r in range(10)]
    assert store.index.n == 10 and store.count_documents() == 10 and store.get_doc("d12") is None
    store.upsert_batch([{"doc_id": f"d{r}", "content": f"c{r}", "embedding": corpus[r], "meta": {}} for r in range(20, 30)])
    for r in (3, 25):
        assert store.retrieve_by_embedding(corpus[r].tolist(), 1)[0][0].doc_id == f"d{r}"
    # the same id twice in one batch: the last one wins, one row
    store.upsert_batch([{"doc_id": "dup", "content": "a", "embedding": corpus[30], "meta": {}},
                        {"doc_id": "dup", "content": "b", "embedding": corpus[31], "meta": {}}])
    assert store.index.n == 21 and store.get_doc("dup").content == "b"
    assert store.retrieve_by_embedding(corpus[31].tolist(), 1)[0][0].doc_id == "dup"
    with pytest.raises(ValueError):
        store.retrieve_by_embedding(corpus[0].tolist(), 5000)  # beyond the per-query result limit: said, not clamped


def test_unknown_doc_level_passes_neither_filter():
    """The reference's filters compare the stored doc_level string (redis_store.py:684, 915-916):
    a row stored with any other level is returned only without a level filter."""
    require_gpu()
    corpus = synthetic.normal_unit_rows(3, 32, seed=8)
    store = B200VectorStore()
    for r, level in enumerate(["child", "parent", "section"]):
        store.upsert(f"d{r}", "x", corpus[r].tolist(), {"doc_level": level})
    q = corpus[2].tolist()
    assert [d.doc_id for d, _ in store.retrieve_by_embedding(q, 3, min_similarity=-1.0)][0] == "d2"
    assert "d2" not in [d.doc_id for d, _ in store.retrieve_by_embedding(q, 3, min_similarity=-1.0, doc_level_filter="child")]
    assert "d2" not in [d.doc_id for d, _ in store.retrieve_by_embedding(q, 3, min_similarity=-1.0, doc_level_filter="parents")]


def test_import_pgvector_bytea_rows():
    """the reference's pgvector side tables (BYTEA packed bits / int8 rows,
    pgvector_store.py:322-355) loaded as they are; results equal an index built from the floats."""
    require_gpu()
    n, dim = 500, 128
    corpus = synthetic.normal_unit_rows(n, dim, seed=12)
    ranges = oracle.calculate_int8_ranges(corpus)
    codes, i8 = oracle.quantize_ubinary(corpus), oracle.quantize_int8(corpus, ranges)
    qcfg = QuantizationConfig(enabled=True, precision="both")
    ref = B200VectorStore(quantization=qcfg, int8_ranges=ranges)
    ref.upsert_batch([{"doc_id": f"p{r}", "content": f"c{r}", "embedding": corpus[r], "meta": {}} for r in range(n)])
    store = B200VectorStore(quantization=qcfg, int8_ranges=ranges)
    doc_rows = [(f"p{r}", f"c{r}", {"doc_level": "child"}, corpus[r].tolist()) for r in range(n)]
    doc_rows.append(("orphan", "no quantised row", {}, corpus[0].tolist()))  # batch-path document: skipped
    binary_rows = [(f"p{r}", memoryview(codes[r].tobytes())) for r in range(n)]
    int8_rows = [(f"p{r}", i8[r].tobytes()) for r in range(n)]
    assert store.import_pgvector_rows(doc_rows, binary_rows, int8_rows) == n
    assert torch.equal(store.index.codes[:n], ref.index.codes[:n]) and torch.equal(store.index.int8[:n], ref.index.int8[:n])
    for qi in (0, 77, 499):
        a = store.retrieve_by_embedding_quantized(corpus[qi].tolist(), 5)
        b = ref.retrieve_by_embedding_quantized(corpus[qi].tolist(), 5)
        assert [(d.doc_id, s) for d, s in a] == [(d.doc_id, s) for d, s in b]
    with pytest.raises(ValueError):  # a short payload is rejected before anything is stored
        store.import_quantized_batch([{"doc_id": "x", "content": "", "binary": b"\x00" * 3, "int8": i8[0].tobytes(),
                                       "embedding": corpus[0]}])
    assert store.index.n == n and not store.has_embedding("x")


def test_batched_retrieval_equals_per_query_agents():
    """hybrid.BatchedRetrieval (one call per phase for a batch of query strings) == the per-query
    agent chain, including the fall-through to the non-empty run (app.py:1234-1239) and the
    orchestrator's first-occurrence merge over sub-queries (orchestrator.py:953-965)."""
    require_gpu()
    from radiant_rag_b200.hybrid import BatchedRetrieval

    emb = HashEmbedder(64)
    store = B200VectorStore()
    for i, text in enumerate(CORPUS_TEXTS):
        store.upsert(f"doc{i:03d}", text, emb.embed_single(text), {"doc_level": "parent" if i % 7 == 3 else "child"})
    bm = PersistentBM25Index(BM25Config(index_path=os.path.join(tempfile.mkdtemp(), "bm25")), store)
    bm.build_from_store()
    rcfg = RetrievalConfig(dense_top_k=8, bm25_top_k=8, fused_top_k=6, rrf_k=60)
    br = BatchedRetrieval(store, bm, emb, rcfg)
    dense, sparse, rrf = DenseRetrievalAgent(store, emb, rcfg), BM25RetrievalAgent(bm, rcfg), RRFAgent(rcfg)
    got = br.search_batch(QUERY_TEXTS, mode="hybrid", top_k=5)
    for qi, query in enumerate(QUERY_TEXTS):
        d = dense.run(query=query, top_k=5).data
        s = sparse.run(query=query, top_k=5).data
        want = rrf.run(runs=[d, s], top_k=5).data[:5] if d and s else (d if d else s)
        assert [(x.doc_id, v) for x, v in got[qi]] == [(x.doc_id, v) for x, v in want], query
    assert [[x.doc_id for x, _ in r] for r in br.search_batch(QUERY_TEXTS[:3], mode="bm25", top_k=4)] == \
        [[x.doc_id for x, _ in sparse.run(query=q, top_k=4).data] for q in QUERY_TEXTS[:3]]
    # one user query expanded into sub-queries
    subs = QUERY_TEXTS[:4]
    d_all, s_all, fused = br.run_retrieval(subs)
    seen, want_d = set(), []
    for q in subs:
        for doc, score in dense.run(query=q).data:
            if doc.doc_id not in seen:
                seen.add(doc.doc_id)
                want_d.append((doc.doc_id, score))
    assert [(x.doc_id, v) for x, v in d_all] == want_d
    want_f = rrf.run(runs=[d_all, s_all]).data
    assert [(x.doc_id, v) for x, v in fused] == [(x.doc_id, v) for x, v in want_f]
