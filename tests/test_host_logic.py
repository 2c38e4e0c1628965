"""CPU: host-side logic of the product package (no kernels run here)."""

import json

import numpy as np
import pytest

from radiant_rag_b200 import synthetic
from radiant_rag_b200.base import BaseVectorStore, StoredDoc, normalize_doc_level
from radiant_rag_b200.bm25_index import BM25Index, _normalize_index_path, _tokenize
from radiant_rag_b200.config import BM25Config, QuantizationConfig, RetrievalConfig
from radiant_rag_b200.index import LanguageTable, code_words, make_tag, tag_predicate
from radiant_rag_b200.sharded import shard_range


def test_tokenizer_matches_reference_golden(golden_dir):
    data = json.loads((golden_dir / "bm25_cases.json").read_text())
    for c in data["tokenizer"]:
        assert _tokenize(c["text"]) == c["tokens"], c["text"]


def _replay(case):
    """Rebuild the product's host index the way the golden case was built."""
    name = case["name"]
    ids, docs = case["doc_ids"], case["doc_tokens"]
    if name in ("ref_test_rebuilt", "zipf_rebuilt"):
        return BM25Index(doc_ids=list(ids), doc_tokens=[list(t) for t in docs], k1=case["k1"], b=case["b"])
    idx = BM25Index(k1=case["k1"], b=case["b"])
    if name == "zipf_incremental_stale_idf":
        for i, t in zip(ids[:100], docs[:100]):
            idx.add_document(i, list(t))
        idx._rebuild_index()  # what the reference's first search does
        for i, t in zip(ids[100:], docs[100:]):
            idx.add_document(i, list(t))
        return idx
    for i, t in zip(ids, docs):
        idx.add_document(i, list(t))
    return idx


def test_bm25_host_tables_match_reference(golden_dir):
    """idf / avgdl / doc_lengths bookkeeping (incremental, stale, rebuilt) equals the
    reference's tables bit for bit - these are the inputs copied to the GPU (R7)."""
    data = json.loads((golden_dir / "bm25_cases.json").read_text())
    for case in data["cases"]:
        if case["name"] == "zipf_after_remove":
            continue
        idx = _replay(case)
        if idx.needs_rebuild and case["name"] != "zipf_incremental_stale_idf":
            idx._rebuild_index()
        assert idx.avgdl == case["avgdl_used"], case["name"]
        assert idx.doc_lengths == case["doc_lengths"], case["name"]
        assert {t: float(v) for t, v in idx.idf.items()} == case["idf_used"], case["name"]


def test_bm25_remove_then_rebuild(golden_dir):
    data = json.loads((golden_dir / "bm25_cases.json").read_text())
    stale = next(c for c in data["cases"] if c["name"] == "zipf_incremental_stale_idf")
    after = next(c for c in data["cases"] if c["name"] == "zipf_after_remove")
    idx = _replay(stale)
    assert idx.remove_document("z003") and idx.remove_document("z100")
    assert not idx.remove_document("nope")
    assert idx.needs_rebuild
    idx._rebuild_index()
    assert idx.doc_ids == after["doc_ids"]
    assert {t: float(v) for t, v in idx.idf.items()} == after["idf_used"]
    assert idx.avgdl == after["avgdl_used"]
    assert idx.doc_id_to_idx == {d: i for i, d in enumerate(after["doc_ids"])}


def test_bm25_add_duplicate_and_serialise_roundtrip():
    idx = BM25Index()
    assert idx.add_document("doc1", ["hello", "world"])
    assert not idx.add_document("doc1", ["other"])
    assert len(idx) == 1 and "doc1" in idx.doc_id_set
    data = idx.to_dict()
    assert data["version"] == 2 and set(data) == {"version", "doc_ids", "doc_tokens", "k1", "b"}
    back = BM25Index.from_dict(json.loads(json.dumps(data)))
    assert back.doc_ids == ["doc1"] and not back.needs_rebuild and back.idf
    empty = BM25Index.from_dict({})
    assert len(empty) == 0 and empty.search(["x"], 3) == []
    assert idx.search([], 10) == []


def test_index_path_normalisation():
    from pathlib import Path

    for name in ("bm25_index", "bm25_index.json.gz", "bm25_index.pkl", "bm25_index.json", "bm25_index.pickle"):
        assert _normalize_index_path(Path("/x") / name) == Path("/x/bm25_index")


def test_config_mirrors_defaults_and_validation():
    r = RetrievalConfig()
    assert (r.dense_top_k, r.bm25_top_k, r.fused_top_k, r.rrf_k, r.min_similarity, r.search_scope) == \
        (10, 10, 15, 60, 0.0, "leaves")
    b = BM25Config()
    assert (b.k1, b.b, b.max_documents, b.auto_save_threshold) == (1.5, 0.75, 100_000, 100)
    q = QuantizationConfig()
    assert (q.enabled, q.precision, q.rescore_multiplier, q.use_rescoring) == (False, "both", 4.0, True)
    with pytest.raises(ValueError):
        QuantizationConfig(precision="fp4")
    with pytest.raises(ValueError):
        QuantizationConfig(rescore_multiplier=0.5)


def test_stored_doc_identity():
    a = StoredDoc("id1", "x", {})
    b = StoredDoc("id1", "y", {"k": 1})
    c = StoredDoc("id2", "x", {})
    assert a == b and a != c and hash(a) == hash(b) and len({a, b, c}) == 2
    assert a != "id1"


def test_default_doc_id_is_sha256_of_content_and_sorted_meta():
    import hashlib

    class S(BaseVectorStore):
        pass

    S.__abstractmethods__ = frozenset()
    s = S()
    want = hashlib.sha256(("text\n" + json.dumps({"a": 1, "b": 2}, sort_keys=True, ensure_ascii=False)).encode()).hexdigest()
    assert s._default_make_doc_id("text", {"b": 2, "a": 1}) == want


def test_doc_level_filter_normalisation_and_tags():
    assert normalize_doc_level("leaves") == "child" and normalize_doc_level("Leaf") == "child"
    assert normalize_doc_level("parents") == "parent" and normalize_doc_level("all") is None
    assert normalize_doc_level(None) is None and normalize_doc_level("weird") is None
    langs = LanguageTable()
    en = langs.id_for("en", create=True)
    de = langs.id_for("de", create=True)
    assert (en, de) == (1, 2) and langs.id_for("en") == 1 and langs.id_for(None) == 0
    child_en = make_tag("child", en)
    parent_de = make_tag("parent", de)
    m, v = tag_predicate("child", 0)
    assert (child_en & m) == v and (parent_de & m) != v
    m, v = tag_predicate("parent", de)
    assert (parent_de & m) == v and (make_tag("parent", en) & m) != v
    m, v = tag_predicate(None, langs.id_for("fr"))  # unknown language matches nothing
    assert (child_en & m) != v and (parent_de & m) != v
    assert tag_predicate(None, 0) == (0, 0)


def test_code_words_and_shards():
    assert [code_words(d) for d in (384, 768, 1024, 100, 32, 1)] == [12, 24, 32, 4, 4, 4]
    for n, w in [(10, 3), (100, 8), (7, 8), (0, 2), (1_000_000, 8)]:
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        for (a, b), (c, d) in zip(ranges, ranges[1:]):
            assert b == c and a <= b
        assert sum(b - a for a, b in ranges) == n


def test_synthetic_is_counter_based():
    full = synthetic.hash_rows_f32(0, 64, 96, seed=5)
    part = synthetic.hash_rows_f32(17, 9, 96, seed=5)
    assert np.array_equal(full[17:26], part)
    assert not np.array_equal(full, synthetic.hash_rows_f32(0, 64, 96, seed=6))
    assert 0.6 < np.linalg.norm(full, axis=1).mean() < 1.5  # power-of-two scale: within sqrt(2) of unit norm
    q = synthetic.hash_query_rows_f32(0, 8, 96, seed=5, n_corpus=64)
    src = synthetic.query_source_row(np.arange(8), 64, 5)
    same = (q[1] == full[src[1]]).mean()
    assert 0.6 < same < 0.9  # odd queries copy ~3/4 of a corpus row
    ptr, toks = synthetic.zipf_corpus(200, 50, seed=1, mean_len=30)
    assert ptr[-1] == toks.size and toks.min() >= 0 and toks.max() < 50
    assert np.array_equal(synthetic.zipf_tokens(int(ptr[10]), int(ptr[11] - ptr[10]), 1, synthetic.zipf_cdf_u32(50)),
                          toks[ptr[10]:ptr[11]])
    counts = np.bincount(toks, minlength=50)
    assert counts[0] > counts[10] > counts[49]


def test_bm25_bank_interleave_is_a_segment_preserving_permutation():
    """Bm25DeviceIndex._bank_interleave only reorders postings INSIDE their (tile, term) segment
    and spreads consecutive postings over row % 16 (torch CPU tensors: no GPU needed)."""
    import torch
    from radiant_rag_b200.bm25_index import Bm25DeviceIndex
    rng = np.random.default_rng(0)
    tile = 512
    keys = []
    for seg in range(60):
        dens = rng.random() ** 2
        rows = np.nonzero(rng.random(tile) < dens)[0]
        keys.append(seg * tile + rows)
    ukey = torch.from_numpy(np.concatenate(keys)).long()
    tf = torch.arange(ukey.numel())
    u2, t2 = Bm25DeviceIndex._bank_interleave(ukey, tf, tile)
    assert torch.equal(torch.sort(u2)[0], ukey)          # a permutation of the postings
    assert torch.equal(ukey[t2], u2)                      # tf travels with its posting
    assert torch.equal(u2 // tile, ukey // tile)          # segments keep their place and extent

    def wavefronts(u):
        seg, row = (u // tile).numpy(), (u % tile).numpy()
        tot = cnt = 0
        for s_ in np.unique(seg):
            r = row[seg == s_]
            for i in range(0, len(r), 16):
                tot += np.bincount(r[i:i + 16] % 16, minlength=16).max()
                cnt += 1
        return tot / cnt

    assert wavefronts(u2) < 0.75 * wavefronts(ukey)       # fewer same-bank accesses per half-warp


def test_agent_results_carry_the_reference_fields():
    """status is an enum with .value, metrics and to_dict() exist (reference base_agent.py:145-184;
    radiant/orchestrator.py:86-88 records result.metrics) - checked without a GPU through the
    error path (a run that is not a list of (doc, score) pairs)."""
    from radiant_rag_b200 import agents
    from radiant_rag_b200.config import RetrievalConfig

    a = agents.RRFAgent(RetrievalConfig())
    r = a.run(runs=[[("not a doc", 1.0)]], correlation_id="abc")
    assert r.success and r.data == [] and r.status.value == "partial" and r.warnings
    assert r.metrics is not None and r.metrics.agent_name == "RRFAgent" and r.metrics.correlation_id == "abc"
    d = r.to_dict()
    assert d["status"] == "partial" and d["metrics"]["agent_category"] == "post_retrieval"
    assert agents.RRFAgent(RetrievalConfig(), enabled=False).run(runs=[]).status.value == "skipped"
    assert a.name == "RRFAgent" and a.category.value == "post_retrieval"


def test_agents_subclass_the_reference_base_agent_when_it_is_importable():
    """Inside a Radiant RAG installation the agents ARE reference BaseAgents (own subprocess: the
    choice is made at import).  Skipped where the reference checkout is absent (the GPU box)."""
    import os
    import subprocess
    import sys

    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "radiant")):
        import pytest
        pytest.skip("reference checkout not present")
    code = (
        "import logging; logging.disable(logging.CRITICAL)\n"
        "from radiant_rag_b200 import agents\n"
        "from radiant_rag_b200.config import RetrievalConfig\n"
        "from radiant.agents.base_agent import BaseAgent, AgentResult, AgentStatus\n"
        "a = agents.RRFAgent(RetrievalConfig())\n"
        "assert agents.HAVE_REFERENCE_AGENTS and isinstance(a, BaseAgent)\n"
        "r = a.run(runs=[[('x', 1.0)]])\n"
        "assert isinstance(r, AgentResult) and r.status is AgentStatus.PARTIAL and r.data == [] and r.metrics\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref, root]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_oracle_helpers_for_full_size_checks():
    """int8_exact_topk_blas == int8_exact_topk; BM25Oracle(only_terms=...) == the full oracle."""
    import oracle
    from oracle.bm25 import BM25Oracle
    from radiant_rag_b200 import synthetic

    rng = np.random.default_rng(0)
    e = rng.integers(-128, 128, (5000, 1024)).astype(np.int8)
    q = rng.integers(-128, 128, (7, 1024)).astype(np.int8)
    e[10] = e[20]
    a, b = oracle.int8_exact_topk(q, e, 10), oracle.int8_exact_topk_blas(q, e, 10, chunk=777)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ptr, toks = synthetic.zipf_corpus(3000, 500, seed=3, mean_len=40)
    full = BM25Oracle(ptr, toks, 500)
    qt = synthetic.zipf_queries(5, 6, 500, seed=3)
    part = BM25Oracle(ptr, toks, 500, only_terms=set(qt.ravel().tolist()))
    for i in range(5):
        r1, r2 = full.search(qt[i].tolist(), 20), part.search(qt[i].tolist(), 20)
        assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1])


def test_quantization_section_from_yaml(tmp_path, monkeypatch):
    """the YAML ``quantization:`` section the reference's loader drops."""
    from radiant_rag_b200.config import QuantizationConfig, quantization_from_yaml

    p = tmp_path / "config.yaml"
    p.write_text("storage:\n  backend: chroma\nredis:\n  quantization:\n    enabled: false\n"
                 "chroma:\n  quantization:\n    enabled: true\n    precision: binary\n    rescore_multiplier: 6.0\n"
                 "    use_rescoring: false\n    bogus: 1\n")
    q = quantization_from_yaml(p)
    assert q == QuantizationConfig(enabled=True, precision="binary", rescore_multiplier=6.0, use_rescoring=False)
    assert quantization_from_yaml(p, backend="redis").enabled is False
    assert quantization_from_yaml({}).enabled is False and quantization_from_yaml({}).rescore_multiplier == 4.0
    monkeypatch.setenv("RADIANT_QUANTIZATION_PRECISION", "int8")
    assert quantization_from_yaml(p).precision == "int8"
    import pytest
    with pytest.raises(ValueError):
        quantization_from_yaml({"redis": {"quantization": {"rescore_multiplier": 0.5}}})


def test_batched_rerank_and_automerge_equal_the_per_query_rule():
    """one cross-encoder call / one parent fetch for a whole batch, results equal
    to the reference's per-query rule (radiant/agents/rerank.py:64-117, automerge.py:65-135)."""
    import pytest
    torch = pytest.importorskip("torch")
    from dataclasses import dataclass

    from radiant_rag_b200.base import StoredDoc
    from radiant_rag_b200.hybrid import automerge_batch, rerank_batch

    class Encoder:
        calls = 0

        def predict(self, pairs, show_progress_bar=False):
            Encoder.calls += 1
            return [float(len(set(q.split()) & set(d.split()))) - 0.01 * len(d) for q, d in pairs]

    class Models:
        cross_encoder = Encoder()

        def rerank(self, query, documents, top_k=None):  # the reference's per-query method
            s = self.cross_encoder.predict([(query, d) for d in documents])
            idx = sorted(((i, float(x)) for i, x in enumerate(s)), key=lambda x: x[1], reverse=True)
            return idx[:top_k] if top_k else idx

    @dataclass
    class Cfg:
        top_k: int = 3
        candidate_multiplier: int = 2
        min_candidates: int = 4
        max_doc_chars: int = 12

    docs = [StoredDoc(f"d{i}", " ".join(["alpha", "beta", "gamma", "delta"][: 1 + i % 4]) + f" x{i}", {}) for i in range(9)]
    queries = ["alpha beta", "gamma", "zzz"]
    lists = [[(d, 1.0) for d in docs[:7]], [(d, 1.0) for d in docs[2:]], []]
    cfg, models = Cfg(), Models()
    Encoder.calls = 0
    got = rerank_batch(models, queries, lists, cfg)
    assert Encoder.calls == 1
    for q, lst, g in zip(queries, lists, got):
        cand = lst[: max(cfg.top_k * cfg.candidate_multiplier, cfg.min_candidates)]
        want = [(cand[i][0], s) for i, s in models.rerank(q, [d.content[: cfg.max_doc_chars] for d, _ in cand], top_k=cfg.top_k)] if cand else []
        assert [(d.doc_id, s) for d, s in g] == [(d.doc_id, s) for d, s in want]

    class Store:
        fetched = []

        def get_doc(self, doc_id):
            Store.fetched.append(doc_id)
            return {"p1": StoredDoc("p1", "short parent", {"doc_level": "parent"}),
                    "p2": StoredDoc("p2", "x" * 500, {"doc_level": "parent"})}.get(doc_id)

    ch = [StoredDoc(f"c{i}", f"child {i}", {"doc_level": "child", "parent_id": "p1" if i < 3 else ("p2" if i < 6 else "p3")})
          for i in range(8)]
    lists = [[(c, 1.0 - 0.1 * i) for i, c in enumerate(ch)], [(ch[0], 0.5), (ch[1], 0.7), (ch[6], 0.9)]]
    got = automerge_batch(Store(), lists, min_children=2, max_parent_chars=100)
    assert sorted(Store.fetched) == ["p1", "p2", "p3"]   # each parent of the batch once
    assert [d.doc_id for d, _ in got[0]] == ["p1", "c3", "c4", "c5", "c6", "c7"] and got[0][0][1] == 1.0
    assert [(d.doc_id, s) for d, s in got[1]] == [("c6", 0.9), ("p1", 0.7)]
