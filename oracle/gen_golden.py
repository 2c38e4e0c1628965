"""Generate tests/golden/* by running the REFERENCE's own Python code (test infrastructure).

Run in the build container only (needs /root/reference, never on the GPU box):

    python -m oracle.gen_golden

What is pinned by real reference output
  bm25_cases.json     BM25Index (radiant/storage/bm25_index.py): add/rebuild/incremental/remove,
                      search scores for every document (top_k = N), idf tables, tokenizer
  rrf_cases.json      RRFAgent (radiant/agents/fusion.py) through run(runs=...)
  rescore_cases.npz   rescore_candidates (radiant/storage/quantization.py:185-222)
  redis_flow.npz      RedisVectorStore (radiant/storage/redis_store.py) over oracle/_fake_redis:
                      _retrieve_by_embedding_linear (R5, with filters and thresholds) and the
                      stage-2 half of retrieve_by_embedding_quantized (int8 key load, float32
                      fallback, rescore, cut, threshold) on BASELINE config 1
  agent_chain.json    DenseRetrievalAgent -> BM25RetrievalAgent -> RRFAgent over an in-memory
                      BaseVectorStore + PersistentBM25Index.build_from_store

What is NOT pinned by the reference (stated in the files): the ubinary/int8 quantiser
(sentence-transformers, absent) and the Hamming stage 1 (no reference code) come from the
oracle's restatement; they are stored so CUDA-vs-oracle parity is checked on fixed vectors.
"""

from __future__ import annotations

import json
import logging
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT))  # ROOT first: the reference also has a top-level `tests` package

import oracle  # noqa: E402
from oracle import _fake_redis  # noqa: E402
from radiant_rag_b200 import synthetic  # noqa: E402
from tests.helpers import CORPUS_TEXTS, QUERY_TEXTS, HashEmbedder  # noqa: E402


def _install_stubs() -> None:
    """A dict-backed ``redis`` and the restated sentence-transformers quantiser, so the
    reference's RedisVectorStore and its quantised code path can execute here."""
    _fake_redis.install()
    st = types.ModuleType("sentence_transformers")
    stq = types.ModuleType("sentence_transformers.quantization")

    def quantize_embeddings(embeddings, precision="float32", ranges=None, calibration_embeddings=None):
        emb = np.asarray(embeddings, dtype=np.float32)
        if precision == "ubinary":
            return oracle.quantize_ubinary(emb)
        if precision == "int8":
            if ranges is None:
                ranges = oracle.calculate_int8_ranges(emb)
            return oracle.quantize_int8(emb, ranges)
        raise ValueError(precision)

    stq.quantize_embeddings = quantize_embeddings
    st.quantization = stq
    sys.modules["sentence_transformers"] = st
    sys.modules["sentence_transformers.quantization"] = stq


def gen_bm25() -> None:
    from radiant.storage.bm25_index import BM25Index, _tokenize

    cases = []

    def full_results(index, tokens):
        res = index.search(list(tokens), top_k=max(len(index), 1))
        return [[d, float(s)] for d, s in res]

    def snapshot(name, index, queries, extra=None):
        case = {
            "name": name,
            "k1": index.k1,
            "b": index.b,
            "doc_ids": list(index.doc_ids),
            "doc_tokens": [list(t) for t in index.doc_tokens],
            "avgdl": float(index.avgdl),
            "doc_lengths": list(index.doc_lengths),
            "idf": {t: float(v) for t, v in index.idf.items()},
            "needs_rebuild": bool(index.needs_rebuild),
            "queries": [{"tokens": list(q), "full": full_results(index, q)} for q in queries],
        }
        # full_results may have triggered a rebuild; record the tables actually used
        case["idf_used"] = {t: float(v) for t, v in index.idf.items()}
        case["avgdl_used"] = float(index.avgdl)
        if extra:
            case.update(extra)
        cases.append(case)

    # 1. the documents of reference tests/test_all.py:415-429, incremental adds
    idx = BM25Index()
    idx.add_document("doc1", ["python", "programming", "language"])
    idx.add_document("doc2", ["java", "programming", "language"])
    idx.add_document("doc3", ["python", "snake", "animal"])
    snapshot("ref_test_incremental", idx,
             [["python"], ["programming", "language"], ["python", "python", "snake"], ["nonexistent"], []])

    # 2. same documents after a serialise -> restore (idf rebuilt, tests/test_all.py:619-647)
    restored = BM25Index.from_dict(idx.to_dict())
    snapshot("ref_test_rebuilt", restored,
             [["python"], ["programming", "language"], ["python", "python", "snake"]])

    # 3. Zipf documents, rebuilt tables
    rng = np.random.default_rng(11)
    vocab = [f"t{i:03d}" for i in range(120)]
    p = 1.0 / np.arange(1, len(vocab) + 1)
    p /= p.sum()
    docs = []
    for d in range(160):
        n = int(rng.integers(3, 40))
        docs.append([vocab[i] for i in rng.choice(len(vocab), size=n, p=p)])
    ids = [f"z{d:03d}" for d in range(len(docs))]
    z = BM25Index(doc_ids=list(ids), doc_tokens=[list(t) for t in docs], k1=1.2, b=0.6)
    queries = [[vocab[i] for i in rng.choice(len(vocab), size=int(rng.integers(1, 9)), p=p)]
               for _ in range(24)]
    queries.append(["t000", "t000", "t001", "zzz", "t119"])
    snapshot("zipf_rebuilt", z, queries)

    # 4. the same documents through incremental add_document (stale idf), default k1/b
    #    A fresh BM25Index has needs_rebuild=True, so its first search rebuilds every idf;
    #    the stale-idf state needs adds AFTER a search (needs_rebuild False by then).
    inc = BM25Index()
    for i, t in zip(ids[:100], docs[:100]):
        inc.add_document(i, list(t))
    inc.search(queries[0], top_k=5)  # rebuild happens here
    for i, t in zip(ids[100:], docs[100:]):
        inc.add_document(i, list(t))
    snapshot("zipf_incremental_stale_idf", inc, queries[:12])

    # 5. removal forces a rebuild on the next search
    inc.remove_document("z003")
    inc.remove_document("z100")
    snapshot("zipf_after_remove", inc, queries[:8])

    # 6. many exact ties: identical documents
    tie = BM25Index()
    for d in range(12):
        tie.add_document(f"tie{d:02d}", ["alpha", "beta"] if d % 3 else ["alpha", "gamma", "gamma"])
    snapshot("ties", tie, [["alpha"], ["gamma", "alpha"], ["beta"]])

    tok_cases = []
    for text in ["Hello, World! This is a test.", "test@email.com, user's data!", "",
                 "version 2.0 released in 2024", "café_x 2.0", "ÀÉÎ õü  tab\tsep\nnew",
                 "a bb ccc d ee", "İstanbul ǅ ß", "x1y2z3-abc__def"]:
        tok_cases.append({"text": text, "tokens": _tokenize(text)})

    (GOLD / "bm25_cases.json").write_text(json.dumps({"cases": cases, "tokenizer": tok_cases},
                                                     ensure_ascii=False))
    print("bm25_cases.json:", len(cases), "cases")


def gen_rrf() -> None:
    from radiant.agents.fusion import RRFAgent
    from radiant.config import RetrievalConfig
    from radiant.storage.base import StoredDoc

    logging.disable(logging.CRITICAL)
    rng = np.random.default_rng(5)
    cases = []

    def run_case(runs, top_k, rrf_k, cfg_top=15, cfg_rrf=60):
        agent = RRFAgent(RetrievalConfig(fused_top_k=cfg_top, rrf_k=cfg_rrf))
        doc_runs = [[(StoredDoc(doc_id=f"d{int(i)}", content="", meta={}), 1.0) for i in run] for run in runs]
        kwargs = {"runs": doc_runs}
        if top_k is not None:
            kwargs["top_k"] = top_k
        if rrf_k is not None:
            kwargs["rrf_k"] = rrf_k
        res = agent.run(**kwargs)
        assert res.success
        cases.append({
            "runs": [[int(i) for i in run] for run in runs], "top_k": top_k, "rrf_k": rrf_k,
            "cfg_top": cfg_top, "cfg_rrf": cfg_rrf,
            "ids": [int(d.doc_id[1:]) for d, _ in res.data], "scores": [float(s) for _, s in res.data],
        })

    # known-answer cases of reference tests/test_all.py:1336-1388
    run_case([[1, 2], [2, 3]], None, None, cfg_top=5)
    run_case([[], []], None, None, cfg_top=5)
    run_case([[0, 1, 2, 3, 4]], None, None, cfg_top=2)
    # tie: first-insertion order wins: [[z,b],[a,b]] -> b, z, a
    run_case([[26, 2], [1, 2]], 10, None)
    # random runs: overlap, duplicates inside a run, 1-4 runs, different lengths and constants
    for _ in range(40):
        n_runs = int(rng.integers(1, 5))
        runs = []
        for _r in range(n_runs):
            ln = int(rng.integers(0, 30))
            run = rng.choice(60, size=ln, replace=bool(rng.integers(0, 2))) if ln else np.array([], int)
            runs.append([int(x) for x in run])
        run_case(runs, int(rng.integers(1, 25)), [None, 60, 1, 10, 1000][int(rng.integers(0, 5))])
    # config-3 shape: dense top-100 + bm25 top-100, fused top-10
    for _ in range(6):
        a = rng.permutation(400)[:100]
        b = np.concatenate([rng.permutation(a)[:40], rng.permutation(np.arange(400, 700))[:60]])
        rng.shuffle(b)
        run_case([[int(x) for x in a], [int(x) for x in b]], 10, 60)
    logging.disable(logging.NOTSET)
    (GOLD / "rrf_cases.json").write_text(json.dumps({"cases": cases}))
    print("rrf_cases.json:", len(cases), "cases")


def gen_rescore() -> None:
    from radiant.storage.quantization import rescore_candidates

    rng = np.random.default_rng(7)
    out = {}
    n_cases = 0
    for dim, c, kind in [(384, 40, "i8"), (384, 40, "f32"), (768, 200, "f32"), (1024, 40, "i8"),
                         (100, 17, "f32"), (36, 9, "i8")]:
        q = rng.standard_normal(dim).astype(np.float32)
        q /= np.linalg.norm(q)
        if kind == "i8":
            rows = rng.integers(-128, 128, size=(c, dim)).astype(np.int8)
            rows[3] = rows[1]  # exact tie: stable order must keep candidate order
        else:
            rows = rng.standard_normal((c, dim)).astype(np.float32)
            rows /= np.linalg.norm(rows, axis=1, keepdims=True)
            rows[5] = rows[2]
        ids = [f"c{i}" for i in range(c)]
        res = rescore_candidates(q, [r for r in rows], ids)
        out[f"q_{n_cases}"] = q
        out[f"rows_{n_cases}"] = rows
        out[f"order_{n_cases}"] = np.asarray([int(d[1:]) for d, _ in res], dtype=np.int64)
        out[f"scores_{n_cases}"] = np.asarray([s for _, s in res], dtype=np.float64)
        n_cases += 1
    out["n_cases"] = np.asarray(n_cases)
    np.savez_compressed(GOLD / "rescore_cases.npz", **out)
    print("rescore_cases.npz:", n_cases, "cases")


def gen_redis_flow() -> None:
    """BASELINE config 1 through the reference's RedisVectorStore on the fake redis."""
    from radiant.config import QuantizationConfig, RedisConfig
    import radiant.storage.quantization as refq
    from radiant.storage.redis_store import RedisVectorStore
    from radiant.storage.base import StoredDoc

    assert refq.QUANTIZATION_AVAILABLE, "stub quantiser not picked up"
    n, dim, nq, top_k, mult = 10_000, 384, 64, 10, 4.0
    corpus = synthetic.normal_unit_rows(n, dim, seed=0)
    queries = synthetic.normal_unit_rows(nq, dim, seed=1000)
    ranges = oracle.calculate_int8_ranges(corpus)
    levels = np.where((np.arange(n) * 2654435761 % 97) < 48, "parent", "child")

    tmp = tempfile.mkdtemp()
    rfile = os.path.join(tmp, "int8_ranges.npy")
    np.save(rfile, ranges)
    _fake_redis.reset()
    cfg = RedisConfig(quantization=QuantizationConfig(enabled=True, precision="both",
                                                      rescore_multiplier=mult, int8_ranges_file=rfile))
    store = RedisVectorStore(cfg)
    logging.disable(logging.CRITICAL)
    for r in range(n):
        meta = {"doc_level": str(levels[r])}
        # a few documents lose their int8 side key later -> float32 fallback (redis_store.py:829-840)
        store.upsert(f"{r:06d}", f"doc {r}", corpus[r].tolist(), meta)
    no_int8 = [7, 123, 4567, 9999]
    for r in no_int8:
        store._r.delete(store._int8_doc_key(f"{r:06d}"))

    out = {"n": n, "dim": dim, "nq": nq, "top_k": top_k, "mult": mult, "seed_corpus": 0,
           "seed_queries": 1000, "no_int8_rows": np.asarray(no_int8), "levels_parent": levels == "parent"}

    # (a) R5: the real linear scan, three filter/threshold variants, first 16 queries
    lin = {}
    for tag, kwargs in [("all", {}), ("child", {"doc_level_filter": "child"}),
                        ("parent_min", {"doc_level_filter": "parents", "min_similarity": 0.12})]:
        ids = np.full((16, top_k), -1, dtype=np.int64)
        sc = np.zeros((16, top_k), dtype=np.float64)
        for qi in range(16):
            res = store._retrieve_by_embedding_linear(queries[qi].tolist(), top_k,
                                                      kwargs.get("min_similarity", 0.0),
                                                      kwargs.get("doc_level_filter"))
            for j, (d, s) in enumerate(res):
                ids[qi, j] = int(d.doc_id)
                sc[qi, j] = s
        lin[tag] = (ids, sc)
        out[f"linear_{tag}_ids"] = ids
        out[f"linear_{tag}_scores"] = sc

    # (b) stage 1 by the oracle's Hamming search (NOT reference code), stage 2 by the REAL
    #     retrieve_by_embedding_quantized (int8 load, fallback, rescore_candidates, cut, filter)
    codes = oracle.quantize_ubinary(corpus)
    qcodes = oracle.quantize_ubinary(queries)
    for tag, valid, min_sim in [("all", None, 0.0), ("child", levels == "child", 0.0),
                                ("all_min", None, 0.25)]:
        cand_k = int(top_k * mult)
        dist, cand = oracle.hamming_topk(codes, qcodes, cand_k, valid=valid)
        ids = np.full((nq, top_k), -1, dtype=np.int64)
        sc = np.zeros((nq, top_k), dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int64)
        for qi in range(nq):
            stage1 = [(StoredDoc(doc_id=f"{int(r):06d}", content="", meta={}), 0.0) for r in cand[qi] if r >= 0]
            store.retrieve_by_embedding = lambda *a, _s=stage1, **k: list(_s)  # stage-1 injection
            res = store.retrieve_by_embedding_quantized(queries[qi].tolist(), top_k, min_similarity=min_sim)
            cnt[qi] = len(res)
            for j, (d, s) in enumerate(res):
                ids[qi, j] = int(d.doc_id)
                sc[qi, j] = s
        del store.retrieve_by_embedding
        out[f"flow_{tag}_cand"] = cand
        out[f"flow_{tag}_dist"] = dist
        out[f"flow_{tag}_ids"] = ids
        out[f"flow_{tag}_scores"] = sc
        out[f"flow_{tag}_count"] = cnt
    logging.disable(logging.NOTSET)
    # checksums of the oracle's (unpinned) quantiser output on this corpus
    i8 = oracle.quantize_int8(corpus, ranges)
    out["codes_crc"] = np.asarray([int(codes.astype(np.uint64).sum()), int(np.bitwise_xor.reduce(codes.view(np.uint64).ravel()))], dtype=np.uint64)
    out["int8_sum"] = np.asarray(int(i8.astype(np.int64).sum()))
    np.savez_compressed(GOLD / "redis_flow.npz", **out)
    print("redis_flow.npz written")


def gen_agent_chain() -> None:
    from radiant.agents.bm25 import BM25RetrievalAgent
    from radiant.agents.dense import DenseRetrievalAgent
    from radiant.agents.fusion import RRFAgent
    from radiant.config import BM25Config, RetrievalConfig
    from radiant.storage.base import BaseVectorStore, StoredDoc
    from radiant.storage.bm25_index import PersistentBM25Index

    class MemStore(BaseVectorStore):
        """Minimal in-memory reference-API store: exact cosine as redis_store.py:863-952."""

        def __init__(self):
            self.docs, self.emb = {}, {}

        def ping(self): return True
        def make_doc_id(self, content, meta=None): return self._default_make_doc_id(content, meta)
        def upsert(self, doc_id, content, embedding, meta=None):
            self.docs[doc_id] = StoredDoc(doc_id, content, dict(meta or {}))
            self.emb[doc_id] = np.asarray(embedding, dtype=np.float32)
        def upsert_doc_only(self, doc_id, content, meta=None):
            self.docs[doc_id] = StoredDoc(doc_id, content, dict(meta or {}))
        def upsert_batch(self, documents): [self.upsert(**d) for d in documents]; return len(documents)
        def upsert_doc_only_batch(self, documents): [self.upsert_doc_only(**d) for d in documents]; return len(documents)
        def get_doc(self, doc_id): return self.docs.get(doc_id)
        def has_embedding(self, doc_id): return doc_id in self.emb
        def delete_doc(self, doc_id): return self.docs.pop(doc_id, None) is not None
        def list_doc_ids(self, pattern="*", limit=10_000): return list(self.docs)[:limit]
        def list_doc_ids_with_embeddings(self, limit=10_000): return list(self.emb)[:limit]
        def get_index_info(self): return {}
        def drop_index(self, delete_documents=False): return True
        def count_documents(self): return len(self.docs)

        def retrieve_by_embedding(self, query_embedding, top_k, min_similarity=0.0, ef_runtime=None,
                                  language_filter=None, doc_level_filter=None):
            q = np.asarray(query_embedding, dtype=np.float32)
            q = q / np.linalg.norm(q)
            scored = []
            for doc_id, e in self.emb.items():
                level = self.docs[doc_id].meta.get("doc_level", "child")
                if doc_level_filter in ("child", "leaves", "leaf") and level != "child":
                    continue
                if doc_level_filter in ("parent", "parents") and level != "parent":
                    continue
                s = float(np.dot(q, e / np.linalg.norm(e)))
                if s >= min_similarity:
                    scored.append((self.docs[doc_id], s))
            scored.sort(key=lambda x: x[1], reverse=True)
            return scored[:top_k]

    logging.disable(logging.CRITICAL)
    emb = HashEmbedder(64)
    store = MemStore()
    for i, text in enumerate(CORPUS_TEXTS):
        meta = {"doc_level": "parent" if i % 7 == 3 else "child"}
        store.upsert(f"doc{i:03d}", text, emb.embed_single(text), meta)
    tmp = tempfile.mkdtemp()
    bm = PersistentBM25Index(BM25Config(index_path=os.path.join(tmp, "bm25")), store)
    bm.build_from_store()
    rcfg = RetrievalConfig(dense_top_k=8, bm25_top_k=8, fused_top_k=6, rrf_k=60)
    dense = DenseRetrievalAgent(store, emb, rcfg)
    sparse = BM25RetrievalAgent(bm, rcfg)
    rrf = RRFAgent(rcfg)
    cases = []
    for qtext in QUERY_TEXTS:
        d = dense.run(query=qtext).data
        s = sparse.run(query=qtext).data
        f = rrf.run(runs=[d, s]).data
        from radiant.storage.bm25_index import _tokenize as _tok
        full = bm.index.search(_tok(qtext), top_k=len(CORPUS_TEXTS))
        cases.append({
            "query": qtext,
            # every matching document: lets the test apply the canonical (score desc, row asc)
            # order where the reference's argpartition/argsort leaves ties unspecified
            "bm25_full": [[doc_id, float(v)] for doc_id, v in full],
            "dense": [[x.doc_id, float(v)] for x, v in d],
            "bm25": [[x.doc_id, float(v)] for x, v in s],
            "fused": [[x.doc_id, float(v)] for x, v in f],
        })
    logging.disable(logging.NOTSET)
    (GOLD / "agent_chain.json").write_text(json.dumps({"cases": cases}, ensure_ascii=False))
    print("agent_chain.json:", len(cases), "queries")


def main() -> None:
    if not REF.exists():
        raise SystemExit("/root/reference not present: goldens can only be regenerated in the build container")
    GOLD.mkdir(parents=True, exist_ok=True)
    _install_stubs()
    gen_bm25()
    gen_rrf()
    gen_rescore()
    gen_agent_chain()
    gen_redis_flow()


if __name__ == "__main__":
    main()
