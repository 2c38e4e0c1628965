"""CPU: the oracle against the goldens produced by the reference's own code
(oracle/gen_golden.py).  These are the checks that PIN the oracle."""

import json

import numpy as np
import pytest

import oracle
from oracle.bm25 import BM25Oracle
from radiant_rag_b200 import synthetic

REL = 1e-5  # north-star tolerance for float32 rescoring
ABS_FLOOR = 1e-6  # |delta| <= REL*|s| + ABS_FLOOR*|q||d|  


def _load_json(golden_dir, name):
    return json.loads((golden_dir / name).read_text())


# ---------------------------------------------------------------- BM25 (R6/R7/R8)
def _oracle_from_case(case):
    """Oracle index with the idf/avgdl tables the reference actually used."""
    docs = case["doc_tokens"]
    orc, vocab = BM25Oracle.from_token_lists(docs, case["k1"], case["b"])
    idf = np.zeros(len(vocab))
    known = np.zeros(len(vocab), dtype=bool)
    for t, tid in vocab.items():
        if t in case["idf_used"]:
            idf[tid] = case["idf_used"][t]
            known[tid] = True
    orc.idf, orc.known, orc.avgdl = idf, known, case["avgdl_used"]
    return orc, vocab


def test_bm25_scores_bit_exact(golden_dir):
    data = _load_json(golden_dir, "bm25_cases.json")
    n_checked = 0
    for case in data["cases"]:
        orc, vocab = _oracle_from_case(case)
        row_of = {d: i for i, d in enumerate(case["doc_ids"])}
        for q in case["queries"]:
            tids = [vocab.get(t, -1) for t in q["tokens"]]
            s = orc.scores(tids)
            ref = {row_of[d]: v for d, v in q["full"]}
            got = {int(r): float(s[r]) for r in np.nonzero(s > 0)[0]}
            assert got.keys() == ref.keys(), (case["name"], q["tokens"])
            for r, v in ref.items():
                assert got[r] == v, (case["name"], q["tokens"], r)  # bit-exact float64
                n_checked += 1
            # canonical order is a valid ordering of the reference's scores
            rows, sc = orc.search(tids, top_k=5)
            assert list(sc) == sorted(sc, reverse=True)
    assert n_checked > 500


def test_bm25_rebuilt_idf_formula(golden_dir):
    """For rebuilt indices the oracle's own idf/avgdl equal the reference's tables."""
    data = _load_json(golden_dir, "bm25_cases.json")
    for case in data["cases"]:
        if case["name"] not in ("ref_test_rebuilt", "zipf_rebuilt", "zipf_after_remove"):
            continue
        orc, vocab = BM25Oracle.from_token_lists(case["doc_tokens"], case["k1"], case["b"])
        assert orc.avgdl == case["avgdl_used"]
        for t, tid in vocab.items():
            assert orc.idf[tid] == case["idf_used"][t], (case["name"], t)


def test_bm25_stale_idf_quirk_present(golden_dir):
    """Incremental adds leave other terms' idf stale: the golden for the
    incremental index must differ from a rebuild, else the case pins nothing."""
    data = _load_json(golden_dir, "bm25_cases.json")
    inc = next(c for c in data["cases"] if c["name"] == "zipf_incremental_stale_idf")
    orc, vocab = BM25Oracle.from_token_lists(inc["doc_tokens"], inc["k1"], inc["b"])
    diffs = sum(1 for t, tid in vocab.items() if orc.idf[tid] != inc["idf_used"][t])
    assert diffs > 10


def test_tokenizer(golden_dir):
    data = _load_json(golden_dir, "bm25_cases.json")
    for c in data["tokenizer"]:
        assert oracle.tokenize(c["text"]) == c["tokens"], c["text"]


# ---------------------------------------------------------------- RRF (R10)
def test_rrf_exact(golden_dir):
    data = _load_json(golden_dir, "rrf_cases.json")
    for c in data["cases"]:
        k = c["top_k"] or c["cfg_top"]
        rk = c["rrf_k"] or c["cfg_rrf"]
        ids, sc = oracle.rrf_fuse(c["runs"], k, rk)
        assert ids.tolist() == c["ids"], c
        assert sc.tolist() == c["scores"], c  # bit-exact float64


def test_rrf_known_answers(golden_dir):
    cases = _load_json(golden_dir, "rrf_cases.json")["cases"]
    assert cases[0]["ids"][0] == 2          # doc in both runs first (ref tests/test_all.py:1336)
    assert cases[1]["ids"] == []            # empty runs
    assert len(cases[2]["ids"]) == 2        # top_k cut
    assert cases[3]["ids"] == [2, 26, 1]    # ties keep first-insertion order


# ---------------------------------------------------------------- rescoring (R3)
def _close(a, b, scale=1.0):
    return abs(a - b) <= REL * abs(b) + ABS_FLOOR * scale


def test_rescore_against_reference(golden_dir):
    z = np.load(golden_dir / "rescore_cases.npz")
    for i in range(int(z["n_cases"])):
        q, rows = z[f"q_{i}"], z[f"rows_{i}"]
        ref_order, ref_scores = z[f"order_{i}"], z[f"scores_{i}"]
        ids = np.arange(rows.shape[0])
        scale = float(np.linalg.norm(q)) * float(np.abs(rows.astype(np.float64)).max()) * np.sqrt(rows.shape[1])
        for exact in (False, True):
            got_ids, got_s = oracle.rescore_f32(q, rows, ids, min_similarity=float("-inf"), exact=exact)
            ref_by_id = dict(zip(ref_order.tolist(), ref_scores.tolist()))
            assert sorted(got_ids.tolist()) == sorted(ref_order.tolist())
            for d, s in zip(got_ids.tolist(), got_s.tolist()):
                assert _close(s, ref_by_id[d], scale), (i, exact, d, s, ref_by_id[d])
            # order agrees wherever the reference's scores are separated by more than the tolerance
            for a, b in zip(got_ids.tolist(), ref_order.tolist()):
                if a != b:
                    assert _close(ref_by_id[a], ref_by_id[b], scale)
        # the stable tie (duplicated row) keeps candidate order in the reference
        pos = {d: p for p, d in enumerate(ref_order.tolist())}
        dup = (1, 3) if rows.dtype == np.int8 else (2, 5)
        assert pos[dup[0]] < pos[dup[1]]


# ---------------------------------------------------------------- redis flow (R5, R3, R12)
@pytest.fixture(scope="module")
def cfg1():
    corpus = synthetic.normal_unit_rows(10_000, 384, seed=0)
    queries = synthetic.normal_unit_rows(64, 384, seed=1000)
    return corpus, queries


def test_linear_scan_against_reference(golden_dir, cfg1):
    z = np.load(golden_dir / "redis_flow.npz")
    corpus, queries = cfg1
    parent = z["levels_parent"]
    variants = {"all": (None, 0.0), "child": (~parent, 0.0), "parent_min": (parent, 0.12)}
    for tag, (valid, min_sim) in variants.items():
        ref_ids, ref_sc = z[f"linear_{tag}_ids"], z[f"linear_{tag}_scores"]
        for qi in range(4):  # the non-exact path is a Python loop over 10k rows
            rows, sc = oracle.exact_cosine_topk(queries[qi], corpus, 10, min_sim, valid, exact=False)
            m = int((ref_ids[qi] >= 0).sum())
            assert rows.tolist() == ref_ids[qi, :m].tolist(), (tag, qi)
            np.testing.assert_allclose(sc, ref_sc[qi, :m], rtol=REL, atol=ABS_FLOOR)
        for qi in range(16):
            rows, sc = oracle.exact_cosine_topk(queries[qi], corpus, 10, min_sim, valid, exact=True)
            m = int((ref_ids[qi] >= 0).sum())
            assert len(rows) == m
            np.testing.assert_allclose(sc, ref_sc[qi, :m], rtol=REL, atol=ABS_FLOOR)
            assert rows.tolist() == ref_ids[qi, :m].tolist(), (tag, qi)


def test_two_stage_flow_against_reference(golden_dir, cfg1):
    z = np.load(golden_dir / "redis_flow.npz")
    corpus, queries = cfg1
    ranges = oracle.calculate_int8_ranges(corpus)
    codes = oracle.quantize_ubinary(corpus)
    i8 = oracle.quantize_int8(corpus, ranges)
    assert int(i8.astype(np.int64).sum()) == int(z["int8_sum"])
    assert int(codes.astype(np.uint64).sum()) == int(z["codes_crc"][0])
    no_int8 = z["no_int8_rows"]
    parent = z["levels_parent"]
    for tag, valid, min_sim in [("all", None, 0.0), ("child", ~parent, 0.0), ("all_min", None, 0.25)]:
        dist, cand = oracle.hamming_topk(codes, oracle.quantize_ubinary(queries), 40, valid=valid)
        assert np.array_equal(cand, z[f"flow_{tag}_cand"])
        assert np.array_equal(dist, z[f"flow_{tag}_dist"])
        for qi in range(queries.shape[0]):
            ids = cand[qi][cand[qi] >= 0]
            # int8 rows preferred, float32 for the rows whose int8 key is missing
            rows = i8[ids].astype(np.float32)
            for j, r in enumerate(ids):
                if r in no_int8:
                    rows[j] = corpus[r]
            got_ids, got_s = oracle.rescore_f32(queries[qi], rows, ids, top_k=10, min_similarity=min_sim,
                                                exact=True)
            m = int(z[f"flow_{tag}_count"][qi])
            assert len(got_ids) == m, (tag, qi)
            ref_ids, ref_s = z[f"flow_{tag}_ids"][qi, :m], z[f"flow_{tag}_scores"][qi, :m]
            np.testing.assert_allclose(got_s, ref_s, rtol=REL, atol=1e-4)
            assert got_ids.tolist() == ref_ids.tolist(), (tag, qi)


# ---------------------------------------------------------------- quantiser restatement (R1/R2, unpinned)
def test_quantizer_shapes_like_reference_validation_tool():
    """The only checks the reference holds: tools/validate_quantization.py:142,159-160,169-170."""
    emb = synthetic.normal_unit_rows(5, 384, seed=3)
    b = oracle.quantize_ubinary(emb)
    assert b.dtype == np.uint8 and b.shape == (5, 48)
    r = oracle.calculate_int8_ranges(emb)
    i8 = oracle.quantize_int8(emb, r)
    assert i8.dtype == np.int8 and i8.shape == (5, 384)
    assert i8.min() >= -128 and i8.max() <= 127
    # MSB-first packing and strict > 0
    v = np.zeros((1, 16), np.float32)
    v[0, 0] = 1.0
    v[0, 9] = 2.0
    v[0, 3] = -1.0
    assert oracle.quantize_ubinary(v).tolist() == [[0x80, 0x40]]


def test_hamming_topk_tie_rule():
    codes = np.zeros((6, 16), np.uint8)
    codes[2, 0] = 0x01
    codes[4, 5] = 0x03
    q = np.zeros((1, 16), np.uint8)
    d, i = oracle.hamming_topk(codes, q, 4)
    assert d.tolist() == [[0, 0, 0, 0]] and i.tolist() == [[0, 1, 3, 5]]
    d, i = oracle.hamming_topk(codes, q, 8)
    assert i.tolist() == [[0, 1, 3, 5, 2, 4, -1, -1]]
    assert d[0, 4] == 1 and d[0, 5] == 2 and d[0, 6] == np.iinfo(np.int32).max
