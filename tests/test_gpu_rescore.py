"""GPU parity: R3 rescoring, R5 exact cosine scan, config-4 int8 exact search."""

import numpy as np
import pytest
import torch

import oracle
from radiant_rag_b200 import quantization as q
from radiant_rag_b200 import synthetic
from tests.gpu_util import (ABS_FLOOR, REL, assert_lists_match_tie_aware, build_index, require_gpu,
                            ulp_diff_f32)

pytestmark = pytest.mark.gpu


def test_rescore_candidates_against_reference_golden(golden_dir):
    """The reference's own rescore_candidates output (tests/golden/rescore_cases.npz)."""
    require_gpu()
    z = np.load(golden_dir / "rescore_cases.npz")
    for i in range(int(z["n_cases"])):
        qv, rows = z[f"q_{i}"], z[f"rows_{i}"]
        ref_order, ref_scores = z[f"order_{i}"].tolist(), z[f"scores_{i}"].tolist()
        got = q.rescore_candidates(qv, [r for r in rows], [f"c{j}" for j in range(rows.shape[0])])
        got_ids = [int(d[1:]) for d, _ in got]
        got_s = [s for _, s in got]
        scale = float(np.abs(rows.astype(np.float64)).max()) * np.sqrt(rows.shape[1])
        assert_lists_match_tie_aware(got_ids, got_s, ref_order, ref_scores, floor=ABS_FLOOR * scale, ctx=i)
        pos = {d: p for p, d in enumerate(got_ids)}
        dup = (1, 3) if rows.dtype == np.int8 else (2, 5)
        assert pos[dup[0]] < pos[dup[1]], "exact ties keep candidate order (stable sort)"
        # against the correctly rounded oracle: same order, scores within 1 float32 ulp
        o_ids, o_s = oracle.rescore_f32(qv, rows, np.arange(rows.shape[0]), min_similarity=float("-inf"), exact=True)
        assert got_ids == o_ids.tolist(), i
        assert ulp_diff_f32(np.asarray(got_s, np.float32), o_s).max() <= 1, i


@pytest.mark.parametrize("dim,use_int8", [(384, True), (768, False), (1024, True), (100, False), (36, True)])
def test_batched_rescore_cut_and_threshold(dim, use_int8):
    require_gpu()
    n, nq, c, top_k = 6000, 37, 40, 10
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=dim)
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=dim, n_corpus=n)
    idx, ranges = build_index(corpus, int8=use_int8, f32=True, row_base=500)
    rows = oracle.quantize_int8(corpus, ranges) if use_int8 else corpus
    rng = np.random.default_rng(0)
    cand = rng.integers(0, n, size=(nq, c)).astype(np.int64) + 500
    cand[0, 5] = -1            # missing candidate
    cand[1, :] = -1            # nothing to rescore
    cand[2, 7] = cand[2, 3]    # the same row twice: equal scores, candidate order kept
    cand[3, 9] = 499           # below this shard's row range: skipped
    cand_d = torch.from_numpy(cand).cuda()
    qd = torch.from_numpy(queries).cuda()
    for min_sim in (float("-inf"), 0.0, 0.05):
        score, ids, count = idx.rescore(qd, cand_d, top_k, min_sim, prefer_int8=use_int8)
        torch.cuda.synchronize()
        for qi in range(nq):
            local = cand[qi] - 500
            ok = (cand[qi] >= 0) & (local >= 0) & (local < n)
            w_ids, w_s = oracle.rescore_f32(queries[qi], rows[np.where(ok, local, 0)],
                                            np.where(ok, cand[qi], -1), top_k, min_sim, exact=True)
            m = int(count[qi])
            assert m == len(w_ids), (qi, min_sim)
            assert ids[qi, :m].cpu().tolist() == w_ids.tolist(), (qi, min_sim)
            assert ulp_diff_f32(score[qi, :m].cpu().numpy(), w_s).max(initial=0) <= 1
            assert (ids[qi, m:].cpu().numpy() == -1).all()


def test_int8_symmetric_rescore_exact():
    require_gpu()
    n, dim, nq, c = 5000, 1024, 16, 40
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=4)
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=4, n_corpus=n)
    idx, ranges = build_index(corpus, int8=True, f32=False)
    i8 = oracle.quantize_int8(corpus, ranges)
    q8 = oracle.quantize_int8_symmetric_query(queries, ranges)
    q8d = idx.quantize_int8_queries(queries)
    assert np.array_equal(q8d.cpu().numpy(), q8)
    cand = np.random.default_rng(1).integers(0, n, size=(nq, c)).astype(np.int64)
    cand[0, 1] = cand[0, 0]
    score, ids, count = idx.rescore_int8_symmetric(q8d, torch.from_numpy(cand).cuda(), 10)
    for qi in range(nq):
        w_ids, w_s = oracle.rescore_i8_exact(q8[qi], i8[cand[qi]], cand[qi], 10)
        assert ids[qi].cpu().tolist() == w_ids.tolist()
        assert score[qi].cpu().tolist() == w_s.tolist()  # int32, bit-exact


def test_exact_cosine_against_reference_linear_scan(golden_dir):
    """R5: the reference's RedisVectorStore._retrieve_by_embedding_linear output."""
    require_gpu()
    z = np.load(golden_dir / "redis_flow.npz")
    corpus = synthetic.normal_unit_rows(10_000, 384, seed=0)
    queries = synthetic.normal_unit_rows(64, 384, seed=1000)
    tags = np.where(z["levels_parent"], 2, 1).astype(np.uint8)
    idx, _ = build_index(corpus, int8=False, f32=True, tags=tags)
    for tag, (m, v, min_sim) in {"all": (0, 0, 0.0), "child": (3, 1, 0.0), "parent_min": (3, 2, 0.12)}.items():
        ids, score, count = idx.search_exact(queries[:16], 10, min_sim, m, v)
        ref_ids, ref_s = z[f"linear_{tag}_ids"], z[f"linear_{tag}_scores"]
        for qi in range(16):
            mref = int((ref_ids[qi] >= 0).sum())
            assert int(count[qi]) == mref, (tag, qi)
            assert_lists_match_tie_aware(ids[qi, :mref].cpu().tolist(), score[qi, :mref].cpu().tolist(),
                                         ref_ids[qi, :mref].tolist(), ref_s[qi, :mref].tolist(), ctx=(tag, qi))


@pytest.mark.parametrize("n,dim,nq", [(3000, 128, 9), (20_000, 768, 3), (777, 100, 17),
                                      (50_001, 384, 1),    # single query (the reference's call shape), 13 select chunks
                                      (601, 1024, 5)])     # widest rows the index takes, n not a multiple of 4
def test_exact_cosine_vs_oracle(n, dim, nq):
    require_gpu()
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=n)
    corpus[5] = 0.0  # zero-norm row is skipped
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=n, n_corpus=n)
    if nq > 2:
        queries[2] = 0.0  # zero-norm query returns nothing
    idx, _ = build_index(corpus, int8=False, f32=True)
    ids, score, count = idx.search_exact(queries, 12, 0.0)
    for qi in range(nq):
        w_ids, w_s = oracle.exact_cosine_topk(queries[qi], corpus, 12, 0.0, exact=True)
        m = int(count[qi])
        assert m == len(w_ids), qi
        assert_lists_match_tie_aware(ids[qi, :m].cpu().tolist(), score[qi, :m].cpu().tolist(),
                                     w_ids.tolist(), w_s.tolist(), rel=1e-6, floor=1e-7, ctx=qi)
    if nq > 2:
        assert int(count[2]) == 0


@pytest.mark.parametrize("n,dim,nq,k", [(4000, 1024, 11, 10), (50_000, 256, 4, 100), (300, 100, 3, 10)])
def test_int8_exact_search_bit_exact(n, dim, nq, k):
    require_gpu()
    rng = np.random.default_rng(n)
    emb = rng.integers(-128, 128, size=(n, dim)).astype(np.int8)
    emb[10] = emb[3]  # exact score ties -> row asc
    qs = rng.integers(-128, 128, size=(nq, dim)).astype(np.int8)
    from radiant_rag_b200.index import DenseIndex
    idx = DenseIndex(dim, device=0, store_int8=True, store_f32=False,
                     int8_ranges=np.stack([-np.ones(dim, np.float32), np.ones(dim, np.float32)]))
    idx._reserve(n)
    idx.int8[:n].copy_(torch.from_numpy(emb))
    idx.tags[:n].fill_(1)
    idx.n = n
    ids, score = idx.search_int8_exact(qs, k)
    w_ids, w_s = oracle.int8_exact_topk(qs, emb, k)
    assert np.array_equal(ids.cpu().numpy(), w_ids)
    assert np.array_equal(score.cpu().numpy(), w_s)


@pytest.mark.parametrize("dim,use_int8,c", [(768, False, 400), (1024, True, 40), (1024, False, 200)])
def test_bulk_copy_ring_scoring_stress(dim, use_int8, c):
    """The shared-memory-staged scoring kernel (one elected thread bulk-copies candidate rows into a
    ring, seven warps consume) at the BASELINE batch shapes: every score equals the correctly
    rounded oracle value and the register-staged fused kernel bit for bit, on rows spread over a
    300k-row index (copies complete out of order) with invalid candidates mixed in."""
    require_gpu()
    n, nq = 300_000, 1024
    from radiant_rag_b200 import _lib
    from radiant_rag_b200.index import DenseIndex, _stream, synth_query_rows_device, synth_rows_device
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    idx = DenseIndex(dim, device=0, store_int8=use_int8, store_f32=not use_int8, int8_ranges=ranges if use_int8 else None,
                     row_base=1000, capacity=n)
    for lo in range(0, n, 100_000):
        idx.add(synth_rows_device(1000 + lo, 100_000, dim, seed=5))
    queries = synth_query_rows_device(0, nq, dim, seed=5, n_corpus=n)
    g = torch.Generator(device="cuda").manual_seed(3)
    cand = torch.randint(0, n, (nq, c), generator=g, device="cuda", dtype=torch.int64) + 1000
    cand[5, 3] = -1
    cand[6, :] = -1
    cand[7, 1] = 999          # not owned by this shard
    cand[8, 2] = 1000 + n     # not owned either
    for _ in range(3):
        ring = idx.score_candidates(queries, cand, prefer_int8=use_int8)
    # the fused register-staged kernel with top_k = c returns every valid score, sorted
    f_score = torch.empty((nq, c), dtype=torch.float32, device="cuda")
    f_idx = torch.empty((nq, c), dtype=torch.int64, device="cuda")
    f_cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    rows, dt = idx.rescore_source(use_int8)
    _lib.call("rr_rescore_f32", queries.data_ptr(), nq, dim, rows.data_ptr(), dt, idx.n, idx.row_base, cand.data_ptr(),
              c, c, float("-inf"), f_score.data_ptr(), f_idx.data_ptr(), f_cnt.data_ptr(), _stream())
    torch.cuda.synchronize()
    ring_h, cand_h = ring.cpu().numpy(), cand.cpu().numpy()
    f_s, f_c = f_score.cpu().numpy(), f_cnt.cpu().numpy()
    for qi in range(nq):
        valid = ring_h[qi] != -np.inf
        local = cand_h[qi] - 1000
        assert np.array_equal(valid, (cand_h[qi] >= 0) & (local >= 0) & (local < n)), qi
        assert np.array_equal(np.sort(ring_h[qi][valid])[::-1], f_s[qi, : f_c[qi]]), qi
    qh = queries.cpu().numpy()
    for qi in (0, 5, 7, 8, 500, 1023):
        local = cand_h[qi] - 1000
        ok = (cand_h[qi] >= 0) & (local >= 0) & (local < n)
        src = rows[torch.from_numpy(np.where(ok, local, 0)).cuda()].cpu().numpy()
        want = (src.astype(np.float64) @ qh[qi].astype(np.float64)).astype(np.float32)
        assert np.array_equal(ring_h[qi][ok], want[ok]), qi
