"""GPU parity: R5 for batches - the tensor-core exact float32 scan (rr_exact_search_f32_tc:
TF32 filter + float64 refine) against the oracle and, bit for bit, against the CUDA-core scan."""

import numpy as np
import pytest
import torch

import oracle
from radiant_rag_b200 import synthetic
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
from tests.gpu_util import assert_lists_match_tie_aware, build_index, require_gpu

pytestmark = pytest.mark.gpu


def _same(a, b):
    return all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("n,dim,nq,k", [(40_000, 768, 33, 10),     # one query block, 24 K blocks
                                        (150_000, 384, 130, 100),  # two query blocks (128 + 2), deep lists
                                        (5_000, 1024, 8, 5),       # fewer row tiles than SMs, widest rows
                                        (20_011, 96, 17, 12)])     # 3 K blocks, ragged last tile
def test_exact_tc_equals_cuda_core_scan_and_oracle(n, dim, nq, k):
    require_gpu()
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=n)
    corpus[5] = 0.0  # zero-norm row is skipped
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=n, n_corpus=n)
    queries[2] = 0.0  # zero-norm query returns nothing
    idx, _ = build_index(corpus, int8=False, f32=True)
    assert idx.search_exact(queries, k, 0.0, use_tc=True)[0].shape == (nq, k)
    tc = idx.search_exact(queries, k, 0.0, use_tc=True)
    assert idx.last_tc_redone == 0
    cc = idx.search_exact(queries, k, 0.0, use_tc=False)
    assert _same(tc, cc), "ids, float32 scores and counts are bit-identical to rr_exact_search_f32"
    ids, score, count = tc
    assert int(count[2]) == 0
    for qi in list(range(min(nq, 6))) + [nq - 1]:
        w_ids, w_s = oracle.exact_cosine_topk(queries[qi], corpus, k, 0.0, exact=True)
        m = int(count[qi])
        assert m == len(w_ids), qi
        assert_lists_match_tie_aware(ids[qi, :m].cpu().tolist(), score[qi, :m].cpu().tolist(),
                                     w_ids.tolist(), w_s.tolist(), rel=1e-6, floor=1e-7, ctx=qi)


def test_exact_tc_against_reference_linear_scan(golden_dir):
    """The reference's RedisVectorStore._retrieve_by_embedding_linear output, filters and threshold."""
    require_gpu()
    z = np.load(golden_dir / "redis_flow.npz")
    corpus = synthetic.normal_unit_rows(10_000, 384, seed=0)
    queries = synthetic.normal_unit_rows(64, 384, seed=1000)
    tags = np.where(z["levels_parent"], 2, 1).astype(np.uint8)
    idx, _ = build_index(corpus, int8=False, f32=True, tags=tags)
    for tag, (m, v, min_sim) in {"all": (0, 0, 0.0), "child": (3, 1, 0.0), "parent_min": (3, 2, 0.12)}.items():
        got = idx.search_exact(queries[:16], 10, min_sim, m, v, use_tc=True)
        assert _same(got, idx.search_exact(queries[:16], 10, min_sim, m, v, use_tc=False)), tag
        ids, score, count = got
        ref_ids, ref_s = z[f"linear_{tag}_ids"], z[f"linear_{tag}_scores"]
        for qi in range(16):
            mref = int((ref_ids[qi] >= 0).sum())
            assert int(count[qi]) == mref, (tag, qi)
            assert_lists_match_tie_aware(ids[qi, :mref].cpu().tolist(), score[qi, :mref].cpu().tolist(),
                                         ref_ids[qi, :mref].tolist(), ref_s[qi, :mref].tolist(), ctx=(tag, qi))


def test_exact_tc_negative_scores_and_unnormalised_rows():
    """Rows of very different lengths and queries whose best cosines are negative: the bound is on
    the cosine, not on the dot product."""
    require_gpu()
    rng = np.random.default_rng(3)
    n, dim, nq = 30_000, 256, 16
    base = rng.standard_normal(dim).astype(np.float32)
    corpus = (base[None, :] * 0.9 + rng.standard_normal((n, dim)).astype(np.float32) * 0.6)
    corpus *= (10.0 ** rng.uniform(-3, 3, size=(n, 1))).astype(np.float32)
    queries = np.concatenate([-base[None, :] + 0.3 * rng.standard_normal((nq // 2, dim)).astype(np.float32),
                              rng.standard_normal((nq // 2, dim)).astype(np.float32) * 50.0]).astype(np.float32)
    idx, _ = build_index(corpus, int8=False, f32=True)
    tc = idx.search_exact(queries, 10, -1.0, use_tc=True)
    assert _same(tc, idx.search_exact(queries, 10, -1.0, use_tc=False))
    assert float(tc[1][0, 0]) < 0.0  # the anti-aligned queries only have negative cosines
    assert int(tc[2].min()) == 10


def test_exact_tc_clustered_duplicates_redo_only_flagged_queries():
    """Thousands of near-copies of one row stored next to each other overflow a CTA's list segment for
    the queries that match them: flagged, redone on the CUDA-core path, results identical."""
    require_gpu()
    n, dim, nq = 60_000, 128, 12
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=9)
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=9, n_corpus=n)
    corpus[10_000:10_000 + 128 * 148] = queries[1][None, :] * np.float32(0.5)  # one full tile per CTA, all tied
    idx, _ = build_index(corpus, int8=False, f32=True)
    tc = idx.search_exact(queries, 10, 0.0, use_tc=True)
    assert 1 <= idx.last_tc_redone < nq
    assert _same(tc, idx.search_exact(queries, 10, 0.0, use_tc=False))
    assert tc[0][1].cpu().tolist() == list(range(10_000, 10_010))  # ties resolve to the lowest rows
    # unchecked call (graph replay): the event is counted instead
    idx.tc_overflow_reset()
    idx.search_exact(queries, 10, 0.0, use_tc=True, check_overflow=False)
    assert idx.tc_overflow_total() > 0
    idx.tc_overflow_reset()


def test_exact_tc_row_updates_refresh_the_cached_norms():
    require_gpu()
    n, dim, nq = 20_000, 128, 9
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=21)
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=21, n_corpus=n)
    idx, _ = build_index(corpus, int8=False, f32=True)
    before = idx.search_exact(queries, 5, 0.0, use_tc=True)
    idx.set_row(123, queries[0] * np.float32(1e4), 1)      # now the best match of query 0, with a huge norm
    idx.add(queries[3:4] * np.float32(1e-3))               # appended row: best match of query 3, tiny norm
    after = idx.search_exact(queries, 5, 0.0, use_tc=True)
    assert _same(after, idx.search_exact(queries, 5, 0.0, use_tc=False))
    assert int(after[0][0, 0]) == 123 and int(after[0][3, 0]) == n
    assert not torch.equal(before[0], after[0])


def test_exact_tc_row_deletes_keep_the_cached_norms_consistent():
    """delete_row moves the last row into the hole; a later add reuses the freed slot at the end.  The cached
    1/|row| values must follow (a stale one would silently steer the filter wrong)."""
    require_gpu()
    n, dim, nq = 20_000, 128, 9
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=23)
    corpus[n - 1] *= np.float32(1e3)   # the row that will move has a very different norm from the one it replaces
    queries = synthetic.hash_query_rows_f32(0, nq, dim, seed=23, n_corpus=n)
    idx, _ = build_index(corpus, int8=False, f32=True)
    idx.search_exact(queries, 5, 0.0, use_tc=True)          # norms cached for all n rows
    assert idx.delete_row(77) == n - 1 and idx.n == n - 1
    got = idx.search_exact(queries, 5, 0.0, use_tc=True)
    assert _same(got, idx.search_exact(queries, 5, 0.0, use_tc=False))
    idx.add(queries[4:5] * np.float32(1e-4))                 # lands in slot n - 1: tiny norm, best match of query 4
    got = idx.search_exact(queries, 5, 0.0, use_tc=True)
    assert _same(got, idx.search_exact(queries, 5, 0.0, use_tc=False))
    assert int(got[0][4, 0]) == n - 1
    assert idx.delete_row(idx.n - 1) is None and idx.n == n - 1   # removing the last row moves nothing


def test_exact_tc_at_size_properties():
    """1M x 768 (BASELINE config 2/3 corpus), 64 queries, device-generated: bit-identical to the
    CUDA-core scan, and every query finds the row it was derived from first."""
    require_gpu()
    n, dim, nq = 1_000_000, 768, 64
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=True, capacity=n)
    for lo in range(0, n, 125_000):
        idx.add(synth_rows_device(lo, 125_000, dim, 1))
    qs = synth_query_rows_device(0, nq, dim, 1, n)
    tc = idx.search_exact(qs, 10, 0.0, use_tc=True)
    assert idx.last_tc_redone == 0
    assert _same(tc, idx.search_exact(qs, 10, 0.0, use_tc=False))
    src = synthetic.query_source_row(np.arange(nq), n, 1)
    for qi in range(1, nq, 2):  # odd queries copy 3/4 of a corpus row: that row is the nearest
        assert int(tc[0][qi, 0]) == int(src[qi]), qi
    rows = idx.f32[:n].cpu().numpy()
    for qi in (0, 17, 63):
        w_ids, w_s = oracle.exact_cosine_topk(qs[qi].cpu().numpy(), rows, 10, 0.0, exact=False)
        assert_lists_match_tie_aware(tc[0][qi].cpu().tolist(), tc[1][qi].cpu().tolist(), w_ids.tolist(), w_s.tolist(),
                                     rel=1e-5, floor=1e-6, ctx=qi)
