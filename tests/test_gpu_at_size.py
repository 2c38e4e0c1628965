"""GPU parity AT SIZE: BASELINE configs 3, 4 and 5 (one-GPU shard shape) against oracle/ on a
sample of the batch.  Nothing here compares one product path with another: the expected values
come from the NumPy oracle over the same inputs (corpus arrays are read back from the device
where regenerating them on the CPU would take minutes; the generator itself is checked against
its NumPy restatement on a sub-sample of the rows)."""

import numpy as np
import pytest
import torch

import oracle
from oracle.bm25 import BM25Oracle
from radiant_rag_b200 import _lib, synthetic
from radiant_rag_b200.bm25_index import Bm25DeviceIndex, synth_zipf_corpus_device
from radiant_rag_b200.hybrid import GraphedHybridSearch, HybridSearch
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
from tests.gpu_util import REL, ABS_FLOOR, assert_lists_match_tie_aware, require_gpu

pytestmark = pytest.mark.gpu


def _dense_oracle(index, queries_host, sample, cand_k, top_k, prefer_int8):
    """Two-stage oracle for sampled queries over the index's own packed codes (read back once)
    with the candidate rows fetched from the device: -> [(rows, scores)] per sampled query."""
    n = index.n
    codes = index.codes[:n].cpu().numpy()[:, : (index.dim + 7) // 8]
    rows_src = index.int8 if (prefer_int8 and index.int8 is not None) else index.f32
    out = []
    for qi in sample:
        qcode = oracle.quantize_ubinary(queries_host[qi:qi + 1])
        _d, cand = oracle.hamming_topk(codes, qcode, cand_k)
        ids = cand[0][cand[0] >= 0]
        rows = rows_src[torch.from_numpy(ids).to(index.device)].cpu().numpy()
        out.append(oracle.rescore_f32(queries_host[qi], rows, ids, top_k=top_k, min_similarity=0.0, exact=True))
    return out


@pytest.fixture(scope="module")
def config3():
    """BASELINE config 3: 1M docs, 50k vocabulary, Zipf terms, ~200 tokens; 1M x 768 dense rows."""
    require_gpu()
    n_docs, v, mean_len, nq, qlen, seed, dim = 1_000_000, 50_000, 200, 1024, 8, 2, 768
    ptr, toks = synth_zipf_corpus_device(n_docs, v, seed, mean_len, device=0)
    bm = Bm25DeviceIndex.build(ptr, toks, v, None, None, 1.5, 0.75, device=0)  # built by the product from tokens
    qt = synthetic.zipf_queries(nq, qlen, v, seed)
    index = DenseIndex(dim, device=0, store_int8=False, store_f32=True, capacity=n_docs)
    for lo in range(0, n_docs, 125_000):
        index.add(synth_rows_device(lo, min(125_000, n_docs - lo), dim, seed))
    queries = synth_query_rows_device(0, nq, dim, seed, n_docs)
    torch.cuda.synchronize()
    return dict(n_docs=n_docs, v=v, ptr=ptr.cpu().numpy(), toks=toks.cpu().numpy(), bm=bm, qt=qt,
                index=index, queries=queries, seed=seed, mean_len=mean_len)


def test_config3_generator_matches_numpy_on_subsample(config3):
    c = config3
    m = 2000
    ptr_h, toks_h = synthetic.zipf_corpus(m, c["v"], c["seed"], c["mean_len"])
    assert np.array_equal(c["ptr"][: m + 1], ptr_h)
    assert np.array_equal(c["toks"][: ptr_h[-1]], toks_h)
    # a shard that starts in the middle of the corpus owns the same global token positions
    lo = 777_000
    ptr_s, toks_s = synth_zipf_corpus_device(1000, c["v"], c["seed"], c["mean_len"], device=0, row_start=lo)
    a, b = c["ptr"][lo], c["ptr"][lo + 1000]
    assert np.array_equal(ptr_s.cpu().numpy(), c["ptr"][lo: lo + 1001] - a)
    assert np.array_equal(toks_s.cpu().numpy(), c["toks"][a:b])


def test_config3_bm25_top100_bit_exact_vs_oracle(config3):
    """1M docs / 1024 queries x 8 tokens / top-100 through the batched filter-and-refine path;
    ids and float64 scores == BM25Oracle over the same tokens on a sample of the queries."""
    c = config3
    bm, qt = c["bm"], c["qt"]
    assert bm.uses_fast_path(qt.shape[0], 100)
    assert bm.n_postings > 140_000_000
    idx, score, count = bm.search_batch(qt, 100)
    torch.cuda.synchronize()
    assert bm.last_flagged <= 8, bm.last_flagged  # the exactness check rarely fails on this data
    sample = list(range(0, 1024, 43))
    need = set(int(t) for qi in sample for t in qt[qi])
    orc = BM25Oracle(c["ptr"], c["toks"], c["v"], only_terms=need)
    for qi in sample:
        rows, sc = orc.search(qt[qi].tolist(), 100)
        m = int(count[qi])
        assert m == rows.size, qi
        assert idx[qi, :m].cpu().tolist() == rows.tolist(), qi
        assert score[qi, :m].cpu().tolist() == sc.tolist(), qi  # float64 ==
    # the exact kernel agrees with the oracle as well (it is the fallback of flagged queries)
    sub = qt[sample[:6]]
    e_idx, e_score, e_count = bm.search_batch(sub, 100, exact=True)
    for j, qi in enumerate(sample[:6]):
        rows, sc = orc.search(qt[qi].tolist(), 100)
        m = int(e_count[j])
        assert e_idx[j, :m].cpu().tolist() == rows.tolist() and e_score[j, :m].cpu().tolist() == sc.tolist()
    c["orc"] = orc
    c["orc_sample"] = sample


def test_config3_hybrid_top10_vs_oracle(config3):
    """The product's one-call hybrid step (dense top-100 + BM25 top-100 -> RRF top-10), eager and
    as a CUDA-graph replay, against dense oracle + BM25Oracle + rrf_fuse on sampled queries."""
    c = config3
    if "orc" not in c:
        pytest.skip("needs the oracle built by the BM25 test")
    hybrid = HybridSearch(c["index"], c["bm"], rescore_multiplier=4.0, prefer_int8=False, overlap=False)
    qt_d = torch.from_numpy(c["qt"]).cuda()
    res = hybrid.search_batch(c["queries"], qt_d, top_k=10, dense_top_k=100, bm25_top_k=100, rrf_k=60)
    torch.cuda.synchronize()
    sample = c["orc_sample"][:12]
    qh = c["queries"].cpu().numpy()
    dense_want = _dense_oracle(c["index"], qh, sample, 400, 100, prefer_int8=False)
    for j, qi in enumerate(sample):
        d_ids, d_s = dense_want[j]
        m = int(res.dense_count[qi])
        assert_lists_match_tie_aware(res.dense_idx[qi, :m].cpu().tolist(), res.dense_score[qi, :m].cpu().tolist(),
                                     d_ids.tolist(), d_s.tolist(), ctx=f"dense q{qi}")
        b_rows, _ = c["orc"].search(c["qt"][qi].tolist(), 100)
        # RRF over the PRODUCT's dense order (float32 near-ties may swap neighbours) and the oracle's BM25 list
        ids, sc = oracle.rrf_fuse([res.dense_idx[qi, :m].cpu().tolist(), b_rows.tolist()], 10, 60)
        fm = int(res.count[qi])
        assert res.idx[qi, :fm].cpu().tolist() == ids.tolist(), qi
        assert res.score[qi, :fm].cpu().tolist() == sc.tolist(), qi
    # graph replay == eager
    g = GraphedHybridSearch(hybrid, 1024, 768, 8, top_k=10, dense_top_k=100, bm25_top_k=100, rrf_k=60)
    out = g(c["queries"].cpu().pin_memory(), torch.from_numpy(c["qt"]).pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(out.idx, res.idx) and torch.equal(out.score, res.score) and torch.equal(out.count, res.count)
    assert hybrid.unchecked_events() == 0
    # the BM25 half on a second stream next to the dense half (parallel graph branches): same bits
    hybrid2 = HybridSearch(c["index"], c["bm"], rescore_multiplier=4.0, prefer_int8=False, overlap=True)
    eager = hybrid2.search_batch(c["queries"], qt_d, top_k=10, dense_top_k=100, bm25_top_k=100, rrf_k=60, check=False)
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(eager.tensors(), res.tensors()))
    g2 = GraphedHybridSearch(hybrid2, 1024, 768, 8, top_k=10, dense_top_k=100, bm25_top_k=100, rrf_k=60)
    for _ in range(3):
        out2 = g2(c["queries"].cpu().pin_memory(), torch.from_numpy(c["qt"]).pin_memory())
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(out2.tensors(), res.tensors()))
    assert hybrid2.unchecked_events() == 0


def test_config4_int8_exact_at_size_vs_oracle():
    """BASELINE config 4 shard shape: 2.5M x 1024 int8 rows, 4096 queries, exact int32 top-10 on the
    tensor cores vs the oracle's exact int32 scores on 32 queries."""
    require_gpu()
    n, dim, nq, seed, top_k = 2_500_000, 1024, 4096, 3, 10
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    index = DenseIndex(dim, device=0, store_int8=True, store_f32=False, int8_ranges=ranges, capacity=n)
    for a in range(0, n, 250_000):
        index.add(synth_rows_device(a, min(250_000, n - a), dim, seed))
    q8 = index.quantize_int8_queries(synth_query_rows_device(0, nq, dim, seed, n))
    idx, score = index.search_int8_exact(q8, top_k)
    torch.cuda.synchronize()
    # the generator and the int8 quantiser on a sub-sample of the rows
    sub = synthetic.hash_rows_f32(1_234_000, 512, dim, seed)
    assert np.array_equal(index.int8[1_234_000: 1_234_512].cpu().numpy(), oracle.quantize_int8(sub, ranges))
    sample = np.arange(0, nq, 128)
    want_r, want_s = oracle.int8_exact_topk_blas(q8[torch.from_numpy(sample).cuda()].cpu().numpy(),
                                                 index.int8[:n].cpu().numpy(), top_k)
    assert np.array_equal(idx[torch.from_numpy(sample).cuda()].cpu().numpy(), want_r)
    assert np.array_equal(score[torch.from_numpy(sample).cuda()].cpu().numpy(), want_s)


def test_config5_shard_shape_two_stage_vs_oracle():
    """One GPU's share of BASELINE config 5: 12.5M x 1024 packed codes + int8 rows, 8192 queries,
    k' = 40, int8 rescoring, top-10 - against the oracle's two-stage flow on sampled queries at
    the full 12.5M rows, and the generator / quantisers against NumPy on a sub-sample."""
    require_gpu()
    n, dim, nq, seed, top_k = 12_500_000, 1024, 8192, 4, 10
    row_base = 25_000_000  # shard 2 of 8
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    index = DenseIndex(dim, device=0, store_int8=True, store_f32=False, int8_ranges=ranges, row_base=row_base,
                       capacity=n)
    for a in range(0, n, 500_000):
        index.add(synth_rows_device(row_base + a, min(500_000, n - a), dim, seed))
    queries = synth_query_rows_device(0, nq, dim, seed, 100_000_000)
    idx, score, count = index.search_quantized(queries, top_k, rescore_multiplier=4.0)
    torch.cuda.synchronize()
    sub = synthetic.hash_rows_f32(row_base + 7_000_000, 256, dim, seed)
    assert np.array_equal(index.codes[7_000_000: 7_000_256].cpu().numpy()[:, : dim // 8], oracle.quantize_ubinary(sub))
    assert np.array_equal(index.int8[7_000_000: 7_000_256].cpu().numpy(), oracle.quantize_int8(sub, ranges))
    sample = list(range(1, nq, 683))  # 12 queries, odd ones have a planted neighbour somewhere in the 100M rows
    qh = queries.cpu().numpy()
    codes = index.codes[:n].cpu().numpy()
    for qi in sample:
        qcode = oracle.quantize_ubinary(qh[qi:qi + 1])
        _d, cand = oracle.hamming_topk(codes, qcode, 40)
        ids = cand[0][cand[0] >= 0]
        rows = index.int8[torch.from_numpy(ids).cuda()].cpu().numpy()
        w_ids, w_s = oracle.rescore_f32(qh[qi], rows, ids + row_base, top_k=top_k, min_similarity=0.0, exact=True)
        m = int(count[qi])
        assert_lists_match_tie_aware(idx[qi, :m].cpu().tolist(), score[qi, :m].cpu().tolist(),
                                     w_ids.tolist(), w_s.tolist(), ctx=f"q{qi}")
