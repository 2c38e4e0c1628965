"""GPU parity: the tensor-core (tcgen05 kind::i8) formulation of the batched stage-1 search
and of the exact int8 search returns exactly what the CUDA-core kernels and the oracle do."""

import numpy as np
import pytest
import torch

import oracle
from radiant_rag_b200 import _lib, synthetic
from radiant_rag_b200.index import DenseIndex, _stream, synth_query_rows_device, synth_rows_device
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def _keys_to_scores(keys: np.ndarray) -> np.ndarray:
    s = ((~keys.astype(np.uint32)) ^ np.uint32(0x80000000)).view(np.int32).astype(np.int64)
    return np.where(keys == 0xFFFFFFFF, np.iinfo(np.int64).min, s)


@pytest.mark.parametrize("n,dim,q", [(128, 128, 16), (1000, 256, 5), (4096 + 37, 768, 130), (20_000, 1024, 256),
                                     (300, 384, 128)])
def test_tc_gemm_scores_exact(n, dim, q):
    """Every int8 x int8 -> int32 dot product coming out of TMEM equals NumPy's."""
    require_gpu()
    _lib.init(0)
    rng = np.random.default_rng(n + dim)
    emb = rng.integers(-128, 128, size=(n, dim)).astype(np.int8)
    qs = rng.integers(-128, 128, size=(q, dim)).astype(np.int8)
    ld = (n + 127) // 128 * 128
    e, qq = torch.from_numpy(emb).cuda(), torch.from_numpy(qs).cuda()
    out = torch.empty((q, ld), dtype=torch.int32, device="cuda")
    _lib.call("rr_tc_dense_keys", e.data_ptr(), n, dim, qq.data_ptr(), q, out.data_ptr(), _stream())
    torch.cuda.synchronize()
    got = _keys_to_scores(out.cpu().numpy().view(np.uint32))
    want = qs.astype(np.int64) @ emb.astype(np.int64).T
    assert np.array_equal(got[:, :n], want), np.argwhere(got[:, :n] != want)[:5]
    assert (got[:, n:] == np.iinfo(np.int64).min).all()


def test_unpack_pm1():
    require_gpu()
    x = synthetic.hash_rows_f32(0, 300, 384, seed=2)
    idx = DenseIndex(384, device=0, store_int8=False, store_f32=False)
    _qf, qc = idx.quantize_queries(x)
    out = torch.empty((300, 384), dtype=torch.int8, device="cuda")
    _lib.call("rr_unpack_codes_pm1", qc.data_ptr(), 300, idx.words * 4, 384, out.data_ptr(), _stream())
    assert np.array_equal(out.cpu().numpy(), np.where(x > 0, 1, -1).astype(np.int8))


def _both_paths(corpus, queries, k, tags=None, mask=0, value=0, row_base=0):
    dim = corpus.shape[1]
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False, row_base=row_base)
    idx.add(corpus, tags)
    _qf, qc = idx.quantize_queries(queries)
    d_tc, i_tc = idx.hamming_topk(qc, k, mask, value, use_tc=True, check_overflow=False)
    d_pc, i_pc = idx.hamming_topk(qc, k, mask, value, use_tc=False)
    torch.cuda.synchronize()
    return idx, (d_tc.cpu().numpy(), i_tc.cpu().numpy()), (d_pc.cpu().numpy(), i_pc.cpu().numpy())


@pytest.mark.parametrize("n,dim,q,k", [
    (10_000, 384, 64, 40),       # BASELINE config 1 shape
    (50_000, 768, 256, 200),     # config 2 width / batch / k'
    (70_000, 1024, 130, 40),     # config 5 width, ragged query block
    (5_000, 128, 17, 10),
    (4_500, 256, 300, 1000),     # k at the limit, 3 query blocks
    (200, 128, 16, 50),          # smaller than one sample
    (30_000, 100, 40, 17),       # dim not a multiple of 32 (padded to 128 bits)
    (9_000, 640, 33, 64),        # words = 20 -> 5 K blocks
    (12_000, 896, 20, 30),       # 7 K blocks: a full and a partial tensor-memory ring stage
    (6_000, 512, 48, 25),        # 4 K blocks: exactly one ring stage per tile
])
def test_tc_hamming_equals_popc_and_oracle(n, dim, q, k):
    require_gpu()
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=n + dim + 1)
    queries = synthetic.hash_query_rows_f32(0, q, dim, seed=n + dim + 1, n_corpus=n)
    idx, tc, pc = _both_paths(corpus, queries, k)
    assert idx.tc_overflow_total() == 0
    want_d, want_i = oracle.hamming_topk(oracle.quantize_ubinary(corpus), oracle.quantize_ubinary(queries), k)
    bad = np.nonzero((tc[0] != want_d).any(1) | (tc[1] != want_i).any(1))[0]
    assert bad.size == 0, (bad[:5], tc[0][bad[0]][:8], want_d[bad[0]][:8], tc[1][bad[0]][:8], want_i[bad[0]][:8])
    assert np.array_equal(pc[0], want_d) and np.array_equal(pc[1], want_i)


def test_tc_hamming_filters_row_base_and_ties():
    require_gpu()
    n, dim = 30_000, 256
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=8)
    queries = synthetic.hash_query_rows_f32(0, 48, dim, seed=8, n_corpus=n)
    tags = np.where(np.arange(n) % 3 == 0, 2, 1).astype(np.uint8)
    for mask, value in [(3, 1), (3, 2), (0xFC, 63 << 2)]:
        idx, tc, pc = _both_paths(corpus, queries, 40, tags, mask, value, row_base=5_000_000)
        assert np.array_equal(tc[0], pc[0]) and np.array_equal(tc[1], pc[1]), (mask, value)
        assert idx.tc_overflow_total() == 0
    # 128-bit codes over 30k rows: massive distance ties, resolved by row id
    c2 = synthetic.hash_rows_f32(0, n, 128, seed=9)
    q2 = synthetic.hash_query_rows_f32(0, 32, 128, seed=9, n_corpus=n)
    idx, tc, pc = _both_paths(c2, q2, 100)
    assert np.array_equal(tc[0], pc[0]) and np.array_equal(tc[1], pc[1])


def test_tc_overflow_falls_back_to_exact_path():
    """All rows identical: every row ties with the sampled bound, the filtered lists overflow,
    the counter says so and hamming_topk(check_overflow=True) redoes the call on the POPC path."""
    require_gpu()
    n, dim = 200_000, 128
    corpus = np.tile(np.linspace(-1, 1, dim, dtype=np.float32), (n, 1))
    queries = np.tile(np.linspace(-1, 1, dim, dtype=np.float32), (20, 1))
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False)
    idx.add(corpus)
    _qf, qc = idx.quantize_queries(queries)
    idx.hamming_topk(qc, 10, use_tc=True, check_overflow=False)
    assert idx.tc_overflow_total() > 0
    d, i = idx.hamming_topk(qc, 10, use_tc=True, check_overflow=True)
    assert (i.cpu().numpy() == np.arange(10)[None, :]).all() and (d.cpu().numpy() == 0).all()


def test_tc_clustered_rows_redo_only_overflowed_queries():
    """Real embeddings cluster: 4000 near-duplicates stored next to each other (one ingest batch)
    that the strided sample mostly misses.  The queries aimed at the cluster overflow their
    candidate lists; a checked call redoes exactly those on the POPC path and every result equals
    the oracle, while the other queries keep their tensor-core answer."""
    require_gpu()
    rng = np.random.default_rng(11)
    n, dim, nq, k = 300_000, 256, 32, 100
    corpus = rng.standard_normal((n, dim)).astype(np.float32)
    center = rng.standard_normal(dim).astype(np.float32)
    lo = 123_456
    corpus[lo: lo + 4000] = center + 0.05 * rng.standard_normal((4000, dim)).astype(np.float32)
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    queries[[3, 17, 30]] = center + 0.05 * rng.standard_normal((3, dim)).astype(np.float32)
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False)
    idx.add(corpus)
    _qf, qc = idx.quantize_queries(queries)
    d, i = idx.hamming_topk(qc, k, use_tc=True, check_overflow=True)
    want_d, want_i = oracle.hamming_topk(oracle.quantize_ubinary(corpus), oracle.quantize_ubinary(queries), k)
    assert np.array_equal(d.cpu().numpy(), want_d) and np.array_equal(i.cpu().numpy(), want_i)
    assert idx.last_tc_redone <= 6, idx.last_tc_redone  # the clustered queries, not the whole batch
    # the two-stage entry point redoes the same queries and still returns the oracle's lists
    idx2 = DenseIndex(dim, device=0, store_int8=False, store_f32=True)
    idx2.add(corpus)
    got_i, got_s, got_c = idx2.search_quantized(queries, 10, rescore_multiplier=10.0, prefer_int8=False)
    want = oracle.two_stage_search(queries, oracle.quantize_ubinary(corpus), corpus, 10, 10.0, exact=True)
    for qi, (w_ids, _w_s) in enumerate(want):
        assert got_i[qi, : int(got_c[qi])].cpu().tolist() == w_ids.tolist(), qi


@pytest.mark.parametrize("n,dim,nq,k", [(8000, 1024, 40, 10), (50_000, 256, 130, 100), (4096, 128, 8, 10)])
def test_tc_int8_exact_search(n, dim, nq, k):
    require_gpu()
    rng = np.random.default_rng(n)
    emb = rng.integers(-128, 128, size=(n, dim)).astype(np.int8)
    emb[10] = emb[3]
    qs = rng.integers(-128, 128, size=(nq, dim)).astype(np.int8)
    idx = DenseIndex(dim, device=0, store_int8=True, store_f32=False,
                     int8_ranges=np.stack([-np.ones(dim, np.float32), np.ones(dim, np.float32)]))
    idx._reserve(n)
    idx.int8[:n].copy_(torch.from_numpy(emb))
    idx.tags[:n].fill_(1)
    idx.n = n
    w_ids, w_s = oracle.int8_exact_topk(qs, emb, k)
    for use_tc in (True, False):
        ids, score = idx.search_int8_exact(qs, k, use_tc=use_tc)
        assert np.array_equal(ids.cpu().numpy(), w_ids), use_tc
        assert np.array_equal(score.cpu().numpy(), w_s), use_tc


def test_tc_config2_full_size():
    """1M x 768, 256 queries, k'=200 on device-generated data: tensor-core == POPC path."""
    require_gpu()
    n, dim, q, k = 1_000_000, 768, 256, 200
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False, capacity=n)
    for lo in range(0, n, 250_000):
        idx.add(synth_rows_device(lo, 250_000, dim, seed=1))
    queries = synth_query_rows_device(0, q, dim, seed=1, n_corpus=n)
    _qf, qc = idx.quantize_queries(queries)
    d_tc, i_tc = idx.hamming_topk(qc, k, use_tc=True, check_overflow=False)
    d_pc, i_pc = idx.hamming_topk(qc, k, use_tc=False)
    assert idx.tc_overflow_total() == 0
    assert torch.equal(d_tc, d_pc) and torch.equal(i_tc, i_pc)
