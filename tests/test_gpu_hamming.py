"""GPU parity: R4 exact Hamming top-k - distances, ids and tie order bit-exact vs the oracle."""

import numpy as np
import pytest
import torch

import oracle
from radiant_rag_b200 import synthetic
from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
from tests.gpu_util import build_index, require_gpu

pytestmark = pytest.mark.gpu


def _check(corpus, queries, k, tags=None, mask=0, value=0, row_base=0):
    idx, _ = build_index(corpus, int8=False, f32=False, tags=tags, row_base=row_base)
    _qf, qc = idx.quantize_queries(queries)
    dist, rows = idx.hamming_topk(qc, k, mask, value)
    torch.cuda.synchronize()
    valid = None if tags is None or mask == 0 else ((tags & mask) == value)
    want_d, want_i = oracle.hamming_topk(oracle.quantize_ubinary(corpus), oracle.quantize_ubinary(queries), k, valid)
    want_i = np.where(want_i >= 0, want_i + row_base, -1)
    got_d, got_i = dist.cpu().numpy(), rows.cpu().numpy()
    bad = np.nonzero((got_d != want_d).any(1) | (got_i != want_i).any(1))[0]
    assert bad.size == 0, (f"queries {bad[:5]} differ", got_d[bad[0]][:12], want_d[bad[0]][:12],
                           got_i[bad[0]][:12], want_i[bad[0]][:12])


@pytest.mark.parametrize("n,dim,q,k", [
    (10_000, 384, 64, 40),      # BASELINE config 1
    (5_000, 768, 33, 200),      # config-2 width and k', ragged query tile
    (70_000, 1024, 5, 40),      # config-5 width
    (3_000, 32, 7, 10),         # 32-bit codes: massive ties
    (100, 128, 3, 200),         # k > n
    (1, 128, 2, 5),
    (300_000, 768, 1, 1000),    # single query, largest k
    (40_000, 100, 9, 17),       # dim not a multiple of 32
    (9_000, 640, 130, 64),      # several query tiles (words = 20)
])
def test_hamming_topk_matches_oracle(n, dim, q, k):
    require_gpu()
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=n + dim)
    queries = synthetic.hash_query_rows_f32(0, q, dim, seed=n + dim, n_corpus=n)
    _check(corpus, queries, k)


def test_empty_index_and_all_ties():
    require_gpu()
    idx = DenseIndex(128, device=0, store_int8=False, store_f32=False)
    _qf, qc = idx.quantize_queries(np.ones((3, 128), np.float32))
    d, i = idx.hamming_topk(qc, 7)
    assert (i.cpu().numpy() == -1).all() and (d.cpu().numpy() == np.iinfo(np.int32).max).all()
    corpus = np.tile(np.linspace(-1, 1, 256, dtype=np.float32), (5000, 1))  # every row identical
    _check(corpus, corpus[:2] * -1.0, 40)
    _check(corpus, corpus[:2], 40)


def test_filters_and_row_base():
    require_gpu()
    n, dim = 20_000, 384
    corpus = synthetic.hash_rows_f32(0, n, dim, seed=1)
    queries = synthetic.hash_query_rows_f32(0, 20, dim, seed=1, n_corpus=n)
    tags = np.where(np.arange(n) % 3 == 0, 2, 1).astype(np.uint8) | (np.arange(n) % 5 << 2).astype(np.uint8)
    _check(corpus, queries, 40, tags, 0x03, 1)            # doc_level == child
    _check(corpus, queries, 40, tags, 0x03, 2)            # parent
    _check(corpus, queries, 40, tags, 0xFF, 2 | (3 << 2))  # parent AND language 3
    _check(corpus, queries, 40, tags, 0xFC, 63 << 2)      # unknown language: nothing passes
    _check(corpus, queries, 25, None, 0, 0, row_base=1_000_000_000)


def test_config1_golden_candidates(golden_dir):
    require_gpu()
    z = np.load(golden_dir / "redis_flow.npz")
    corpus = synthetic.normal_unit_rows(10_000, 384, seed=0)
    queries = synthetic.normal_unit_rows(64, 384, seed=1000)
    parent = z["levels_parent"]
    tags = np.where(parent, 2, 1).astype(np.uint8)
    idx, _ = build_index(corpus, int8=False, f32=False, tags=tags)
    _qf, qc = idx.quantize_queries(queries)
    for tag, (m, v) in {"all": (0, 0), "child": (3, 1)}.items():
        d, i = idx.hamming_topk(qc, 40, m, v)
        assert np.array_equal(i.cpu().numpy(), z[f"flow_{tag}_cand"]), tag
        assert np.array_equal(d.cpu().numpy(), z[f"flow_{tag}_dist"]), tag


def test_config2_full_size_properties():
    """1M x 768, 256 queries, k'=200 (BASELINE config 2) generated on device.  Size-independent
    properties for every query + a full oracle comparison for a sample of queries."""
    require_gpu()
    n, dim, q, k = 1_000_000, 768, 256, 200
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False, capacity=n)
    step = 250_000
    for lo in range(0, n, step):
        idx.add(synth_rows_device(lo, step, dim, seed=1))
    queries = synth_query_rows_device(0, q, dim, seed=1, n_corpus=n)
    _qf, qc = idx.quantize_queries(queries)
    dist, rows = idx.hamming_topk(qc, k)
    torch.cuda.synchronize()
    d, r = dist.cpu().numpy().astype(np.int64), rows.cpu().numpy()
    assert (r >= 0).all() and (r < n).all()
    key = (d << 32) | r
    assert (np.diff(key, axis=1) > 0).all(), "lists must be strictly increasing in (dist, row)"
    codes = idx.codes[:n].cpu().numpy()[:, : dim // 8]
    qcodes = qc.cpu().numpy()[:, : dim // 8]
    for qi in range(q):  # returned distances are the true distances of the returned rows
        true = oracle.hamming_distances(codes[r[qi]], qcodes[qi])
        assert np.array_equal(true, d[qi]), qi
    for qi in list(range(0, q, 37)) + [1, q - 1]:  # exactness: nothing better was left out
        full = oracle.hamming_distances(codes, qcodes[qi]).astype(np.int64)
        fkey = np.sort((full << 32) | np.arange(n))[:k]
        assert np.array_equal(fkey, key[qi]), qi
    # odd queries copy 3/4 of a corpus row: that row must be the nearest neighbour
    src = synthetic.query_source_row(np.arange(q), n, 1)
    assert (r[1::2, 0] == src[1::2]).mean() > 0.95
