"""GPU parity: R1/R2 quantisers and the synthetic generators, bit-exact vs the oracle."""

import numpy as np
import pytest
import torch

import oracle
from radiant_rag_b200 import _lib, synthetic
from radiant_rag_b200.index import (DenseIndex, _stream, synth_query_rows_device, synth_rows_device)
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def _special_rows(n, dim, seed):
    x = np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)
    x[0, :] = 0.0
    x[1, ::2] = -0.0
    if dim > 3:
        x[2, 3] = np.nan
        x[2, 1] = np.inf
        x[2, 2] = -np.inf
    x[3 % n, :] = np.float32(1e-45)  # denormal > 0
    return x


@pytest.mark.parametrize("dim", [384, 768, 1024, 100, 36, 8, 1000, 1])
def test_ubinary_bit_exact(dim):
    require_gpu()
    x = _special_rows(257, dim, dim)
    idx = DenseIndex(dim, device=0, store_int8=False, store_f32=False)
    _qf, qc = idx.quantize_queries(x)
    got = qc.cpu().numpy()
    want = oracle.quantize_ubinary(x)
    nbytes = want.shape[1]
    assert got.shape == (257, idx.words * 4)
    assert np.array_equal(got[:, :nbytes], want)
    assert not got[:, nbytes:].any()  # padding bits are zero


@pytest.mark.parametrize("dim", [384, 1024, 100, 7])
def test_int8_bit_exact(dim):
    require_gpu()
    from radiant_rag_b200 import quantization as q

    x = np.random.default_rng(dim).standard_normal((300, dim)).astype(np.float32)
    ranges = oracle.calculate_int8_ranges(x)
    assert np.array_equal(q.quantize_embeddings(x, "int8", ranges), oracle.quantize_int8(x, ranges))
    # out-of-range saturates, degenerate range (hi == lo) gives 0 - both as the oracle
    y = x * 3.0
    r2 = ranges.copy()
    r2[1, 0] = r2[0, 0]
    assert np.array_equal(q.quantize_embeddings(y, "int8", r2), oracle.quantize_int8(y, r2))
    assert np.array_equal(q.quantize_embeddings(x, "uint8", ranges).astype(np.int16) - 128,
                          oracle.quantize_int8(x, ranges).astype(np.int16))
    # ranges=None: batch min/max, as sentence-transformers
    assert np.array_equal(q.quantize_embeddings(x, "int8"), oracle.quantize_int8(x, ranges))


def test_module_level_api_like_reference_validation_tool():
    """tools/validate_quantization.py:119-188 of the reference, on the GPU path."""
    require_gpu()
    from radiant_rag_b200 import quantization as q

    assert q.get_binary_dimension(384) == 48
    emb = synthetic.normal_unit_rows(5, 384, seed=3)
    b = q.quantize_embeddings(emb, precision="ubinary")
    assert b.dtype == np.uint8 and b.shape == (5, 48)
    assert np.array_equal(b, oracle.quantize_ubinary(emb))
    sb = q.quantize_embeddings(emb, precision="binary")
    assert sb.dtype == np.int8 and np.array_equal(sb.astype(np.int16) + 128, b.astype(np.int16))
    data = q.embedding_to_bytes(b[0])
    assert np.array_equal(q.bytes_to_embedding(data, np.uint8, (48,)), b[0])
    ranges = q.calculate_int8_ranges(emb)
    i8 = q.quantize_embeddings(emb, precision="int8", ranges=ranges)
    assert i8.dtype == np.int8 and i8.shape == (5, 384)
    res = q.rescore_candidates(emb[0], [i8[j] for j in range(5)], [f"d{j}" for j in range(5)])
    assert len(res) == 5 and all(isinstance(s, float) for _, s in res)
    assert [s for _, s in res] == sorted((s for _, s in res), reverse=True)
    assert q.rescore_candidates(emb[0], [], []) == []


def test_synthetic_device_equals_numpy():
    require_gpu()
    for dim in (96, 768):
        a = synth_rows_device(1000, 77, dim, seed=12).cpu().numpy()
        assert np.array_equal(a, synthetic.hash_rows_f32(1000, 77, dim, seed=12))
        b = synth_query_rows_device(3, 21, dim, seed=12, n_corpus=5000).cpu().numpy()
        assert np.array_equal(b, synthetic.hash_query_rows_f32(3, 21, dim, seed=12, n_corpus=5000))
    n = 500
    lens = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.call("rr_synth_doc_lengths", lens.data_ptr(), 40, n, 9, 200, _stream())
    assert np.array_equal(lens.cpu().numpy(), synthetic.doc_lengths(40, n, 9, 200).astype(np.int32))
    cdf = synthetic.zipf_cdf_u32(1000)
    cdf_d = torch.from_numpy(cdf.view(np.int32)).cuda()
    toks = torch.empty(4000, dtype=torch.int32, device="cuda")
    _lib.call("rr_synth_zipf_tokens", toks.data_ptr(), 123, 4000, 9, cdf_d.data_ptr(), 1000, _stream())
    assert np.array_equal(toks.cpu().numpy(), synthetic.zipf_tokens(123, 4000, 9, cdf))
