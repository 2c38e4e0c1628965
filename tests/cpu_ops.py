"""Test double for the shard compute ops: the oracle on CPU tensors.  Lets the
world_size-2 gloo tests exercise the row-sharding / allgather / merge host logic
of radiant_rag_b200.sharded without a GPU.  NOT part of the product."""

import numpy as np
import torch

import oracle


class CpuShardOps:
    def __init__(self, corpus_f32, rescore_rows, row_base):
        self.corpus = corpus_f32
        self.codes = oracle.quantize_ubinary(corpus_f32) if len(corpus_f32) else np.zeros((0, corpus_f32.shape[1] // 8), np.uint8)
        self.rows = rescore_rows
        self.row_base = row_base

    def quantize_queries(self, queries):
        q = np.asarray(queries, dtype=np.float32)
        return torch.from_numpy(q), torch.from_numpy(oracle.quantize_ubinary(q))

    def hamming_topk(self, qcodes, k, tag_mask=0, tag_value=0, check_overflow=True):
        d, i = oracle.hamming_topk(self.codes, qcodes.numpy(), k)
        i = np.where(i >= 0, i + self.row_base, -1)
        return torch.from_numpy(d), torch.from_numpy(i)

    def merge_hamming(self, dist_all, idx_all, k):
        d, i = dist_all.numpy().astype(np.int64), idx_all.numpy()
        q = d.shape[0]
        out_d = np.full((q, k), np.iinfo(np.int32).max, np.int32)
        out_i = np.full((q, k), -1, np.int64)
        for r in range(q):
            ok = i[r] >= 0
            key = np.sort((d[r][ok] << 40) | i[r][ok])[:k]
            out_d[r, : key.size] = key >> 40
            out_i[r, : key.size] = key & ((1 << 40) - 1)
        return torch.from_numpy(out_d), torch.from_numpy(out_i)

    def score_candidates(self, queries_f32, cand_idx, prefer_int8=True):
        q = queries_f32.numpy()
        c = cand_idx.numpy()
        out = np.full(c.shape, -np.inf, np.float32)
        local = c - self.row_base
        own = (c >= 0) & (local >= 0) & (local < len(self.rows))
        for r in range(c.shape[0]):
            rows = self.rows[local[r][own[r]]].astype(np.float64)
            out[r, own[r]] = (rows @ q[r].astype(np.float64)).astype(np.float32)
        return torch.from_numpy(out)

    def rank_scored(self, scores, cand_idx, top_k, min_similarity):
        s, c = scores.numpy(), cand_idx.numpy()
        q = s.shape[0]
        out_i = np.full((q, top_k), -1, np.int64)
        out_s = np.zeros((q, top_k), np.float32)
        out_c = np.zeros(q, np.int32)
        for r in range(q):
            ok = np.nonzero((c[r] >= 0) & np.isfinite(s[r]))[0]
            order = ok[np.argsort(-s[r][ok].astype(np.float64), kind="stable")][:top_k]
            order = order[s[r][order].astype(np.float64) >= min_similarity]
            out_i[r, : order.size] = c[r][order]
            out_s[r, : order.size] = s[r][order]
            out_c[r] = order.size
        return torch.from_numpy(out_i), torch.from_numpy(out_s), torch.from_numpy(out_c)

    def pack_hamming(self, dist_loc, idx_loc):
        d, i = dist_loc.numpy().astype(np.int64), idx_loc.numpy()
        return torch.from_numpy(np.where(i >= 0, (d << 40) | i, -1))

    def merge_hamming_gathered(self, keys_all, k):
        keys = keys_all.numpy()                      # [G, Q, k']
        g, q, k_in = keys.shape
        out_d = np.full((q, k), 0x7FFFFFFF, np.int32)
        out_i = np.full((q, k), -1, np.int64)
        for r in range(q):
            kk = np.sort(keys[:, r, :].reshape(-1))
            kk = kk[kk >= 0][:k]                     # key order == (dist asc, row asc)
            out_d[r, : kk.size] = (kk >> 40).astype(np.int32)
            out_i[r, : kk.size] = kk & ((1 << 40) - 1)
        return torch.from_numpy(out_d), torch.from_numpy(out_i)

    def search_int8_exact(self, queries_i8, top_k, tag_mask=0, tag_value=0):
        r, sc = oracle.int8_exact_topk(np.asarray(queries_i8), self.rows, top_k)
        r = np.where(r >= 0, r + self.row_base, -1)
        return torch.from_numpy(r), torch.from_numpy(sc)

    def merge_scores_i32(self, score_all, idx_all, k):
        s, i = score_all.numpy(), idx_all.numpy()
        q = s.shape[0]
        out_i = np.full((q, k), -1, np.int64)
        out_s = np.full((q, k), np.iinfo(np.int32).min, np.int32)
        for r in range(q):
            ok = np.nonzero(i[r] >= 0)[0]
            order = ok[np.lexsort((i[r][ok], -s[r][ok].astype(np.int64)))][:k]
            out_i[r, : order.size] = i[r][order]
            out_s[r, : order.size] = s[r][order]
        return torch.from_numpy(out_i), torch.from_numpy(out_s)

    def merge_scores_f64(self, score_all, idx_all, k):
        s, i = score_all.numpy(), idx_all.numpy()
        q = s.shape[0]
        out_i = np.full((q, k), -1, np.int64)
        out_s = np.zeros((q, k), np.float64)
        out_c = np.zeros(q, np.int32)
        for r in range(q):
            ok = np.nonzero(i[r] >= 0)[0]
            order = ok[np.lexsort((i[r][ok], -s[r][ok]))][:k]
            out_i[r, : order.size] = i[r][order]
            out_s[r, : order.size] = s[r][order]
            out_c[r] = order.size
        return torch.from_numpy(out_i), torch.from_numpy(out_s), torch.from_numpy(out_c)


class CpuBm25Shard:
    """Oracle BM25 over the documents of one shard with GLOBAL idf / avgdl."""

    def __init__(self, orc_shard, row_base):
        self.orc = orc_shard
        self.row_base = row_base

    def search_batch(self, q_terms, k, check=True):
        qt = np.asarray(q_terms)
        q = qt.shape[0]
        idx = np.full((q, k), -1, np.int64)
        sc = np.zeros((q, k), np.float64)
        cnt = np.zeros(q, np.int32)
        for r in range(q):
            rows, s = self.orc.search(qt[r].tolist(), k)
            idx[r, : rows.size] = rows + self.row_base
            sc[r, : rows.size] = s
            cnt[r] = rows.size
        return torch.from_numpy(idx), torch.from_numpy(sc), torch.from_numpy(cnt)
