/*
 * radiant_rag_b200.h - C ABI of the B200 retrieval hot path (librr_b200.so).
 *
 * The reference (dshipley71/radiant-rag) is pure Python and has no FFI; the
 * boundary is a set of duck-typed Python interfaces.  Each entry
 * point below replaces the arithmetic of one reference function; the Python
 * wrappers in radiant-rag_b200/ keep the reference's class and method names
 * and call these through ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *     _host; the library never allocates on the hot path (workspace is passed in,
 *     sized by the matching *_workspace_bytes call);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the
 *     call returns without synchronising;
 *   - return 0 on success, a negative rr_status on failure; rr_last_error() gives the
 *     message of the calling thread's last failure.  No C++ exception crosses the ABI;
 *   - re-entrant: the reference calls dense and BM25 retrieval from two threads
 *     (radiant/orchestrator.py:994-998); there is no global mutable state besides the
 *     thread-local error string, the thread-local state of the two timing aids and
 *     per-device attributes set once in rr_init.
 *   - row ids are int64 positions in the index (insertion order); missing result
 *     slots hold row -1.
 */
#ifndef RADIANT_RAG_B200_H
#define RADIANT_RAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_ABI_VERSION 2

typedef enum rr_status {
  RR_OK = 0,
  RR_ERR_INVALID = -1,   /* bad argument (null pointer, unsupported size) */
  RR_ERR_CUDA = -2,      /* CUDA runtime error, text in rr_last_error()   */
  RR_ERR_WORKSPACE = -3, /* workspace too small                           */
  RR_ERR_NO_DEVICE = -4  /* no sm_100 device                              */
} rr_status;

typedef enum rr_dtype { RR_F32 = 0, RR_I8 = 1 } rr_dtype;

#define RR_MAX_WORDS 32 /* packed code width limit: 32 x u32 = 1024 bits   */
#define RR_MAX_K 1024   /* largest top-k any entry point accepts           */

int rr_abi_version(void);
const char* rr_last_error(void);
/* Select the device, check it is sm_100, opt kernels into large shared memory. */
int rr_init(int device);
int rr_sm_count(void);

/* ---- R1: quantize_embeddings(emb, "ubinary")
 * replaces radiant/storage/quantization.py:74-108 -> sentence_transformers
 * np.packbits(emb > 0): bit = (x > 0), dim 8b in the MSB of byte b.
 * codes rows are `code_stride` bytes apart (>= ceil(dim/8), multiple of 16 for the
 * scan); padding bytes are written as zero. */
int rr_quantize_ubinary(const float* emb, int64_t n, int32_t dim, uint8_t* codes,
                        int32_t code_stride, void* stream);

/* ---- R2: quantize_embeddings(emb, "int8", ranges)
 * replaces the same call site; ranges f32 [2, dim] (min row, max row) as produced by
 * calculate_int8_ranges (quantization.py:159-182).  float32 (x-lo)/((hi-lo)/255)-128,
 * truncation toward zero, saturated to [-128,127]. */
int rr_quantize_int8(const float* emb, int64_t n, int32_t dim, const float* ranges,
                     int8_t* out, void* stream);

/* ---- R4 + R12: stage-1 candidate search of retrieve_by_embedding_quantized
 * (radiant/storage/redis_store.py:799-809, chroma_store.py:588-619,
 * pgvector_store.py:794-802; documented as Hamming in
 * docs/BINARY_QUANTIZATION_README.md:84-100).
 * codes  u32 [n, words]   packed sign bits (words multiple of 4, <= RR_MAX_WORDS)
 * tags   u8  [n] or NULL  row passes when (tags[row] & tag_mask) == tag_value
 * qcodes u32 [q, words]
 * out_dist i32 [q, k], out_idx i64 [q, k]: exact top-k by (dist asc, row asc);
 * out_idx = row + row_base (shard offset); missing slots (INT32_MAX, -1). */
size_t rr_hamming_topk_workspace_bytes(int64_t n, int32_t words, int32_t q, int32_t k);
int rr_hamming_topk(const uint32_t* codes, int64_t n, int32_t words, const uint8_t* tags,
                    uint8_t tag_mask, uint8_t tag_value, const uint32_t* qcodes, int32_t q,
                    int32_t k, int64_t row_base, int32_t* out_dist, int64_t* out_idx,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- Tensor-core (tcgen05 kind::i8, TMEM accumulators, TMA-fed) formulation of the BATCHED
 * stage-1 search.  The packed codes are read as they are (1 bit per dimension from HBM) and
 * expanded on chip to a 0/255 unsigned-int8 operand in tensor memory; with +-1 int8 queries
 * hamming = popc(q) - dot / 255 exactly.
 * Same arguments and bit-identical results as rr_hamming_topk, except:
 *   q_pm1     i8 [q, 32*words]  the query codes expanded by rr_unpack_codes_pm1(qcodes, q,
 *             4*words, 32*words, ...);
 *   *overflow device u32, zeroed by the caller; incremented when a query's bounded candidate
 *             list overflowed (the results of that query are then not exact);
 *   overflow_flags device u8 [q] or NULL: 1 for every query whose list overflowed - the caller
 *             redoes exactly those queries with rr_hamming_topk;
 *   workspace from rr_tc_search_workspace_bytes. */
int rr_unpack_codes_pm1(const uint8_t* codes, int64_t n, int32_t code_stride, int32_t dim,
                        int8_t* out, void* stream);
size_t rr_tc_search_workspace_bytes(int64_t n, int32_t q, int32_t k);
int rr_hamming_topk_tc(const uint32_t* codes, int64_t n, int32_t words, const uint8_t* tags,
                       uint8_t tag_mask, uint8_t tag_value, const int8_t* q_pm1, int32_t q,
                       int32_t k, int64_t row_base, int32_t* out_dist, int64_t* out_idx,
                       uint32_t* overflow, uint8_t* overflow_flags, void* workspace,
                       size_t workspace_bytes, void* stream);
/* BASELINE config 4 on tensor cores: exact int8 x int8 -> int32 search, (score desc, row asc);
 * same contract as rr_int8_search_topk, workspace from rr_tc_search_workspace_bytes. */
int rr_int8_search_topk_tc(const int8_t* emb, int64_t n, int32_t dim, const uint8_t* tags,
                           uint8_t tag_mask, uint8_t tag_value, const int8_t* queries_i8,
                           int32_t q, int32_t top_k, int64_t row_base, int32_t* out_score,
                           int64_t* out_idx, uint32_t* overflow, uint8_t* overflow_flags,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Test/debug: raw tensor-core scores as order-preserving keys, u32 [q][ceil(n/128)*128]
 * (key = ~(score ^ 0x80000000), 0xFFFFFFFF = padded row). */
int rr_tc_dense_keys(const int8_t* emb, int64_t n, int32_t dim, const int8_t* queries_i8, int32_t q,
                     uint32_t* out_keys, void* stream);
/* Measurement aid (bench.py roofline): when enabled, rr_hamming_topk_tc / rr_int8_search_topk_tc
 * bracket each of their four kernels with CUDA events on the launching stream;
 * rr_tc_last_timing_ms waits for the last call and returns the durations in ms:
 * out_ms[0] sample pass, [1] tau, [2] filter pass (the tcgen05 kernel), [3] list select.
 * Not for use inside stream capture. */
int rr_tc_timing(int32_t enable);
int rr_tc_last_timing_ms(float* out_ms);

/* ---- R3: rescore_candidates (radiant/storage/quantization.py:185-222) plus the
 * caller's cut/filter (radiant/storage/redis_store.py:850-854).
 * queries f32 [q, dim]; emb rows [n, dim] of emb_dtype (RR_I8 rows are cast to f32,
 * exactly as `.astype(np.float32)` does); cand_idx i64 [q, c] in stage-1 order,
 * entries < 0 or outside [row_base, row_base + n) are skipped.
 * score = dot in float64 rounded once to float32; order (score desc, stage-1 order
 * asc); the first top_k are kept, then those with score >= min_similarity.
 * out_score f32 [q, top_k], out_idx i64 [q, top_k] (-1 padded), out_count i32 [q]. */
int rr_rescore_f32(const float* queries, int32_t q, int32_t dim, const void* emb,
                   int32_t emb_dtype, int64_t n, int64_t row_base, const int64_t* cand_idx,
                   int32_t c, int32_t top_k, double min_similarity, float* out_score,
                   int64_t* out_idx, int32_t* out_count, void* stream);

/* Score-only variant used by the row-sharded path: out_score f32 [q, c] holds the
 * score of each candidate this shard owns and -inf for the others (no sort). */
int rr_score_candidates_f32(const float* queries, int32_t q, int32_t dim, const void* emb,
                            int32_t emb_dtype, int64_t n, int64_t row_base,
                            const int64_t* cand_idx, int32_t c, float* out_score, void* stream);

/* Order already-scored candidates: (score desc, position asc), cut, filter. */
int rr_rank_scored_f32(const float* scores, const int64_t* cand_idx, int32_t q, int32_t c,
                       int32_t top_k, double min_similarity, float* out_score,
                       int64_t* out_idx, int32_t* out_count, void* stream);

/* ---- north-star extension: symmetric int8 x int8 -> int32 rescoring (bit-exact).
 * Not in the reference.  Order (score desc, stage-1 order asc). */
int rr_rescore_i8(const int8_t* queries_i8, int32_t q, int32_t dim, const int8_t* emb,
                  int64_t n, int64_t row_base, const int64_t* cand_idx, int32_t c,
                  int32_t top_k, int32_t* out_score, int64_t* out_idx, int32_t* out_count,
                  void* stream);

/* ---- R5: exact float32 cosine scan, RedisVectorStore._retrieve_by_embedding_linear
 * (radiant/storage/redis_store.py:863-952): normalise query and row, dot, keep
 * >= min_similarity, (score desc, row asc), top_k.  Zero-norm rows are skipped. */
size_t rr_exact_search_f32_workspace_bytes(int64_t n, int32_t q, int32_t k);
int rr_exact_search_f32(const float* emb, int64_t n, int32_t dim, const uint8_t* tags,
                        uint8_t tag_mask, uint8_t tag_value, const float* queries, int32_t q,
                        int32_t top_k, double min_similarity, int64_t row_base,
                        float* out_score, int64_t* out_idx, int32_t* out_count,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- R5 for BATCHES (csrc/tc_exact.cu): the results of rr_exact_search_f32, bit for bit, with the
 * row x query products on the tensor cores.  TF32 cosines (tcgen05.mma kind::tf32 straight from the
 * float32 rows, read from HBM once per 128 queries) against a sampled bound minus a rigorous error
 * margin keep a few hundred rows per query; their exact scores are then recomputed in
 * rr_exact_search_f32's own float64 operation order and selected (score desc, row asc).
 * row_inv_norm f32 [n] = 1 / |row| (0 for zero rows) from rr_row_inv_norms_f32; it only steers the
 * filter, the exact norms are recomputed for the rows that are kept.
 * Needs dim % 32 == 0, 16-byte aligned rows and queries, 128 <= n < 2^31 and a corpus large enough
 * for top_k (rr_exact_search_f32_tc_supported).  A query whose candidate list outgrew its
 * workspace segment (clustered duplicates, selective tag filters) raises *overflow (accumulated,
 * never reset here) and overflow_flags[query] (u8 [q], may be NULL; written for every query): its
 * rows in the outputs are not guaranteed and the caller redoes it with rr_exact_search_f32. */
int rr_row_inv_norms_f32(const float* emb, int64_t n, int32_t dim, float* out, void* stream);
int rr_exact_search_f32_tc_supported(int64_t n, int32_t dim, int32_t q, int32_t top_k);
size_t rr_exact_search_f32_tc_workspace_bytes(int64_t n, int32_t q, int32_t top_k);
int rr_exact_search_f32_tc(const float* emb, const float* row_inv_norm, int64_t n, int32_t dim,
                           const uint8_t* tags, uint8_t tag_mask, uint8_t tag_value,
                           const float* queries, int32_t q, int32_t top_k, double min_similarity,
                           int64_t row_base, float* out_score, int64_t* out_idx, int32_t* out_count,
                           uint32_t* overflow, uint8_t* overflow_flags, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- BASELINE config 4: exact int8 x int8 -> int32 search, (score desc, row asc). */
size_t rr_int8_search_topk_workspace_bytes(int64_t n, int32_t q, int32_t k);
int rr_int8_search_topk(const int8_t* emb, int64_t n, int32_t dim, const uint8_t* tags,
                        uint8_t tag_mask, uint8_t tag_value, const int8_t* queries_i8,
                        int32_t q, int32_t top_k, int64_t row_base, int32_t* out_score,
                        int64_t* out_idx, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- R6: BM25Index.search (radiant/storage/bm25_index.py:218-270).
 * The index is a tile-sharded inverted CSR built by rr_bm25_impacts + host code:
 *   tile t owns rows [t*tile_docs, (t+1)*tile_docs);
 *   tile_term_ptr i64 [n_tiles, n_terms+1] offsets into post_row / post_impact;
 *   post_row u32 (row inside the shard), post_impact f64 = idf_t * (tf*(k1+1)) /
 *   (tf + k1*((1-b) + (b*len)/avgdl)) evaluated in the reference's operation order.
 * q_terms i32 [q, q_len] term ids in query-token order (-1 = unknown/padding).
 * Scores accumulate in float64 in token order, repeats included; results
 * (score desc, row asc), score > 0 only.
 * out_score f64 [q, k], out_idx i64 [q, k] (-1 padded), out_count i32 [q]. */
size_t rr_bm25_topk_workspace_bytes(int32_t n_tiles, int32_t q, int32_t k);
int rr_bm25_topk(const int64_t* tile_term_ptr, const uint32_t* post_row,
                 const double* post_impact, int32_t n_tiles, int32_t tile_docs,
                 int32_t n_terms, int64_t n_docs, const int32_t* q_terms, int32_t q,
                 int32_t q_len, int32_t k, int64_t row_base, double* out_score,
                 int64_t* out_idx, int32_t* out_count, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ---- R6 for BATCHES: the same results as rr_bm25_topk, bit for bit, by filter-and-refine with
 * the postings shared across the batch (csrc/bm25_fast.cu): float32 scores of every (query,
 * document) against a sampled bound keep a few hundred candidates per query, whose float64
 * scores are then recomputed in the reference's operation order.  Needs, on top of the
 * rr_bm25_topk index: rows ASCENDING inside every (tile, term) segment, tile_docs a multiple of
 * 128 and <= 1024, every impact finite and in [2^-100, 2^100], and for the n_head <= 64 terms
 * chosen as "head" terms
 *   post_pack u64 [P]                        per posting, in post_row order: row-in-tile << 32 |
 *                                            round(impact * 2^fx_shift) as u32; fx_shift chosen so
 *                                            that 64 * max impact * 2^fx_shift <= 2^30
 *   head_max  f32 [n_head]                   upper bound of the term's FLOAT16-rounded impact over all docs
 *   head_slot i32 [n_terms]                  slot of the term, -1 for all other terms
 *   head_imp  f64 [n_tiles, n_head, tile_docs] the term's impact per document of the tile,
 *                                            0.0 where the document lacks the term (read by the refine)
 *   head_imp_f16 f16 [n_tiles, n_head, tile_docs] the same values rounded to float16 (every value
 *                                            finite, i.e. < 65504; 16-byte aligned): what the filter
 *                                            stages into shared memory
 * (head terms keep their sparse postings too).
 * inexact_flags u8 [q] or NULL: 1 for a query whose exactness check failed (candidate list
 * overflow, or the k-th candidate score does not clear the bound) - its output row is NOT
 * guaranteed and the caller must redo it with rr_bm25_topk; *inexact_counter (device u32 or
 * NULL) is incremented once per flagged query. */
size_t rr_bm25_fast_workspace_bytes(int32_t n_tiles, int32_t tile_docs, int64_t n_docs, int32_t q,
                                    int32_t k);
/* Largest n_head the filter kernel's shared memory holds for this tile size (-1: tile size not
 * supported by rr_bm25_topk_fast). */
int rr_bm25_fast_max_head(int32_t tile_docs);
int rr_bm25_topk_fast(const int64_t* tile_term_ptr, const uint32_t* post_row,
                      const double* post_impact, const uint64_t* post_pack, int32_t fx_shift,
                      const int32_t* head_slot, const double* head_imp, const void* head_imp_f16,
                      const float* head_max, int32_t n_head, int32_t n_tiles, int32_t tile_docs, int32_t n_terms,
                      int64_t n_docs, const int32_t* q_terms, int32_t q, int32_t q_len, int32_t k,
                      int64_t row_base, double* out_score, int64_t* out_idx, int32_t* out_count,
                      uint8_t* inexact_flags, uint32_t* inexact_counter, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Measurement aid, as rr_tc_timing: out_ms[0] sample pass, [1] tau, [2] filter pass, [3] refine.
 * The timing state of both aids is per host thread.  Not for use inside stream capture. */
int rr_bm25_timing(int32_t enable);
int rr_bm25_last_timing_ms(float* out_ms);

/* R7 helper: per-posting impact in the reference's float64 operation order
 * (bm25_index.py:252-255).  post_tf i32, post_len i32 (length of the posting's
 * document), post_idf f64 (idf of the posting's term, copied from the host index). */
int rr_bm25_impacts(const int32_t* post_tf, const int32_t* post_len, const double* post_idf,
                    int64_t n_post, double k1, double b, double avgdl, double* post_impact,
                    void* stream);

/* ---- R10: RRFAgent._execute (radiant/agents/fusion.py:61-102).
 * run_idx i64 [q, total_len] (device): the runs of one query concatenated in run order;
 * run_off_host i32 [n_runs+1] (HOST pointer) segment bounds shared by all queries,
 * run_off_host[0] == 0, total_len = run_off_host[n_runs] <= 4096, n_runs <= 16.
 * A run shorter than its segment is padded with -1 at its tail.  Doc ids < 2^32.
 * score[id] accumulates 1.0/(rrf_k + rank) in float64 in run order (rank = 1-based
 * position in the run); output order (score desc, first-insertion position asc);
 * out_idx i64 [q, k] (-1 padded), out_score f64 [q, k], out_count i32 [q] or NULL. */
int rr_rrf_fuse(const int64_t* run_idx, const int32_t* run_off_host, int32_t n_runs, int32_t q,
                double rrf_k, int32_t k, int64_t* out_idx, double* out_score,
                int32_t* out_count, void* stream);
/* Same fusion with every run in its own buffer (no concatenation copy): run r of query i is
 * run_ptrs_host[r] + i * run_len_host[r], i64 [q, run_len_host[r]] on the device; the two arrays
 * of n_runs entries are HOST arrays.  This is what the batched hybrid step uses: the dense and
 * BM25 top-k lists are fused where their kernels left them. */
int rr_rrf_fuse_runs(const int64_t* const* run_ptrs_host, const int32_t* run_len_host,
                     int32_t n_runs, int32_t q, double rrf_k, int32_t k, int64_t* out_idx,
                     double* out_score, int32_t* out_count, void* stream);

/* ---- merge of per-shard candidate lists after the NCCL allgather.
 * Inputs are [q, n_in] lists (n_in = shards * k); idx < 0 marks padding; idx < 2^32
 * (2^40 for the Hamming merge).  Outputs [q, k], padded with idx = -1. */
/* (dist asc, idx asc) */
int rr_merge_hamming(const int32_t* in_dist, const int64_t* in_idx, int32_t q, int32_t n_in,
                     int32_t k, int32_t* out_dist, int64_t* out_idx, void* stream);
/* One-collective form of the Hamming exchange: rr_pack_hamming turns a shard's (dist, global row)
 * lists into int64 keys (dist << 40 | row, -1 = padding) so that ONE all_gather moves them;
 * rr_merge_hamming_gathered merges the gathered buffer in its native layout
 * in_keys i64 [n_shards][q][k_in] (no transpose) by (dist asc, row asc) -> [q, k].  Every shard's k_in
 * keys must be in ASCENDING order with the padding last - the order rr_hamming_topk / rr_hamming_topk_tc
 * return and rr_pack_hamming keeps (the merge is a tree of sorted-list merges). */
int rr_pack_hamming(const int32_t* dist, const int64_t* idx, int64_t n, int64_t* out_keys,
                    void* stream);
int rr_merge_hamming_gathered(const int64_t* in_keys, int32_t n_shards, int32_t q, int32_t k_in,
                              int32_t k, int32_t* out_dist, int64_t* out_idx, void* stream);
/* (score desc, idx asc), float64 scores (BM25) */
int rr_merge_scores_f64(const double* in_score, const int64_t* in_idx, int32_t q,
                        int32_t n_in, int32_t k, double* out_score, int64_t* out_idx,
                        int32_t* out_count, void* stream);
/* One-collective form of the BM25 exchange: every shard writes its lists into ONE buffer of 8-byte
 * words [2][q][k_in] (plane 0 = float64 scores, plane 1 = global rows, -1 = padding), ONE all_gather
 * moves it and this merges the gathered [n_shards][2][q][k_in] buffer where it lies.  Every shard's list must
 * be in (score desc, row asc) order with the padding last, as rr_bm25_topk / rr_bm25_topk_fast return it. */
int rr_merge_scores_f64_gathered(const int64_t* in_words, int32_t n_shards, int32_t q, int32_t k_in,
                                 int32_t k, double* out_score, int64_t* out_idx, int32_t* out_count,
                                 void* stream);
/* (score desc, idx asc), int32 scores (config 4) */
int rr_merge_scores_i32(const int32_t* in_score, const int64_t* in_idx, int32_t q,
                        int32_t n_in, int32_t k, int32_t* out_score, int64_t* out_idx,
                        void* stream);

/* ---- synthetic data (bench / parity only; bit-identical to synthetic.py).
 * value = irwin_hall4(splitmix64(row * dim + col, seed)) * 2^-shift. */
int rr_synth_rows_f32(float* out, int64_t row_start, int64_t n_rows, int32_t dim,
                      uint64_t seed, int32_t shift, void* stream);
int rr_synth_query_rows_f32(float* out, int64_t q_start, int64_t n_q, int32_t dim,
                            uint64_t seed, int64_t n_corpus, int32_t shift, void* stream);
int rr_synth_doc_lengths(int32_t* out, int64_t row_start, int64_t n, uint64_t seed,
                         int32_t mean_len, void* stream);
int rr_synth_zipf_tokens(int32_t* out, int64_t pos_start, int64_t n, uint64_t seed,
                         const uint32_t* cdf, int32_t n_terms, void* stream);

/* ---- peak probes (measurement aids for bench.py / tools/peak_probe.py, not on the product path).
 * Each call times its own kernel with CUDA events on `stream` (best of 5) and synchronises.
 *   rr_probe_popc    32-bit POPC instructions per second over a full grid;
 *   rr_probe_smem    bytes per second of 128-bit shared-memory loads over a full grid;
 *   rr_probe_gather  bytes per second of a random-row gather (see below);
 *   rr_probe_i8_mma  int8 operations per second (2 per MAC) of back-to-back
 *                    tcgen05.mma.kind::i8 from resident operands, one CTA per SM:
 *                    mode 0 = M128 N256 K32, A and B in shared memory; 1 = M128 N128, both in shared
 *                    memory; 2 = M128 N128, A in tensor memory (the batched Hamming scan's form). */
int rr_probe_popc(int32_t iters, double* out_popc32_per_s_host, void* stream);
/* rr_probe_gather (csrc/rescore.cu): the candidate gather of rr_score_candidates_f32 with the
 * arithmetic left out - bytes per second the memory system delivers for q * c random rows of this
 * index through the same bulk-copy ring (best of 5).  scratch_scores f32 [q, c] is overwritten. */
int rr_probe_gather(const void* emb, int32_t emb_dtype, int64_t n, int32_t dim, const int64_t* cand_idx,
                    int32_t q, int32_t c, float* scratch_scores, double* out_bytes_per_s_host, void* stream);
int rr_probe_smem(int32_t iters, double* out_bytes_per_s_host, void* stream);
int rr_probe_i8_mma(int32_t mode, int32_t iters, double* out_ops_per_s_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADIANT_RAG_B200_H */
