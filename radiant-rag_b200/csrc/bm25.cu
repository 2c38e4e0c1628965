// R6 / R7: BM25 scoring and top-k over a tile-sharded inverted CSR.
//
// Replaces BM25Index.search (reference radiant/storage/bm25_index.py:218-270), which
// walks every document with list.count per query token.  Arithmetic kept bit-exact:
//   impact(t, d) = idf_t * ((tf*(k1+1)) / (tf + k1*((1-b) + (b*len_d)/avgdl)))   [float64,
//                  no FMA, the reference's operation order, bm25_index.py:252-255]
//   score[d]    += impact(t, d)   for each query token t in query order, repeats included.
// impact is query-independent, so rr_bm25_impacts evaluates it once per posting at
// index-build time and the query kernel only adds.
//
// Layout in HBM: the document space is cut into tiles of `tile_docs` rows; postings are
// stored tile-major (tile, term, row) as SoA post_row u32 / post_impact f64 with a
// [n_tiles, n_terms+1] offset table.  A CTA owns one (query, tile): float64
// accumulators for the tile live in shared memory, the CTA walks the tile's segment of
// each query term in token order (a document occurs at most once per term, so the adds
// of one term never collide and the order of additions per document is the token
// order, as in the reference), then block_select_sorted takes the tile's top-k by
// (score desc, row asc).  CTAs are ordered query-fastest so all queries of a batch hit
// one tile's postings while they are L2-resident: HBM traffic ~ one pass over the
// index per batch, the rest is L2 -> SM traffic.
// Algorithmic bytes per query: sum over query tokens of df(t) * 12 B.
#include "common.cuh"
#include "merge.cuh"

namespace rr {

#ifndef RR_BM_THREADS
#define RR_BM_THREADS 256
#endif
#ifndef RR_BM_UNROLL
#define RR_BM_UNROLL 4
#endif
constexpr int BM_THREADS = RR_BM_THREADS;
constexpr int BM_UNROLL = RR_BM_UNROLL;
constexpr int BM_TERMS = 32;  // query tokens whose segment bounds are fetched together

struct Bm25Args {
  const long long* tile_term_ptr;
  const u32* post_row;
  const double* post_impact;
  int n_tiles;
  int tile_docs;
  int n_terms;
  long long n_docs;
  const int* q_terms;
  int q;
  int q_len;
  int k;
  int cap;
  u64* part_k1;  // [q][n_tiles][k]
  u32* part_k2;
  u64* qbound;   // [q] best tile-local k-th key seen so far (K1_INVALID = none yet)
};

__global__ void __launch_bounds__(BM_THREADS) bm25_tile_kernel(const Bm25Args a) {
  extern __shared__ __align__(16) unsigned char bm_smem[];
  double* acc = reinterpret_cast<double*>(bm_smem);             // [tile_docs]
  u64* s_k1 = reinterpret_cast<u64*>(acc + a.tile_docs);        // [cap]
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + a.cap);             // [cap]
  __shared__ SelectScratch<BM_THREADS> sc;
  const int tile = blockIdx.x / a.q;
  const int qi = blockIdx.x % a.q;
  const long long tile_lo = (long long)tile * a.tile_docs;
  const int rows_here = (int)min((long long)a.tile_docs, a.n_docs - tile_lo);

  for (int i = threadIdx.x; i < a.tile_docs; i += BM_THREADS) acc[i] = 0.0;
  __syncthreads();

  const long long* ptr = a.tile_term_ptr + (size_t)tile * (a.n_terms + 1);
  const int* terms = a.q_terms + (size_t)qi * a.q_len;
  // The walk is a chain of dependent global loads (term id -> segment bounds -> postings ->
  // accumulate -> barrier), and a (query, tile) CTA has only a handful of short segments, so
  // the chain latency, not bandwidth, is what a CTA waits on.  Two measures: the segment
  // bounds of up to BM_TERMS query tokens are fetched by that many threads at once, and the
  // postings are software-pipelined one batch ahead across segment boundaries (loads of the
  // next batch are in flight while the current one is added; only the adds are ordered).
  __shared__ long long s_lo[BM_TERMS], s_hi[BM_TERMS];
  constexpr int BATCH = BM_UNROLL * BM_THREADS;
  for (int j0 = 0; j0 < a.q_len; j0 += BM_TERMS) {
    const int nt = min(BM_TERMS, a.q_len - j0);
    __syncthreads();  // previous block of terms fully accumulated; s_lo/s_hi reusable
    if (threadIdx.x < nt) {
      const int t = terms[j0 + threadIdx.x];
      long long lo = 0, hi = 0;
      if (t >= 0 && t < a.n_terms) {  // unknown token: skipped (bm25_index.py:238-239)
        lo = __ldg(ptr + t);
        hi = __ldg(ptr + t + 1);
      }
      s_lo[threadIdx.x] = lo;
      s_hi[threadIdx.x] = hi;
    }
    __syncthreads();
    // cursor = (term j, first posting of the batch); CTA-uniform
    int j = 0;
    long long base = s_lo[0];
    while (j < nt && base >= s_hi[j]) {
      ++j;
      if (j < nt) base = s_lo[j];
    }
    u32 row_c[BM_UNROLL], row_n[BM_UNROLL];
    double imp_c[BM_UNROLL], imp_n[BM_UNROLL];
    auto load = [&](int jj, long long bb, u32* r, double* im) {
      const long long hi = s_hi[jj];
#pragma unroll
      for (int u = 0; u < BM_UNROLL; ++u) {
        const long long pu = bb + threadIdx.x + (long long)u * BM_THREADS;
        const bool ok = pu < hi;
        r[u] = ok ? __ldg(a.post_row + pu) : 0xFFFFFFFFu;
        im[u] = ok ? __ldg(a.post_impact + pu) : 0.0;
      }
    };
    if (j < nt) load(j, base, row_c, imp_c);
    while (j < nt) {
      int jn = j;
      long long bn = base + BATCH;
      while (jn < nt && bn >= s_hi[jn]) {
        ++jn;
        if (jn < nt) bn = s_lo[jn];
      }
      if (jn < nt) load(jn, bn, row_n, imp_n);
      // a document occurs at most once in a term's segment: adds inside a term never collide
#pragma unroll
      for (int u = 0; u < BM_UNROLL; ++u) {
        if (row_c[u] != 0xFFFFFFFFu) {
          const int r = (int)((long long)row_c[u] - tile_lo);
          acc[r] = __dadd_rn(acc[r], imp_c[u]);
        }
      }
      if (jn != j) __syncthreads();  // next batch belongs to a later query token
#pragma unroll
      for (int u = 0; u < BM_UNROLL; ++u) {
        row_c[u] = row_n[u];
        imp_c[u] = imp_n[u];
      }
      j = jn;
      base = bn;
    }
  }
  __syncthreads();

  auto get = [&](long long i, u64& x, u32& y) {
    const double s = acc[i];
    x = (s > 0.0) ? ~f64_orderable(s) : K1_INVALID;  // score <= 0 dropped (bm25_index.py:267)
    y = (u32)(tile_lo + i);
  };
  // Cross-tile bound: once any tile of this query has k positive scores, its k-th key is
  // an upper bound on the query's final k-th key, and a document whose key is strictly
  // larger can never reach the merged top-k.  Tiles are scheduled query-fastest, so for
  // batches wider than one wave of CTAs every tile but the first sees a bound, keeps
  // ~k candidates and finishes with one pass and a small sort instead of a radix select.
  __shared__ u64 s_bound;
  if (threadIdx.x == 0) s_bound = *reinterpret_cast<volatile u64*>(a.qbound + qi);
  __syncthreads();
  const u64 bound = s_bound;
  int m = -1;
  bool sorted = true;
  if (bound != K1_INVALID) {
    if (threadIdx.x == 0) sc.count = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < rows_here; i += BM_THREADS) {
      u64 x;
      u32 y;
      get(i, x, y);
      if (x <= bound) {
        const int slot = atomicAdd(&sc.count, 1);
        if (slot < a.cap) {
          s_k1[slot] = x;
          s_k2[slot] = y;
        }
      }
    }
    __syncthreads();
    const int cnt = sc.count;
    if (cnt <= a.k) {
      // all of them go to the merge, which does not need its input sorted
      m = cnt;
      sorted = false;
    } else if (cnt <= a.cap) {
      int p = 1;
      while (p < cnt) p <<= 1;
      for (int i = cnt + threadIdx.x; i < p; i += BM_THREADS) {
        s_k1[i] = K1_INVALID;
        s_k2[i] = K2_INVALID;
      }
      __syncthreads();
      block_bitonic_sort_pairs<BM_THREADS>(s_k1, s_k2, p);
      m = a.k;
    }
    __syncthreads();
  }
  if (m < 0) m = block_select_sorted<BM_THREADS, true>(get, rows_here, a.k, s_k1, s_k2, a.cap, sc);
  if (sorted && m == a.k && threadIdx.x == 0)
    atomicMin(reinterpret_cast<unsigned long long*>(a.qbound + qi), (unsigned long long)s_k1[a.k - 1]);
  const size_t o = ((size_t)qi * a.n_tiles + tile) * a.k;
  for (int j = threadIdx.x; j < a.k; j += BM_THREADS) {
    a.part_k1[o + j] = (j < m) ? s_k1[j] : K1_INVALID;
    a.part_k2[o + j] = (j < m) ? s_k2[j] : K2_INVALID;
  }
}

__global__ void __launch_bounds__(256)
    bm25_impacts_kernel(const int* __restrict__ tf, const int* __restrict__ len,
                        const double* __restrict__ idf, long long n, double k1, double k1p1,
                        double omb, double b, double avgdl, double* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double f = (double)tf[i];
    const double num = __dmul_rn(f, k1p1);
    const double norm = __dadd_rn(omb, __ddiv_rn(__dmul_rn(b, (double)len[i]), avgdl));
    const double den = __dadd_rn(f, __dmul_rn(k1, norm));
    out[i] = __dmul_rn(idf[i], __ddiv_rn(num, den));
  }
}

}  // namespace rr

using namespace rr;

extern "C" size_t rr_bm25_topk_workspace_bytes(int32_t n_tiles, int32_t q, int32_t k) {
  if (n_tiles <= 0 || q <= 0 || k <= 0) return 256;
  const size_t e = (size_t)n_tiles * q * k;
  return align_up(e * 8, 256) + align_up(e * 4, 256) + align_up((size_t)q * 8, 256) + 256;
}

extern "C" int rr_bm25_impacts(const int32_t* post_tf, const int32_t* post_len,
                               const double* post_idf, int64_t n_post, double k1, double b,
                               double avgdl, double* post_impact, void* stream) {
  RR_CHECK_ARG(n_post >= 0, "negative size");
  if (n_post == 0) return RR_OK;
  RR_CHECK_ARG(post_tf && post_len && post_idf && post_impact, "null pointer");
  long long blocks = (n_post + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  // (k1 + 1) and (1 - b) are evaluated on the host in double exactly as Python does
  const double k1p1 = k1 + 1;
  const double omb = 1 - b;
  bm25_impacts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      post_tf, post_len, post_idf, n_post, k1, k1p1, omb, b, avgdl, post_impact);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

__global__ void fill_missing_f64_kernel(double* s, long long* idx, int* count, long long total,
                                        int q) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    s[i] = 0.0;
    idx[i] = -1;
  }
  if (count && i < q) count[i] = 0;
}

extern "C" int rr_bm25_topk(const int64_t* tile_term_ptr, const uint32_t* post_row,
                            const double* post_impact, int32_t n_tiles, int32_t tile_docs,
                            int32_t n_terms, int64_t n_docs, const int32_t* q_terms, int32_t q,
                            int32_t q_len, int32_t k, int64_t row_base, double* out_score,
                            int64_t* out_idx, int32_t* out_count, void* workspace,
                            size_t workspace_bytes, void* stream) {
  RR_CHECK_ARG(q >= 0 && n_docs >= 0 && q_len >= 0, "negative size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(out_score && out_idx, "null pointer");
  RR_CHECK_ARG(k >= 1 && k <= RR_MAX_K, "k out of range");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_docs == 0 || n_tiles == 0 || q_len == 0 || n_terms == 0) {
    const long long total = (long long)q * k;
    const long long threads = total > q ? total : q;
    fill_missing_f64_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        out_score, (long long*)out_idx, out_count, total, q);
    RR_LAUNCH_CHECK();
    return RR_OK;
  }
  RR_CHECK_ARG(tile_term_ptr && post_row && post_impact && q_terms, "null pointer");
  RR_CHECK_ARG(tile_docs >= 32 && tile_docs <= 16384, "tile_docs must be in [32, 16384]");
  RR_CHECK_ARG((long long)n_tiles * tile_docs >= n_docs, "tiles do not cover n_docs");
  RR_CHECK_ARG((long long)n_tiles * q < (1LL << 31), "grid too large");
  const size_t need = rr_bm25_topk_workspace_bytes(n_tiles, q, k);
  if (!workspace || workspace_bytes < need) {
    set_error("rr_bm25_topk: workspace %zu < %zu", workspace_bytes, need);
    return RR_ERR_WORKSPACE;
  }
  Bm25Args a;
  a.tile_term_ptr = (const long long*)tile_term_ptr;
  a.post_row = post_row;
  a.post_impact = post_impact;
  a.n_tiles = n_tiles;
  a.tile_docs = tile_docs;
  a.n_terms = n_terms;
  a.n_docs = n_docs;
  a.q_terms = q_terms;
  a.q = q;
  a.q_len = q_len;
  a.k = k;
  a.cap = merge_cap(k);
  const size_t e = (size_t)n_tiles * q * k;
  a.part_k1 = (u64*)workspace;
  a.part_k2 = (u32*)((char*)workspace + align_up(e * 8, 256));
  a.qbound = (u64*)((char*)workspace + align_up(e * 8, 256) + align_up(e * 4, 256));
  RR_CUDA(cudaMemsetAsync(a.qbound, 0xFF, (size_t)q * 8, st));
  const size_t smem = (size_t)tile_docs * 8 + (size_t)a.cap * 12;
  RR_CUDA(cudaFuncSetAttribute(bm25_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  bm25_tile_kernel<<<(unsigned)((long long)n_tiles * q), BM_THREADS, smem, st>>>(a);
  RR_LAUNCH_CHECK();
  MergeArgs m;
  m.k1 = a.part_k1;
  m.k2 = a.part_k2;
  m.n_in = (long long)n_tiles * k;
  m.k = k;
  m.cap = 0;
  m.row_base = row_base;
  m.out_a = out_score;
  m.out_idx = (long long*)out_idx;
  m.out_count = out_count;
  return launch_merge_pairs<MERGE_F64_DESC>(m, q, st);
}
