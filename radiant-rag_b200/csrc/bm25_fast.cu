// R6, batched: BM25 top-k as filter-and-refine, with the postings shared across the query batch.
//
// Replaces BM25Index.search (reference radiant/storage/bm25_index.py:218-270) for BATCHES of
// queries.  rr_bm25_topk (bm25.cu) walks every (query, tile) pair's postings out of L2 into
// float64 accumulators and is bound by that traffic (sum_t df(t) * 12 B per query, nothing
// shared between queries).  This path returns the SAME bits in a fraction of the time:
//
//   phase A  (bm25_fast_kernel)  float32 scores for every (query, document), in any order.
//            A persistent CTA per SM owns a tile of <= 1024 documents at a time and serves ALL
//            queries of the batch from it: the dense "head" terms of the tile (the <= 32 terms
//            with the largest document frequency - most of the posting mass under Zipf) are
//            staged ONCE per tile into shared memory as float32 columns and then read by every
//            query that contains them; the sparse "tail" terms of a query are scattered into a
//            per-warp fixed-point accumulator with shared-memory atomics.  One warp = one query;
//            the score of a document lives in a register and is only compared with a bound.
//            When the head terms of a query cannot reach the bound on their own (uq < tau_q, the
//            usual case: frequent terms carry little idf) the head columns are read only for the
//            groups of documents whose tail sums come within uq of the bound (MaxScore-style
//            pruning, but inside a filter whose survivors are verified exactly).
//            pass 1 (SAMPLE) runs on every s-th tile and keeps one maximum per lane;
//            tau_q = the k'-th largest of them is a score that >= k' documents reach;
//            pass 2 (FILTER) keeps the documents with score >= tau_q (about s * k' per query).
//   phase B  (bm25_refine_kernel) exact float64 scores of the survivors in the reference's
//            operation order (query-token order, __dadd_rn, impacts as rr_bm25_impacts built
//            them), exact top-k by (score desc, row asc), score > 0.
//
// Exactness.  All impacts are positive.  Tail impacts enter the filter score as fixed-point integers
// (round(impact * 2^s), accumulated with native integer shared-memory atomics), head impacts as
// FLOAT16 (twice as many head columns fit in shared memory, and the 30 terms that follow the first 30
// carry half of the remaining postings under Zipf), summed in float32, so |A - E| <= eps * E + eps_abs
// with eps = (q_len + 4) * 2^-24 + 2^-10 and eps_abs = (q_len + 1) * 2^-(s+1) + q_len * 2^-24 for the
// filter score A and the reference's float64 score E of any document.  A document outside the survivor list has A < tau, hence E < tau * (1 + eps) + eps_abs.
// If the k-th exact score among the survivors exceeds that, no outside document can enter or tie
// the top-k and the result is the reference's, bit for bit.  The refine kernel CHECKS this per query (and list overflow); a query
// that fails is flagged and counted, and the caller redoes it with rr_bm25_topk.  Impacts
// outside [2^-100, 2^100] or non-positive disable this path at index build time (bm25_index.py).
//
// Algorithmic bytes per batch: one pass over the index (postings * 12 B + head columns); per
// query without sharing: sum_t df(t) * 12 B (SURVEY.md 8d reports both).
#include <math.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "select.cuh"
#include "tau.cuh"

namespace rr {

#ifndef RR_BF_WARPS
#define RR_BF_WARPS 32
#endif
constexpr int BF_WARPS = RR_BF_WARPS;  // one query per warp; more resident warps = more posting loads in flight
constexpr int BF_THREADS = BF_WARPS * 32;
constexpr int BF_MAX_HEAD = 64;   // head slot h lives on lane h & 31 (two counters per lane)
constexpr int BF_MAX_TILE = 1024;
#ifndef RR_BF_INFLIGHT
#define RR_BF_INFLIGHT 6
#endif
constexpr int BF_INFLIGHT = RR_BF_INFLIGHT;   // posting loads a lane issues before its first accumulate


struct BfArgs {
  const long long* tile_term_ptr;  // [n_tiles][n_terms + 1]
  const u64* post_pack;     // [P] row-in-tile << 32 | round(impact * 2^fx_shift) (same order as post_row)
  const float* head_max;    // [n_head] largest float32 impact of the head term over all documents
  float fx_inv;             // 2^-fx_shift
  float fx_scale;           // 2^fx_shift
  const int* head_slot;    // [n_terms] slot of a head term, -1 for tail terms
  const double* head_imp;  // [n_tiles][n_head][tile_docs] dense float64 impacts, 0 = absent (refine)
  const __half* head_imp_h;  // the same rounded to float16: what the filter stages into shared memory
  int n_head;
  int n_tiles;
  int tile_docs;
  int n_terms;
  long long n_docs;
  const int* q_terms;  // [q][q_len]
  int q;
  int q_len;
  int stride;        // this pass visits tiles 0, stride, 2*stride, ...
  int n_pass_tiles;
  int n_full;        // the first n_full tiles of the pass (a multiple of the grid) are one work item each;
  int q_groups;      // the queries of each remaining tile are split over q_groups work items, so that the last,
                     // partial round of tiles keeps every SM busy (and small shards with fewer tiles than SMs too)
  int n_items;       // n_full + (n_pass_tiles - n_full) * q_groups
  u32* lane_max;     // SAMPLE out: [q][n_pass_tiles * 32] keys ~orderable(max), 0xFFFFFFFF = none
  const float* tau;  // FILTER in:  [q]
  u32* list_cnt;     // FILTER out: [q]
  u64* list;         // FILTER out: [q][cap]  float bits << 32 | local row
  int cap;
};

struct TokenInfo {
  int hs;        // head slot or -1
  int len;       // tail segment length in this tile (0 for head / unknown tokens)
  long long lo;  // first posting of the segment
};

// lane j describes query token j0 + j for this tile
__device__ __forceinline__ TokenInfo bf_token_info(const BfArgs& a, const long long* ptr, int t) {
  TokenInfo ti;
  ti.hs = -1;
  ti.len = 0;
  ti.lo = 0;
  if (t >= 0 && t < a.n_terms) {  // unknown token: skipped (bm25_index.py:238-239)
    // three independent loads (the bounds of a head term are fetched and dropped: one load
    // latency per query instead of two)
    const int hs = __ldg(a.head_slot + t);
    const long long lo = __ldg(ptr + t);
    const long long hi = __ldg(ptr + t + 1);
    ti.hs = hs;
    ti.lo = lo;
    ti.len = hs < 0 ? (int)(hi - lo) : 0;
  }
  return ti;
}

__device__ __forceinline__ float bf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool SAMPLE>
__global__ void __launch_bounds__(BF_THREADS, 1) bm25_fast_kernel(const BfArgs a) {
  extern __shared__ __align__(16) unsigned char bf_smem[];
  const int T = a.tile_docs;
  __half* cols = reinterpret_cast<__half*>(bf_smem);                         // [n_head][T] float16 head impacts
  u32* tacc_all = reinterpret_cast<u32*>(cols + (size_t)a.n_head * T);       // [BF_WARPS][T] fixed-point tail sums
  __shared__ int s_excl[BF_WARPS][34];         // compacted tail tokens: first flat index, then 2 sentinels
  __shared__ long long s_base[BF_WARPS][32];   // first posting of the segment minus its first flat index
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32* tacc = tacc_all + (size_t)warp * T;
  const int t4 = T >> 2;  // 16-byte groups per column
  const float inv = a.fx_inv;
  const float hmax = (lane < a.n_head) ? __ldg(a.head_max + lane) : 0.0f;            // slot lane
  const float hmax_hi = (32 + lane < a.n_head) ? __ldg(a.head_max + 32 + lane) : 0.0f;  // slot 32 + lane

  for (int w = blockIdx.x; w < a.n_items; w += gridDim.x) {
    int pt = w, grp = 0, ng = 1;
    if (w >= a.n_full) {
      const int r = w - a.n_full;
      pt = a.n_full + r / a.q_groups;
      grp = r - (pt - a.n_full) * a.q_groups;
      ng = a.q_groups;
    }
    const int q_lo = (int)((long long)a.q * grp / ng);
    const int q_hi = (int)((long long)a.q * (grp + 1) / ng);
    const int tile = pt * a.stride;
    const long long tile_lo = (long long)tile * T;
    const int rows_here = (int)min((long long)T, a.n_docs - tile_lo);
    __syncthreads();  // every warp is done with the previous tile's columns
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.head_imp_h + (size_t)tile * a.n_head * T);
      uint4* dst = reinterpret_cast<uint4*>(cols);
      const int n16 = (a.n_head * T) >> 3;  // T is a multiple of 128
      for (int i = threadIdx.x; i < n16; i += BF_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const long long* ptr = a.tile_term_ptr + (size_t)tile * (a.n_terms + 1);
    // the tail accumulator is zeroed once per tile; every query leaves it clean (its dense pass
    // writes zeros back behind the values it reads)
    for (int i = lane; i < t4; i += 32) reinterpret_cast<uint4*>(tacc)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();

    // software pipeline over this warp's queries: token ids two queries ahead, their head
    // slots / segment bounds one query ahead (a chain of dependent loads otherwise)
    int t_n2 = -1, t_n1 = -1;
    TokenInfo ti_n1;
    ti_n1.hs = -1;
    ti_n1.len = 0;
    ti_n1.lo = 0;
    if (q_lo + warp < q_hi) {
      t_n1 = (lane < a.q_len) ? __ldg(a.q_terms + (size_t)(q_lo + warp) * a.q_len + lane) : -1;
      ti_n1 = bf_token_info(a, ptr, t_n1);
    }
    if (q_lo + warp + BF_WARPS < q_hi)
      t_n2 = (lane < a.q_len) ? __ldg(a.q_terms + (size_t)(q_lo + warp + BF_WARPS) * a.q_len + lane) : -1;

    for (int qi = q_lo + warp; qi < q_hi; qi += BF_WARPS) {
      TokenInfo ti = ti_n1;
      const int t_cur_next = t_n2;  // tokens of query qi + BF_WARPS
      if (qi + 2 * BF_WARPS < q_hi)
        t_n2 = (lane < a.q_len) ? __ldg(a.q_terms + (size_t)(qi + 2 * BF_WARPS) * a.q_len + lane) : -1;
      else
        t_n2 = -1;
      if (qi + BF_WARPS < q_hi) ti_n1 = bf_token_info(a, ptr, t_cur_next);
      float tau = 0.0f;
      if (!SAMPLE) tau = __ldg(a.tau + qi);

      int head_cnt = 0, head_cnt_hi = 0;  // lane l: multiplicity of head slots l and 32 + l in this query
      bool sparse = false;
      float uq = 0.0f;
      u32 thr_fx = 0u;  // fixed-point tail sum a document needs before its head terms can matter

      for (int j0 = 0; j0 < a.q_len; j0 += 32) {
        if (j0 > 0) {  // long queries: later chunks are not prefetched
          const int t = (j0 + lane < a.q_len) ? __ldg(a.q_terms + (size_t)qi * a.q_len + j0 + lane) : -1;
          ti = bf_token_info(a, ptr, t);
        }
        // head tokens: count multiplicities per slot
        unsigned hm = __ballot_sync(0xffffffffu, ti.hs >= 0);
        while (hm) {
          const int src = __ffs(hm) - 1;
          hm &= hm - 1;
          const int h = __shfl_sync(0xffffffffu, ti.hs, src);
          if (lane == (h & 31)) {
            if (h < 32)
              ++head_cnt;
            else
              ++head_cnt_hi;
          }
        }
        if (!SAMPLE && a.q_len <= 32) {
          // The head terms of a query add at most uq to any document (their largest impact anywhere).
          // When that alone cannot reach the bound (the usual case: frequent terms carry little
          // idf), the head columns are only read for the 16-document groups whose tail sums come
          // within uq of it.
          uq = bf_warp_sum(fmaf((float)head_cnt_hi, hmax_hi, (float)head_cnt * hmax)) * 1.000002f;
          sparse = tau > 0.0f && uq < tau * 0.999998f;
          if (sparse) thr_fx = (u32)fminf((tau * 0.999998f - uq) * a.fx_scale * 0.99999f, 4.0e9f);
        }
        // tail tokens: all their postings of this tile as ONE flat index space (balanced over the
        // lanes whatever the segment lengths), walked with a per-lane segment cursor; up to BF_INFLIGHT
        // independent 8-byte loads per lane are in flight before the first accumulate.  The sums are
        // fixed-point integers, so the scatter is one native shared-memory atomic per posting.
        // (BF_INFLIGHT: 6 measured best with 32 warps - 1.36 ms against 1.42 at 12 and 1.37 at 4.)
        // (Measured alternatives that were not faster: one segment at a time with lanes striding over
        // its postings - a quarter fewer instructions, the same time: the kernel is bound by the latency of
        // each warp's dependent chain at 8 warps per scheduler, not by issue slots;
        // float accumulators (a compare-and-swap loop per
        // posting), atomic-free read-modify-write of one segment per instruction with a warp barrier
        // in between, scoring only the documents the tail touches - a third of a tile - and staging
        // the postings through shared memory with cp.async, which leaves room for fewer warps.)
        int incl = ti.len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total > 0) {
          const unsigned tm = __ballot_sync(0xffffffffu, ti.len > 0);
          const int nt = __popc(tm);
          if (ti.len > 0) {
            const int r = __popc(tm & ((1u << lane) - 1u));
            s_excl[warp][r] = incl - ti.len;
            s_base[warp][r] = ti.lo - (long long)(incl - ti.len);
          }
          if (lane == 0) {
            s_excl[warp][nt] = total;
            s_excl[warp][nt + 1] = 0x7fffffff;
          }
          __syncwarp();
          int j = 0;
          int nb = s_excl[warp][1];
          long long base = s_base[warp][0];
          for (int f0 = lane; f0 < total; f0 += 32 * BF_INFLIGHT) {
            u64 pk[BF_INFLIGHT];
#pragma unroll
            for (int u = 0; u < BF_INFLIGHT; ++u) {
              const int f = f0 + 32 * u;
              pk[u] = ~0ull;
              if (f < total) {
                while (f >= nb) {  // rare: a lane crosses a segment boundary every few postings
                  ++j;
                  nb = s_excl[warp][j + 1];
                  base = s_base[warp][j];
                }
                pk[u] = __ldg(a.post_pack + base + f);
              }
            }
#pragma unroll
            for (int u = 0; u < BF_INFLIGHT; ++u)
              if (pk[u] != ~0ull) atomicAdd(tacc + (int)(pk[u] >> 32), (u32)pk[u]);
          }
          __syncwarp();
        }
      }
      __syncwarp();

      // dense pass: scores of the documents this lane owns (tail sum + head columns), 16 documents
      // at a time: four float4 accumulators keep the register count low enough for 32 resident warps (<= 64 registers)
      const unsigned hmask = __ballot_sync(0xffffffffu, head_cnt > 0);
      const unsigned hmask_hi = __ballot_sync(0xffffffffu, head_cnt_hi > 0);
      float lmax = 0.0f;
      for (int hb = 0; hb < t4; hb += 128) {
        uint4 tv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = hb + i * 32 + lane;
          tv[i] = make_uint4(0u, 0u, 0u, 0u);
          if (v < t4) {
            tv[i] = reinterpret_cast<const uint4*>(tacc)[v];
            reinterpret_cast<uint4*>(tacc)[v] = make_uint4(0u, 0u, 0u, 0u);  // clean for the next query
          }
        }
        if (!SAMPLE && sparse) {
          // integer test on the raw sums: no document of this group can reach the bound
          bool hit = false;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            hit |= (tv[i].x >= thr_fx) | (tv[i].y >= thr_fx) | (tv[i].z >= thr_fx) | (tv[i].w >= thr_fx);
          if (!__any_sync(0xffffffffu, hit)) continue;
        }
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          acc[i] = make_float4(__uint2float_rn(tv[i].x) * inv, __uint2float_rn(tv[i].y) * inv,
                               __uint2float_rn(tv[i].z) * inv, __uint2float_rn(tv[i].w) * inv);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          unsigned hm = part ? hmask_hi : hmask;
          while (hm) {
            const int hl = __ffs(hm) - 1;
            hm &= hm - 1;
            const float c = (float)__shfl_sync(0xffffffffu, part ? head_cnt_hi : head_cnt, hl);
            const uint2* cp = reinterpret_cast<const uint2*>(cols + (size_t)(32 * part + hl) * T) + hb + lane;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (hb + i * 32 + lane < t4) {
                const uint2 xx = cp[i * 32];  // four float16 impacts: documents 4v .. 4v + 3
                const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&xx.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&xx.y));
                acc[i].x = fmaf(c, lo.x, acc[i].x);
                acc[i].y = fmaf(c, lo.y, acc[i].y);
                acc[i].z = fmaf(c, hi.x, acc[i].z);
                acc[i].w = fmaf(c, hi.y, acc[i].w);
              }
            }
          }
        }
        if (SAMPLE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int d = (hb + i * 32 + lane) * 4;
            if (d + 0 < rows_here) lmax = fmaxf(lmax, acc[i].x);
            if (d + 1 < rows_here) lmax = fmaxf(lmax, acc[i].y);
            if (d + 2 < rows_here) lmax = fmaxf(lmax, acc[i].z);
            if (d + 3 < rows_here) lmax = fmaxf(lmax, acc[i].w);
          }
        } else {
          bool any = false;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float mx = fmaxf(fmaxf(acc[i].x, acc[i].y), fmaxf(acc[i].z, acc[i].w));
            any |= (mx >= tau) && (mx > 0.0f);
          }
          if (any) {  // rare: about s * k' survivors per query over the whole corpus
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int d = (hb + i * 32 + lane) * 4;
              const float v4[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (v4[e] >= tau && v4[e] > 0.0f && d + e < rows_here) {
                  const u32 slot = atomicAdd(a.list_cnt + qi, 1u);
                  if (slot < (u32)a.cap)
                    a.list[(size_t)qi * a.cap + slot] =
                        ((u64)__float_as_uint(v4[e]) << 32) | (u64)(u32)(tile_lo + d + e);
                }
              }
            }
          }
        }
      }
      if (SAMPLE)
        a.lane_max[(size_t)qi * ((size_t)a.n_pass_tiles * 32) + (size_t)pt * 32 + lane] =
            lmax > 0.0f ? ~f32_orderable(lmax) : 0xFFFFFFFFu;
      __syncwarp();  // the zeros written above are visible to the next query's scatter
    }
  }
}

struct BrArgs {
  const long long* tile_term_ptr;
  const u32* post_row;
  const double* post_impact;
  const int* head_slot;
  const double* head_imp;
  int n_head;
  int tile_docs;
  int n_terms;
  const int* q_terms;
  int q_len;
  const u32* list_cnt;
  const u64* list;
  int cap;
  const float* tau;
  int k;
  int sel_cap;
  int chunk;  // candidates whose per-token impacts are staged together
  long long row_base;
  double eps;      // relative error bound of a filter score
  double eps_abs;  // absolute error bound (fixed-point tail sums)
  double* out_score;
  long long* out_idx;
  int* out_count;
  unsigned char* flags;
  u32* counter;
};

constexpr int BR_THREADS = 256;
#ifndef RR_BR_CAND_CAP
#define RR_BR_CAND_CAP 2048
#endif
#ifndef RR_BR_CHUNK
#define RR_BR_CHUNK 256
#endif
constexpr int BR_CAND_CAP = RR_BR_CAND_CAP;  // exact-scored candidates per query

// Phase B.  The survivor list of a query holds every document with float32 score A >= tau_q
// (about s * k' of them).  Only those that can still reach the top-k are scored exactly:
//   A_k   = k-th largest A among the survivors;  candidates C = { A >= thr },
//   thr   = A_k * (1 - 8 eps) - 4 eps_abs
// (a survivor below thr has E < thr * (1 + eps) + eps_abs, while the k best-by-A survivors have
// E >= A_k * (1 - eps) - eps_abs, so it cannot be among the k best exact scores).  Exact scores: one thread
// per (candidate, query token) fetches the impact - dense head column or a binary search in the
// (tile, term) segment - and one thread per candidate adds them in query-token order (__dadd_rn).
__global__ void __launch_bounds__(BR_THREADS) bm25_refine_kernel(const BrArgs a) {
  extern __shared__ __align__(16) unsigned char br_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(br_smem);                    // [sel_cap]
  double* s_es = reinterpret_cast<double*>(s_k1 + a.sel_cap);     // [BR_CAND_CAP] exact scores
  double* s_imp = s_es + BR_CAND_CAP;                             // [chunk][q_len]
  u32* s_k2 = reinterpret_cast<u32*>(s_imp + (size_t)a.chunk * a.q_len);  // [sel_cap]
  u32* s_cand = s_k2 + a.sel_cap;                                 // [BR_CAND_CAP] rows
  int* s_term = reinterpret_cast<int*>(s_cand + BR_CAND_CAP);     // [q_len]
  int* s_hs = s_term + a.q_len;                                   // [q_len]
  __shared__ SelectScratch<BR_THREADS> sc;
  __shared__ int s_nc;
  const int q = blockIdx.x;
  const u32 cnt = a.list_cnt[q];
  const int n_s = (int)min(cnt, (u32)a.cap);
  for (int j = threadIdx.x; j < a.q_len; j += BR_THREADS) {
    int t = a.q_terms[(size_t)q * a.q_len + j];
    int hs = -1;
    if (t >= 0 && t < a.n_terms) hs = a.head_slot[t];
    else t = -1;
    s_term[j] = t;
    s_hs[j] = hs;
  }
  if (threadIdx.x == 0) s_nc = 0;
  const u64* lst = a.list + (size_t)q * a.cap;
  // ---- k-th largest float32 score among the survivors (positive floats order like their bits)
  auto get_a = [&](long long i, u64& x, u32& y) {
    const u64 e = lst[i];
    x = (u64)(~(u32)(e >> 32));
    y = (u32)e;
  };
  const int m1 = block_select_sorted<BR_THREADS, true>(get_a, n_s, a.k, s_k1, s_k2, a.sel_cap, sc);
  double thr = 0.0;  // fewer than k survivors: all of them are candidates
  if (m1 == a.k) {
    thr = (double)__uint_as_float(~(u32)s_k1[a.k - 1]) * (1.0 - 8.0 * a.eps) - 4.0 * a.eps_abs;
    if (thr < 0.0) thr = 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_s; i += BR_THREADS) {
    const u64 e = lst[i];
    if ((double)__uint_as_float((u32)(e >> 32)) >= thr) {
      const int slot = atomicAdd(&s_nc, 1);
      if (slot < BR_CAND_CAP) s_cand[slot] = (u32)e;
    }
  }
  __syncthreads();
  const int n_c_all = s_nc;
  const int n_c = min(n_c_all, BR_CAND_CAP);
  // ---- exact float64 scores of the candidates, `chunk` candidates at a time
  const int T = a.tile_docs;
  for (int c0 = 0; c0 < n_c; c0 += a.chunk) {
    const int nc = min(a.chunk, n_c - c0);
    for (int w = threadIdx.x; w < nc * a.q_len; w += BR_THREADS) {
      const int c = w / a.q_len, j = w - c * a.q_len;
      const int t = s_term[j];
      double imp = 0.0;  // absent term / unknown token: x + 0.0 == x exactly
      if (t >= 0) {
        const u32 row = s_cand[c0 + c];
        const int tile = (int)(row / (u32)T);
        const int r = (int)(row - (u32)tile * (u32)T);
        const int hs = s_hs[j];
        if (hs >= 0) {
          imp = __ldg(a.head_imp + ((size_t)tile * a.n_head + hs) * T + r);
        } else {
          const long long* ptr = a.tile_term_ptr + (size_t)tile * (a.n_terms + 1);
          long long lo = __ldg(ptr + t), hi = __ldg(ptr + t + 1);
          while (lo < hi) {  // rows ascend inside a (tile, term) segment
            const long long mid = (lo + hi) >> 1;
            const u32 rm = __ldg(a.post_row + mid);
            if (rm < row) lo = mid + 1;
            else if (rm > row) hi = mid;
            else {
              imp = __ldg(a.post_impact + mid);
              break;
            }
          }
        }
      }
      s_imp[w] = imp;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nc; c += BR_THREADS) {
      double acc = 0.0;
      for (int j = 0; j < a.q_len; ++j) acc = __dadd_rn(acc, s_imp[c * a.q_len + j]);  // query-token order
      s_es[c0 + c] = acc;
    }
    __syncthreads();
  }
  auto get_e = [&](long long i, u64& x, u32& y) {
    const double s = s_es[i];
    x = (s > 0.0) ? ~f64_orderable(s) : K1_INVALID;  // score <= 0 dropped (bm25_index.py:267)
    y = s_cand[i];
  };
  const int m = block_select_sorted<BR_THREADS, true>(get_e, n_c, a.k, s_k1, s_k2, a.sel_cap, sc);
  for (int j = threadIdx.x; j < a.k; j += BR_THREADS) {
    const size_t o = (size_t)q * a.k + j;
    if (j < m) {
      a.out_score[o] = f64_from_orderable(~s_k1[j]);
      a.out_idx[o] = (long long)s_k2[j] + a.row_base;
    } else {
      a.out_score[o] = 0.0;
      a.out_idx[o] = -1;
    }
  }
  if (threadIdx.x == 0) {
    if (a.out_count) a.out_count[q] = m;
    // every document that was NOT scored exactly has E < bound * (1 + eps) + eps_abs
    const double tau = (double)a.tau[q];
    const double bound = thr > tau ? thr : tau;
    bool ok = cnt <= (u32)a.cap && n_c_all <= BR_CAND_CAP;
    if (bound > 0.0) {
      if (m < a.k) {
        ok = false;
      } else {
        const double theta = f64_from_orderable(~s_k1[a.k - 1]);
        if (!(theta > bound * (1.0 + 2.0 * a.eps) + 2.0 * a.eps_abs)) ok = false;
      }
    }
    if (a.flags) a.flags[q] = ok ? 0 : 1;
    if (!ok && a.counter) atomicAdd(a.counter, 1u);
  }
}

// head columns + one tail accumulator per warp + the static arrays must fit the 227 KB a CTA may use
static inline int bf_max_head(int tile_docs) {
  const long long avail = 232448LL - (long long)BF_WARPS * (34 * 4 + 32 * 8) - 1024 - (long long)BF_WARPS * tile_docs * 4;
  long long h = avail / ((long long)tile_docs * 2);  // float16 columns
  if (h > BF_MAX_HEAD) h = BF_MAX_HEAD;
  return h < 0 ? 0 : (int)h;
}
// Work items of a pass: whole tiles while they fill complete rounds of the grid; the tiles of the last,
// partial round (all tiles of a small shard) are split by queries into g items each, g minimising
//   rounds of items per CTA x (query rounds per item + staging of the tile's columns ~ half a query round).
// (Measured before: 977 tiles on 148 SMs cost exactly what 1036 tiles cost - 7 rounds - and 888 tiles 6/7 of it.)
struct BfItems {
  int n_full, groups, n_items;
};
static inline BfItems bf_plan_items(int n_pass_tiles, int q, int sms) {
  BfItems it;
  it.n_full = (n_pass_tiles / sms) * sms;
  const int rem = n_pass_tiles - it.n_full;
  it.groups = 1;
  if (rem > 0) {
    const int max_g = q / (2 * BF_WARPS) > 1 ? q / (2 * BF_WARPS) : 1;  // at least two queries per warp
    double best = 1e30;
    for (int g = 1; g <= max_g && g <= 64; ++g) {
      const int rounds = (rem * g + sms - 1) / sms;
      const int qr = ((q + g - 1) / g + BF_WARPS - 1) / BF_WARPS;
      const double cost = (double)rounds * ((double)qr + 0.5);
      if (cost < best - 1e-9) {
        best = cost;
        it.groups = g;
      }
    }
  }
  it.n_items = it.n_full + rem * it.groups;
  return it;
}
// the bound is a score that k' DISTINCT sampled documents reach (one per lane), so at least k' >= k
// documents survive the filter; a few extra keep the k-th exact score clear of the bound
static inline int bf_kprime(int k) { return k + 4; }
static inline int bf_list_cap(int k) {
  int c = 4096;
  while (c < 16 * bf_kprime(k)) c <<= 1;
  if (c > 32768) c = 32768;
  return c;
}
// sample stride (0 = no sample pass: every positive score survives)
static inline int bf_stride(int n_tiles, int tile_docs, long long n_docs, int k) {
  const int cap = bf_list_cap(k), kp = bf_kprime(k);
  if (n_docs <= cap) return 0;
  int s = cap / (3 * kp);
  if (s > 8) s = 8;  // (every 16th tile was measured: the sample pass halves, but the looser bound lets twice as
                     //  many documents through the filter pass, whose appends cost more than was saved)
  const long long by_count = (long long)n_tiles * 32 / (4LL * kp);
  if (s > by_count) s = (int)by_count;
  if (s < 1) s = 1;
  return s;
}

struct BfLayout {
  size_t lane_max, tau, list_cnt, list, total;
};
static BfLayout bf_layout(int n_tiles, int tile_docs, long long n_docs, int q, int k) {
  BfLayout l;
  const int s = bf_stride(n_tiles, tile_docs, n_docs, k);
  const size_t n_samp = s ? (size_t)(n_tiles + s - 1) / s : 0;
  const size_t cap = (size_t)bf_list_cap(k);
  size_t off = 0;
  l.lane_max = off;
  off += align_up((size_t)q * n_samp * 32 * 4, 256);
  l.tau = off;
  off += align_up((size_t)q * 4, 256);
  l.list_cnt = off;
  off += align_up((size_t)q * 4, 256);
  l.list = off;
  off += align_up((size_t)q * cap * 8, 256);
  l.total = off + 256;
  return l;
}

// optional per-kernel event timing (rr_bm25_timing), per host thread like rr_tc_timing:
// events 0..4 bracket sample pass, tau, filter pass, refine
static thread_local bool g_bf_timing = false;
static thread_local bool g_bf_timed = false;
static thread_local cudaEvent_t g_bf_ev[5];
static thread_local bool g_bf_ev_ready = false;
static void bf_mark(int i, cudaStream_t st) {
  if (g_bf_timing && g_bf_ev_ready) cudaEventRecord(g_bf_ev[i], st);
}

}  // namespace rr

using namespace rr;

extern "C" int rr_bm25_timing(int32_t enable) {
  if (enable && !g_bf_ev_ready) {
    for (int i = 0; i < 5; ++i) RR_CUDA(cudaEventCreate(&g_bf_ev[i]));
    g_bf_ev_ready = true;
  }
  g_bf_timing = enable != 0;
  if (!g_bf_timing) g_bf_timed = false;
  return RR_OK;
}

extern "C" int rr_bm25_last_timing_ms(float* out_ms) {
  RR_CHECK_ARG(out_ms != nullptr, "null pointer");
  if (!g_bf_timed) {
    set_error("rr_bm25_last_timing_ms: no timed call (rr_bm25_timing(1) first)");
    return RR_ERR_INVALID;
  }
  RR_CUDA(cudaEventSynchronize(g_bf_ev[4]));
  for (int i = 0; i < 4; ++i) RR_CUDA(cudaEventElapsedTime(&out_ms[i], g_bf_ev[i], g_bf_ev[i + 1]));
  return RR_OK;
}

extern "C" int rr_bm25_fast_max_head(int32_t tile_docs) {
  if (tile_docs < 128 || tile_docs > BF_MAX_TILE || tile_docs % 128 != 0) return -1;  // tile unsupported
  return bf_max_head(tile_docs);
}

extern "C" size_t rr_bm25_fast_workspace_bytes(int32_t n_tiles, int32_t tile_docs, int64_t n_docs,
                                               int32_t q, int32_t k) {
  if (n_tiles <= 0 || q <= 0 || k <= 0 || tile_docs <= 0) return 256;
  return bf_layout(n_tiles, tile_docs, n_docs, q, k).total;
}

extern "C" int rr_bm25_topk_fast(const int64_t* tile_term_ptr, const uint32_t* post_row,
                                 const double* post_impact, const uint64_t* post_pack,
                                 int32_t fx_shift, const int32_t* head_slot,
                                 const double* head_imp, const void* head_imp_f16, const float* head_max,
                                 int32_t n_head, int32_t n_tiles,
                                 int32_t tile_docs, int32_t n_terms, int64_t n_docs,
                                 const int32_t* q_terms, int32_t q, int32_t q_len, int32_t k,
                                 int64_t row_base, double* out_score, int64_t* out_idx,
                                 int32_t* out_count, uint8_t* inexact_flags,
                                 uint32_t* inexact_counter, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  RR_CHECK_ARG(q >= 0 && n_docs > 0 && q_len > 0 && n_tiles > 0 && n_terms > 0, "bad size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(out_score && out_idx, "null pointer");
  RR_CHECK_ARG(tile_term_ptr && post_row && post_impact && post_pack && head_slot && q_terms, "null pointer");
  RR_CHECK_ARG(k >= 1 && k <= RR_MAX_K, "k out of range");
  RR_CHECK_ARG(tile_docs >= 128 && tile_docs <= BF_MAX_TILE && tile_docs % 128 == 0,
               "tile_docs must be a multiple of 128 in [128, 1024]");
  RR_CHECK_ARG(n_head >= 0 && n_head <= BF_MAX_HEAD, "n_head must be in [0, 64]");
  RR_CHECK_ARG(n_head == 0 || (head_imp && head_imp_f16 && head_max), "head_imp / head_imp_f16 / head_max is null");
  RR_CHECK_ARG(((uintptr_t)head_imp_f16 & 15) == 0, "head_imp_f16 must be 16-byte aligned");
  RR_CHECK_ARG(fx_shift >= 0 && fx_shift <= 60, "fx_shift out of range");
  RR_CHECK_ARG(q_len <= 64, "rr_bm25_topk_fast takes at most 64 tokens per query (fixed-point headroom)");
  RR_CHECK_ARG((long long)n_tiles * tile_docs >= n_docs, "tiles do not cover n_docs");
  RR_CHECK_ARG(n_docs < (1LL << 32), "n_docs must be < 2^32");
  const BfLayout l = bf_layout(n_tiles, tile_docs, n_docs, q, k);
  if (!workspace || workspace_bytes < l.total) {
    set_error("rr_bm25_topk_fast: workspace %zu < %zu", workspace_bytes, l.total);
    return RR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int stride = bf_stride(n_tiles, tile_docs, n_docs, k);
  const int cap = bf_list_cap(k);
  const int kp = bf_kprime(k);
  BfArgs a;
  a.tile_term_ptr = (const long long*)tile_term_ptr;
  a.post_pack = (const u64*)post_pack;
  a.head_max = head_max;
  a.fx_inv = (float)ldexp(1.0, -fx_shift);
  a.fx_scale = (float)ldexp(1.0, fx_shift);
  a.head_slot = head_slot;
  a.head_imp = head_imp;
  a.head_imp_h = reinterpret_cast<const __half*>(head_imp_f16);
  a.n_head = n_head;
  a.n_tiles = n_tiles;
  a.tile_docs = tile_docs;
  a.n_terms = n_terms;
  a.n_docs = n_docs;
  a.q_terms = q_terms;
  a.q = q;
  a.q_len = q_len;
  a.lane_max = (u32*)(ws + l.lane_max);
  a.tau = (const float*)(ws + l.tau);
  a.list_cnt = (u32*)(ws + l.list_cnt);
  a.list = (u64*)(ws + l.list);
  a.cap = cap;
  const size_t smem = (size_t)n_head * tile_docs * 2 + (size_t)BF_WARPS * tile_docs * 4;
  RR_CHECK_ARG(n_head <= bf_max_head(tile_docs), "too many head columns for this tile size (rr_bm25_fast_max_head)");
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  RR_CUDA(cudaMemsetAsync(ws + l.list_cnt, 0, (size_t)q * 4, st));
  bf_mark(0, st);
  if (stride > 0) {
    a.stride = stride;
    a.n_pass_tiles = (n_tiles + stride - 1) / stride;
    const BfItems it = bf_plan_items(a.n_pass_tiles, q, sms);
    a.n_full = it.n_full;
    a.q_groups = it.groups;
    a.n_items = it.n_items;
    RR_CUDA(cudaFuncSetAttribute(bm25_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = it.n_items < sms ? it.n_items : sms;
    bm25_fast_kernel<true><<<grid, BF_THREADS, smem, st>>>(a);
    RR_LAUNCH_CHECK();
    bf_mark(1, st);
    tau_keys_kernel<true><<<q, TAU_THREADS, 0, st>>>(a.lane_max, (long long)a.n_pass_tiles * 32, kp,
                                                      (void*)(ws + l.tau), 0);
    RR_LAUNCH_CHECK();
  } else {
    bf_mark(1, st);
    RR_CUDA(cudaMemsetAsync(ws + l.tau, 0, (size_t)q * 4, st));
  }
  bf_mark(2, st);
  a.stride = 1;
  a.n_pass_tiles = n_tiles;
  RR_CUDA(cudaFuncSetAttribute(bm25_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    const BfItems it = bf_plan_items(n_tiles, q, sms);
    a.n_full = it.n_full;
    a.q_groups = it.groups;
    a.n_items = it.n_items;
    const int grid = it.n_items < sms ? it.n_items : sms;
    bm25_fast_kernel<false><<<grid, BF_THREADS, smem, st>>>(a);
    RR_LAUNCH_CHECK();
  }
  bf_mark(3, st);
  BrArgs r;
  r.tile_term_ptr = a.tile_term_ptr;
  r.post_row = post_row;
  r.post_impact = post_impact;
  r.head_slot = head_slot;
  r.head_imp = head_imp;
  r.n_head = n_head;
  r.tile_docs = tile_docs;
  r.n_terms = n_terms;
  r.q_terms = q_terms;
  r.q_len = q_len;
  r.list_cnt = a.list_cnt;
  r.list = a.list;
  r.cap = cap;
  r.tau = a.tau;
  r.k = k;
  r.sel_cap = 64;
  while (r.sel_cap < 2 * k) r.sel_cap <<= 1;
  if (r.sel_cap > 2048) r.sel_cap = 2048;
  while (r.sel_cap < k) r.sel_cap <<= 1;
  r.chunk = (int)((32 * 1024) / ((size_t)q_len * 8));
  if (r.chunk > RR_BR_CHUNK) r.chunk = RR_BR_CHUNK;
  if (r.chunk < 1) r.chunk = 1;
  r.row_base = row_base;
  // float32 arithmetic: (q_len + 4) * 2^-24; float16 head impacts: each within 2^-11 (1 + 2^-12) of
  // the float64 value, or within 2^-25 absolutely when subnormal - 2^-10 relative bounds the sum
  r.eps = (double)(q_len + 4) * 5.9604644775390625e-08 + (n_head > 0 ? 9.765625e-04 : 0.0);
  r.eps_abs = (double)(q_len + 1) * ldexp(1.0, -(fx_shift + 1)) + (n_head > 0 ? (double)q_len * ldexp(1.0, -24) : 0.0);
  r.out_score = out_score;
  r.out_idx = (long long*)out_idx;
  r.out_count = out_count;
  r.flags = inexact_flags;
  r.counter = inexact_counter;
  const size_t rsm = (size_t)r.sel_cap * 12 + (size_t)BR_CAND_CAP * 12 + (size_t)r.chunk * q_len * 8 +
                     (size_t)q_len * 8 + 16;
  RR_CUDA(cudaFuncSetAttribute(bm25_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));
  bm25_refine_kernel<<<q, BR_THREADS, rsm, st>>>(r);
  RR_LAUNCH_CHECK();
  bf_mark(4, st);
  g_bf_timed = g_bf_timing && g_bf_ev_ready;
  return RR_OK;
}
