// Stage-1 candidate search: exact Hamming top-k over packed sign bits (R4, R12).
//
// Replaces the candidate search of retrieve_by_embedding_quantized
// (reference radiant/storage/redis_store.py:799-809, chroma_store.py:588-619,
// pgvector_store.py:794-802; documented as a Hamming search in
// docs/BINARY_QUANTIZATION_README.md:84-100).
//
// Layout: codes u32 [n, W] row-major (W = dim/32 rounded up to a multiple of 4 so a
// row is a whole number of 16-byte vectors).  A CTA owns one slab of rows and one
// tile of queries.  Each thread keeps one row in registers (W/4 128-bit loads) and
// walks the query tile, whose packed codes sit in shared memory and are read as
// warp-wide broadcasts; XOR + POPC + IADD3 give the distance.  Per query the CTA
// keeps a shared-memory candidate queue guarded by a running threshold (the k-th
// best key seen so far): a row is appended only when its 32-bit key
// (dist << 21 | row-in-slab) beats the threshold, appends are warp-aggregated, and a
// full queue is cut back to its k best by a warp-level radix select.  The per-slab
// lists are merged by merge_pairs_kernel (select.cuh).
//
// Roofline: algorithmic bytes per launch = n * W * 4 (one pass over the
// codes; the query tiles of one slab run back to back so re-reads hit L2); integer
// work = q * n * W popcounts.  HBM-bound for q <= 3, POPC-pipe-bound above.

#include "common.cuh"
#include "merge.cuh"

namespace rr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_ROW_BITS = 21;  // rows per slab <= 2^21
constexpr u32 SCAN_KEY_MAX = 0xFFFFFFFFu;

struct ScanPlan {
  int q_tile;
  int n_qtiles;
  int cap;
  int slabs;
  long long rows_per_slab;
  size_t smem;
};

static ScanPlan plan_scan(long long n, int words, int q, int k) {
  ScanPlan p;
  const int slack = 2 * SCAN_THREADS;
  p.cap = (int)align_up((size_t)k + slack, 32);
  const size_t per_q = (size_t)words * 4 + (size_t)p.cap * 4 + 8;
  const size_t fixed = SCAN_WARPS * 256 * sizeof(int) + 64;
  const size_t budget = 100 * 1024;
  int qt = (int)((budget - fixed) / per_q);
  if (qt < 1) qt = 1;
  if (qt > 64) qt = 64;
  if (qt > q) qt = q;
  p.n_qtiles = (q + qt - 1) / qt;
  p.q_tile = (q + p.n_qtiles - 1) / p.n_qtiles;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long target = 4LL * sms;  // two resident CTAs per SM, two waves
  long long slabs = (target + p.n_qtiles - 1) / p.n_qtiles;
  const long long max_slabs = (n + 2047) / 2048;  // at least 2048 rows per slab
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  long long rps = (n + slabs - 1) / slabs;
  rps = (long long)align_up((size_t)rps, SCAN_THREADS);
  const long long max_rps = 1LL << SCAN_ROW_BITS;
  if (rps > max_rps) rps = max_rps;
  if (rps < SCAN_THREADS) rps = SCAN_THREADS;
  p.rows_per_slab = rps;
  p.slabs = (int)((n + rps - 1) / rps);
  if (p.slabs < 1) p.slabs = 1;
  p.smem = (size_t)p.q_tile * per_q + fixed + 16 * (size_t)p.q_tile;
  return p;
}

// Cut a queue of n unique u32 keys back to its k smallest (n > k).  One warp.
__device__ __forceinline__ u32 warp_compact_queue(u32* queue, int n, int k, int* hist) {
  const int lane = threadIdx.x & 31;
  u32 prefix = 0, mask = 0;
  int need = k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const u32 key = queue[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1);
    }
    __syncwarp();
    int c[8];
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = hist[lane * 8 + j];
      s += c[j];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int excl = incl - s;
    const bool here = (excl < need) && (incl >= need);
    int bsel = 7, run = excl;
    if (here) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (run + c[j] >= need) {
          bsel = j;
          break;
        }
        run += c[j];
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, here);
    const int src = __ffs(bal) - 1;
    const int bucket = __shfl_sync(0xffffffffu, lane * 8 + bsel, src);
    need = __shfl_sync(0xffffffffu, need - run, src);
    prefix |= ((u32)bucket) << shift;
    mask |= 0xFFu << shift;
    __syncwarp();
  }
  const u32 kth = prefix;
  int w = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const u32 key = (i < n) ? queue[i] : SCAN_KEY_MAX;
    const bool keep = (i < n) && (key <= kth);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) queue[w + __popc(bal & ((1u << lane) - 1u))] = key;
    w += __popc(bal);
    __syncwarp();
  }
  return kth;
}

struct ScanArgs {
  const uint4* codes;
  long long n;
  const uint8_t* tags;
  unsigned tag_mask;
  unsigned tag_value;
  const u32* qcodes;
  int q;
  int k;
  int q_tile;
  int n_qtiles;
  int cap;
  int slabs;
  long long rows_per_slab;
  u64* part_k1;  // [q][slabs][k]
  u32* part_k2;
};

template <int W>
__global__ void __launch_bounds__(SCAN_THREADS, 2) hamming_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int V = W / 4;  // 16-byte vectors per row
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int slab = blockIdx.x / a.n_qtiles;
  const int qt = blockIdx.x % a.n_qtiles;
  const int q0 = qt * a.q_tile;
  const int nq = min(a.q_tile, a.q - q0);

  uint4* sq = reinterpret_cast<uint4*>(smem_raw);                       // [q_tile][V]
  u32* queue = reinterpret_cast<u32*>(sq + (size_t)a.q_tile * V);        // [q_tile][cap]
  u32* thr = queue + (size_t)a.q_tile * a.cap;                           // [q_tile]
  int* cnt = reinterpret_cast<int*>(thr + a.q_tile);                     // [q_tile]
  int* whist = cnt + a.q_tile;                                           // [warps][256]

  for (int i = tid; i < nq * V; i += SCAN_THREADS)
    sq[i] = reinterpret_cast<const uint4*>(a.qcodes)[(size_t)q0 * V + i];
  for (int i = tid; i < nq; i += SCAN_THREADS) {
    thr[i] = SCAN_KEY_MAX;
    cnt[i] = 0;
  }
  __syncthreads();

  const long long row_lo = (long long)slab * a.rows_per_slab;
  const long long row_hi = min(a.n, row_lo + a.rows_per_slab);

  for (long long base = row_lo; base < row_hi; base += SCAN_THREADS) {
    const long long row = base + tid;
    bool valid = row < row_hi;
    if (valid && a.tags != nullptr) valid = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
    uint4 r[V];
    if (valid) {
      const uint4* src = a.codes + (size_t)row * V;
#pragma unroll
      for (int v = 0; v < V; ++v) r[v] = __ldg(src + v);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) r[v] = make_uint4(0, 0, 0, 0);
    }
    const u32 local = (u32)(row - row_lo);
#pragma unroll 2
    for (int qi = 0; qi < nq; ++qi) {
      const uint4* qv = sq + (size_t)qi * V;
      int d = 0;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const uint4 c = qv[v];
        d += __popc(r[v].x ^ c.x) + __popc(r[v].y ^ c.y);
        d += __popc(r[v].z ^ c.z) + __popc(r[v].w ^ c.w);
      }
      const u32 key = ((u32)d << SCAN_ROW_BITS) | local;
      const bool pass = valid && (key < thr[qi]);
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if (bal) {
        int slot = 0;
        if (lane == (__ffs(bal) - 1)) slot = atomicAdd(&cnt[qi], __popc(bal));
        slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
        if (pass) queue[(size_t)qi * a.cap + slot + __popc(bal & ((1u << lane) - 1u))] = key;
      }
    }
    __syncthreads();
    // queues that could overflow in the next batch are cut back to their k best
    for (int qi = warp; qi < nq; qi += SCAN_WARPS) {
      const int c = cnt[qi];
      if (c > a.cap - SCAN_THREADS) {
        const u32 kth = warp_compact_queue(queue + (size_t)qi * a.cap, c, a.k, whist + warp * 256);
        if (lane == 0) {
          thr[qi] = kth;
          cnt[qi] = a.k;
        }
      }
    }
    __syncthreads();
  }

  // final cut and write-out of this slab's list for each query of the tile
  for (int qi = warp; qi < nq; qi += SCAN_WARPS) {
    int c = cnt[qi];
    u32* qu = queue + (size_t)qi * a.cap;
    if (c > a.k) {
      warp_compact_queue(qu, c, a.k, whist + warp * 256);
      c = a.k;
    }
    __syncwarp();
    const size_t o = ((size_t)(q0 + qi) * a.slabs + slab) * a.k;
    for (int j = lane; j < a.k; j += 32) {
      if (j < c) {
        const u32 key = qu[j];
        a.part_k1[o + j] = (u64)(key >> SCAN_ROW_BITS);
        a.part_k2[o + j] = (u32)(row_lo + (key & ((1u << SCAN_ROW_BITS) - 1u)));
      } else {
        a.part_k1[o + j] = K1_INVALID;
        a.part_k2[o + j] = K2_INVALID;
      }
    }
  }
}


__global__ void fill_missing_hamming_kernel(int* dist, long long* idx, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    dist[i] = 0x7fffffff;
    idx[i] = -1;
  }
}

template <int W>
static int launch_scan(const ScanArgs& a, const ScanPlan& p, cudaStream_t st) {
  RR_CUDA(cudaFuncSetAttribute(hamming_scan_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)p.smem));
  hamming_scan_kernel<W><<<p.slabs * p.n_qtiles, SCAN_THREADS, p.smem, st>>>(a);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

}  // namespace rr

using namespace rr;

extern "C" size_t rr_hamming_topk_workspace_bytes(int64_t n, int32_t words, int32_t q, int32_t k) {
  if (n <= 0 || q <= 0 || k <= 0 || words <= 0) return 256;
  const ScanPlan p = plan_scan(n, words, q, k);
  return align_up((size_t)q * p.slabs * k * 8, 256) + align_up((size_t)q * p.slabs * k * 4, 256) + 256;
}

extern "C" int rr_hamming_topk(const uint32_t* codes, int64_t n, int32_t words, const uint8_t* tags,
                               uint8_t tag_mask, uint8_t tag_value, const uint32_t* qcodes,
                               int32_t q, int32_t k, int64_t row_base, int32_t* out_dist,
                               int64_t* out_idx, void* workspace, size_t workspace_bytes,
                               void* stream) {
  RR_CHECK_ARG(q >= 0 && n >= 0, "negative size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(codes || n == 0, "codes is null");
  RR_CHECK_ARG(qcodes && out_dist && out_idx, "null pointer");
  RR_CHECK_ARG(words >= 4 && words <= RR_MAX_WORDS && words % 4 == 0,
               "words must be a multiple of 4 in [4, 32]");
  RR_CHECK_ARG(k >= 1 && k <= RR_MAX_K, "k out of range");
  RR_CHECK_ARG(n < (1LL << 32), "shard larger than 2^32 rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {  // empty shard: every slot is missing
    const long long total = (long long)q * k;
    fill_missing_hamming_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        out_dist, (long long*)out_idx, total);
    RR_LAUNCH_CHECK();
    return RR_OK;
  }
  const size_t need = rr_hamming_topk_workspace_bytes(n, words, q, k);
  if (workspace_bytes < need || !workspace) {
    set_error("rr_hamming_topk: workspace %zu < %zu", workspace_bytes, need);
    return RR_ERR_WORKSPACE;
  }
  const ScanPlan p = plan_scan(n, words, q, k);
  ScanArgs a;
  a.codes = (const uint4*)codes;
  a.n = n;
  a.tags = tags;
  a.tag_mask = tag_mask;
  a.tag_value = tag_value;
  a.qcodes = qcodes;
  a.q = q;
  a.k = k;
  a.q_tile = p.q_tile;
  a.n_qtiles = p.n_qtiles;
  a.cap = p.cap;
  a.slabs = p.slabs;
  a.rows_per_slab = p.rows_per_slab;
  a.part_k1 = (u64*)workspace;
  a.part_k2 = (u32*)((char*)workspace + align_up((size_t)q * p.slabs * k * 8, 256));
  int rc = RR_ERR_INVALID;
  switch (words) {
    case 4: rc = launch_scan<4>(a, p, st); break;
    case 8: rc = launch_scan<8>(a, p, st); break;
    case 12: rc = launch_scan<12>(a, p, st); break;
    case 16: rc = launch_scan<16>(a, p, st); break;
    case 20: rc = launch_scan<20>(a, p, st); break;
    case 24: rc = launch_scan<24>(a, p, st); break;
    case 28: rc = launch_scan<28>(a, p, st); break;
    case 32: rc = launch_scan<32>(a, p, st); break;
    default: break;
  }
  if (rc != RR_OK) return rc;
  MergeArgs m;
  m.k1 = a.part_k1;
  m.k2 = a.part_k2;
  m.n_in = (long long)p.slabs * k;
  m.k = k;
  m.cap = 0;
  m.row_base = row_base;
  m.out_a = out_dist;
  m.out_idx = (long long*)out_idx;
  m.out_count = nullptr;
  return launch_merge_pairs<MERGE_HAMMING>(m, q, st);
}

// ---- 8(e): merges of the allgathered per-shard lists --------------------------------
extern "C" int rr_merge_hamming(const int32_t* in_dist, const int64_t* in_idx, int32_t q,
                                int32_t n_in, int32_t k, int32_t* out_dist, int64_t* out_idx,
                                void* stream) {
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(in_dist && in_idx && out_dist && out_idx, "null pointer");
  RR_CHECK_ARG(q > 0 && n_in > 0 && k >= 1 && k <= RR_MAX_K, "bad size");
  return launch_merge_typed<MERGE_HAMMING_PACKED>(in_dist, (const long long*)in_idx, q, n_in, k,
                                                  out_dist, (long long*)out_idx, nullptr,
                                                  (cudaStream_t)stream);
}

extern "C" int rr_merge_scores_f64(const double* in_score, const int64_t* in_idx, int32_t q,
                                   int32_t n_in, int32_t k, double* out_score, int64_t* out_idx,
                                   int32_t* out_count, void* stream) {
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(in_score && in_idx && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(q > 0 && n_in > 0 && k >= 1 && k <= RR_MAX_K, "bad size");
  return launch_merge_typed<MERGE_F64_DESC>(in_score, (const long long*)in_idx, q, n_in, k,
                                            out_score, (long long*)out_idx, out_count,
                                            (cudaStream_t)stream);
}

extern "C" int rr_merge_scores_i32(const int32_t* in_score, const int64_t* in_idx, int32_t q,
                                   int32_t n_in, int32_t k, int32_t* out_score, int64_t* out_idx,
                                   void* stream) {
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(in_score && in_idx && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(q > 0 && n_in > 0 && k >= 1 && k <= RR_MAX_K, "bad size");
  return launch_merge_typed<MERGE_I32_DESC>(in_score, (const long long*)in_idx, q, n_in, k,
                                            out_score, (long long*)out_idx, nullptr,
                                            (cudaStream_t)stream);
}
