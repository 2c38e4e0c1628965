// Stage-1 candidate search: exact Hamming top-k over packed sign bits (R4, R12).
//
// Replaces the candidate search of retrieve_by_embedding_quantized
// (reference radiant/storage/redis_store.py:799-809, chroma_store.py:588-619,
// pgvector_store.py:794-802; documented as a Hamming search in
// docs/BINARY_QUANTIZATION_README.md:84-100).
//
// Layout: codes u32 [n, W] row-major (W = dim/32 rounded up to a multiple of 4 so a
// row is a whole number of 16-byte vectors).  A CTA owns one slab of rows and one
// tile of queries.
//   * Rows are streamed from HBM into a ring of shared-memory tiles (256 rows each) by
//     the TMA bulk-copy engine (cp.async.bulk + mbarrier complete_tx); the next tiles
//     are in flight while the current one is scored, so the single-query case runs at
//     memory speed.
//   * Each thread pulls one row of the tile into registers with 128-bit loads in a
//     rotated chunk order (lane i reads chunk (j+i) mod V at step j), which is
//     bank-conflict-free on the dense tile without a swizzle, then walks the query
//     tile (packed query codes in shared memory, read with the same rotation).
//   * XOR then a first-level carry-save adder tree (LOP3) folds three words into a
//     "ones" and a "twos" word, so only 2/3 of the POPCs of the naive loop issue on the
//     quarter-rate POPC pipe, which is the pipe that bounds the batched scan.
//   * Per query the CTA keeps a shared-memory candidate queue guarded by a running
//     threshold: a row is appended only when its 32-bit key (dist << 21 | row-in-slab)
//     beats the threshold; appends are warp-aggregated; a full queue is cut back to its
//     k best by a warp-level radix select.  Each cut also publishes (atomicMin) the
//     distance of the CTA's k-th best as a bound every other CTA of that query may prune
//     with - k rows at or below it exist, so nothing above it can be in the answer.
// The per-slab lists are merged by merge_pairs_kernel (merge.cuh).
//
// Roofline: algorithmic bytes per launch = n * W * 4 (one pass over the
// codes; the query tiles of one slab run back to back so re-reads hit L2); integer
// work = q * n * W word XOR+POPC.  HBM-bound for q <= 3, POPC-pipe-bound above.

#include "common.cuh"
#include "merge.cuh"

namespace rr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_ROW_BITS = 21;  // rows per slab <= 2^21
constexpr u32 SCAN_KEY_MAX = 0xFFFFFFFFu;

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

__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  u32 done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

struct Scan2Plan {
  int q_tile, n_qtiles, cap, slabs, stages;
  long long rows_per_slab;
  size_t smem;
};

constexpr int SCAN_TILE_ROWS = SCAN_THREADS;
constexpr size_t SCAN_SMEM_BUDGET = 112 * 1024;  // two CTAs per SM

static size_t scan2_smem(int words, int stages, int q_tile, int cap) {
  return (size_t)stages * SCAN_TILE_ROWS * words * 4 + (size_t)q_tile * ((size_t)words * 4 + (size_t)cap * 4 + 8) +
         SCAN_WARPS * 256 * sizeof(int) + 8 * (size_t)stages + 32;
}

static Scan2Plan plan_scan2(long long n, int words, int q, int k) {
  Scan2Plan p;
  p.cap = (int)align_up((size_t)k + 2 * SCAN_THREADS, 32);
  // memory-bound regime (a few queries): deep ring; compute-bound regime: 2 stages
  p.stages = (q <= 4) ? 4 : 2;
  int qt = 0;
  while (true) {
    const int want = q < 64 ? q : 64;
    qt = want;
    while (qt > 1 && scan2_smem(words, p.stages, qt, p.cap) > SCAN_SMEM_BUDGET) --qt;
    if (scan2_smem(words, p.stages, qt, p.cap) <= SCAN_SMEM_BUDGET || p.stages == 1) break;
    --p.stages;
  }
  p.n_qtiles = (q + qt - 1) / qt;
  p.q_tile = (q + p.n_qtiles - 1) / p.n_qtiles;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  // two resident CTAs per SM; one wave when memory-bound, two when compute-bound
  // whole waves only: slabs * n_qtiles <= target, rounded DOWN (a few CTAs over a wave
  // boundary cost a full extra wave)
  long long target = (q <= 4 ? 2LL : 4LL) * sms;
  long long slabs = target / p.n_qtiles;
  const long long max_slabs = (n + 2047) / 2048;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  const long long max_rps = 1LL << SCAN_ROW_BITS;
  long long rps;
  while (true) {
    rps = (n + slabs - 1) / slabs;
    rps = (long long)align_up((size_t)rps, SCAN_TILE_ROWS);
    if (rps <= max_rps) break;
    ++slabs;  // very large shards: more slabs than one wave, rows per slab capped by the key width
  }
  p.rows_per_slab = rps;
  p.slabs = (int)((n + rps - 1) / rps);
  if (p.slabs < 1) p.slabs = 1;
  p.smem = scan2_smem(words, p.stages, p.q_tile, p.cap);
  return p;
}

struct Scan2Args {
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
  int stages;
  long long rows_per_slab;
  u64* part_k1;  // [q][slabs][k]
  u32* part_k2;
  u32* gbound;   // [q] cross-CTA pruning bound (key space), starts at 0xFFFFFFFF
};

// popcount of W xor-words through one level of 3:2 carry-save adders
template <int W>
__device__ __forceinline__ int csa_popcount(const u32 (&x)[W]) {
  constexpr int T = W / 3;
  int ones = 0, twos = 0;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const u32 a = x[3 * t], b = x[3 * t + 1], c = x[3 * t + 2];
    const u32 s = a ^ b ^ c;                 // LOP3
    const u32 h = (a & b) | (c & (a ^ b));   // LOP3 (majority)
    ones += __popc(s);
    twos += __popc(h);
  }
#pragma unroll
  for (int w = 3 * T; w < W; ++w) ones += __popc(x[w]);
  return ones + 2 * twos;
}

template <int W>
__global__ void __launch_bounds__(SCAN_THREADS, 2) hamming_scan2_kernel(const Scan2Args a) {
  extern __shared__ __align__(128) unsigned char smem2_raw[];
  unsigned char* smem_raw = smem2_raw;
  constexpr int V = W / 4;                          // 16-byte chunks per row
  constexpr int TILE_BYTES = SCAN_TILE_ROWS * W * 4;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int slab = blockIdx.x / a.n_qtiles;
  const int qt = blockIdx.x % a.n_qtiles;
  const int q0 = qt * a.q_tile;
  const int nq = min(a.q_tile, a.q - q0);

  unsigned char* tiles = smem_raw;                                                     // [stages][TILE_BYTES]
  unsigned char* sq = tiles + (size_t)a.stages * TILE_BYTES;                           // [q_tile][W*4]
  u32* queue = reinterpret_cast<u32*>(sq + (size_t)a.q_tile * W * 4);                  // [q_tile][cap]
  u32* thr = queue + (size_t)a.q_tile * a.cap;                                         // [q_tile]
  int* cnt = reinterpret_cast<int*>(thr + a.q_tile);                                   // [q_tile]
  int* whist = cnt + a.q_tile;                                                         // [warps][256]
  u64* mbar = reinterpret_cast<u64*>(align_up_dev((size_t)(whist + SCAN_WARPS * 256), 8));  // [stages]

  const long long row_lo = (long long)slab * a.rows_per_slab;
  const long long row_hi = min(a.n, row_lo + a.rows_per_slab);
  const int ntiles = (int)((row_hi - row_lo + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS);

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(smem_addr(mbar + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < nq * V; i += SCAN_THREADS)
    reinterpret_cast<uint4*>(sq)[i] = reinterpret_cast<const uint4*>(a.qcodes)[(size_t)q0 * V + i];
  for (int i = tid; i < nq; i += SCAN_THREADS) {
    thr[i] = SCAN_KEY_MAX;
    cnt[i] = 0;
  }
  __syncthreads();

  auto issue_tile = [&](int t) {
    const int s = t % a.stages;
    const long long r0 = row_lo + (long long)t * SCAN_TILE_ROWS;
    const long long rows = min((long long)SCAN_TILE_ROWS, row_hi - r0);
    const u32 bytes = (u32)(rows * W * 4);
    const u32 bar = smem_addr(mbar + s);
    mbar_expect_tx(bar, bytes);
    bulk_copy_g2s(smem_addr(tiles + (size_t)s * TILE_BYTES), a.codes + (size_t)r0 * V, bytes, bar);
  };
  if (tid == 0) {
    const int pre = ntiles < a.stages ? ntiles : a.stages;
    for (int t = 0; t < pre; ++t) issue_tile(t);
  }

  // rotated chunk order: at step j this lane handles chunk (j + lane) mod V
  int coff[V];
  {
    const int l = lane % V;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      int c = l + j;
      if (c >= V) c -= V;
      coff[j] = c * 16;
    }
  }

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % a.stages;
    mbar_wait(smem_addr(mbar + s), (u32)((t / a.stages) & 1));
    const long long row = row_lo + (long long)t * SCAN_TILE_ROWS + tid;
    bool valid = row < row_hi;
    if (valid && a.tags != nullptr) valid = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
    uint4 r[V];
    {
      const unsigned char* src = tiles + (size_t)s * TILE_BYTES + (size_t)tid * (W * 4);
#pragma unroll
      for (int j = 0; j < V; ++j) r[j] = *reinterpret_cast<const uint4*>(src + coff[j]);
    }
    const u32 local = (u32)(row - row_lo);
    auto distance = [&](int qi) -> int {
      const unsigned char* qb = sq + (size_t)qi * (W * 4);
      u32 x[W];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const uint4 c = *reinterpret_cast<const uint4*>(qb + coff[j]);
        x[4 * j + 0] = r[j].x ^ c.x;
        x[4 * j + 1] = r[j].y ^ c.y;
        x[4 * j + 2] = r[j].z ^ c.z;
        x[4 * j + 3] = r[j].w ^ c.w;
      }
      return csa_popcount<W>(x);
    };
    auto offer = [&](int qi, int d) {
      const u32 key = ((u32)d << SCAN_ROW_BITS) | local;
      const bool pass = valid && (key < thr[qi]);
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if (bal) {
        int slot = 0;
        if (lane == (__ffs(bal) - 1)) slot = atomicAdd(&cnt[qi], __popc(bal));
        slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
        if (pass) queue[(size_t)qi * a.cap + slot + __popc(bal & ((1u << lane) - 1u))] = key;
      }
    };
    // two independent distance chains per iteration keep the POPC pipe busy across the
    // (rarely taken) append branches
    int qi = 0;
#pragma unroll 1
    for (; qi + 1 < nq; qi += 2) {
      const int d0 = distance(qi);
      const int d1 = distance(qi + 1);
      offer(qi, d0);
      offer(qi + 1, d1);
    }
    if (qi < nq) offer(qi, distance(qi));
    __syncthreads();  // every thread has its row in registers and is done appending
    if (tid == 0 && t + a.stages < ntiles) issue_tile(t + a.stages);  // refill the freed slot
    for (int qi = warp; qi < nq; qi += SCAN_WARPS) {
      const int c = cnt[qi];
      u32 tl = thr[qi];
      if (c > a.cap - SCAN_THREADS) {
        tl = warp_compact_queue(queue + (size_t)qi * a.cap, c, a.k, whist + warp * 256);
        if (lane == 0) {
          cnt[qi] = a.k;
          // k rows with dist <= dist(kth) exist: nothing with a larger dist can be in the answer
          atomicMin(a.gbound + q0 + qi, ((tl >> SCAN_ROW_BITS) + 1u) << SCAN_ROW_BITS);
        }
      }
      if (lane == 0) {
        const u32 g = *reinterpret_cast<volatile u32*>(a.gbound + q0 + qi);
        thr[qi] = tl < g ? tl : g;
      }
    }
    __syncthreads();
  }

  // final cut and write-out of this slab's list for each query of the tile
  for (int qi = warp; qi < nq; qi += SCAN_WARPS) {
    int c = cnt[qi];
    u32* qu = queue + (size_t)qi * a.cap;
    if (c > a.k) {
      const u32 kth = warp_compact_queue(qu, c, a.k, whist + warp * 256);
      c = a.k;
      if (lane == 0) atomicMin(a.gbound + q0 + qi, ((kth >> SCAN_ROW_BITS) + 1u) << SCAN_ROW_BITS);
    }
    __syncwarp();
    const u32 g = *reinterpret_cast<volatile u32*>(a.gbound + q0 + qi);
    const size_t o = ((size_t)(q0 + qi) * a.slabs + slab) * a.k;
    for (int j = lane; j < a.k; j += 32) {
      const u32 key = (j < c) ? qu[j] : SCAN_KEY_MAX;
      if (j < c && key < g) {
        a.part_k1[o + j] = (u64)(key >> SCAN_ROW_BITS);
        a.part_k2[o + j] = (u32)(row_lo + (key & ((1u << SCAN_ROW_BITS) - 1u)));
      } else {
        a.part_k1[o + j] = K1_INVALID;
        a.part_k2[o + j] = K2_INVALID;
      }
    }
  }
}

template <int W>
static int launch_scan2(const Scan2Args& a, const Scan2Plan& p, cudaStream_t st) {
  RR_CUDA(cudaFuncSetAttribute(hamming_scan2_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)p.smem));
  hamming_scan2_kernel<W><<<p.slabs * p.n_qtiles, SCAN_THREADS, p.smem, st>>>(a);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

__global__ void fill_missing_hamming_kernel(int* dist, long long* idx, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    dist[i] = 0x7fffffff;
    idx[i] = -1;
  }
}

}  // namespace rr

using namespace rr;

extern "C" size_t rr_hamming_topk_workspace_bytes(int64_t n, int32_t words, int32_t q, int32_t k) {
  if (n <= 0 || q <= 0 || k <= 0 || words <= 0) return 256;
  const Scan2Plan p2 = plan_scan2(n, words, q, k);
  const size_t slabs = (size_t)p2.slabs;
  return align_up((size_t)q * slabs * k * 8, 256) + align_up((size_t)q * slabs * k * 4, 256) +
         align_up((size_t)q * 4, 256) + 256;
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
  u64* part_k1 = nullptr;
  u32* part_k2 = nullptr;
  int slabs = 0;
  int rc = RR_ERR_INVALID;
  {
    const Scan2Plan p = plan_scan2(n, words, q, k);
    Scan2Args a;
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
    a.slabs = slabs = p.slabs;
    a.stages = p.stages;
    a.rows_per_slab = p.rows_per_slab;
    const size_t o1 = align_up((size_t)q * p.slabs * k * 8, 256);
    const size_t o2 = o1 + align_up((size_t)q * p.slabs * k * 4, 256);
    a.part_k1 = part_k1 = (u64*)workspace;
    a.part_k2 = part_k2 = (u32*)((char*)workspace + o1);
    a.gbound = (u32*)((char*)workspace + o2);
    RR_CUDA(cudaMemsetAsync(a.gbound, 0xFF, (size_t)q * 4, st));
    switch (words) {
      case 4: rc = launch_scan2<4>(a, p, st); break;
      case 8: rc = launch_scan2<8>(a, p, st); break;
      case 12: rc = launch_scan2<12>(a, p, st); break;
      case 16: rc = launch_scan2<16>(a, p, st); break;
      case 20: rc = launch_scan2<20>(a, p, st); break;
      case 24: rc = launch_scan2<24>(a, p, st); break;
      case 28: rc = launch_scan2<28>(a, p, st); break;
      case 32: rc = launch_scan2<32>(a, p, st); break;
      default: break;
    }
  }
  if (rc != RR_OK) return rc;
  MergeArgs m;
  m.k1 = part_k1;
  m.k2 = part_k2;
  m.n_in = (long long)slabs * k;
  m.k = k;
  m.cap = 0;
  m.row_base = row_base;
  m.out_a = out_dist;
  m.out_idx = (long long*)out_idx;
  m.out_count = nullptr;
  return launch_merge_pairs<MERGE_HAMMING>(m, q, st);
}

// ---- 8(e): merges of the allgathered per-shard lists --------------------------------
namespace rr {
// Hamming lists as they come out of ONE all_gather: keys [shards][q][k_in], key = dist << 40 | row,
// negative = padding.  No transpose copy: the CTA of query q walks its k_in entries of every shard.
__global__ void __launch_bounds__(MERGE_THREADS)
    merge_hamming_gathered_kernel(const long long* keys, int n_shards, int q_total, int k_in, int k, int cap,
                                  int* out_dist, long long* out_idx) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + cap);
  __shared__ SelectScratch<MERGE_THREADS> sc;
  const int q = blockIdx.x;
  auto get = [&](long long i, u64& x, u32& y) {
    const int g = (int)(i / k_in);
    const int j = (int)(i - (long long)g * k_in);
    const long long key = keys[((size_t)g * q_total + q) * k_in + j];
    x = key < 0 ? K1_INVALID : (u64)key;
    y = key < 0 ? K2_INVALID : 0u;
  };
  const int m = block_select_sorted<MERGE_THREADS>(get, (long long)n_shards * k_in, k, s_k1, s_k2, cap, sc);
  for (int j = threadIdx.x; j < k; j += MERGE_THREADS) {
    const bool have = j < m;
    merge_write<MERGE_HAMMING_PACKED>(out_dist, out_idx, (size_t)q * k + j, have, have ? s_k1[j] : 0,
                                      have ? s_k2[j] : 0, 0);
  }
}

// The same merge for SORTED shard lists (what rr_hamming_topk* + rr_pack_hamming leave: ascending keys, padding
// last): a tree of bitonic top-P merges instead of a radix select + full sort.  Lists are padded to P = 2^p
// >= k_in entries (+inf keys) and to G2 = 2^g >= n_shards lists in shared memory; one level merges list pairs
// in place: C[i] = min(A[i], B[P-1-i]) holds the P smallest of A u B as a bitonic sequence, log2(P) compare-
// exchange stages sort it.  log2(G2) * (1 + log2(P)) barriers in all (30 for 8 x 512) against several radix
// passes over every entry plus a 45-stage sort.
__global__ void __launch_bounds__(MERGE_THREADS)
    merge_hamming_tree_kernel(const long long* keys, int n_shards, int g2, int q_total, int k_in, int p_len, int k,
                              int* out_dist, long long* out_idx) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s = reinterpret_cast<u64*>(merge_smem);  // [g2][p_len]
  const int q = blockIdx.x;
  const int p_mask = p_len - 1;
  const int p_shift = 31 - __clz(p_len);
  for (int i = threadIdx.x; i < g2 * p_len; i += MERGE_THREADS) {
    const int g = i >> p_shift, j = i & p_mask;
    u64 v = ~0ull;
    if (g < n_shards && j < k_in) {
      const long long key = keys[((size_t)g * q_total + q) * k_in + j];
      if (key >= 0) v = (u64)key;
    }
    s[i] = v;
  }
  __syncthreads();
  const int half = p_len >> 1;
  const int h_shift = p_shift - 1;
  for (int w = 1; w < g2; w <<= 1) {  // lists at distance w merge into the lower one
    const int pairs = g2 / (2 * w);
    for (int i = threadIdx.x; i < pairs * p_len; i += MERGE_THREADS) {
      const int pr = i >> p_shift, j = i & p_mask;
      u64* a_list = s + (size_t)(2 * w * pr) * p_len;
      const u64 a = a_list[j], b = a_list[(size_t)w * p_len + (p_mask - j)];
      a_list[j] = a < b ? a : b;
    }
    __syncthreads();
    for (int st = half; st > 0; st >>= 1) {
      for (int i = threadIdx.x; i < pairs * half; i += MERGE_THREADS) {
        const int pr = i >> h_shift, t = i & (half - 1);
        const int lo = ((t & ~(st - 1)) << 1) | (t & (st - 1));
        u64* a_list = s + (size_t)(2 * w * pr) * p_len;
        const u64 a = a_list[lo], b = a_list[lo + st];
        if (a > b) {
          a_list[lo] = b;
          a_list[lo + st] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += MERGE_THREADS) {
    const u64 v = j < p_len ? s[j] : ~0ull;
    const bool have = v != ~0ull;
    merge_write<MERGE_HAMMING_PACKED>(out_dist, out_idx, (size_t)q * k + j, have, have ? v : 0, 0u, 0);
  }
}

// BM25 lists as the all_gather leaves them: words [n_shards][2][q][k_in], plane 0 = float64 score bits,
// plane 1 = global row (-1 = padding); (score desc, row asc), score > 0 was applied by the shards
__global__ void __launch_bounds__(MERGE_THREADS)
    merge_f64_gathered_kernel(const long long* words, int n_shards, int q_total, int k_in, int k, int cap,
                              double* out_score, long long* out_idx, int* out_count) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + cap);
  __shared__ SelectScratch<MERGE_THREADS> sc;
  const int q = blockIdx.x;
  auto get = [&](long long i, u64& x, u32& y) {
    const int g = (int)(i / k_in);
    const int j = (int)(i - (long long)g * k_in);
    const size_t base = (size_t)g * 2 * q_total * k_in + (size_t)q * k_in + j;
    const long long row = words[base + (size_t)q_total * k_in];
    if (row < 0) {
      x = K1_INVALID;
      y = K2_INVALID;
    } else {
      x = ~f64_orderable(__longlong_as_double(words[base]));
      y = (u32)row;
    }
  };
  const int m = block_select_sorted<MERGE_THREADS>(get, (long long)n_shards * k_in, k, s_k1, s_k2, cap, sc);
  for (int j = threadIdx.x; j < k; j += MERGE_THREADS) {
    const bool have = j < m;
    merge_write<MERGE_F64_DESC>(out_score, out_idx, (size_t)q * k + j, have, have ? s_k1[j] : 0, have ? s_k2[j] : 0, 0);
  }
  if (out_count && threadIdx.x == 0) out_count[q] = m;
}

// The BM25 merge as the same tree of bitonic top-P merges (sorted shard lists: score desc, row asc, padding last),
// on (key1 = ~orderable(score), key2 = row) pairs
__global__ void __launch_bounds__(MERGE_THREADS)
    merge_f64_tree_kernel(const long long* words, int n_shards, int g2, int q_total, int k_in, int p_len, int k,
                          double* out_score, long long* out_idx, int* out_count) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s1 = reinterpret_cast<u64*>(merge_smem);          // [g2][p_len]
  u32* s2 = reinterpret_cast<u32*>(s1 + (size_t)g2 * p_len);  // [g2][p_len]
  const int q = blockIdx.x;
  const int p_mask = p_len - 1;
  const int p_shift = 31 - __clz(p_len);
  for (int i = threadIdx.x; i < g2 * p_len; i += MERGE_THREADS) {
    const int g = i >> p_shift, j = i & p_mask;
    u64 x = K1_INVALID;
    u32 y = K2_INVALID;
    if (g < n_shards && j < k_in) {
      const size_t base = (size_t)g * 2 * q_total * k_in + (size_t)q * k_in + j;
      const long long row = words[base + (size_t)q_total * k_in];
      if (row >= 0) {
        x = ~f64_orderable(__longlong_as_double(words[base]));
        y = (u32)row;
      }
    }
    s1[i] = x;
    s2[i] = y;
  }
  __syncthreads();
  const int half = p_len >> 1;
  const int h_shift = p_shift - 1;
  for (int w = 1; w < g2; w <<= 1) {
    const int pairs = g2 / (2 * w);
    for (int i = threadIdx.x; i < pairs * p_len; i += MERGE_THREADS) {
      const int pr = i >> p_shift, j = i & p_mask;
      const size_t ia = (size_t)(2 * w * pr) * p_len + j;
      const size_t ib = (size_t)(2 * w * pr + w) * p_len + (p_mask - j);
      if (pair_less(s1[ib], s2[ib], s1[ia], s2[ia])) {
        s1[ia] = s1[ib];
        s2[ia] = s2[ib];
      }
    }
    __syncthreads();
    for (int st = half; st > 0; st >>= 1) {
      for (int i = threadIdx.x; i < pairs * half; i += MERGE_THREADS) {
        const int pr = i >> h_shift, t = i & (half - 1);
        const size_t lo = (size_t)(2 * w * pr) * p_len + (((t & ~(st - 1)) << 1) | (t & (st - 1)));
        const size_t hi = lo + st;
        const u64 a1 = s1[lo], b1 = s1[hi];
        const u32 a2 = s2[lo], b2 = s2[hi];
        if (pair_less(b1, b2, a1, a2)) {
          s1[lo] = b1;
          s2[lo] = b2;
          s1[hi] = a1;
          s2[hi] = a2;
        }
      }
      __syncthreads();
    }
  }
  int m = 0;
  for (int j0 = 0; j0 < k; j0 += MERGE_THREADS) {
    const int j = j0 + threadIdx.x;
    const bool have = j < k && j < p_len && s1[j] != K1_INVALID;
    if (j < k) merge_write<MERGE_F64_DESC>(out_score, out_idx, (size_t)q * k + j, have, have ? s1[j] : 0, have ? s2[j] : 0, 0);
    m += __syncthreads_count(have);
  }
  if (out_count && threadIdx.x == 0) out_count[q] = m;
}

__global__ void __launch_bounds__(256)
    pack_hamming_kernel(const int* dist, const long long* idx, long long n, long long* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const long long r = idx[i];
    out[i] = r < 0 ? -1LL : (((long long)dist[i] << 40) | r);
  }
}

}  // namespace rr

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

extern "C" int rr_pack_hamming(const int32_t* dist, const int64_t* idx, int64_t n, int64_t* out_keys,
                               void* stream) {
  RR_CHECK_ARG(n >= 0, "negative size");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(dist && idx && out_keys, "null pointer");
  pack_hamming_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      dist, (const long long*)idx, n, (long long*)out_keys);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_merge_hamming_gathered(const int64_t* in_keys, int32_t n_shards, int32_t q, int32_t k_in,
                                         int32_t k, int32_t* out_dist, int64_t* out_idx, void* stream) {
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(in_keys && out_dist && out_idx, "null pointer");
  RR_CHECK_ARG(q > 0 && n_shards > 0 && k_in > 0 && k >= 1 && k <= RR_MAX_K, "bad size");
  // sorted shard lists (the documented input): bitonic merge tree when it fits shared memory
  int p_len = 2, g2 = 1;
  while (p_len < k_in) p_len <<= 1;
  while (g2 < n_shards) g2 <<= 1;
  const size_t tree_smem = (size_t)g2 * p_len * 8;
  if (k <= p_len && p_len <= 1024 && tree_smem <= 96 * 1024) {
    RR_CUDA(cudaFuncSetAttribute(merge_hamming_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tree_smem));
    merge_hamming_tree_kernel<<<q, MERGE_THREADS, tree_smem, (cudaStream_t)stream>>>(
        (const long long*)in_keys, n_shards, g2, q, k_in, p_len, k, out_dist, (long long*)out_idx);
    RR_LAUNCH_CHECK();
    return RR_OK;
  }
  const int cap = merge_cap(k);
  merge_hamming_gathered_kernel<<<q, MERGE_THREADS, (size_t)cap * 12, (cudaStream_t)stream>>>(
      (const long long*)in_keys, n_shards, q, k_in, k, cap, out_dist, (long long*)out_idx);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_merge_scores_f64_gathered(const int64_t* in_words, int32_t n_shards, int32_t q, int32_t k_in,
                                            int32_t k, double* out_score, int64_t* out_idx,
                                            int32_t* out_count, void* stream) {
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(in_words && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(q > 0 && n_shards > 0 && k_in > 0 && k >= 1 && k <= RR_MAX_K, "bad size");
  int p_len = 2, g2 = 1;
  while (p_len < k_in) p_len <<= 1;
  while (g2 < n_shards) g2 <<= 1;
  const size_t tree_smem = (size_t)g2 * p_len * 12;
  if (k <= p_len && p_len <= 1024 && tree_smem <= 96 * 1024) {  // sorted shard lists: bitonic merge tree
    RR_CUDA(cudaFuncSetAttribute(merge_f64_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tree_smem));
    merge_f64_tree_kernel<<<q, MERGE_THREADS, tree_smem, (cudaStream_t)stream>>>(
        (const long long*)in_words, n_shards, g2, q, k_in, p_len, k, out_score, (long long*)out_idx, out_count);
    RR_LAUNCH_CHECK();
    return RR_OK;
  }
  const int cap = merge_cap(k);
  merge_f64_gathered_kernel<<<q, MERGE_THREADS, (size_t)cap * 12, (cudaStream_t)stream>>>(
      (const long long*)in_words, n_shards, q, k_in, k, cap, out_score, (long long*)out_idx, out_count);
  RR_LAUNCH_CHECK();
  return RR_OK;
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
