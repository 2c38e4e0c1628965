// R3: stage-2 rescoring of the stage-1 candidates.
//
// Replaces rescore_candidates (reference radiant/storage/quantization.py:185-222) and
// the caller's cut + threshold (radiant/storage/redis_store.py:850-854):
//   score_c = float32(dot(q_f32, float32(row_c)));  stable sort by score desc;
//   keep the first top_k, then those with score >= min_similarity.
//
// One CTA per query.  Candidate rows are different for every query, so this is a
// batched gather + GEMV (HBM-bound), not a GEMM: each warp streams one
// candidate row with 128-bit loads, accumulates the products in float64 (every
// f32*f32 and f32*int8 product is exact in double, so the sum is the correctly
// rounded dot product up to one final rounding to float32), and the CTA orders the
// (score, position) keys with a shared-memory bitonic sort.
// Algorithmic bytes: q * c * dim * sizeof(row element) + q * dim * 4.
#include <math.h>

#include "common.cuh"
#include "select.cuh"
#include "tc_ptx.cuh"

namespace rr {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef RR_RS_ROWS
#define RR_RS_ROWS 2
#endif
#ifndef RR_RS_UNROLL
#define RR_RS_UNROLL 3
#endif
constexpr int RS_ROWS = RR_RS_ROWS;      // candidate rows a warp scores per iteration
constexpr int RS_UNROLL = RR_RS_UNROLL;  // column groups (of 32 float4) whose loads are issued together

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot(q (shared, f32), row (global f32)) in float64, fixed summation order
__device__ __forceinline__ double dot_f32_row(const float* sq, const float* row, int dim, int lane) {
  double acc = 0.0;
  if ((dim & 3) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    for (int v = lane; v < (dim >> 2); v += 32) {
      const float4 e = __ldg(r4 + v);
      const float4 w = q4[v];
      acc += (double)w.x * (double)e.x;
      acc += (double)w.y * (double)e.y;
      acc += (double)w.z * (double)e.z;
      acc += (double)w.w * (double)e.w;
    }
  } else {
    for (int d = lane; d < dim; d += 32) acc += (double)sq[d] * (double)__ldg(row + d);
  }
  return warp_sum_f64(acc);
}

__device__ __forceinline__ double dot_i8_row(const float* sq, const int8_t* row, int dim, int lane) {
  double acc = 0.0;
  if ((dim & 3) == 0) {
    // lane-contiguous 4-byte groups: 128 B per warp request, query read as float4 (conflict-free)
    const int* r4 = reinterpret_cast<const int*>(row);
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    for (int v = lane; v < (dim >> 2); v += 32) {
      const int e = __ldg(r4 + v);
      const float4 w = q4[v];
      acc += (double)w.x * (double)(int)(signed char)(e & 0xFF);
      acc += (double)w.y * (double)(int)(signed char)((e >> 8) & 0xFF);
      acc += (double)w.z * (double)(int)(signed char)((e >> 16) & 0xFF);
      acc += (double)w.w * (double)(int)(signed char)((e >> 24) & 0xFF);
    }
  } else {
    for (int d = lane; d < dim; d += 32) acc += (double)sq[d] * (double)(int)__ldg(row + d);
  }
  return warp_sum_f64(acc);
}

struct RescoreArgs {
  const float* queries;
  int dim;
  const void* emb;
  long long n;
  long long row_base;
  const long long* cand_idx;
  int c;
  int top_k;
  double min_similarity;
  float* out_score;
  long long* out_idx;
  int* out_count;
};

// Bitonic sort of p u64 keys ascending in shared memory.
__device__ __forceinline__ void block_bitonic_sort_u64(u64* keys, int p) {
  for (int size = 2; size <= p; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (p >> 1); t += RS_THREADS) {
        const int i = ((t / stride) * (stride << 1)) + (t % stride);
        const int j = i + stride;
        const bool asc = ((i & size) == 0);
        const u64 x = keys[i], y = keys[j];
        if ((x > y) == asc) {
          keys[i] = y;
          keys[j] = x;
        }
      }
      __syncthreads();
    }
  }
}

// keys[] hold (~orderable(score) << 32 | position) for valid candidates; sorts them and
// writes the reference's cut + filter.
__device__ __forceinline__ void rank_and_write_f32(u64* keys, int c, int p, const long long* cand,
                                                   int top_k, double min_sim, float* out_score,
                                                   long long* out_idx, int* out_count, int* s_cnt) {
  for (int i = c + threadIdx.x; i < p; i += RS_THREADS) keys[i] = K1_INVALID;
  if (threadIdx.x == 0) *s_cnt = 0;
  __syncthreads();
  block_bitonic_sort_u64(keys, p);
  // scores are descending, so the entries >= min_sim form a prefix of the first top_k
  for (int j = threadIdx.x; j < top_k; j += RS_THREADS) {
    bool have = false;
    float s = 0.0f;
    long long id = -1;
    if (j < c) {
      const u64 key = keys[j];
      if (key != K1_INVALID) {
        s = f32_from_orderable(~(u32)(key >> 32));
        if ((double)s >= min_sim) {
          have = true;
          id = cand[(u32)key];
        }
      }
    }
    out_score[j] = have ? s : 0.0f;
    out_idx[j] = have ? id : -1;
    if (have) atomicAdd(s_cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0 && out_count) *out_count = *s_cnt;
}

// MODE 0: full rescore (score, sort, cut, filter).  MODE 1: scores only.
template <int EMB, int MODE>
__global__ void __launch_bounds__(RS_THREADS) rescore_f32_kernel(const RescoreArgs a, int p) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  float* sq = reinterpret_cast<float*>(rs_smem);
  u64* keys = reinterpret_cast<u64*>(rs_smem + align_up_dev((size_t)a.dim * 4));
  __shared__ int s_cnt;
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < a.dim; d += RS_THREADS) sq[d] = a.queries[(size_t)q * a.dim + d];
  __syncthreads();
  const long long* cand = a.cand_idx + (size_t)q * a.c;
  // RS_ROWS candidate rows per warp iteration: the 128-bit loads of all of them are in flight
  // together, which is what the latency-bound random row gather needs
  // MODE 1 may split a query's candidates over gridDim.y CTAs (more rows in flight per SM)
  const int gw = blockIdx.y * RS_WARPS + warp, tw = gridDim.y * RS_WARPS;
  for (int cb = gw; cb < a.c; cb += RS_ROWS * tw) {
    int ci[RS_ROWS];
    long long loc[RS_ROWS];
    bool ok[RS_ROWS];
    double acc[RS_ROWS];
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
      ci[r] = cb + r * tw;
      const long long idx = (ci[r] < a.c) ? cand[ci[r]] : -1;
      loc[r] = idx - a.row_base;
      ok[r] = idx >= 0 && loc[r] >= 0 && loc[r] < a.n;
      acc[r] = 0.0;
    }
    if (EMB == RR_F32 && (a.dim & 3) == 0) {
      const float4* rp[RS_ROWS];
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r)
        rp[r] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.emb) + (size_t)(ok[r] ? loc[r] : 0) * a.dim);
      const float4* q4 = reinterpret_cast<const float4*>(sq);
      // RS_UNROLL column groups per step: all RS_ROWS * RS_UNROLL 128-bit loads are issued before the
      // first add (the gather is latency-bound); the adds keep their ascending column order per row
      const int nv = a.dim >> 2;
      for (int vb = lane; vb < nv; vb += 32 * RS_UNROLL) {
        float4 e[RS_UNROLL][RS_ROWS];
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) {
          const int v = vb + 32 * u;
#pragma unroll
          for (int r = 0; r < RS_ROWS; ++r)
            e[u][r] = (ok[r] && v < nv) ? __ldg(rp[r] + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) {
          const int v = vb + 32 * u;
          if (v < nv) {
            const float4 w = q4[v];
            const double wx = (double)w.x, wy = (double)w.y, wz = (double)w.z, ww = (double)w.w;
#pragma unroll
            for (int r = 0; r < RS_ROWS; ++r) {  // same order of additions per row as one row at a time
              acc[r] += wx * (double)e[u][r].x;
              acc[r] += wy * (double)e[u][r].y;
              acc[r] += wz * (double)e[u][r].z;
              acc[r] += ww * (double)e[u][r].w;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r) acc[r] = warp_sum_f64(acc[r]);
    } else {
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r)
        if (ok[r])
          acc[r] = (EMB == RR_F32)
                       ? dot_f32_row(sq, reinterpret_cast<const float*>(a.emb) + (size_t)loc[r] * a.dim, a.dim, lane)
                       : dot_i8_row(sq, reinterpret_cast<const int8_t*>(a.emb) + (size_t)loc[r] * a.dim, a.dim, lane);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r) {
        if (ci[r] >= a.c) continue;
        const float sc = ok[r] ? (float)acc[r] : -INFINITY;
        if (MODE == 0)
          keys[ci[r]] = ok[r] ? (((u64)(~f32_orderable(sc)) << 32) | (u64)(u32)ci[r]) : K1_INVALID;
        else
          a.out_score[(size_t)q * a.c + ci[r]] = sc;
      }
    }
  }
  if (MODE == 0) {
    __syncthreads();
    rank_and_write_f32(keys, a.c, p, cand, a.top_k, a.min_similarity,
                       a.out_score + (size_t)q * a.top_k, a.out_idx + (size_t)q * a.top_k,
                       a.out_count ? a.out_count + q : nullptr, &s_cnt);
  }
}

// Batched scoring with the candidate rows STAGED through shared memory by the bulk-copy engine.
// The gather is a random read of contiguous rows (3 KB float32 / 1 KB int8 at the BASELINE
// shapes); what limits it is the number of bytes a CTA keeps in flight, and with register-staged
// loads that is bounded by registers (the kernel above reaches ~40 % of HBM).  Here one elected
// thread issues cp.async.bulk (global -> shared, mbarrier complete_tx) for up to `stages` rows ahead
// of the seven consumer warps; bytes in flight are bounded by shared memory only.  A consumer warp
// owns a row (and, the ring depth being a multiple of seven, always the same ring slots): same lane-strided float64 accumulation and the same shuffle tree as
// dot_f32_row / dot_i8_row, so the scores are bit-identical to the register-staged kernel.
constexpr int RG_THREADS = 256;
constexpr int RG_CONSUMERS = RG_THREADS / 32 - 1;
constexpr int RG_MAX_STAGES = 32;

// The [q][c] candidate slots are cut into EQUAL contiguous ranges, one per CTA, regardless of query
// boundaries (a CTA keeps the <= nqv query vectors its range touches in shared memory): with two
// resident CTAs per SM and exactly 2 x SMs x m ranges there is no partial last wave.
// COPY_ONLY (rr_probe_gather): the consumers release every row unread - the kernel then measures what
// the memory system delivers for this access pattern (random rows of row_bytes).
template <int EMB, bool COPY_ONLY>
__global__ void __launch_bounds__(RG_THREADS) rescore_ring_kernel(const RescoreArgs a, int stages, int row_bytes, int per,
                                                                   int nqv, long long total) {
  extern __shared__ __align__(128) unsigned char rg_smem[];
  unsigned char* ring = rg_smem;                                               // [stages][row_bytes]
  float* sq = reinterpret_cast<float*>(ring + (size_t)stages * row_bytes);     // [nqv][dim4 * 4]
  const int dimp = (int)align_up_dev((size_t)a.dim, 4);
  long long* s_loc = reinterpret_cast<long long*>(sq + (size_t)nqv * dimp);    // [per]
  int* s_slot = reinterpret_cast<int*>(s_loc + per);                           // [per] flattened slot q * c + ci
  __shared__ __align__(8) u64 full[RG_MAX_STAGES];
  __shared__ __align__(8) u64 empty[RG_MAX_STAGES];
  __shared__ int s_nvalid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long lo = (long long)blockIdx.x * per;
  const long long hi = min(total, lo + per);
  if (lo >= hi) return;
  const int q_first = (int)(lo / a.c);
  const int q_last = (int)((hi - 1) / a.c);
  if (threadIdx.x == 0) {
    s_nvalid = 0;
    for (int s = 0; s < stages; ++s) {
      tc_mbar_init(tc_smem(full + s), 1);
      tc_mbar_init(tc_smem(empty + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (q_last - q_first + 1) * a.dim; i += RG_THREADS) {
    const int j = i / a.dim, d = i - j * a.dim;
    sq[(size_t)j * dimp + d] = a.queries[(size_t)(q_first + j) * a.dim + d];
  }
  __syncthreads();
  for (long long sl = lo + threadIdx.x; sl < hi; sl += RG_THREADS) {
    const long long idx = a.cand_idx[sl];
    const long long loc = idx - a.row_base;
    if (idx >= 0 && loc >= 0 && loc < a.n) {
      const int slot = atomicAdd(&s_nvalid, 1);  // rows are independent: their order is free
      s_loc[slot] = loc;
      s_slot[slot] = (int)sl;
    } else {
      a.out_score[sl] = -INFINITY;  // not owned by this shard / padding
    }
  }
  __syncthreads();
  const int nv = s_nvalid;
  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* base = reinterpret_cast<const unsigned char*>(a.emb);
      for (int j = 0; j < nv; ++j) {
        const int s = j % stages;
        const int f = j / stages;
        if (f > 0) tc_mbar_wait(tc_smem(empty + s), (u32)(f - 1) & 1u);
        tc_mbar_expect_tx(tc_smem(full + s), (u32)row_bytes);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(tc_smem(ring + (size_t)s * row_bytes)), "l"(base + (size_t)s_loc[j] * row_bytes), "r"((u32)row_bytes),
              "r"(tc_smem(full + s)) : "memory");
      }
    }
  } else {
    const int nv4 = a.dim >> 2;
    for (int j = warp - 1; j < nv; j += RG_CONSUMERS) {
      const int s = j % stages;
      const int slot = s_slot[j];
      const float4* q4 = reinterpret_cast<const float4*>(sq + (size_t)(slot / a.c - q_first) * dimp);
      tc_mbar_wait(tc_smem(full + s), (u32)(j / stages) & 1u);
      double acc = 0.0;
      if (COPY_ONLY) {
        acc = (double)reinterpret_cast<const float*>(ring + (size_t)s * row_bytes)[lane];
      } else if (EMB == RR_F32) {
        const float4* r4 = reinterpret_cast<const float4*>(ring + (size_t)s * row_bytes);
        for (int v = lane; v < nv4; v += 32) {
          const float4 e = r4[v];
          const float4 w = q4[v];
          acc += (double)w.x * (double)e.x;
          acc += (double)w.y * (double)e.y;
          acc += (double)w.z * (double)e.z;
          acc += (double)w.w * (double)e.w;
        }
      } else {
        const int* r4 = reinterpret_cast<const int*>(ring + (size_t)s * row_bytes);
        for (int v = lane; v < nv4; v += 32) {
          const int e = r4[v];
          const float4 w = q4[v];
          acc += (double)w.x * (double)(int)(signed char)(e & 0xFF);
          acc += (double)w.y * (double)(int)(signed char)((e >> 8) & 0xFF);
          acc += (double)w.z * (double)(int)(signed char)((e >> 16) & 0xFF);
          acc += (double)w.w * (double)(int)(signed char)((e >> 24) & 0xFF);
        }
      }
      __syncwarp();  // every lane has read its part of the slot
      if (lane == 0) tc_mbar_arrive(tc_smem(empty + s));
      acc = warp_sum_f64(acc);
      if (lane == 0) a.out_score[slot] = (float)acc;
    }
  }
}

__global__ void __launch_bounds__(RS_THREADS)
    rank_scored_f32_kernel(const float* scores, const long long* cand_idx, int c, int p, int top_k,
                           double min_sim, float* out_score, long long* out_idx, int* out_count) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  u64* keys = reinterpret_cast<u64*>(rs_smem);
  __shared__ int s_cnt;
  const int q = blockIdx.x;
  const long long* cand = cand_idx + (size_t)q * c;
  for (int ci = threadIdx.x; ci < c; ci += RS_THREADS) {
    const float s = scores[(size_t)q * c + ci];
    const bool valid = cand[ci] >= 0 && s != -INFINITY;
    keys[ci] = valid ? (((u64)(~f32_orderable(s)) << 32) | (u64)(u32)ci) : K1_INVALID;
  }
  __syncthreads();
  rank_and_write_f32(keys, c, p, cand, top_k, min_sim, out_score + (size_t)q * top_k,
                     out_idx + (size_t)q * top_k, out_count ? out_count + q : nullptr, &s_cnt);
}

// symmetric int8 x int8 -> int32 (extension mode), exact
__global__ void __launch_bounds__(RS_THREADS)
    rescore_i8_kernel(const int8_t* queries, int dim, const int8_t* emb, long long n,
                      long long row_base, const long long* cand_idx, int c, int p, int top_k,
                      int* out_score, long long* out_idx, int* out_count) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  int8_t* sq = reinterpret_cast<int8_t*>(rs_smem);
  u64* keys = reinterpret_cast<u64*>(rs_smem + align_up_dev((size_t)dim));
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < dim; d += RS_THREADS) sq[d] = queries[(size_t)q * dim + d];
  __syncthreads();
  const long long* cand = cand_idx + (size_t)q * c;
  for (int ci = warp; ci < c; ci += RS_WARPS) {
    const long long idx = cand[ci];
    const long long local = idx - row_base;
    const bool valid = idx >= 0 && local >= 0 && local < n;
    int acc = 0;
    if (valid) {
      const int8_t* row = emb + (size_t)local * dim;
      if ((dim & 15) == 0) {
        const int4* r16 = reinterpret_cast<const int4*>(row);
        const int4* q16 = reinterpret_cast<const int4*>(sq);
        for (int v = lane; v < (dim >> 4); v += 32) {
          const int4 e = __ldg(r16 + v);
          const int4 w = q16[v];
          acc = __dp4a(e.x, w.x, acc);
          acc = __dp4a(e.y, w.y, acc);
          acc = __dp4a(e.z, w.z, acc);
          acc = __dp4a(e.w, w.w, acc);
        }
      } else {
        for (int d = lane; d < dim; d += 32) acc += (int)row[d] * (int)sq[d];
      }
      acc = warp_sum_i32(acc);
    }
    if (lane == 0)
      keys[ci] = valid ? (((u64)(~i32_orderable(acc)) << 32) | (u64)(u32)ci) : K1_INVALID;
  }
  for (int i = c + threadIdx.x; i < p; i += RS_THREADS) keys[i] = K1_INVALID;
  __syncthreads();
  block_bitonic_sort_u64(keys, p);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < top_k; j += RS_THREADS) {
    const size_t o = (size_t)q * top_k + j;
    bool have = false;
    if (j < c) {
      const u64 key = keys[j];
      if (key != K1_INVALID) {
        have = true;
        out_score[o] = i32_from_orderable(~(u32)(key >> 32));
        out_idx[o] = cand[(u32)key];
        atomicAdd(&s_cnt, 1);
      }
    }
    if (!have) {
      out_score[o] = (int)0x80000000;
      out_idx[o] = -1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && out_count) out_count[q] = s_cnt;
}

static int pow2_at_least(int c) {
  int p = 1;
  while (p < c) p <<= 1;
  return p;
}

}  // namespace rr

using namespace rr;

static int check_rescore_args(int32_t q, int32_t dim, int32_t c, int32_t top_k) {
  RR_CHECK_ARG(q >= 0 && dim > 0, "bad size");
  RR_CHECK_ARG(c >= 1 && c <= RR_MAX_K, "candidate count out of range");
  RR_CHECK_ARG(top_k >= 1 && top_k <= RR_MAX_K, "top_k out of range");
  RR_CHECK_ARG(dim <= 8192, "dim > 8192 unsupported");
  return RR_OK;
}

extern "C" int rr_rescore_f32(const float* queries, int32_t q, int32_t dim, const void* emb,
                              int32_t emb_dtype, int64_t n, int64_t row_base,
                              const int64_t* cand_idx, int32_t c, int32_t top_k,
                              double min_similarity, float* out_score, int64_t* out_idx,
                              int32_t* out_count, void* stream) {
  int rc = check_rescore_args(q, dim, c, top_k);
  if (rc != RR_OK) return rc;
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(queries && cand_idx && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(emb || n == 0, "emb is null");
  RR_CHECK_ARG(emb_dtype == RR_F32 || emb_dtype == RR_I8, "emb_dtype must be RR_F32 or RR_I8");
  RescoreArgs a{queries, dim, emb, n, row_base, (const long long*)cand_idx, c, top_k,
                min_similarity, out_score, (long long*)out_idx, out_count};
  const int p = pow2_at_least(c);
  const size_t smem = align_up((size_t)dim * 4, 16) + (size_t)p * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (emb_dtype == RR_F32)
    rescore_f32_kernel<RR_F32, 0><<<q, RS_THREADS, smem, st>>>(a, p);
  else
    rescore_f32_kernel<RR_I8, 0><<<q, RS_THREADS, smem, st>>>(a, p);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

// ring launch shared by rr_score_candidates_f32 and rr_probe_gather; returns RR_OK, or -1 when the
// shape does not fit the ring (the caller then uses the register-staged kernel)
template <bool COPY_ONLY>
static int launch_rescore_ring(const RescoreArgs& a, int q, int emb_dtype, cudaStream_t st) {
  const int dim = a.dim, c = a.c;
  const int row_bytes = emb_dtype == RR_F32 ? dim * 4 : dim;
  if (!((dim & 3) == 0 && (row_bytes & 15) == 0 && ((size_t)a.emb & 15) == 0 && a.n > 0)) return -1;
  // equal slot ranges, 2 x SMs x m of them (two CTAs per SM are resident): m is the smallest
  // multiplier for which a range touches <= 8 queries and holds <= 2048 slots
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const long long total = (long long)q * c;
  const int dimp = (int)align_up((size_t)dim, 4);
  int per = 0, nqv = 0;
  for (int m = 1; m <= 64; ++m) {
    const long long n_cta = 2LL * sms * m;
    per = (int)((total + n_cta - 1) / n_cta);
    if (per < 4 * RG_CONSUMERS) per = 4 * RG_CONSUMERS;  // tiny calls: fewer, fuller CTAs
    nqv = (per - 1) / c + 2;
    if (nqv > q) nqv = q;
    if (nqv <= 8 && per <= 2048) break;
  }
  const long long n_cta = (total + per - 1) / per;
  const size_t fixed = (size_t)nqv * dimp * 4 + (size_t)per * 12 + 16;
  // a ring slot must always be consumed by the SAME warp (row j -> slot j % stages, warp j % 7):
  // a parity wait is only meaningful for a waiter that has seen the previous phase complete,
  // so the ring depth is a multiple of the consumer count
  int stages = fixed < 100 * 1024 ? (int)((100 * 1024 - fixed) / row_bytes) : 0;
  if (stages > RG_MAX_STAGES) stages = RG_MAX_STAGES;
  stages = stages / RG_CONSUMERS * RG_CONSUMERS;
  if (!(stages >= RG_CONSUMERS && nqv <= 8 && per <= 2048 && total < (1LL << 31))) return -1;
  const size_t rsm = (size_t)stages * row_bytes + fixed;
  if (emb_dtype == RR_F32) {
    RR_CUDA(cudaFuncSetAttribute(rescore_ring_kernel<RR_F32, COPY_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));
    rescore_ring_kernel<RR_F32, COPY_ONLY><<<(unsigned)n_cta, RG_THREADS, rsm, st>>>(a, stages, row_bytes, per, nqv, total);
  } else {
    RR_CUDA(cudaFuncSetAttribute(rescore_ring_kernel<RR_I8, COPY_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));
    rescore_ring_kernel<RR_I8, COPY_ONLY><<<(unsigned)n_cta, RG_THREADS, rsm, st>>>(a, stages, row_bytes, per, nqv, total);
  }
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_score_candidates_f32(const float* queries, int32_t q, int32_t dim, const void* emb,
                                       int32_t emb_dtype, int64_t n, int64_t row_base,
                                       const int64_t* cand_idx, int32_t c, float* out_score,
                                       void* stream) {
  int rc = check_rescore_args(q, dim, c, 1);
  if (rc != RR_OK) return rc;
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(queries && cand_idx && out_score, "null pointer");
  RR_CHECK_ARG(emb || n == 0, "emb is null");
  RR_CHECK_ARG(emb_dtype == RR_F32 || emb_dtype == RR_I8, "emb_dtype must be RR_F32 or RR_I8");
  RescoreArgs a{queries, dim, emb, n, row_base, (const long long*)cand_idx, c, 1,
                0.0, out_score, nullptr, nullptr};
  cudaStream_t st = (cudaStream_t)stream;
  // rows staged through shared memory by the bulk-copy engine when they are 16-byte multiples
  // (float32: dim % 4 == 0, int8: dim % 16 == 0) at 16-byte aligned addresses
  rc = launch_rescore_ring<false>(a, q, emb_dtype, st);
  if (rc != -1) return rc;
  const size_t smem = align_up((size_t)dim * 4, 16) + 8;
  // enough CTAs for ~4 per SM, at least one warp iteration of work each
  int split = (4 * 148 + q - 1) / q;
  const int max_split = (c + RS_ROWS * RS_WARPS - 1) / (RS_ROWS * RS_WARPS);
  if (split > max_split) split = max_split;
  if (split > 8) split = 8;
  if (split < 1) split = 1;
  dim3 grid(q, split);
  if (emb_dtype == RR_F32)
    rescore_f32_kernel<RR_F32, 1><<<grid, RS_THREADS, smem, st>>>(a, 1);
  else
    rescore_f32_kernel<RR_I8, 1><<<grid, RS_THREADS, smem, st>>>(a, 1);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_rank_scored_f32(const float* scores, const int64_t* cand_idx, int32_t q, int32_t c,
                                  int32_t top_k, double min_similarity, float* out_score,
                                  int64_t* out_idx, int32_t* out_count, void* stream) {
  int rc = check_rescore_args(q, 1, c, top_k);
  if (rc != RR_OK) return rc;
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(scores && cand_idx && out_score && out_idx, "null pointer");
  const int p = pow2_at_least(c);
  rank_scored_f32_kernel<<<q, RS_THREADS, (size_t)p * 8, (cudaStream_t)stream>>>(
      scores, (const long long*)cand_idx, c, p, top_k, min_similarity, out_score,
      (long long*)out_idx, out_count);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_rescore_i8(const int8_t* queries_i8, int32_t q, int32_t dim, const int8_t* emb,
                             int64_t n, int64_t row_base, const int64_t* cand_idx, int32_t c,
                             int32_t top_k, int32_t* out_score, int64_t* out_idx,
                             int32_t* out_count, void* stream) {
  int rc = check_rescore_args(q, dim, c, top_k);
  if (rc != RR_OK) return rc;
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(queries_i8 && cand_idx && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(emb || n == 0, "emb is null");
  const int p = pow2_at_least(c);
  const size_t smem = align_up((size_t)dim, 16) + (size_t)p * 8;
  rescore_i8_kernel<<<q, RS_THREADS, smem, (cudaStream_t)stream>>>(
      queries_i8, dim, emb, n, row_base, (const long long*)cand_idx, c, p, top_k, out_score,
      (long long*)out_idx, out_count);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

// Measurement aid (tools/peak_probe.py): the gather of rr_score_candidates_f32 with the arithmetic
// left out - bytes per second the memory system delivers for q * c random rows of this index.
extern "C" int rr_probe_gather(const void* emb, int32_t emb_dtype, int64_t n, int32_t dim, const int64_t* cand_idx,
                               int32_t q, int32_t c, float* scratch_scores, double* out_bytes_per_s, void* stream) {
  RR_CHECK_ARG(emb && cand_idx && scratch_scores && out_bytes_per_s && q > 0 && c > 0 && n > 0, "bad argument");
  RR_CHECK_ARG(emb_dtype == RR_F32 || emb_dtype == RR_I8, "emb_dtype must be RR_F32 or RR_I8");
  cudaStream_t st = (cudaStream_t)stream;
  // the query vectors are loaded but never used: any readable buffer of q * dim floats will do
  RescoreArgs a{reinterpret_cast<const float*>(emb), dim, emb, n, 0, (const long long*)cand_idx, c, 1,
                0.0, scratch_scores, nullptr, nullptr};
  RR_CHECK_ARG((long long)q * dim <= n * (long long)(emb_dtype == RR_F32 ? dim : dim / 4), "index too small to stand in for the queries");
  cudaEvent_t e0, e1;
  RR_CUDA(cudaEventCreate(&e0));
  RR_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    RR_CUDA(cudaEventRecord(e0, st));
    int rc = launch_rescore_ring<true>(a, q, emb_dtype, st);
    if (rc != RR_OK) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      if (rc == -1) set_error("rr_probe_gather: shape does not fit the ring kernel");
      return rc == -1 ? RR_ERR_INVALID : rc;
    }
    RR_CUDA(cudaEventRecord(e1, st));
    RR_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *out_bytes_per_s = (double)q * c * (emb_dtype == RR_F32 ? dim * 4.0 : (double)dim) / (best * 1e-3);
  return RR_OK;
}
