// R5 and BASELINE config 4: exact (non-quantised-stage) scans.
//
// rr_exact_search_f32 replaces RedisVectorStore._retrieve_by_embedding_linear
// (reference radiant/storage/redis_store.py:863-952): cosine of the query against
// every row, rows with zero norm skipped, doc_level filter, keep >= min_similarity,
// order (score desc, row asc), top_k.
// rr_int8_search_topk is the exact int8 x int8 -> int32 search of config 4 with the
// same ordering.
//
// First-correct design (round 1): a warp streams one row with 128-bit loads and keeps
// EX_QT query accumulators in registers, so every row is read from HBM once per tile
// of EX_QT queries; scores are written as 32-bit order-preserving keys [q, n] and the
// per-query top-k is taken by block_select_sorted.  float32 rows accumulate in
// float64 (correctly rounded cosine); int8 rows use DP4A (exact).
// Algorithmic bytes per launch: n * dim * sizeof(elem) + 4 * n * EX_QT written.
#include <math.h>

#include "common.cuh"
#include "merge.cuh"

namespace rr {

constexpr int EX_THREADS = 256;
constexpr int EX_WARPS = EX_THREADS / 32;
constexpr int EX_QT = 8;
constexpr u32 KEY32_INVALID = 0xFFFFFFFFu;

struct ExactArgs {
  const void* emb;
  long long n;
  int dim;
  const uint8_t* tags;
  unsigned tag_mask, tag_value;
  const void* queries;
  int q;
  double min_similarity;
  u32* keys;  // [q][n]
};

__global__ void __launch_bounds__(EX_THREADS) exact_f32_scores_kernel(const ExactArgs a) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  float* sq = reinterpret_cast<float*>(ex_smem);  // [EX_QT][dim]
  __shared__ double s_qnorm[EX_QT];
  const int q0 = blockIdx.y * EX_QT;
  const int nq = min(EX_QT, a.q - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* queries = reinterpret_cast<const float*>(a.queries);
  for (int i = threadIdx.x; i < EX_QT * a.dim; i += EX_THREADS) {
    const int qi = i / a.dim;
    sq[i] = (qi < nq) ? queries[(size_t)(q0 + qi) * a.dim + (i % a.dim)] : 0.0f;
  }
  __syncthreads();
  if (warp < EX_QT) {
    double s = 0.0;
    for (int d = lane; d < a.dim; d += 32) {
      const double v = (double)sq[warp * a.dim + d];
      s += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qnorm[warp] = sqrt(s);
  }
  __syncthreads();
  const float* emb = reinterpret_cast<const float*>(a.emb);
  const bool vec = (a.dim & 3) == 0;
  const long long warps_total = (long long)gridDim.x * EX_WARPS;
  for (long long row = (long long)blockIdx.x * EX_WARPS + warp; row < a.n; row += warps_total) {
    bool valid = true;
    if (a.tags) valid = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
    double acc[EX_QT];
    double nn = 0.0;
#pragma unroll
    for (int j = 0; j < EX_QT; ++j) acc[j] = 0.0;
    if (valid) {
      const float* r = emb + (size_t)row * a.dim;
      if (vec) {
        const float4* r4 = reinterpret_cast<const float4*>(r);
        for (int v = lane; v < (a.dim >> 2); v += 32) {
          const float4 e = __ldg(r4 + v);
          const double ex = e.x, ey = e.y, ez = e.z, ew = e.w;
          nn += ex * ex;
          nn += ey * ey;
          nn += ez * ez;
          nn += ew * ew;
#pragma unroll
          for (int j = 0; j < EX_QT; ++j) {
            const float4 w = reinterpret_cast<const float4*>(sq + j * a.dim)[v];
            acc[j] += (double)w.x * ex;
            acc[j] += (double)w.y * ey;
            acc[j] += (double)w.z * ez;
            acc[j] += (double)w.w * ew;
          }
        }
      } else {
        for (int d = lane; d < a.dim; d += 32) {
          const double e = (double)__ldg(r + d);
          nn += e * e;
#pragma unroll
          for (int j = 0; j < EX_QT; ++j) acc[j] += (double)sq[j * a.dim + d] * e;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
#pragma unroll
      for (int j = 0; j < EX_QT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane < nq) {
      double dotv = 0.0;
#pragma unroll
      for (int j = 0; j < EX_QT; ++j)
        if (j == lane) dotv = acc[j];
      const double qn = s_qnorm[lane];
      u32 key = KEY32_INVALID;
      if (valid && nn > 0.0 && qn > 0.0) {
        const float s = (float)(dotv / (sqrt(nn) * qn));
        if ((double)s >= a.min_similarity) key = ~f32_orderable(s);
      }
      a.keys[(size_t)(q0 + lane) * a.n + row] = key;
    }
  }
}

__global__ void __launch_bounds__(EX_THREADS) exact_i8_scores_kernel(const ExactArgs a) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  int8_t* sq = reinterpret_cast<int8_t*>(ex_smem);  // [EX_QT][dim]
  const int q0 = blockIdx.y * EX_QT;
  const int nq = min(EX_QT, a.q - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int8_t* queries = reinterpret_cast<const int8_t*>(a.queries);
  for (int i = threadIdx.x; i < EX_QT * a.dim; i += EX_THREADS) {
    const int qi = i / a.dim;
    sq[i] = (qi < nq) ? queries[(size_t)(q0 + qi) * a.dim + (i % a.dim)] : (int8_t)0;
  }
  __syncthreads();
  const int8_t* emb = reinterpret_cast<const int8_t*>(a.emb);
  const bool vec = (a.dim & 3) == 0;
  const long long warps_total = (long long)gridDim.x * EX_WARPS;
  for (long long row = (long long)blockIdx.x * EX_WARPS + warp; row < a.n; row += warps_total) {
    bool valid = true;
    if (a.tags) valid = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
    int acc[EX_QT];
#pragma unroll
    for (int j = 0; j < EX_QT; ++j) acc[j] = 0;
    if (valid) {
      const int8_t* r = emb + (size_t)row * a.dim;
      if (vec) {
        const int* r4 = reinterpret_cast<const int*>(r);
        for (int v = lane; v < (a.dim >> 2); v += 32) {
          const int e = __ldg(r4 + v);
#pragma unroll
          for (int j = 0; j < EX_QT; ++j)
            acc[j] = __dp4a(e, reinterpret_cast<const int*>(sq + j * a.dim)[v], acc[j]);
        }
      } else {
        for (int d = lane; d < a.dim; d += 32) {
          const int e = (int)__ldg(r + d);
#pragma unroll
          for (int j = 0; j < EX_QT; ++j) acc[j] += e * (int)sq[j * a.dim + d];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < EX_QT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane < nq) {
      int dotv = 0;
#pragma unroll
      for (int j = 0; j < EX_QT; ++j)
        if (j == lane) dotv = acc[j];
      a.keys[(size_t)(q0 + lane) * a.n + row] = valid ? ~i32_orderable(dotv) : KEY32_INVALID;
    }
  }
}

// per-query top-k over the 32-bit key array
template <int MODE>
__global__ void __launch_bounds__(MERGE_THREADS)
    select_keys32_kernel(const u32* keys, long long n, int k, int cap, long long row_base, void* out_a,
                         long long* out_idx, int* out_count) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + cap);
  __shared__ SelectScratch<MERGE_THREADS> sc;
  const int q = blockIdx.x;
  const u32* kq = keys + (size_t)q * n;
  auto get = [&](long long i, u64& x, u32& y) {
    const u32 key = kq[i];
    x = (key == KEY32_INVALID) ? K1_INVALID : (u64)key;
    y = (u32)i;
  };
  const int m = block_select_sorted<MERGE_THREADS>(get, n, k, s_k1, s_k2, cap, sc);
  for (int j = threadIdx.x; j < k; j += MERGE_THREADS) {
    const bool have = j < m;
    merge_write<MODE>(out_a, out_idx, (size_t)q * k + j, have, have ? s_k1[j] : 0, have ? s_k2[j] : 0,
                      row_base);
  }
  if (out_count && threadIdx.x == 0) out_count[q] = m;
}

static int exact_grid_x(long long n) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long g = (n + EX_WARPS - 1) / EX_WARPS;
  const long long cap = (long long)sms * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace rr

using namespace rr;

extern "C" size_t rr_exact_search_f32_workspace_bytes(int64_t n, int32_t q, int32_t k) {
  (void)k;
  if (n <= 0 || q <= 0) return 256;
  return align_up((size_t)n * (size_t)q * 4, 256) + 256;
}

extern "C" size_t rr_int8_search_topk_workspace_bytes(int64_t n, int32_t q, int32_t k) {
  return rr_exact_search_f32_workspace_bytes(n, q, k);
}

__global__ void fill_missing_f32_kernel(float* s, long long* idx, int* count, long long total, int q) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    s[i] = 0.0f;
    idx[i] = -1;
  }
  if (count && i < q) count[i] = 0;
}
__global__ void fill_missing_i32_kernel(int* s, long long* idx, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    s[i] = (int)0x80000000;
    idx[i] = -1;
  }
}

extern "C" int rr_exact_search_f32(const float* emb, int64_t n, int32_t dim, const uint8_t* tags,
                                   uint8_t tag_mask, uint8_t tag_value, const float* queries,
                                   int32_t q, int32_t top_k, double min_similarity,
                                   int64_t row_base, float* out_score, int64_t* out_idx,
                                   int32_t* out_count, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  RR_CHECK_ARG(q >= 0 && n >= 0 && dim > 0, "bad size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(queries && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(emb || n == 0, "emb is null");
  RR_CHECK_ARG(top_k >= 1 && top_k <= RR_MAX_K, "top_k out of range");
  RR_CHECK_ARG(dim <= 4096, "dim > 4096 unsupported");
  RR_CHECK_ARG(n < (1LL << 32), "shard larger than 2^32 rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    const long long total = (long long)q * top_k;
    const long long threads = total > q ? total : q;
    fill_missing_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        out_score, (long long*)out_idx, out_count, total, q);
    RR_LAUNCH_CHECK();
    return RR_OK;
  }
  const size_t need = rr_exact_search_f32_workspace_bytes(n, q, top_k);
  if (!workspace || workspace_bytes < need) {
    set_error("rr_exact_search_f32: workspace %zu < %zu", workspace_bytes, need);
    return RR_ERR_WORKSPACE;
  }
  ExactArgs a{emb, n, dim, tags, tag_mask, tag_value, queries, q, min_similarity, (u32*)workspace};
  const size_t smem = (size_t)EX_QT * dim * 4;
  RR_CUDA(cudaFuncSetAttribute(exact_f32_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  dim3 grid(exact_grid_x(n), (q + EX_QT - 1) / EX_QT);
  exact_f32_scores_kernel<<<grid, EX_THREADS, smem, st>>>(a);
  RR_LAUNCH_CHECK();
  const int cap = merge_cap(top_k);
  select_keys32_kernel<MERGE_F32_DESC><<<q, MERGE_THREADS, (size_t)cap * 12, st>>>(
      (const u32*)workspace, n, top_k, cap, row_base, out_score, (long long*)out_idx, out_count);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_int8_search_topk(const int8_t* emb, int64_t n, int32_t dim, const uint8_t* tags,
                                   uint8_t tag_mask, uint8_t tag_value, const int8_t* queries_i8,
                                   int32_t q, int32_t top_k, int64_t row_base, int32_t* out_score,
                                   int64_t* out_idx, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  RR_CHECK_ARG(q >= 0 && n >= 0 && dim > 0, "bad size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(queries_i8 && out_score && out_idx, "null pointer");
  RR_CHECK_ARG(emb || n == 0, "emb is null");
  RR_CHECK_ARG(top_k >= 1 && top_k <= RR_MAX_K, "top_k out of range");
  RR_CHECK_ARG(dim <= 16384, "dim > 16384 unsupported");
  RR_CHECK_ARG(n < (1LL << 32), "shard larger than 2^32 rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    const long long total = (long long)q * top_k;
    fill_missing_i32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out_score,
                                                                           (long long*)out_idx, total);
    RR_LAUNCH_CHECK();
    return RR_OK;
  }
  const size_t need = rr_int8_search_topk_workspace_bytes(n, q, top_k);
  if (!workspace || workspace_bytes < need) {
    set_error("rr_int8_search_topk: workspace %zu < %zu", workspace_bytes, need);
    return RR_ERR_WORKSPACE;
  }
  ExactArgs a{emb, n, dim, tags, tag_mask, tag_value, queries_i8, q, 0.0, (u32*)workspace};
  const size_t smem = (size_t)EX_QT * dim;
  RR_CUDA(cudaFuncSetAttribute(exact_i8_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  dim3 grid(exact_grid_x(n), (q + EX_QT - 1) / EX_QT);
  exact_i8_scores_kernel<<<grid, EX_THREADS, smem, st>>>(a);
  RR_LAUNCH_CHECK();
  const int cap = merge_cap(top_k);
  select_keys32_kernel<MERGE_I32_DESC><<<q, MERGE_THREADS, (size_t)cap * 12, st>>>(
      (const u32*)workspace, n, top_k, cap, row_base, out_score, (long long*)out_idx, nullptr);
  RR_LAUNCH_CHECK();
  return RR_OK;
}
