// R5 and BASELINE config 4: exact (non-quantised-stage) scans.
//
// rr_exact_search_f32 replaces RedisVectorStore._retrieve_by_embedding_linear
// (reference radiant/storage/redis_store.py:863-952): cosine of the query against
// every row, rows with zero norm skipped, doc_level filter, keep >= min_similarity,
// order (score desc, row asc), top_k.
// rr_int8_search_topk is the exact int8 x int8 -> int32 search of config 4 with the
// same ordering.
//
// Score kernels: a warp streams R rows at a time with 128-bit loads (R independent loads
// in flight per lane) against QT queries held in shared memory as float64, so a row is
// read from HBM once per QT queries and the query operands are read from shared memory
// once per R rows.  float32 rows accumulate in float64 (correctly rounded cosine, no
// f32->f64 conversion of the query in the loop; the fused multiply-adds are spelled out because
// tc_exact.cu's refine kernel restates this exact operation order for its candidates); int8 rows use DP4A (exact).  Scores go
// out as 32-bit order-preserving keys [q, n].
// Selection is two-level: (chunk, query) CTAs take the top-k of <= SEL_CHUNK_MAX keys each
// (block_select_sorted with the per-thread-minimum bound shortcut), then one CTA per
// query merges the chunk winners.  Single-query latency is therefore one pass over the
// rows at HBM speed plus two short launches.
// Algorithmic bytes per launch: n * dim * sizeof(elem) per QT queries + 8 * n * q (keys).
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

// dim % 4 == 0 and 16-byte aligned rows.  Shared memory: the queries as float64, split
// into the (x,y) and (z,w) halves of each float4 column group so that consecutive lanes
// read consecutive 16-byte words (conflict-free LDS.128).
template <int QT, int R>
__global__ void __launch_bounds__(EX_THREADS) exact_f32_scores_vec_kernel(const ExactArgs a) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  const int dim4 = a.dim >> 2;
  double2* sq_lo = reinterpret_cast<double2*>(ex_smem);  // [QT][dim4]
  double2* sq_hi = sq_lo + (size_t)QT * dim4;             // [QT][dim4]
  __shared__ double s_qnorm[QT];
  const int q0 = blockIdx.y * QT;
  const int nq = min(QT, a.q - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4* queries = reinterpret_cast<const float4*>(a.queries);
  for (int i = threadIdx.x; i < QT * dim4; i += EX_THREADS) {
    const int qi = i / dim4, v = i - qi * dim4;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qi < nq) w = queries[(size_t)(q0 + qi) * dim4 + v];
    sq_lo[i] = make_double2((double)w.x, (double)w.y);
    sq_hi[i] = make_double2((double)w.z, (double)w.w);
  }
  __syncthreads();
  for (int j = warp; j < QT; j += EX_WARPS) {
    double s = 0.0;
    for (int v = lane; v < dim4; v += 32) {
      const double2 lo = sq_lo[j * dim4 + v], hi = sq_hi[j * dim4 + v];
      s = __fma_rn(lo.x, lo.x, s);
      s = __fma_rn(lo.y, lo.y, s);
      s = __fma_rn(hi.x, hi.x, s);
      s = __fma_rn(hi.y, hi.y, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qnorm[j] = sqrt(s);
  }
  __syncthreads();
  const float4* emb = reinterpret_cast<const float4*>(a.emb);
  const long long groups = (a.n + R - 1) / R;
  const long long warps_total = (long long)gridDim.x * EX_WARPS;
  for (long long g = (long long)blockIdx.x * EX_WARPS + warp; g < groups; g += warps_total) {
    const long long row0 = g * R;
    bool valid[R];
    const float4* rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
      valid[r] = row < a.n;
      if (valid[r] && a.tags) valid[r] = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
      rp[r] = emb + (size_t)(valid[r] ? row : 0) * dim4;
    }
    double acc[QT][R];
    double nn[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      nn[r] = 0.0;
#pragma unroll
      for (int j = 0; j < QT; ++j) acc[j][r] = 0.0;
    }
    for (int v = lane; v < dim4; v += 32) {
      float4 e[R];
#pragma unroll
      for (int r = 0; r < R; ++r) e[r] = valid[r] ? __ldg(rp[r] + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      double ex[R], ey[R], ez[R], ew[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ex[r] = (double)e[r].x;
        ey[r] = (double)e[r].y;
        ez[r] = (double)e[r].z;
        ew[r] = (double)e[r].w;
        nn[r] = __fma_rn(ex[r], ex[r], nn[r]);
        nn[r] = __fma_rn(ey[r], ey[r], nn[r]);
        nn[r] = __fma_rn(ez[r], ez[r], nn[r]);
        nn[r] = __fma_rn(ew[r], ew[r], nn[r]);
      }
#pragma unroll
      for (int j = 0; j < QT; ++j) {
        const double2 lo = sq_lo[j * dim4 + v], hi = sq_hi[j * dim4 + v];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[j][r] = __fma_rn(lo.x, ex[r], acc[j][r]);
          acc[j][r] = __fma_rn(lo.y, ey[r], acc[j][r]);
          acc[j][r] = __fma_rn(hi.x, ez[r], acc[j][r]);
          acc[j][r] = __fma_rn(hi.y, ew[r], acc[j][r]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        nn[r] += __shfl_xor_sync(0xffffffffu, nn[r], o);
#pragma unroll
        for (int j = 0; j < QT; ++j) acc[j][r] += __shfl_xor_sync(0xffffffffu, acc[j][r], o);
      }
    }
    // lane = r * QT + j writes the key of (query j, row r)
    if (lane < QT * R) {
      const int r_me = lane / QT, j_me = lane % QT;
      double dotv = 0.0, nnv = 0.0;
      bool ok = false;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r == r_me) {
          nnv = nn[r];
          ok = valid[r];
#pragma unroll
          for (int j = 0; j < QT; ++j)
            if (j == j_me) dotv = acc[j][r];
        }
      const long long row = row0 + r_me;
      if (j_me < nq && row < a.n) {
        const double qn = s_qnorm[j_me];
        u32 key = KEY32_INVALID;
        if (ok && nnv > 0.0 && qn > 0.0) {
          const float s = (float)(dotv / (sqrt(nnv) * qn));
          if ((double)s >= a.min_similarity) key = ~f32_orderable(s);
        }
        a.keys[(size_t)(q0 + j_me) * a.n + row] = key;
      }
    }
  }
}

// any dim / alignment (scalar loads)
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

// level 1 of the per-query top-k over the 32-bit key array: CTA (chunk, query) leaves its
// chunk's k best as (key, row) pairs in part_k1/part_k2 [q][n_chunks][k]
__global__ void __launch_bounds__(MERGE_THREADS)
    select_keys32_chunk_kernel(const u32* keys, long long n, long long chunk, int n_chunks, int k, int cap,
                               u64* part_k1, u32* part_k2) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + cap);
  __shared__ SelectScratch<MERGE_THREADS> sc;
  const int q = blockIdx.y;
  const long long lo = (long long)blockIdx.x * chunk;
  const long long len = min(chunk, n - lo);
  const u32* kq = keys + (size_t)q * n + lo;
  auto get = [&](long long i, u64& x, u32& y) {
    const u32 key = kq[i];
    x = (key == KEY32_INVALID) ? K1_INVALID : (u64)key;
    y = (u32)(lo + i);
  };
  const int m = block_select_sorted<MERGE_THREADS, true>(get, len, k, s_k1, s_k2, cap, sc);
  const size_t o = ((size_t)q * n_chunks + blockIdx.x) * k;
  for (int j = threadIdx.x; j < k; j += MERGE_THREADS) {
    part_k1[o + j] = (j < m) ? s_k1[j] : K1_INVALID;
    part_k2[o + j] = (j < m) ? s_k2[j] : K2_INVALID;
  }
}

constexpr long long SEL_CHUNK_MIN = 4096;
constexpr int SEL_CHUNKS_MAX = 1024;

static long long select_chunk_rows(long long n) {
  long long c = (n + SEL_CHUNKS_MAX - 1) / SEL_CHUNKS_MAX;
  if (c < SEL_CHUNK_MIN) c = SEL_CHUNK_MIN;
  return (c + 255) / 256 * 256;
}
static int select_n_chunks(long long n) {
  const long long c = select_chunk_rows(n);
  return (int)((n + c - 1) / c);
}

// keys [q][n] -> out (score, idx, count); `part` holds [q][n_chunks][k] pairs
template <int MODE>
static int select_keys32(const u32* keys, long long n, int q, int k, long long row_base, void* part,
                         void* out_a, long long* out_idx, int* out_count, cudaStream_t st) {
  const int n_chunks = select_n_chunks(n);
  const long long chunk = select_chunk_rows(n);
  const size_t e = (size_t)q * n_chunks * k;
  u64* part_k1 = (u64*)part;
  u32* part_k2 = (u32*)((char*)part + align_up(e * 8, 256));
  const int cap = merge_cap(k);
  dim3 grid(n_chunks, q);
  select_keys32_chunk_kernel<<<grid, MERGE_THREADS, (size_t)cap * 12, st>>>(keys, n, chunk, n_chunks, k, cap,
                                                                          part_k1, part_k2);
  RR_LAUNCH_CHECK();
  MergeArgs m;
  m.k1 = part_k1;
  m.k2 = part_k2;
  m.n_in = (long long)n_chunks * k;
  m.k = k;
  m.cap = 0;
  m.row_base = row_base;
  m.out_a = out_a;
  m.out_idx = out_idx;
  m.out_count = out_count;
  return launch_merge_pairs<MODE>(m, q, st);
}

template <int QT, int R>
static int launch_exact_f32_vec(const ExactArgs& a, int grid_x, cudaStream_t st) {
  const size_t smem = (size_t)QT * a.dim * 8;
  RR_CUDA(cudaFuncSetAttribute(exact_f32_scores_vec_kernel<QT, R>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(grid_x, (a.q + QT - 1) / QT);
  exact_f32_scores_vec_kernel<QT, R><<<grid, EX_THREADS, smem, st>>>(a);
  RR_LAUNCH_CHECK();
  return RR_OK;
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
  if (n <= 0 || q <= 0 || k <= 0) return 256;
  const size_t e = (size_t)q * select_n_chunks(n) * k;
  return align_up((size_t)n * (size_t)q * 4, 256) + align_up(e * 8, 256) + align_up(e * 4, 256) + 256;
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
  RR_CHECK_ARG(q <= 65535, "more than 65535 queries per call");
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
  const bool vec = (dim & 3) == 0 && (((uintptr_t)emb | (uintptr_t)queries) & 15) == 0;
  // queries per pass: as many as the batch needs, while the float64 copy fits 96 KB
  int qt = q >= 8 ? 8 : q >= 2 ? 4 : 1;  // (a QT=2 build measured slower than QT=4 on B200)
  while (qt > 1 && (size_t)qt * dim * 8 > 96 * 1024) qt = qt == 8 ? 4 : 1;
  if (vec && (size_t)qt * dim * 8 <= 96 * 1024) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int rc;
    if (qt == 8) {
      long long gx = (n / 2 + EX_WARPS - 1) / EX_WARPS;
      rc = launch_exact_f32_vec<8, 2>(a, (int)max(1LL, min(gx, (long long)sms * 4)), st);
    } else {
      long long gx = (n / 4 + EX_WARPS - 1) / EX_WARPS;
      const int g = (int)max(1LL, min(gx, (long long)sms * 4));
      rc = qt == 4 ? launch_exact_f32_vec<4, 4>(a, g, st) : launch_exact_f32_vec<1, 4>(a, g, st);
    }
    if (rc != RR_OK) return rc;
  } else {
    const size_t smem = (size_t)EX_QT * dim * 4;
    RR_CUDA(cudaFuncSetAttribute(exact_f32_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    dim3 grid(exact_grid_x(n), (q + EX_QT - 1) / EX_QT);
    exact_f32_scores_kernel<<<grid, EX_THREADS, smem, st>>>(a);
    RR_LAUNCH_CHECK();
  }
  void* part = (char*)workspace + align_up((size_t)n * (size_t)q * 4, 256);
  return select_keys32<MERGE_F32_DESC>((const u32*)workspace, n, q, top_k, row_base, part, out_score,
                                       (long long*)out_idx, out_count, st);
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
  RR_CHECK_ARG(q <= 65535, "more than 65535 queries per call");
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
  void* part = (char*)workspace + align_up((size_t)n * (size_t)q * 4, 256);
  return select_keys32<MERGE_I32_DESC>((const u32*)workspace, n, q, top_k, row_base, part, out_score,
                                       (long long*)out_idx, nullptr, st);
}
