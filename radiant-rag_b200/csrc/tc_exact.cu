// R5 for BATCHES: the exact float32 cosine scan (reference
// radiant/storage/redis_store.py:863-952, `_retrieve_by_embedding_linear`) with the
// row x query products on the tensor cores, results bit-identical to rr_exact_search_f32.
//
// rr_exact_search_f32 accumulates every (query, row) product in float64 on the CUDA cores: one
// pass over the rows per 8 queries and FP64-pipe bound from 4 queries up (11 ms for 64 queries
// over 1M x 768).  Here the scan is filter-and-refine, like the batched BM25 path:
//   pass 0  (EPI_COLMAX) TF32 cosines of a strided sample of row tiles, one maximum per
//           (CTA, row slot, query); tau_keys_kernel takes a value tau_q that >= k sampled rows reach.
//   pass 1  (EPI_FILTER) TF32 cosines of EVERY row against up to 128 queries per CTA
//           (tcgen05.mma kind::tf32, M = 128 rows, N = 128 queries, K = 8; float32 rows and queries
//           go from HBM / L2 through a TMA ring straight into the MMA, no conversion pass); rows with
//           approx >= tau_q - 2*delta are appended to the query's list.  HBM-bound: the rows are
//           read once per 128 queries.
//   pass 2  (tx_refine_kernel) the EXACT score of every listed row, in exactly the operation order
//           of exact_f32_scores_vec_kernel (exact.cu: float64 fused multiply-adds lane-strided over
//           float4 groups, xor-shuffle tree, one rounding to float32), min_similarity applied.
//   pass 3  tc_select_lists_kernel: (score desc, row asc) top-k.
// Exactness: |approx - cos| <= delta for every row (TF32 drops 13 mantissa bits of each factor:
// relative 2^-10 each, so |sum - dot| <= (2^-9 + 2^-20) * sum|a_i b_i| <= 2^-9 * |a||b| by
// Cauchy-Schwarz, plus float32 accumulation; delta = 2^-8 leaves a factor 2).  k sampled rows have
// approx >= tau, hence cos >= tau - delta, so every row of the true top-k has cos >= tau - delta
// and approx >= tau - 2*delta: it is listed, and refined exactly.  A list that outgrows its
// segment raises the overflow counter and the query's flag; the caller redoes those queries with
// rr_exact_search_f32.
// Algorithmic bytes per launch of the filter pass: n * dim * 4 per 128 queries (+ the same from L2
// for the query tile, which is re-streamed per row tile because 128 x dim float32 does not fit in
// shared memory next to the row ring).
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "merge.cuh"
#include "tau.cuh"
#include "tc_lists.cuh"
#include "tc_ptx.cuh"

namespace rr {

constexpr int TX_BM = 128;                          // rows per tile (TMEM lanes)
constexpr int TX_BN = 128;                          // queries per CTA (TMEM columns per accumulator)
constexpr int TX_KF = 32;                           // floats of K per ring stage = one 128-byte swizzle row
constexpr int TX_TILE_BYTES = TX_BM * TX_KF * 4;    // 16 KB per operand per stage
constexpr int TX_STAGE_BYTES = 2 * TX_TILE_BYTES;   // rows + queries
constexpr int TX_STAGES = 6;
constexpr int TX_EPI_GROUPS = 2;                    // groups of 4 epilogue warps; group g owns columns [64g, 64g+64)
constexpr int TX_EPI_CHUNKS = TX_BN / 32 / TX_EPI_GROUPS;
constexpr int TX_THREADS = 64 + 128 * TX_EPI_GROUPS;  // producer, MMA issuer, epilogue warps
constexpr u32 TX_IDESC = tc_idesc_tf32(TX_BM, TX_BN);
constexpr float TX_DELTA = 1.0f / 256.0f;           // bound on |TF32 cosine - cosine| (see header)
constexpr int TX_EPI_FILTER = 0, TX_EPI_COLMAX = 2;

struct TxArgs {
  long long n;            // rows
  int q;                  // queries
  int kb;                 // K blocks of 32 floats (dim / 32)
  long long tile_stride;  // launched tile i covers row tile i * tile_stride
  long long n_tiles;      // launched tiles
  const uint8_t* tags;
  unsigned tag_mask, tag_value;
  const float* inv_norm;  // [n] 1 / |row| (0 for zero rows)
  const float* qnorm;     // [q] |query|
  const int* tau;         // [q] filter pass: sampled bound, i32-orderable encoding of the float (INT_MIN = none)
  u32* keys;              // sample pass: [q][gridDim.x * 128] keys ~orderable(max approx * |q|)
  u32* cnt;               // [q][gridDim.x] list-segment lengths
  u32* list_row;          // [q][gridDim.x][cap_cta]
  int cap_cta;
};

template <int EPI>
__global__ void __launch_bounds__(TX_THREADS, 1)
    tc_tf32_scan_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const TxArgs a) {
  extern __shared__ __align__(1024) unsigned char tx_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up_dev((size_t)tx_smem_raw, 1024));
  unsigned char* ring = base;  // [stages][rows 16 KB | queries 16 KB]
  float* thr = reinterpret_cast<float*>(ring + (size_t)TX_STAGES * TX_STAGE_BYTES);  // [128]
  u32* s_cnt = reinterpret_cast<u32*>(thr + TX_BN);                                   // [128]
  u64* full = reinterpret_cast<u64*>(s_cnt + TX_BN);  // [stages]
  u64* empty = full + TX_STAGES;                      // [stages]
  u64* tmem_full = empty + TX_STAGES;                 // [2]
  u64* tmem_empty = tmem_full + 2;                    // [2]
  u32* tmem_ptr = reinterpret_cast<u32*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * TX_BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TX_STAGES; ++s) {
      tc_mbar_init(tc_smem(full + s), 1);
      tc_mbar_init(tc_smem(empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc_mbar_init(tc_smem(tmem_full + s), 1);
      tc_mbar_init(tc_smem(tmem_empty + s), 128 * TX_EPI_GROUPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr u32 tmem_cols = 2 * TX_BN;  // two float32 accumulators
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(tmem_ptr)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2 && warp < 6) {
    const int t = threadIdx.x - 64;  // 0..127
    const int qq = q0 + t;
    float th = __int_as_float(0x7f800000);  // +inf: padding columns and zero queries list nothing
    if (EPI == TX_EPI_FILTER && qq < a.q) {
      const float qn = a.qnorm[qq];
      if (qn > 0.0f && qn < __int_as_float(0x7f800000)) {
        const int te = a.tau[qq];
        const float tau = te == (int)0x80000000 ? -__int_as_float(0x7f800000) : f32_from_orderable(i32_orderable(te));
        // sample and filter values are approx * |q|: both sides of the bound carry the factor
        th = tau - 2.0f * TX_DELTA * qn - fabsf(tau) * (1.0f / 262144.0f);
      }
    }
    thr[t] = th;
    s_cnt[t] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *reinterpret_cast<volatile u32*>(tmem_ptr);

  if (warp == 0) {
    // ===================== producer: one row box and one query box per K block =====================
    if (lane == 0) {
      u32 s = 0, ph = 0;
      for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x) {
        const long long row0 = i * a.tile_stride * TX_BM;
        for (int kb = 0; kb < a.kb; ++kb) {
          tc_mbar_wait(tc_smem(empty + s), ph ^ 1u);
          tc_mbar_expect_tx(tc_smem(full + s), TX_STAGE_BYTES);
          unsigned char* st = ring + (size_t)s * TX_STAGE_BYTES;
          tc_tma_load_2d(tc_smem(st), &map_a, kb * TX_KF, (int)row0, tc_smem(full + s));
          tc_tma_load_2d(tc_smem(st + TX_TILE_BYTES), &map_b, kb * TX_KF, q0, tc_smem(full + s));
          if (++s == TX_STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const u64 desc0 = tc_smem_desc(tc_smem(ring));
    const u32 full0 = tc_smem(full), empty0 = tc_smem(empty);
    u32 stage = 0, phase = 0, tcount = 0;
    for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x, ++tcount) {
      const u32 as = tcount & 1u;
      tc_mbar_wait(tc_smem(tmem_empty + as), ((tcount >> 1) & 1u) ^ 1u);  // the epilogue has read this accumulator
      tc_fence_after();
      const u32 d_tmem = tmem_base + as * TX_BN;
      for (int kb = 0; kb < a.kb; ++kb) {
        tc_mbar_wait(full0 + stage * 8, phase);
        tc_fence_after();
        const u64 ad = desc0 + (u64)(stage * (TX_STAGE_BYTES >> 4));
        const u64 bd = ad + (u64)(TX_TILE_BYTES >> 4);
        tc_mma_tf32_ss_elect(d_tmem, ad, bd, TX_IDESC, kb > 0 ? 1u : 0u);  // 8 floats = 32 bytes of K each
        tc_mma_tf32_ss_elect(d_tmem, ad + 2, bd + 2, TX_IDESC, 1u);
        tc_mma_tf32_ss_elect(d_tmem, ad + 4, bd + 4, TX_IDESC, 1u);
        tc_mma_tf32_ss_elect(d_tmem, ad + 6, bd + 6, TX_IDESC, 1u);
        tc_commit_elect(empty0 + stage * 8);  // frees the stage when these MMAs have read it
        if (++stage == TX_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      tc_commit_elect(tc_smem(tmem_full + as));  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int lq = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int c0 = ((warp - 2) >> 2) * TX_EPI_CHUNKS;
    float colmax[EPI == TX_EPI_COLMAX ? TX_EPI_CHUNKS * 32 : 1];
#pragma unroll
    for (int j = 0; j < (EPI == TX_EPI_COLMAX ? TX_EPI_CHUNKS * 32 : 1); ++j) colmax[j] = -__int_as_float(0x7f800000);
    u32 tcount = 0;
    for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x, ++tcount) {
      const u32 as = tcount & 1u;
      const long long row = i * a.tile_stride * TX_BM + lq * 32 + lane;
      bool valid = row < a.n;
      if (valid && a.tags != nullptr) valid = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
      const float inv = valid ? __ldg(a.inv_norm + row) : 0.0f;
      tc_mbar_wait(tc_smem(tmem_full + as), (tcount >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < TX_EPI_CHUNKS; ++cc) {
        const int c = c0 + cc;
        u32 v[32];
        const u32 taddr = tmem_base + ((u32)(lq * 32) << 16) + as * TX_BN + c * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (EPI == TX_EPI_COLMAX) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) colmax[cc * 32 + j] = fmaxf(colmax[cc * 32 + j], __uint_as_float(v[j]) * inv);
          }
        } else {
          u32 hit = 0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t4 = reinterpret_cast<const float4*>(thr + c * 32)[j4];
            hit |= (__uint_as_float(v[4 * j4 + 0]) * inv >= t4.x ? 1u : 0u) << (4 * j4 + 0);
            hit |= (__uint_as_float(v[4 * j4 + 1]) * inv >= t4.y ? 1u : 0u) << (4 * j4 + 1);
            hit |= (__uint_as_float(v[4 * j4 + 2]) * inv >= t4.z ? 1u : 0u) << (4 * j4 + 2);
            hit |= (__uint_as_float(v[4 * j4 + 3]) * inv >= t4.w ? 1u : 0u) << (4 * j4 + 3);
          }
          if (!valid) hit = 0;
          while (hit) {  // rare: about 3 * k * stride rows per query reach the bound
            const int j = __ffs(hit) - 1;
            hit &= hit - 1;
            const u32 slot = atomicAdd(s_cnt + c * 32 + j, 1u);
            if (slot < (u32)a.cap_cta)
              a.list_row[((size_t)(q0 + c * 32 + j) * gridDim.x + blockIdx.x) * a.cap_cta + slot] = (u32)row;
          }
        }
      }
      tc_fence_before();
      tc_mbar_arrive(tc_smem(tmem_empty + as));
    }
    if (EPI == TX_EPI_COLMAX) {
      const size_t ld = (size_t)gridDim.x * TX_BM;
#pragma unroll
      for (int cc = 0; cc < TX_EPI_CHUNKS; ++cc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qq = q0 + (c0 + cc) * 32 + j;
          const float mx = colmax[cc * 32 + j];
          if (qq < a.q)
            a.keys[(size_t)qq * ld + (size_t)blockIdx.x * TX_BM + lq * 32 + lane] =
                mx == -__int_as_float(0x7f800000) ? 0xFFFFFFFFu : ~f32_orderable(mx);
        }
      }
    } else {
      asm volatile("bar.sync 1, %0;" ::"n"(128 * TX_EPI_GROUPS) : "memory");  // the epilogue warps only
      const int t = threadIdx.x - 64;
      if (t < TX_BN && q0 + t < a.q) a.cnt[(size_t)(q0 + t) * gridDim.x + blockIdx.x] = s_cnt[t];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// 1 / |row| in float32 (from the float64 sum), one warp per row
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ emb, long long n, int dim,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * 8;
  const int dim4 = dim >> 2;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += warps) {
    const float4* r4 = reinterpret_cast<const float4*>(emb + (size_t)row * dim);
    double s = 0.0;
    for (int v = lane; v < dim4; v += 32) {
      const float4 e = __ldg(r4 + v);
      s += (double)e.x * e.x + (double)e.y * e.y + (double)e.z * e.z + (double)e.w * e.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.0f;
  }
}

// |query| in float32, one warp per query
__global__ void __launch_bounds__(32) tx_qnorm_kernel(const float* __restrict__ queries, int dim, float* __restrict__ out) {
  const float* r = queries + (size_t)blockIdx.x * dim;
  double s = 0.0;
  for (int d = threadIdx.x; d < dim; d += 32) s += (double)r[d] * r[d];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)sqrt(s);
}

// Exact scores of the listed rows: CTA (segment, query), one warp per row.  The operation order is
// that of exact_f32_scores_vec_kernel (exact.cu) - per lane, float4 groups v = lane, lane + 32, ...;
// fused multiply-adds x, y, z, w into the dot and into the squared norm; xor-shuffle tree 16..1;
// dot / (sqrt(nn) * |q|) rounded once to float32 - so the key equals that kernel's, bit for bit.
constexpr int TXR_THREADS = 256;
__global__ void __launch_bounds__(TXR_THREADS)
    tx_refine_kernel(const float* __restrict__ emb, int dim, const float* __restrict__ queries, const u32* __restrict__ cnt,
                     const u32* __restrict__ list_row, u32* __restrict__ list_key, int cap_cta, double min_similarity) {
  extern __shared__ __align__(16) unsigned char txr_smem[];
  float4* sq4 = reinterpret_cast<float4*>(txr_smem);  // [dim / 4]
  const int q = blockIdx.y, seg = blockIdx.x, n_cta = gridDim.x;
  u32 c = cnt[(size_t)q * n_cta + seg];
  if (c == 0) return;
  if (c > (u32)cap_cta) c = (u32)cap_cta;
  const int dim4 = dim >> 2;
  const float4* qrow = reinterpret_cast<const float4*>(queries + (size_t)q * dim);
  for (int v = threadIdx.x; v < dim4; v += TXR_THREADS) sq4[v] = qrow[v];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double qs = 0.0;
  for (int v = lane; v < dim4; v += 32) {
    const float4 w = sq4[v];
    const double wx = (double)w.x, wy = (double)w.y, wz = (double)w.z, ww = (double)w.w;
    qs = __fma_rn(wx, wx, qs);
    qs = __fma_rn(wy, wy, qs);
    qs = __fma_rn(wz, wz, qs);
    qs = __fma_rn(ww, ww, qs);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qs += __shfl_xor_sync(0xffffffffu, qs, o);
  const double qn = sqrt(qs);
  const size_t lbase = ((size_t)q * n_cta + seg) * cap_cta;
  for (u32 j = warp; j < c; j += TXR_THREADS / 32) {
    const u32 row = list_row[lbase + j];
    const float4* r4 = reinterpret_cast<const float4*>(emb + (size_t)row * dim);
    double acc = 0.0, nn = 0.0;
#pragma unroll 4
    for (int v = lane; v < dim4; v += 32) {
      const float4 e = __ldg(r4 + v);
      const float4 w = sq4[v];
      const double ex = (double)e.x, ey = (double)e.y, ez = (double)e.z, ew = (double)e.w;
      nn = __fma_rn(ex, ex, nn);
      nn = __fma_rn(ey, ey, nn);
      nn = __fma_rn(ez, ez, nn);
      nn = __fma_rn(ew, ew, nn);
      acc = __fma_rn((double)w.x, ex, acc);
      acc = __fma_rn((double)w.y, ey, acc);
      acc = __fma_rn((double)w.z, ez, acc);
      acc = __fma_rn((double)w.w, ew, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) {
      u32 key = 0xFFFFFFFFu;
      if (nn > 0.0 && qn > 0.0) {
        const float s = (float)(acc / (sqrt(nn) * qn));
        if ((double)s >= min_similarity) key = ~f32_orderable(s);
      }
      list_key[lbase + j] = key;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*TxEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TxEncodeTiledFn tx_encode_fn() {
  static TxEncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (TxEncodeTiledFn)p;
  return fn;
}

// rows x dim float32 matrix, K-major; box = 128 rows x 32 floats (128 bytes), SWIZZLE_128B;
// rows past the end read as zeros
static int tx_make_map(CUtensorMap* map, const void* ptr, long long rows, int dim) {
  TxEncodeTiledFn fn = tx_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return RR_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim * 4};
  cuuint32_t box[2] = {(cuuint32_t)TX_KF, (cuuint32_t)TX_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld dim=%d", (int)r, rows, dim);
    return RR_ERR_CUDA;
  }
  return RR_OK;
}

struct TxPlan {
  bool ok;
  long long tiles, sample_tiles, keys_per_q;
  int stride, sample_ctas, qblocks, ctas_x, cap_cta;
  size_t off_keys, off_tau, off_qnorm, off_cnt, off_key, off_row, total;
};

// The sample pass reads n / stride rows per query block; the rows that later reach the bound
// (about 3 * k * stride per query: the 2*delta margin roughly triples them on embedding-like
// data) are re-read by the refine.  Both are HBM traffic, so stride ~ sqrt(n / (3 k q_block)).
static TxPlan tx_plan(long long n, int q, int k) {
  TxPlan p;
  p.ok = false;
  p.tiles = (n + TX_BM - 1) / TX_BM;
  p.qblocks = (q + TX_BN - 1) / TX_BN;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  p.ctas_x = sms / p.qblocks;
  if (p.ctas_x < 1) p.ctas_x = 1;
  if (p.ctas_x > p.tiles) p.ctas_x = (int)p.tiles;
  const int qb = q < TX_BN ? q : TX_BN;
  const double want = sqrt((double)n / (3.0 * (double)k * (double)qb));
  int s = 1;
  while (s < 128 && (double)s * 1.4142 < want) s <<= 1;
  // one key per (CTA, row slot): the sample must offer at least 4k slots
  for (;; s >>= 1) {
    p.stride = s;
    p.sample_tiles = (p.tiles + s - 1) / s;
    p.sample_ctas = (int)(p.sample_tiles < p.ctas_x ? p.sample_tiles : p.ctas_x);
    if ((long long)p.sample_ctas * TX_BM >= 4LL * k) break;
    if (s == 1) return p;
  }
  p.keys_per_q = (long long)p.sample_ctas * TX_BM;
  const long long expect = 3LL * k * p.stride;
  long long cap = 2 * (expect / p.ctas_x + 1) + 64;
  const long long rows_per_cta = ((p.tiles + p.ctas_x - 1) / p.ctas_x) * TX_BM;
  if (cap > rows_per_cta) cap = rows_per_cta;
  p.cap_cta = (int)cap;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  p.off_keys = take((size_t)q * p.keys_per_q * 4);
  p.off_tau = take((size_t)q * 4);
  p.off_qnorm = take((size_t)q * 4);
  p.off_cnt = take((size_t)q * p.ctas_x * 4);
  p.off_key = take((size_t)q * p.ctas_x * p.cap_cta * 4);
  p.off_row = take((size_t)q * p.ctas_x * p.cap_cta * 4);
  p.total = o + 256;
  p.ok = true;
  return p;
}

static bool tx_shape_ok(long long n, int dim, int q, int k) {
  return n >= TX_BM && n < (1LL << 31) && dim >= TX_KF && dim % TX_KF == 0 && dim <= 4096 && q >= 1 && q <= 65535 &&
         k >= 1 && k <= RR_MAX_K;
}

constexpr size_t TX_SMEM_BYTES = 1024 + (size_t)TX_STAGES * TX_STAGE_BYTES + 2 * TX_BN * 4 + (2 * TX_STAGES + 4) * 8 + 64;

}  // namespace rr

using namespace rr;

extern "C" int rr_row_inv_norms_f32(const float* emb, int64_t n, int32_t dim, float* out, void* stream) {
  RR_CHECK_ARG(n >= 0 && dim > 0 && dim % 4 == 0, "bad size (dim must be a multiple of 4)");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(emb && out, "null pointer");
  RR_CHECK_ARG(((uintptr_t)emb & 15) == 0, "rows must be 16-byte aligned");
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (n + 7) / 8;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  row_inv_norm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(emb, n, dim, out);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_exact_search_f32_tc_supported(int64_t n, int32_t dim, int32_t q, int32_t top_k) {
  if (!tx_shape_ok(n, dim, q, top_k)) return 0;
  return tx_plan(n, q, top_k).ok ? 1 : 0;
}

extern "C" size_t rr_exact_search_f32_tc_workspace_bytes(int64_t n, int32_t q, int32_t top_k) {
  if (n <= 0 || q <= 0 || top_k <= 0) return 256;
  const TxPlan p = tx_plan(n, q, top_k);
  return p.ok ? p.total : 256;
}

extern "C" int rr_exact_search_f32_tc(const float* emb, const float* row_inv_norm, int64_t n, int32_t dim,
                                      const uint8_t* tags, uint8_t tag_mask, uint8_t tag_value,
                                      const float* queries, int32_t q, int32_t top_k, double min_similarity,
                                      int64_t row_base, float* out_score, int64_t* out_idx, int32_t* out_count,
                                      uint32_t* overflow, uint8_t* overflow_flags, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  RR_CHECK_ARG(tx_shape_ok(n, dim, q, top_k),
               "tensor-core exact scan needs n in [128, 2^31), dim a multiple of 32 <= 4096, q <= 65535, top_k <= RR_MAX_K");
  RR_CHECK_ARG(emb && row_inv_norm && queries && out_score && out_idx && out_count && overflow, "null pointer");
  RR_CHECK_ARG((((uintptr_t)emb | (uintptr_t)queries) & 15) == 0, "rows and queries must be 16-byte aligned");
  const TxPlan p = tx_plan(n, q, top_k);
  RR_CHECK_ARG(p.ok, "corpus too small for this top_k on the tensor-core path (rr_exact_search_f32_tc_supported)");
  if (!workspace || workspace_bytes < p.total) {
    set_error("rr_exact_search_f32_tc: workspace %zu < %zu", workspace_bytes, p.total);
    return RR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  CUtensorMap map_a, map_b;
  int rc = tx_make_map(&map_a, emb, n, dim);
  if (rc != RR_OK) return rc;
  rc = tx_make_map(&map_b, queries, q, dim);
  if (rc != RR_OK) return rc;
  RR_CUDA(cudaFuncSetAttribute(tc_tf32_scan_kernel<TX_EPI_COLMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)TX_SMEM_BYTES));
  RR_CUDA(cudaFuncSetAttribute(tc_tf32_scan_kernel<TX_EPI_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)TX_SMEM_BYTES));
  TxArgs a;
  a.n = n;
  a.q = q;
  a.kb = dim / TX_KF;
  a.tags = tags;
  a.tag_mask = tag_mask;
  a.tag_value = tag_value;
  a.inv_norm = row_inv_norm;
  a.qnorm = (const float*)(w + p.off_qnorm);
  a.tau = nullptr;
  a.keys = (u32*)(w + p.off_keys);
  a.cnt = (u32*)(w + p.off_cnt);
  a.list_row = (u32*)(w + p.off_row);
  a.cap_cta = p.cap_cta;
  tx_qnorm_kernel<<<q, 32, 0, st>>>(queries, dim, (float*)(w + p.off_qnorm));
  RR_LAUNCH_CHECK();
  // ---- pass 0: sample
  a.tile_stride = p.stride;
  a.n_tiles = p.sample_tiles;
  {
    dim3 grid((unsigned)p.sample_ctas, p.qblocks);
    tc_tf32_scan_kernel<TX_EPI_COLMAX><<<grid, TX_THREADS, TX_SMEM_BYTES, st>>>(map_a, map_b, a);
    RR_LAUNCH_CHECK();
  }
  tau_keys_kernel<false><<<q, TAU_THREADS, 0, st>>>(a.keys, p.keys_per_q, top_k, (void*)(w + p.off_tau), 1);
  RR_LAUNCH_CHECK();
  // ---- pass 1: filter over all rows
  a.tile_stride = 1;
  a.n_tiles = p.tiles;
  a.tau = (const int*)(w + p.off_tau);
  {
    dim3 grid((unsigned)p.ctas_x, p.qblocks);
    tc_tf32_scan_kernel<TX_EPI_FILTER><<<grid, TX_THREADS, TX_SMEM_BYTES, st>>>(map_a, map_b, a);
    RR_LAUNCH_CHECK();
  }
  // ---- pass 2: exact scores of the listed rows
  {
    dim3 grid((unsigned)p.ctas_x, q);
    tx_refine_kernel<<<grid, TXR_THREADS, (size_t)dim * 4, st>>>(emb, dim, queries, a.cnt, a.list_row,
                                                                 (u32*)(w + p.off_key), p.cap_cta, min_similarity);
    RR_LAUNCH_CHECK();
  }
  // ---- pass 3: (score desc, row asc) top-k of each query's segments
  const int kcap = merge_cap(top_k);
  const size_t list_smem = (size_t)kcap * 12 + (size_t)LIST_STAGE_CAP * 8;
  RR_CUDA(cudaFuncSetAttribute(tc_select_lists_kernel<MERGE_F32_DESC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)list_smem));
  tc_select_lists_kernel<MERGE_F32_DESC><<<q, LIST_THREADS, list_smem, st>>>(
      a.cnt, (const int*)(w + p.off_key), a.list_row, p.ctas_x, p.cap_cta, top_k, kcap, dim, nullptr, nullptr, row_base,
      out_score, (long long*)out_idx, out_count, overflow, overflow_flags, 1);
  RR_LAUNCH_CHECK();
  return RR_OK;
}
