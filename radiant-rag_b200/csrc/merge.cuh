// Per-query merge kernels built on block_select_sorted (select.cuh).
//
// merge_pairs_kernel  : input already encoded as (u64 key1, u32 key2) pairs, e.g. the
//                       per-slab lists of the Hamming scan or the per-tile lists of BM25.
// merge_typed_kernel  : input as typed (value, int64 idx) lists, e.g. the lists that
//                       come back from the NCCL allgather of the row-sharded path
//                       (SURVEY.md 8e).
// Both write (value, idx) sorted by the canonical order and pad with idx = -1.
#pragma once

#include "common.cuh"
#include "select.cuh"

namespace rr {

constexpr int MERGE_THREADS = 256;

enum MergeMode {
  MERGE_HAMMING = 0,         // key1 = dist, key2 = row            -> (i32 dist, i64 idx)
  MERGE_F64_DESC = 1,        // key1 = ~orderable(f64), key2 = row  -> (f64 score, i64 idx)
  MERGE_I32_DESC = 2,        // key1 = ~orderable(i32), key2 = row  -> (i32 score, i64 idx)
  MERGE_HAMMING_PACKED = 3,  // key1 = dist << 40 | idx             -> (i32 dist, i64 idx)
  MERGE_F32_DESC = 4         // key1 = ~orderable(f32), key2 = row  -> (f32 score, i64 idx)
};

struct MergeArgs {
  const u64* k1;  // [q][n_in]
  const u32* k2;  // [q][n_in] or nullptr (key2 = 0)
  long long n_in;
  int k;
  int cap;
  long long row_base;
  void* out_a;
  long long* out_idx;
  int* out_count;
};

template <int MODE>
__device__ __forceinline__ void merge_write(void* out_a, long long* out_idx, size_t o, bool have,
                                            u64 x, u32 y, long long row_base) {
  if (have) {
    if (MODE == MERGE_HAMMING) {
      reinterpret_cast<int*>(out_a)[o] = (int)x;
      out_idx[o] = (long long)y + row_base;
    } else if (MODE == MERGE_HAMMING_PACKED) {
      reinterpret_cast<int*>(out_a)[o] = (int)(x >> 40);
      out_idx[o] = (long long)(x & ((1ull << 40) - 1ull));
    } else if (MODE == MERGE_F64_DESC) {
      reinterpret_cast<double*>(out_a)[o] = f64_from_orderable(~x);
      out_idx[o] = (long long)y + row_base;
    } else if (MODE == MERGE_F32_DESC) {
      reinterpret_cast<float*>(out_a)[o] = f32_from_orderable((u32)(~x));
      out_idx[o] = (long long)y + row_base;
    } else {
      reinterpret_cast<int*>(out_a)[o] = i32_from_orderable((u32)(~x));
      out_idx[o] = (long long)y + row_base;
    }
  } else {
    if (MODE == MERGE_HAMMING || MODE == MERGE_HAMMING_PACKED) {
      reinterpret_cast<int*>(out_a)[o] = 0x7fffffff;
    } else if (MODE == MERGE_F64_DESC) {
      reinterpret_cast<double*>(out_a)[o] = 0.0;
    } else if (MODE == MERGE_F32_DESC) {
      reinterpret_cast<float*>(out_a)[o] = 0.0f;
    } else {
      reinterpret_cast<int*>(out_a)[o] = (int)0x80000000;
    }
    out_idx[o] = -1;
  }
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) merge_pairs_kernel(const MergeArgs a) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + a.cap);
  __shared__ SelectScratch<THREADS> sc;
  const int q = blockIdx.x;
  const u64* k1 = a.k1 + (size_t)q * a.n_in;
  const u32* k2 = a.k2 ? a.k2 + (size_t)q * a.n_in : nullptr;
  auto get = [&](long long i, u64& x, u32& y) {
    x = k1[i];
    y = k2 ? k2[i] : 0u;
  };
  const int m = block_select_sorted<THREADS>(get, a.n_in, a.k, s_k1, s_k2, a.cap, sc);
  for (int j = threadIdx.x; j < a.k; j += THREADS) {
    const bool have = j < m;
    merge_write<MODE>(a.out_a, a.out_idx, (size_t)q * a.k + j, have, have ? s_k1[j] : 0,
                      have ? s_k2[j] : 0, a.row_base);
  }
  if (a.out_count && threadIdx.x == 0) a.out_count[q] = m;
}

static inline int merge_cap(int k) {
  int c = 64;
  while (c < 2 * k) c <<= 1;
  if (c > 2048) c = 2048;
  while (c < k) c <<= 1;
  return c;
}

// Few queries and long lists: one CTA per query is all the parallelism there is, so use
// the widest block; otherwise 256 threads and several CTAs per SM.
template <int MODE>
static int launch_merge_pairs(MergeArgs a, int q, cudaStream_t st) {
  a.cap = merge_cap(a.k);
  if (q <= 64 && a.n_in >= 8192)
    merge_pairs_kernel<MODE, 1024><<<q, 1024, (size_t)a.cap * 12, st>>>(a);
  else
    merge_pairs_kernel<MODE, MERGE_THREADS><<<q, MERGE_THREADS, (size_t)a.cap * 12, st>>>(a);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

template <int MODE>
__global__ void __launch_bounds__(MERGE_THREADS)
    merge_typed_kernel(const void* in_a, const long long* in_idx, long long n_in, int k, int cap,
                       void* out_a, long long* out_idx, int* out_count) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + cap);
  __shared__ SelectScratch<MERGE_THREADS> sc;
  const int q = blockIdx.x;
  const long long* idx = in_idx + (size_t)q * n_in;
  const size_t base = (size_t)q * n_in;
  auto get = [&](long long i, u64& x, u32& y) {
    const long long r = idx[i];
    if (r < 0) {
      x = K1_INVALID;
      y = K2_INVALID;
      return;
    }
    if (MODE == MERGE_HAMMING_PACKED) {
      x = ((u64)(u32) reinterpret_cast<const int*>(in_a)[base + i] << 40) | (u64)r;
      y = 0;
    } else if (MODE == MERGE_F64_DESC) {
      x = ~f64_orderable(reinterpret_cast<const double*>(in_a)[base + i]);
      y = (u32)r;
    } else if (MODE == MERGE_F32_DESC) {
      x = (u64)(~f32_orderable(reinterpret_cast<const float*>(in_a)[base + i]));
      y = (u32)r;
    } else {
      x = (u64)(~i32_orderable(reinterpret_cast<const int*>(in_a)[base + i]));
      y = (u32)r;
    }
  };
  const int m = block_select_sorted<MERGE_THREADS>(get, n_in, k, s_k1, s_k2, cap, sc);
  for (int j = threadIdx.x; j < k; j += MERGE_THREADS) {
    const bool have = j < m;
    merge_write<MODE>(out_a, out_idx, (size_t)q * k + j, have, have ? s_k1[j] : 0,
                      have ? s_k2[j] : 0, 0);
  }
  if (out_count && threadIdx.x == 0) out_count[q] = m;
}

template <int MODE>
static int launch_merge_typed(const void* in_a, const long long* in_idx, int q, long long n_in, int k,
                              void* out_a, long long* out_idx, int* out_count, cudaStream_t st) {
  const int cap = merge_cap(k);
  merge_typed_kernel<MODE><<<q, MERGE_THREADS, (size_t)cap * 12, st>>>(in_a, in_idx, n_in, k, cap,
                                                                      out_a, out_idx, out_count);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

}  // namespace rr
