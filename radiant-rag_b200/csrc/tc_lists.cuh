// Per-query selection over the list segments the tensor-core filter passes leave behind
// (tc_search.cu: int32 scores / Hamming; tc_exact.cu: float32 cosine keys).
#pragma once

#include "common.cuh"
#include "merge.cuh"

namespace rr {

// thresholds are clamped so that bias + score cannot wrap (|score| <= 127 * 127 * 1024 < 2^25);
// INT_MIN ("keep everything") therefore becomes -2^30
__host__ __device__ __forceinline__ int tc_tau_eff(int tau) { return tau < -(1 << 30) ? -(1 << 30) : tau; }
constexpr int TC_PACKED_SCALE = 255;  // packed-mode score = 255 * (popc(q) - hamming)

// exact top-k of each query's filtered list segments, (score desc, row asc)
constexpr int LIST_THREADS = 512;
constexpr int LIST_STAGE_CAP = 6144;  // list entries of one query staged in shared memory (48 KB)
template <int MODE>  // MERGE_I32_DESC: scores as is; MERGE_HAMMING: dist = popc(q) - score / hamming_scale;
                     // MERGE_F32_DESC: list_score holds ~orderable(float32) keys (0xFFFFFFFF = dropped by the refine)
__global__ void __launch_bounds__(LIST_THREADS, 2)
    tc_select_lists_kernel(const u32* cnt, const int* list_score, const u32* list_row, int n_cta, int cap_cta,
                           int k, int kcap, int dim, const int8_t* q_pm1, const int* tau, long long row_base,
                           void* out_a, long long* out_idx, int* out_count, unsigned* overflow,
                           unsigned char* overflow_flags, int hamming_scale) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u64* s_k1 = reinterpret_cast<u64*>(merge_smem);
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + kcap);
  u32* st_key = s_k2 + kcap;               // [LIST_STAGE_CAP] compacted list of this query
  u32* st_row = st_key + LIST_STAGE_CAP;   // [LIST_STAGE_CAP]
  __shared__ SelectScratch<LIST_THREADS> sc;
  __shared__ u32 s_seg[256];  // n_cta <= 148
  __shared__ u32 s_off[256];
  __shared__ u32 s_total;
  __shared__ int s_qpop;
  __shared__ int s_ovf;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) s_ovf = 0;
  if (MODE == MERGE_HAMMING) {
    if (threadIdx.x == 0) s_qpop = 0;
    __syncthreads();
    int c = 0;
    for (int d = threadIdx.x; d < dim; d += LIST_THREADS) c += q_pm1[(size_t)q * dim + d] > 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_qpop, c);
  }
  for (int i = threadIdx.x; i < n_cta; i += LIST_THREADS) {
    const u32 c = cnt[(size_t)q * n_cta + i];
    if (c > (u32)cap_cta) {
      atomicAdd(overflow, 1u);
      s_ovf = 1;  // this query's result is not guaranteed: the caller redoes it on the exact path
    }
    s_seg[i] = c < (u32)cap_cta ? c : (u32)cap_cta;
  }
  __syncthreads();
  if (overflow_flags && threadIdx.x == 0) overflow_flags[q] = (unsigned char)s_ovf;
  const int* ls = list_score + (size_t)q * n_cta * cap_cta;
  const u32* lr = list_row + (size_t)q * n_cta * cap_cta;
  // i / cap_cta by multiply-high where that is exact (i * cap_cta < 2^32 for every slot)
  const u32 magic = ((u64)n_cta * cap_cta * cap_cta < 0x100000000ull)
                        ? (u32)((0x100000000ull + (u32)cap_cta - 1) / (u32)cap_cta) : 0u;
  auto get = [&](long long i, u64& x, u32& y) {
    const int seg = magic ? (int)__umulhi((u32)i, magic) : (int)(i / cap_cta);
    const int j = (int)i - seg * cap_cta;
    if ((u32)j < s_seg[seg]) {
      if (MODE == MERGE_F32_DESC) {
        const u32 key = (u32)ls[i];
        x = key == 0xFFFFFFFFu ? K1_INVALID : (u64)key;
      } else {
        x = (u64)(~i32_orderable(ls[i]));
      }
      y = lr[i];
    } else {
      x = K1_INVALID;
      y = K2_INVALID;
    }
  };
  // The segments are sparse (a third full on average) and every pass over them is a chain of
  // dependent L2 loads, so they are first compacted into shared memory, one warp per segment
  // with coalesced loads, and the selection runs over the dense copy.
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    u32 run = 0;
    for (int base = 0; base < n_cta; base += 32) {
      const u32 c = (base + lane < n_cta) ? s_seg[base + lane] : 0u;
      u32 incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (base + lane < n_cta) s_off[base + lane] = run + incl - c;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_total = run;
  }
  __syncthreads();
  const u32 total = s_total;
  int m;
  if (total <= (u32)LIST_STAGE_CAP) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int seg = warp; seg < n_cta; seg += LIST_THREADS / 32) {
      const u32 c = s_seg[seg], o = s_off[seg];
      const size_t src = (size_t)seg * cap_cta;
      for (u32 j = lane; j < c; j += 32) {
        st_key[o + j] = MODE == MERGE_F32_DESC ? (u32)ls[src + j] : ~i32_orderable(ls[src + j]);
        st_row[o + j] = lr[src + j];
      }
    }
    __syncthreads();
    auto get_dense = [&](long long i, u64& x, u32& y) {
      x = (MODE == MERGE_F32_DESC && st_key[i] == 0xFFFFFFFFu) ? K1_INVALID : (u64)st_key[i];
      y = st_row[i];
    };
    m = block_select_sorted<LIST_THREADS, true>(get_dense, (long long)total, k, s_k1, s_k2, kcap, sc);
  } else {
    m = block_select_sorted<LIST_THREADS, true>(get, (long long)n_cta * cap_cta, k, s_k1, s_k2, kcap, sc);
  }
  for (int j = threadIdx.x; j < k; j += LIST_THREADS) {
    const size_t o = (size_t)q * k + j;
    if (MODE == MERGE_F32_DESC) {
      reinterpret_cast<float*>(out_a)[o] = j < m ? f32_from_orderable(~(u32)s_k1[j]) : 0.0f;
      out_idx[o] = j < m ? (long long)s_k2[j] + row_base : -1;
    } else if (j < m) {
      const int s = i32_from_orderable((u32)(~s_k1[j])) + tc_tau_eff(tau[q]);  // lists hold score - tau_eff
      reinterpret_cast<int*>(out_a)[o] = (MODE == MERGE_HAMMING) ? s_qpop - s / hamming_scale : s;
      out_idx[o] = (long long)s_k2[j] + row_base;
    } else {
      reinterpret_cast<int*>(out_a)[o] = (MODE == MERGE_HAMMING) ? 0x7fffffff : (int)0x80000000;
      out_idx[o] = -1;
    }
  }
  if (out_count && threadIdx.x == 0) out_count[q] = m;
}

}  // namespace rr
