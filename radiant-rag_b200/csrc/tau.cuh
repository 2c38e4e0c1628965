// Per-query bound from a sample of order-preserving keys, shared by the tensor-core Hamming /
// int8 scan (int32 scores) and the batched BM25 filter (float32 scores).
#pragma once

#include "common.cuh"
#include "select.cuh"

namespace rr {

// key = ~orderable(score) (smaller key = better score), 0xFFFFFFFF = no sample in this slot
template <bool FLOAT_OUT>
__device__ __forceinline__ void tau_store(void* out, int q, u32 key, bool have) {
  if (FLOAT_OUT)
    reinterpret_cast<float*>(out)[q] = have ? f32_from_orderable(~key) : 0.0f;   // 0 = no bound
  else
    reinterpret_cast<int*>(out)[q] = have ? i32_from_orderable(~key) : (int)0x80000000;
}

// tau_q = a score that at least k sample rows reach (INT_MIN when the sample holds fewer than
// k valid rows).  Keys are ~orderable(score).  One 1024-thread CTA per query: one pass for the
// per-thread minima and a sort of those (k <= 256), else the exact k-th smallest key by an
// MSB-first byte radix select over the bytes that vary.
constexpr int TAU_THREADS = 1024;
template <bool FLOAT_OUT>
__global__ void __launch_bounds__(TAU_THREADS) tau_keys_kernel(const u32* keys, long long n, int k, void* tau_out,
                                                                int allow_minima) {
  __shared__ SelectScratch<TAU_THREADS> sc;
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint4* kq = reinterpret_cast<const uint4*>(keys + (size_t)q * n);  // n is a multiple of 128
  const long long n4 = n >> 2;
  u32 vor = 0, vand = ~0u, tmin = 0xFFFFFFFFu;
  int vcnt = 0;
  for (long long i = tid; i < n4; i += TAU_THREADS) {
    const uint4 v = kq[i];
    const u32 e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (e[j] != 0xFFFFFFFFu) {
        vor |= e[j];
        vand &= e[j];
        ++vcnt;
        tmin = min(tmin, e[j]);
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vor |= __shfl_xor_sync(0xffffffffu, vor, o);
    vand &= __shfl_xor_sync(0xffffffffu, vand, o);
    vcnt += __shfl_xor_sync(0xffffffffu, vcnt, o);
  }
  if (lane == 0) {
    sc.red_or[warp] = vor;
    sc.red_and[warp] = vand;
    sc.red_cnt[warp] = vcnt;
  }
  __syncthreads();
  if (tid == 0) {
    u64 o = 0, a2 = ~0ull;
    int c = 0;
    for (int w = 0; w < TAU_THREADS / 32; ++w) {
      o |= sc.red_or[w];
      a2 &= sc.red_and[w];
      c += sc.red_cnt[w];
    }
    sc.b_or = o;
    sc.b_and = a2;
    sc.b_valid = c;
  }
  __syncthreads();
  if (sc.b_valid < k) {
    if (tid == 0) tau_store<FLOAT_OUT>(tau_out, q, 0xFFFFFFFFu, false);
    return;
  }
  // Any score that at least k sample rows reach is a valid bound, it need not be the exact
  // k-th: the k-th smallest of the per-thread minimum keys is one (the k smallest minima
  // belong to k different rows) and costs one ranking of 1024 keys instead of the radix
  // passes; it lets ~10% more rows through the filter pass than the exact k-th would.
  // (allow_minima = 0: go straight to the exact k-th - cheaper when a thread holds only a few keys,
  //  as for the BM25 sample: the single-warp bisection below costs more than the radix passes there)
  if (allow_minima && k <= TAU_THREADS / 4) {
    const u64 bound = block_kth_smallest<TAU_THREADS>(tmin == 0xFFFFFFFFu ? K1_INVALID : (u64)tmin, k - 1, sc.tmin,
                                                       &sc.kth);
    if (bound != K1_INVALID) {
      if (tid == 0) tau_store<FLOAT_OUT>(tau_out, q, (u32)bound, true);
      return;
    }
    __syncthreads();
  }
  const u32 v_or = (u32)sc.b_or;
  const u32 diff = v_or ^ (u32)sc.b_and;
  u32 prefix = 0, mask = 0;
  int need = k;
  for (int byte = 3; byte >= 0; --byte) {
    const u32 bm = 0xFFu << (8 * byte);
    if ((diff & bm) == 0) {
      prefix |= v_or & bm;
      mask |= bm;
      continue;
    }
    for (int i = tid; i < 256; i += TAU_THREADS) sc.hist[i] = 0;
    __syncthreads();
    for (long long i = tid; i < n4; i += TAU_THREADS) {
      const uint4 v = kq[i];
      const u32 e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e[j] != 0xFFFFFFFFu && (e[j] & mask) == prefix) atomicAdd(&sc.hist[(e[j] >> (8 * byte)) & 0xFF], 1);
    }
    __syncthreads();
    select_find_bucket<TAU_THREADS>(sc, need);
    __syncthreads();
    prefix |= ((u32)sc.b_bucket) << (8 * byte);
    mask |= bm;
    need = sc.b_need;
  }
  if (tid == 0) tau_store<FLOAT_OUT>(tau_out, q, prefix, true);
}

}  // namespace rr
