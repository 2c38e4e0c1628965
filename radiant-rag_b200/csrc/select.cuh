// Block-wide exact top-k selection on (u64 key1 asc, u32 key2 asc) pairs.
//
// Used by every merge step of the hot path: per-slab Hamming candidates, per-tile
// BM25 candidates, exact-scan scores and the post-allgather shard merge
// (SURVEY.md 8e).  MSB-first byte-wise radix select with two shortcuts:
//   * bytes on which all valid keys agree are skipped (one OR/AND pre-pass);
//   * as soon as "definitely selected" + "still undecided" fits the shared-memory
//     staging area the radix passes stop and a bitonic sort finishes the job.
// Ties on key1 are resolved by key2 (row id), so results are deterministic and
// match the canonical tie rules of SURVEY.md 8(a).
#pragma once

#include "common.cuh"

namespace rr {

template <int THREADS>
struct SelectScratch {
  int hist[256];
  u64 red_or[THREADS / 32];
  u64 red_and[THREADS / 32];
  int red_cnt[THREADS / 32];
  u64 b_or;
  u64 b_and;
  int b_valid;
  int b_bucket;
  int b_need;
  int b_ceq;
  int count;
  u64 kth;            // result slot of block_kth_smallest
  u64 tmin[THREADS];  // per-thread minimum key1 (bound shortcut)
  u32 tidx[THREADS];
};

// One radix pass worth of bucket search: warp 0 scans hist[256] for the bucket that
// holds the `need`-th element; results go to sc.b_bucket / b_need / b_ceq.
template <int THREADS>
__device__ __forceinline__ void select_find_bucket(SelectScratch<THREADS>& sc, int need) {
  const int tid = threadIdx.x;
  if (tid < 32) {
    const int lane = tid;
    int c[8];
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = sc.hist[lane * 8 + j];
      s += c[j];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int excl = incl - s;
    if (excl < need && incl >= need) {
      int run = excl;
      int bsel = 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (run + c[j] >= need) {
          bsel = j;
          break;
        }
        run += c[j];
      }
      sc.b_bucket = lane * 8 + bsel;
      sc.b_need = need - run;
      sc.b_ceq = c[bsel];
    }
  }
}

// Bitonic sort of p (power of two) pairs in shared memory, ascending (k1, k2).
template <int THREADS>
__device__ __forceinline__ void block_bitonic_sort_pairs(u64* s_k1, u32* s_k2, int p) {
  for (int size = 2; size <= p; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (p >> 1); t += THREADS) {
        const int i = ((t / stride) * (stride << 1)) + (t % stride);
        const int j = i + stride;
        const bool asc = ((i & size) == 0);
        const u64 a1 = s_k1[i], b1 = s_k1[j];
        const u32 a2 = s_k2[i], b2 = s_k2[j];
        const bool a_gt_b = pair_less(b1, b2, a1, a2);
        if (a_gt_b == asc) {
          s_k1[i] = b1;
          s_k2[i] = b2;
          s_k1[j] = a1;
          s_k2[j] = a2;
        }
      }
      __syncthreads();
    }
  }
}

// k-th smallest value (rank 0-based, with multiplicity) of one u64 key per thread.
// Each warp sorts its 32 keys in registers (shuffle bitonic network, no block barrier) and
// parks the sorted run in shared memory; then ONE warp bisects the value range, lane w
// counting the elements <= mid in run w by binary search.  Two block barriers and a few
// thousand warp instructions in total, against ~50 barriers for a block-wide bitonic sort.
// `runs` is [THREADS] u64 scratch, `result` one u64 in shared memory.
template <int THREADS>
__device__ __forceinline__ u64 block_kth_smallest(u64 key, int rank_wanted, u64* runs, u64* result) {
  static_assert(THREADS % 32 == 0 && THREADS <= 1024, "one lane per sorted run");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64 x = key;
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const u64 y = __shfl_xor_sync(0xffffffffu, x, stride);
      const bool up = (lane & size) == 0;       // this block of the network sorts ascending
      const bool lower = (lane & stride) == 0;  // this lane keeps the smaller element when ascending
      x = ((lower == up) == (y < x)) ? y : x;
    }
  }
  runs[tid] = x;
  __syncthreads();
  if (warp == 0) {
    constexpr int RUNS = THREADS / 32;
    const bool have = lane < RUNS;
    const u64* r = runs + (have ? lane : 0) * 32;
    u64 lo = have ? r[0] : ~0ull, hi = have ? r[31] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u64 l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = l2 < lo ? l2 : lo;
      hi = h2 > hi ? h2 : hi;
    }
    while (lo < hi) {  // smallest value v with #(keys <= v) > rank_wanted
      const u64 mid = lo + ((hi - lo) >> 1);
      int cnt = 0;
      if (have) {
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
          if (r[cnt + step - 1] <= mid) cnt += step;
        if (cnt == 31 && r[31] <= mid) cnt = 32;
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt > rank_wanted) hi = mid;
      else lo = mid + 1;
    }
    if (lane == 0) *result = lo;
  }
  __syncthreads();
  return *result;
}

// Select the min(k, #valid) smallest pairs among elements 0..n-1 produced by
// get(i, k1, k2) (k1 == K1_INVALID marks a hole) and leave them SORTED at
// s_k1[0..ret), s_k2[0..ret).  `cap` is the capacity of s_k1/s_k2, a power of two
// >= k.  All threads of the block must call this; returns the same value to all.
//
// TRY_BOUND (for keys with few ties and n >> k, e.g. float scores): the k-th smallest of
// the per-thread minimum keys is the key of an element that has >= k elements at or
// below it, so one compare-and-gather pass against it keeps ~k candidates and the radix
// passes (whose shared-memory histogram serialises when keys share their high bytes)
// are skipped.  If ties push the candidate count past `cap` the radix path runs as usual.
template <int THREADS, bool TRY_BOUND = false, class Get>
__device__ int block_select_sorted(Get get, long long n, int k, u64* s_k1, u32* s_k2, int cap,
                                   SelectScratch<THREADS>& sc) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  // ---- pass 0: count valid, OR / AND of valid keys
  u64 vor = 0, vand = ~0ull;
  u64 tmin = K1_INVALID;
  int vcnt = 0;
  for (long long i = tid; i < n; i += THREADS) {
    u64 a;
    u32 b;
    get(i, a, b);
    if (a != K1_INVALID) {
      vor |= a;
      vand &= a;
      ++vcnt;
      if (TRY_BOUND) tmin = a < tmin ? a : tmin;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vor |= __shfl_xor_sync(0xffffffffu, vor, o);
    vand &= __shfl_xor_sync(0xffffffffu, vand, o);
    vcnt += __shfl_xor_sync(0xffffffffu, vcnt, o);
  }
  __syncthreads();  // protect scratch reuse across back-to-back calls
  if (lane == 0) {
    sc.red_or[warp] = vor;
    sc.red_and[warp] = vand;
    sc.red_cnt[warp] = vcnt;
  }
  __syncthreads();
  if (tid == 0) {
    u64 o = 0, a = ~0ull;
    int c = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
      o |= sc.red_or[w];
      a &= sc.red_and[w];
      c += sc.red_cnt[w];
    }
    sc.b_or = o;
    sc.b_and = a;
    sc.b_valid = c;
    sc.count = 0;
  }
  __syncthreads();
  const int valid = sc.b_valid;
  const int kk = valid < k ? valid : k;
  if (kk == 0) return 0;
  const u64 v_or = sc.b_or;
  const u64 diff = v_or ^ sc.b_and;

  if (TRY_BOUND && valid > cap && k <= THREADS / 2) {
    const u64 bound = block_kth_smallest<THREADS>(tmin, kk - 1, sc.tmin, &sc.kth);
    if (bound != K1_INVALID) {
      for (long long i = tid; i < n; i += THREADS) {
        u64 a;
        u32 b;
        get(i, a, b);
        if (a <= bound) {
          const int slot = atomicAdd(&sc.count, 1);
          if (slot < cap) {
            s_k1[slot] = a;
            s_k2[slot] = b;
          }
        }
      }
      __syncthreads();
      const int cnt = sc.count;
      if (cnt <= cap) {
        int p = 1;
        while (p < cnt) p <<= 1;
        for (int i = cnt + tid; i < p; i += THREADS) {
          s_k1[i] = K1_INVALID;
          s_k2[i] = K2_INVALID;
        }
        __syncthreads();
        block_bitonic_sort_pairs<THREADS>(s_k1, s_k2, p);
        return cnt < kk ? cnt : kk;
      }
      __syncthreads();
      if (tid == 0) sc.count = 0;
      __syncthreads();
    }
  }

  // ---- radix select on key1
  u64 prefix = 0, mask = 0;
  int need = kk;     // how many of the undecided (prefix-matching) elements we still need
  int ceq = valid;   // how many elements match the prefix so far
  bool k1_done = false;
  {
    int byte = 7;
    while (true) {
      if ((kk - need) + ceq <= cap) break;  // everything undecided fits: sort finishes it
      if (byte < 0) {
        k1_done = true;
        break;
      }
      const u64 bm = 0xFFull << (8 * byte);
      if ((diff & bm) == 0) {
        prefix |= (v_or & bm);
        mask |= bm;
        --byte;
        continue;
      }
      for (int i = tid; i < 256; i += THREADS) sc.hist[i] = 0;
      __syncthreads();
      for (long long i = tid; i < n; i += THREADS) {
        u64 a;
        u32 b;
        get(i, a, b);
        if (a != K1_INVALID && (a & mask) == prefix) atomicAdd(&sc.hist[(int)((a >> (8 * byte)) & 0xFF)], 1);
      }
      __syncthreads();
      select_find_bucket<THREADS>(sc, need);
      __syncthreads();
      prefix |= ((u64)sc.b_bucket) << (8 * byte);
      mask |= bm;
      need = sc.b_need;
      ceq = sc.b_ceq;
      --byte;
    }
  }

  // ---- radix select on key2 among key1 ties (only when they do not fit)
  u32 p2 = 0, m2 = 0;
  if (k1_done) {
    int byte = 3;
    while (byte >= 0 && (kk - need) + ceq > cap) {
      const u32 bm = 0xFFu << (8 * byte);
      for (int i = tid; i < 256; i += THREADS) sc.hist[i] = 0;
      __syncthreads();
      for (long long i = tid; i < n; i += THREADS) {
        u64 a;
        u32 b;
        get(i, a, b);
        if (a == prefix && (b & m2) == p2) atomicAdd(&sc.hist[(int)((b >> (8 * byte)) & 0xFF)], 1);
      }
      __syncthreads();
      select_find_bucket<THREADS>(sc, need);
      __syncthreads();
      p2 |= ((u32)sc.b_bucket) << (8 * byte);
      m2 |= bm;
      need = sc.b_need;
      ceq = sc.b_ceq;
      --byte;
    }
  }

  // ---- gather decided + undecided
  for (long long i = tid; i < n; i += THREADS) {
    u64 a;
    u32 b;
    get(i, a, b);
    if (a == K1_INVALID) continue;
    const u64 am = a & mask;
    bool take = false;
    if (am < prefix) {
      take = true;
    } else if (am == prefix) {
      take = !k1_done || ((b & m2) <= p2);
    }
    if (take) {
      const int slot = atomicAdd(&sc.count, 1);
      if (slot < cap) {
        s_k1[slot] = a;
        s_k2[slot] = b;
      }
    }
  }
  __syncthreads();
  int cnt = sc.count < cap ? sc.count : cap;
  int p = 1;
  while (p < cnt) p <<= 1;
  for (int i = cnt + tid; i < p; i += THREADS) {
    s_k1[i] = K1_INVALID;
    s_k2[i] = K2_INVALID;
  }
  __syncthreads();
  block_bitonic_sort_pairs<THREADS>(s_k1, s_k2, p);
  return cnt < kk ? cnt : kk;
}

}  // namespace rr
