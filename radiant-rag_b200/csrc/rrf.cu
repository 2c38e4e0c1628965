// R10: Reciprocal Rank Fusion as a segmented merge-and-fuse kernel.
//
// Replaces RRFAgent._execute (reference radiant/agents/fusion.py:61-102):
//   for run in runs: for rank, doc in enumerate(run, 1):
//       score[doc] = score.get(doc, 0.0) + 1.0 / (rrf_k + rank)          (float64)
//   fused = list(score.items())            # first-insertion order
//   fused.sort(key=score, reverse=True)    # stable: ties keep first-insertion order
//   return fused[:k]
// A document that occurs twice in one run is counted twice, as in the reference.
//
// One CTA per query.  The runs of a query arrive concatenated (run_off gives the
// segment bounds, -1 pads a short run at its tail).  Entries are sorted by
// (doc id, position); the head of each doc segment then adds its reciprocals IN
// POSITION ORDER - the same order of float64 additions as the reference's dict
// accumulation - and emits (score desc, first position asc) keys, which a second
// shared-memory bitonic sort orders.  Launch/latency-bound; bytes = q * L * 8 in,
// q * k * 16 out.
#include "common.cuh"
#include "select.cuh"

namespace rr {

constexpr int RRF_THREADS = 256;
constexpr int RRF_MAX_LEN = 4096;
constexpr int RRF_MAX_RUNS = 16;

__device__ __forceinline__ void rrf_bitonic_u64(u64* keys, int p) {
  for (int size = 2; size <= p; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (p >> 1); t += RRF_THREADS) {
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

struct RrfArgs {
  const long long* run_ptr[RRF_MAX_RUNS];  // run r of query q: run_ptr[r] + q * run_stride[r]
  int run_stride[RRF_MAX_RUNS];
  int run_off[RRF_MAX_RUNS + 1];  // position of run r inside the concatenated list
  int n_runs;
  int len;
  int p;  // power of two >= len
  double rrf_k;
  int k;
  long long* out_idx;
  double* out_score;
  int* out_count;
};

__global__ void __launch_bounds__(RRF_THREADS) rrf_fuse_kernel(const RrfArgs a) {
  extern __shared__ __align__(16) unsigned char rrf_smem[];
  u64* key_a = reinterpret_cast<u64*>(rrf_smem);  // [p] (id << 32 | pos), later reused
  u64* s_k1 = key_a + a.p;                        // [p] ~orderable(score)
  u32* s_k2 = reinterpret_cast<u32*>(s_k1 + a.p);  // [p] first position
  __shared__ int s_groups;
  const int q = blockIdx.x;
  auto entry = [&](int pos) -> long long {
    int r = 0;
    while (r + 1 < a.n_runs && pos >= a.run_off[r + 1]) ++r;
    return a.run_ptr[r][(size_t)q * a.run_stride[r] + (pos - a.run_off[r])];
  };

  for (int pos = threadIdx.x; pos < a.p; pos += RRF_THREADS) {
    u64 key = K1_INVALID;
    if (pos < a.len) {
      const long long id = entry(pos);
      if (id >= 0) key = ((u64)(u32)id << 32) | (u64)(u32)pos;
    }
    key_a[pos] = key;
    s_k1[pos] = K1_INVALID;
    s_k2[pos] = K2_INVALID;
  }
  if (threadIdx.x == 0) s_groups = 0;
  __syncthreads();
  rrf_bitonic_u64(key_a, a.p);

  // segment heads accumulate their group's reciprocals in position order
  for (int i = threadIdx.x; i < a.len; i += RRF_THREADS) {
    const u64 key = key_a[i];
    if (key == K1_INVALID) continue;
    const u32 id = (u32)(key >> 32);
    if (i > 0 && (u32)(key_a[i - 1] >> 32) == id) continue;  // not a head
    double score = 0.0;
    for (int j = i; j < a.len; ++j) {
      const u64 kj = key_a[j];
      if (kj == K1_INVALID || (u32)(kj >> 32) != id) break;
      const int pos = (int)(u32)kj;
      int r = 0;
      while (r + 1 < a.n_runs && pos >= a.run_off[r + 1]) ++r;
      const int rank = pos - a.run_off[r] + 1;
      score = __dadd_rn(score, __ddiv_rn(1.0, __dadd_rn(a.rrf_k, (double)rank)));
    }
    const int slot = atomicAdd(&s_groups, 1);
    s_k1[slot] = ~f64_orderable(score);
    s_k2[slot] = (u32)key;  // first position of this doc: unique, so slot order is irrelevant
  }
  __syncthreads();
  const int groups = s_groups;
  int p2 = 1;
  while (p2 < groups) p2 <<= 1;
  block_bitonic_sort_pairs<RRF_THREADS>(s_k1, s_k2, p2);
  const int m = groups < a.k ? groups : a.k;
  for (int j = threadIdx.x; j < a.k; j += RRF_THREADS) {
    const size_t o = (size_t)q * a.k + j;
    if (j < m) {
      a.out_idx[o] = entry((int)s_k2[j]);
      a.out_score[o] = f64_from_orderable(~s_k1[j]);
    } else {
      a.out_idx[o] = -1;
      a.out_score[o] = 0.0;
    }
  }
  if (a.out_count && threadIdx.x == 0) a.out_count[q] = m;
}

}  // namespace rr

using namespace rr;

static int launch_rrf(RrfArgs& a, int32_t n_runs, int32_t q, double rrf_k, int32_t k,
                      int64_t* out_idx, double* out_score, int32_t* out_count, void* stream);

extern "C" int rr_rrf_fuse(const int64_t* run_idx, const int32_t* run_off_host, int32_t n_runs,
                           int32_t q, double rrf_k, int32_t k, int64_t* out_idx, double* out_score,
                           int32_t* out_count, void* stream) {
  RR_CHECK_ARG(q >= 0, "negative size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(run_idx && run_off_host && out_idx && out_score, "null pointer");
  RR_CHECK_ARG(n_runs >= 1 && n_runs <= RRF_MAX_RUNS, "n_runs must be in [1, 16]");
  RR_CHECK_ARG(k >= 1 && k <= RR_MAX_K, "k out of range");
  RrfArgs a;
  for (int r = 0; r <= n_runs; ++r) a.run_off[r] = run_off_host[r];
  for (int r = 0; r < n_runs; ++r)
    RR_CHECK_ARG(a.run_off[r + 1] >= a.run_off[r], "run_off must be non-decreasing");
  RR_CHECK_ARG(a.run_off[0] == 0, "run_off[0] must be 0");
  for (int r = 0; r < n_runs; ++r) {
    a.run_ptr[r] = (const long long*)run_idx + a.run_off[r];
    a.run_stride[r] = a.run_off[n_runs];
  }
  return launch_rrf(a, n_runs, q, rrf_k, k, out_idx, out_score, out_count, stream);
}

extern "C" int rr_rrf_fuse_runs(const int64_t* const* run_ptrs_host, const int32_t* run_len_host,
                                int32_t n_runs, int32_t q, double rrf_k, int32_t k, int64_t* out_idx,
                                double* out_score, int32_t* out_count, void* stream) {
  RR_CHECK_ARG(q >= 0, "negative size");
  if (q == 0) return RR_OK;
  RR_CHECK_ARG(run_ptrs_host && run_len_host && out_idx && out_score, "null pointer");
  RR_CHECK_ARG(n_runs >= 1 && n_runs <= RRF_MAX_RUNS, "n_runs must be in [1, 16]");
  RR_CHECK_ARG(k >= 1 && k <= RR_MAX_K, "k out of range");
  RrfArgs a;
  a.run_off[0] = 0;
  for (int r = 0; r < n_runs; ++r) {
    RR_CHECK_ARG(run_len_host[r] >= 0, "negative run length");
    RR_CHECK_ARG(run_ptrs_host[r] || run_len_host[r] == 0, "null run pointer");
    a.run_ptr[r] = (const long long*)run_ptrs_host[r];
    a.run_stride[r] = run_len_host[r];
    a.run_off[r + 1] = a.run_off[r] + run_len_host[r];
  }
  return launch_rrf(a, n_runs, q, rrf_k, k, out_idx, out_score, out_count, stream);
}

static int launch_rrf(RrfArgs& a, int32_t n_runs, int32_t q, double rrf_k, int32_t k,
                      int64_t* out_idx, double* out_score, int32_t* out_count, void* stream) {
  a.len = a.run_off[n_runs];
  RR_CHECK_ARG(a.len >= 1 && a.len <= RRF_MAX_LEN, "total run length must be in [1, 4096]");
  a.n_runs = n_runs;
  a.p = 1;
  while (a.p < a.len) a.p <<= 1;
  a.rrf_k = rrf_k;
  a.k = k;
  a.out_idx = (long long*)out_idx;
  a.out_score = out_score;
  a.out_count = out_count;
  const size_t smem = (size_t)a.p * 20;
  RR_CUDA(cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  rrf_fuse_kernel<<<q, RRF_THREADS, smem, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK();
  return RR_OK;
}
