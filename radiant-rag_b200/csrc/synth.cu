// Synthetic corpora generated on device (bench / parity tests only).
// Bit-identical to radiant-rag_b200/synthetic.py: every element is a pure function of
// (seed, row, column) through the splitmix64 finaliser and integer arithmetic, so a
// 100M-row shard never has to cross PCIe and any sub-range can be regenerated on the
// CPU for a parity check.
#include "common.cuh"

namespace rr {

constexpr u64 SEED_QUERY_SALT = 0x51ED270Bull;
constexpr u64 SEED_MIX_SALT = 0x2545F491ull;
constexpr u64 SEED_LEN_SALT = 0x0D15EA5Eull;
constexpr u64 SEED_TOK_SALT = 0x7F4A7C15ull;

__device__ __forceinline__ u64 rr_splitmix64(u64 counter, u64 seed) {
  u64 z = counter + seed * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ int irwin_hall4(u64 z) {
  return (int)((z & 0xFFFF) + ((z >> 16) & 0xFFFF) + ((z >> 32) & 0xFFFF) + ((z >> 48) & 0xFFFF)) -
         131070;
}

__global__ void __launch_bounds__(256) synth_rows_kernel(float* out, long long row_start,
                                                         long long total, int dim, u64 seed,
                                                         float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const u64 ctr = (u64)row_start * (u64)dim + (u64)i;
    out[i] = (float)irwin_hall4(rr_splitmix64(ctr, seed)) * scale;
  }
}

__global__ void __launch_bounds__(256) synth_query_rows_kernel(float* out, long long q_start,
                                                               long long total, int dim, u64 seed,
                                                               long long n_corpus, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const u64 qi = (u64)q_start + (u64)(i / dim);
    const u64 col = (u64)(i % dim);
    const u64 ctr = qi * (u64)dim + col;
    int s = irwin_hall4(rr_splitmix64(ctr, seed ^ SEED_QUERY_SALT));
    if (qi & 1ull) {
      const u64 zm = rr_splitmix64(ctr, seed ^ SEED_MIX_SALT);
      if ((zm & 3ull) != 0ull) {
        const u64 src = rr_splitmix64(qi, seed ^ SEED_MIX_SALT) % (u64)(n_corpus > 0 ? n_corpus : 1);
        s = irwin_hall4(rr_splitmix64(src * (u64)dim + col, seed));
      }
    }
    out[i] = (float)s * scale;
  }
}

__global__ void __launch_bounds__(256) synth_doc_len_kernel(int* out, long long row_start,
                                                            long long n, u64 seed, int mean_len) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long s = (long long)irwin_hall4(rr_splitmix64((u64)(row_start + i), seed ^ SEED_LEN_SALT));
    long long len = (long long)mean_len + ((s * 49) >> 17);
    out[i] = (int)(len < 1 ? 1 : len);
  }
}

__global__ void __launch_bounds__(256) synth_tokens_kernel(int* out, long long pos_start, long long n,
                                                           u64 seed, const u32* cdf, int n_terms) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u32 u = (u32)(rr_splitmix64((u64)(pos_start + i), seed ^ SEED_TOK_SALT) >> 32);
    int lo = 0, hi = n_terms;  // first index with cdf[index] >= u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(cdf + mid) < u) lo = mid + 1;
      else hi = mid;
    }
    out[i] = lo;
  }
}

static unsigned blocks_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace rr

using namespace rr;

extern "C" int rr_synth_rows_f32(float* out, int64_t row_start, int64_t n_rows, int32_t dim,
                                 uint64_t seed, int32_t shift, void* stream) {
  RR_CHECK_ARG(n_rows >= 0 && dim > 0 && shift >= 0 && shift < 64, "bad size");
  if (n_rows == 0) return RR_OK;
  RR_CHECK_ARG(out, "null pointer");
  const long long total = (long long)n_rows * dim;
  synth_rows_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      out, row_start, total, dim, seed, ldexpf(1.0f, -shift));
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_synth_query_rows_f32(float* out, int64_t q_start, int64_t n_q, int32_t dim,
                                       uint64_t seed, int64_t n_corpus, int32_t shift,
                                       void* stream) {
  RR_CHECK_ARG(n_q >= 0 && dim > 0 && shift >= 0 && shift < 64, "bad size");
  if (n_q == 0) return RR_OK;
  RR_CHECK_ARG(out, "null pointer");
  const long long total = (long long)n_q * dim;
  synth_query_rows_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      out, q_start, total, dim, seed, n_corpus, ldexpf(1.0f, -shift));
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_synth_doc_lengths(int32_t* out, int64_t row_start, int64_t n, uint64_t seed,
                                    int32_t mean_len, void* stream) {
  RR_CHECK_ARG(n >= 0, "bad size");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(out, "null pointer");
  synth_doc_len_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(out, row_start, n, seed,
                                                                      mean_len);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_synth_zipf_tokens(int32_t* out, int64_t pos_start, int64_t n, uint64_t seed,
                                    const uint32_t* cdf, int32_t n_terms, void* stream) {
  RR_CHECK_ARG(n >= 0 && n_terms > 0, "bad size");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(out && cdf, "null pointer");
  synth_tokens_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(out, pos_start, n, seed, cdf,
                                                                     n_terms);
  RR_LAUNCH_CHECK();
  return RR_OK;
}
