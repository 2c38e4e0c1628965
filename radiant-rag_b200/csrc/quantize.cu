// R1 / R2: embedding quantisation on device.
//
// Replaces quantize_embeddings (reference radiant/storage/quantization.py:74-108),
// which forwards to sentence_transformers.quantization.quantize_embeddings:
//   "ubinary": np.packbits(emb > 0).reshape(N, -1)      (dim 8b is the MSB of byte b)
//   "int8"   : ((emb - lo) / ((hi - lo) / 255) - 128).astype(int8), float32 arithmetic
// Both kernels are pure streaming (HBM-bound): 4 B read per dimension, 1/8 B or 1 B
// written.
#include "common.cuh"

namespace rr {

// One warp per (row, 32-dim word): coalesced 128-byte read, ballot packs the signs.
__global__ void __launch_bounds__(256) quantize_ubinary_kernel(const float* __restrict__ emb,
                                                               long long n, int dim,
                                                               uint8_t* __restrict__ codes,
                                                               int stride_bytes) {
  const int words = stride_bytes >> 2;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long tasks = n * words;
  for (long long t = warp_global; t < tasks; t += warps_total) {
    const long long row = t / words;
    const int w = (int)(t % words);
    const int d = w * 32 + lane;
    float x = 0.0f;
    if (d < dim) x = __ldg(emb + row * dim + d);
    const unsigned bal = __ballot_sync(0xffffffffu, x > 0.0f);
    if (lane == 0) {
      // ballot bit l is dimension 32w + l; packbits wants dimension 8b in the MSB of byte b
      const unsigned word = __byte_perm(__brev(bal), 0, 0x0123);
      *reinterpret_cast<unsigned*>(codes + row * stride_bytes + 4 * w) = word;
    }
  }
}

__global__ void __launch_bounds__(256) quantize_int8_kernel(const float* __restrict__ emb,
                                                            long long total, int dim,
                                                            const float* __restrict__ ranges,
                                                            int8_t* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int d = (int)(i % dim);
    const float lo = __ldg(ranges + d);
    const float hi = __ldg(ranges + dim + d);
    const float step = __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
    float v = __fsub_rn(__fdiv_rn(__fsub_rn(__ldg(emb + i), lo), step), 128.0f);
    if (v != v) v = 0.0f;  // NaN (0/0 when hi == lo) -> 0, as the oracle
    v = fminf(fmaxf(v, -128.0f), 127.0f);
    out[i] = (int8_t)(int)truncf(v);
  }
}

}  // namespace rr

extern "C" int rr_quantize_ubinary(const float* emb, int64_t n, int32_t dim, uint8_t* codes,
                                   int32_t code_stride, void* stream) {
  RR_CHECK_ARG(n >= 0 && dim > 0, "bad size");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(emb && codes, "null pointer");
  RR_CHECK_ARG(code_stride % 4 == 0 && code_stride * 8 >= dim, "code_stride must be a multiple of 4 covering dim");
  const long long tasks = (long long)n * (code_stride / 4);
  long long blocks = (tasks + 7) / 8;  // 8 warps per block
  const long long max_blocks = 148LL * 32;
  if (blocks > max_blocks) blocks = max_blocks;
  rr::quantize_ubinary_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(emb, n, dim, codes,
                                                                                  code_stride);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_quantize_int8(const float* emb, int64_t n, int32_t dim, const float* ranges,
                                int8_t* out, void* stream) {
  RR_CHECK_ARG(n >= 0 && dim > 0, "bad size");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(emb && ranges && out, "null pointer");
  const long long total = (long long)n * dim;
  long long blocks = (total + 255) / 256;
  const long long max_blocks = 148LL * 32;
  if (blocks > max_blocks) blocks = max_blocks;
  rr::quantize_int8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(emb, total, dim, ranges,
                                                                               out);
  RR_LAUNCH_CHECK();
  return RR_OK;
}
