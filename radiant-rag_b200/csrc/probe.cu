// Peak probes (measurement aids, not on the product path): the denominators of bench.py's
// roofline lines are MEASURED on the box instead of assumed (tools/peak_probe.py).
//
//   rr_probe_popc    : every thread of a full grid issues independent POPC chains -> popc32/s
//   rr_probe_i8_mma  : one CTA per SM, one warp issues tcgen05.mma.kind::i8 back to back from
//                      resident shared-memory / tensor-memory operands (no loads, no epilogue)
//                      -> int8 op/s of the tensor pipe at the clock the box actually runs
//   rr_probe_smem    : every warp streams 128-bit shared-memory loads -> shared-memory B/s
// Each call times its kernel with CUDA events on the given stream and synchronises.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace rr {

__global__ void __launch_bounds__(256) probe_popc_kernel(int iters, u32* sink) {
  // 8 independent dependent chains x = popc(x) ^ k per thread: one POPC + one LOP3 per step, so
  // the figure is a LOWER bound of the POPC pipe when it is not the narrower of the two
  u32 x[8], k[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[j] = threadIdx.x * 2654435761u + blockIdx.x * 40503u + j * 0x9E3779B9u;
    k[j] = x[j] * 0x85EBCA6Bu + sink[1];  // not a compile-time constant
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __popc(x[j]) ^ k[j];
  }
  u32 s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 0xFFFFFFFFu) sink[0] = s;  // practically never: keeps the chains alive
}

__global__ void __launch_bounds__(512) probe_smem_kernel(int iters, float* sink) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  float4* s4 = reinterpret_cast<float4*>(ps_smem);
  const int n4 = 65536 / 16;  // 64 KB window: the 8 loads of an iteration hit 8 different addresses
  for (int i = threadIdx.x; i < n4; i += blockDim.x) s4[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int p = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = s4[(p + u * 512) & (n4 - 1)];
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
      a.w += v.w;
    }
    p += 64;
  }
  if (a.x + a.y + a.z + a.w == -1.0f) sink[0] = a.x;
}

// mode 0: A and B from shared memory, N = 256;  1: SS, N = 128;  2: A from tensor memory, N = 128
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe_mma_kernel(int iters) {
  extern __shared__ __align__(1024) unsigned char pm_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up_dev((size_t)pm_raw, 1024));
  constexpr int N = MODE == 0 ? 256 : 128;
  unsigned char* sa = base;            // 128 rows x 128 B (K = 128 int8), SWIZZLE_128B
  unsigned char* sb = base + 16384;    // N rows x 128 B
  __shared__ u64 bar;
  __shared__ u32 tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<u32*>(base)[i] = 0x01FF01FFu;
  if (threadIdx.x == 0) {
    tc_mbar_init(tc_smem(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy fills -> MMA reads
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *reinterpret_cast<volatile u32*>(&tmem_ptr);
  if (warp == 0) {
    constexpr u32 idesc = tc_idesc_i8(128, N, MODE != 2, true);
    const u64 ad = tc_smem_desc(tc_smem(sa)), bd = tc_smem_desc(tc_smem(sb));
    const u32 d_tmem = tmem_base;            // N int32 columns
    const u32 a_tmem = tmem_base + 256;      // 32 columns per 128 bytes of K (mode 2; contents irrelevant)
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 2)
          tc_mma_i8_ts_elect(d_tmem, a_tmem + 8 * j, bd + 2 * j, idesc, 1u);
        else
          tc_mma_i8_ss_elect(d_tmem, ad + 2 * j, bd + 2 * j, idesc, 1u);
      }
    }
    tc_commit_elect(tc_smem(&bar));
    tc_mbar_wait(tc_smem(&bar), 0);
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <class Launch>
static int probe_time(Launch launch, cudaStream_t st, float* ms) {
  cudaEvent_t e0, e1;
  RR_CUDA(cudaEventCreate(&e0));
  RR_CUDA(cudaEventCreate(&e1));
  launch();  // warm-up
  RR_CUDA(cudaStreamSynchronize(st));
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    RR_CUDA(cudaEventRecord(e0, st));
    launch();
    RR_CUDA(cudaEventRecord(e1, st));
    RR_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    RR_CUDA(cudaEventElapsedTime(&t, e0, e1));
    if (t < best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  RR_LAUNCH_CHECK();
  *ms = best;
  return RR_OK;
}

}  // namespace rr

using namespace rr;

extern "C" int rr_probe_popc(int32_t iters, double* out_popc32_per_s, void* stream) {
  RR_CHECK_ARG(iters > 0 && out_popc32_per_s, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  u32* sink = nullptr;
  RR_CUDA(cudaMalloc(&sink, 256));
  RR_CUDA(cudaMemsetAsync(sink, 0, 256, st));
  int sms = sm_count() > 0 ? sm_count() : 148;
  const int blocks = sms * 8;
  float ms = 0.f;
  int rc = probe_time([&] { probe_popc_kernel<<<blocks, 256, 0, st>>>(iters, sink); }, st, &ms);
  cudaFree(sink);
  if (rc != RR_OK) return rc;
  *out_popc32_per_s = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
  return RR_OK;
}

extern "C" int rr_probe_smem(int32_t iters, double* out_bytes_per_s, void* stream) {
  RR_CHECK_ARG(iters > 0 && out_bytes_per_s, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* sink = nullptr;
  RR_CUDA(cudaMalloc(&sink, 256));
  int sms = sm_count() > 0 ? sm_count() : 148;
  const int blocks = sms * 2;
  float ms = 0.f;
  RR_CUDA(cudaFuncSetAttribute(probe_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  int rc = probe_time([&] { probe_smem_kernel<<<blocks, 512, 65536, st>>>(iters, sink); }, st, &ms);
  cudaFree(sink);
  if (rc != RR_OK) return rc;
  *out_bytes_per_s = (double)blocks * 512.0 * 8.0 * 16.0 * iters / (ms * 1e-3);
  return RR_OK;
}

extern "C" int rr_probe_i8_mma(int32_t mode, int32_t iters, double* out_ops_per_s, void* stream) {
  RR_CHECK_ARG(iters > 0 && out_ops_per_s && mode >= 0 && mode <= 2, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int sms = sm_count() > 0 ? sm_count() : 148;
  const size_t smem = 1024 + 16384 + 256 * 128;
  float ms = 0.f;
  int rc;
  if (mode == 0) {
    RR_CUDA(cudaFuncSetAttribute(probe_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rc = probe_time([&] { probe_mma_kernel<0><<<sms, 128, smem, st>>>(iters); }, st, &ms);
  } else if (mode == 1) {
    RR_CUDA(cudaFuncSetAttribute(probe_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rc = probe_time([&] { probe_mma_kernel<1><<<sms, 128, smem, st>>>(iters); }, st, &ms);
  } else {
    RR_CUDA(cudaFuncSetAttribute(probe_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rc = probe_time([&] { probe_mma_kernel<2><<<sms, 128, smem, st>>>(iters); }, st, &ms);
  }
  if (rc != RR_OK) return rc;
  const double n = mode == 0 ? 256.0 : 128.0;
  *out_ops_per_s = 2.0 * 128.0 * n * 32.0 * 4.0 * iters * sms / (ms * 1e-3);
  return RR_OK;
}
