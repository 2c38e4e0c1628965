// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a only).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace rr {

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ u32 tc_smem(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(u32 bar, u32 parity) {
  u32 done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// for waiters that are far ahead of their partner (the copy producer): back off between
// polls so the spin does not take issue slots from the warps sharing the scheduler
__device__ __forceinline__ void tc_mbar_wait_backoff(u32 bar, u32 parity) {
  u32 done;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
__device__ __forceinline__ void tc_tma_load_2d(u32 dst, const CUtensorMap* map, int c0, int c1, u32 bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u32 bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(u32 d_tmem, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// Warp-uniform issue helpers: every lane executes the asm with identical (uniform) operands
// and elect.sync picks the single issuing lane, so the compiler keeps descriptors and
// addresses in uniform registers instead of broadcasting them out of a divergent branch
// (the MMA warp's own instruction latency is what paces short MMAs).
__device__ __forceinline__ void tc_mma_i8_ss_elect(u32 d_tmem, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, pa;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_mma_i8_ts_elect(u32 d_tmem, u32 a_tmem, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, pa;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// kind::tf32: A and B are float32 words in shared memory (K-major, SWIZZLE_128B: 32 floats per row
// of the swizzle atom, K = 8 per instruction); the low 13 mantissa bits are ignored, D is float32
__device__ __forceinline__ void tc_mma_tf32_ss_elect(u32 d_tmem, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, pa;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(u32 bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar) : "memory");
}
// A operand from tensor memory (lane = row, four int8 K-elements per 32-bit column)
__device__ __forceinline__ void tc_mma_i8_ts(u32 d_tmem, u32 a_tmem, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B (8-row group) |
// version 1 (Blackwell) | layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ u64 tc_smem_desc(u32 saddr) {
  return (u64)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::i8: D = S32 (bits 4-5 = 2), A format at bit 7 and B format at
// bit 10 (0 = unsigned int8, 1 = signed int8), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
// kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (format 2), both K-major
__host__ __device__ constexpr u32 tc_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((u32)(n >> 3) << 17) | ((u32)(m >> 4) << 24);
}
__host__ __device__ constexpr u32 tc_idesc_i8(int m, int n, bool a_signed, bool b_signed) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((u32)(n >> 3) << 17) |
         ((u32)(m >> 4) << 24);
}

}  // namespace rr
