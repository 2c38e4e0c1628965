// Shared device/host helpers for librr_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/radiant_rag_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

namespace rr {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();
int max_smem_optin();

#define RR_CHECK_ARG(cond, msg)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      rr::set_error("%s: %s", __func__, msg);           \
      return RR_ERR_INVALID;                            \
    }                                                   \
  } while (0)

#define RR_CUDA(call)                                                             \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      rr::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
      return RR_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

#define RR_LAUNCH_CHECK()                                                            \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      rr::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return RR_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ __forceinline__ size_t align_up_dev(size_t x, size_t a = 16) { return (x + a - 1) / a * a; }

// ---- key encodings -------------------------------------------------------------
constexpr u64 K1_INVALID = ~0ull;
constexpr u32 K2_INVALID = ~0u;

// float -> u32 whose unsigned order equals the float order (-inf < ... < +inf).
__device__ __forceinline__ u32 f32_orderable(float f) {
  u32 b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(u32 o) {
  u32 b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
__device__ __forceinline__ u64 f64_orderable(double d) {
  u64 b = (u64)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_orderable(u64 o) {
  u64 b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)b);
}
// int32 -> u32 preserving order
__device__ __forceinline__ u32 i32_orderable(int v) { return (u32)v ^ 0x80000000u; }
__device__ __forceinline__ int i32_from_orderable(u32 o) { return (int)(o ^ 0x80000000u); }

__device__ __forceinline__ bool pair_less(u64 a1, u32 a2, u64 b1, u32 b2) {
  return (a1 < b1) || (a1 == b1 && a2 < b2);
}

}  // namespace rr
