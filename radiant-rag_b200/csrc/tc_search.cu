// Tensor-core exact search: int8 x int8 -> int32 scores on tcgen05 with a fused filter
// epilogue, for (a) BASELINE config 4 (exact int8 search, "tensor-core rescoring") and
// (b) the BATCHED Hamming scan, where the packed sign bits of the rows are expanded ON CHIP
// to a 0 / 255 unsigned-int8 operand (HBM traffic stays at 1 bit per dimension), the
// queries are +-1 int8, and  hamming = popc(q) - dot / 255  exactly (SURVEY.md section 7
// H1b: above ~3 queries per pass the POPC pipe, not HBM, bounds the popcount formulation).
//
// Kernel (one CTA per SM, 448 threads, warp-specialised):
//   warp 0        TMA producer: the CTA's 128 queries (B operand, K-major, SWIZZLE_128B) are
//                 loaded once and stay resident in shared memory.  int8 mode: corpus tiles of
//                 128 rows x 128 bytes of K (A operand) stream through an mbarrier ring of TMA
//                 boxes.  Packed mode: each 128-row tile of packed codes is ONE contiguous
//                 bulk copy into a double buffer.
//   warps 6-9     (packed mode) expand the tile's sign bits (multiply + PRMT sign-replicate,
//                 5 instructions per 8 dims) and store them with tcgen05.st into an A-operand
//                 ring that lives in TENSOR MEMORY (TC_KBPS K blocks per stage), then signal
//                 the MMA warp.
//   warp 1        allocates TMEM and issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=128,
//                 K=32 per instruction); two 128x128 int32 accumulators in TMEM are
//                 double-buffered against the epilogue.
//   warps 2-5,    epilogue, two warps per TMEM lane quarter (64 query columns each):
//   10-13         tcgen05.ld 32x32b.x32 brings a thread's row of 32 query scores to registers.
//                 Filter mode: the accumulators start at -tau_q (written by these warps), so a
//                 hit is a clear sign bit; hits are appended to the query's list.
// Exact top-k without materialising Q x N scores (164 GB in config 4):
//   pass 0  (EPI_COLMAX) the same kernel over a strided sample of row tiles keeps the best score
//           per (CTA, row slot, query); tc_tau_kernel takes a score tau_q that at least k of
//           those slots reach.  At least k rows score >= tau_q in the full set, so
//   pass 1  (EPI_FILTER) over all rows keeps exactly the rows with score >= tau_q
//           (about stride*k per query) and
//   pass 2  tc_select_lists_kernel takes the exact top-k by (score desc, row asc).
// A list that outgrows its capacity raises an overflow counter and the caller falls back
// to the CUDA-core path (never observed on the synthetic corpora; guards adversarial data).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "merge.cuh"
#include "tc_ptx.cuh"
#include "tau.cuh"
#include "tc_lists.cuh"

namespace rr {

constexpr int TC_BM = 128;          // corpus rows per MMA tile (TMEM lanes)
constexpr int TC_BN = 128;          // queries per CTA (TMEM columns per accumulator)
constexpr int TC_BK = 128;          // bytes of K per stage = one 128B swizzle atom row
constexpr int TC_MAX_STAGES = 8;
#ifndef RR_TC_EPI_GROUPS
#define RR_TC_EPI_GROUPS 2
#endif
#ifndef RR_TC_UNPACK_GROUPS
#define RR_TC_UNPACK_GROUPS 1
#endif
constexpr int TC_UNPACK_GROUPS = RR_TC_UNPACK_GROUPS;  // groups of 4 unpack warps; group g expands the K blocks with it % groups == g
constexpr int TC_EPI_GROUPS = RR_TC_EPI_GROUPS;     // groups of 4 epilogue warps; group h owns column chunks [2h, 2h+2)
constexpr int TC_EPI_CHUNKS = TC_BN / 32 / TC_EPI_GROUPS;
constexpr int TC_THREADS = 192 + 128 * TC_UNPACK_GROUPS + 128 * (TC_EPI_GROUPS - 1);  // producer, MMA, 4 epilogue,
                                                                                   // unpack, 4 more epilogue warps
constexpr int TC_TILE_BYTES = TC_BM * TC_BK;  // 16 KB
constexpr int TC_MAX_KB = 8;                  // dim <= 1024
#ifndef RR_TC_KBPS
#define RR_TC_KBPS 4
#endif
constexpr int TC_KBPS = RR_TC_KBPS;           // K blocks per tensor-memory ring stage (1, 2 or 4)

// cute::UMMA::InstrDescriptor for kind::i8: D = S32, A = B = signed int8, both K-major,
// N = 128 (>>3 at bit 17), M = 128 (>>4 at bit 24).
constexpr u32 TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((u32)(TC_BN >> 3) << 17) | ((u32)(TC_BM >> 4) << 24);
// packed mode: A = unsigned int8 (format 0), B = signed int8
constexpr u32 TC_IDESC_UA = (2u << 4) | (0u << 7) | (1u << 10) | ((u32)(TC_BN >> 3) << 17) | ((u32)(TC_BM >> 4) << 24);

struct TcArgs {
  long long n;          // corpus rows
  int q;                // queries
  int kb;               // K blocks of 128 bytes (dim / 128)
  long long tile_stride;  // launched tile i covers corpus tile i * tile_stride
  long long n_tiles;      // launched tiles
  const uint8_t* tags;
  unsigned tag_mask, tag_value;
  const int* tau;       // [q] thresholds or nullptr (keep everything; dense mode)
  int dense;            // 1: write keys [q][n_tiles*128]; 0: filtered lists
  u32* dense_keys;      // dense mode output (order-preserving ~score keys, 0xFFFFFFFF invalid)
  u32* cnt;             // [q][n_cta] list-segment lengths (written once per CTA at exit)
  int* list_score;      // [q][n_cta][cap_cta]: every CTA appends to its own segment, so the
  u32* list_row;        //   slot counters live in shared memory (no global atomics)
  int cap_cta;
  int stages;           // A-operand ring depth (<= TC_MAX_STAGES)
  int packed;           // 1: A is built in shared memory from packed sign bits (packed_codes)
  const uint8_t* packed_codes;  // [n][kb * 16] np.packbits rows (packed mode)
  int debug;            // timing experiments only (RR_TC_DEBUG bit flags): 1 = unpackers skip expand + store,
                        // 2 = MMA warp skips the MMAs, 4 = epilogue skips the TMEM reads (results are garbage)
};

// Expand 32 packed sign bits (np.packbits order: dim 8b+j is bit 7-j of byte b, bytes in
// little-endian word order) to 32 bytes 0xFF (bit set) / 0x00, as eight 32-bit words.
// Per source byte: one multiply parks bits 7,6,5,4 (resp. 3,2,1,0) on the MSBs of the four
// result bytes (the shifted copies b<<0,9,18,27 resp. b<<4,13,22,31 do not overlap), and PRMT
// in sign-replicate mode turns each MSB into a whole byte: 5 instructions per 8 dims.
// The MMA reads A as UNSIGNED int8, so a set bit is 255 and
//   score = sum_i A_i * q_i = 255 * (#(a=1,q=1) - #(a=1,q=0)) = 255 * (popc(q) - hamming(a, q)).
__device__ __forceinline__ u32 tc_prmt(u32 a, u32 b, u32 sel) {
  u32 d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ void tc_expand32(u32 w, u32* out) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const u32 b = (k == 3) ? (w >> 24) : tc_prmt(w, 0u, 0x4440u | (u32)k);  // byte k, zero-extended
    out[2 * k] = tc_prmt(b * 0x08040201u, 0u, 0xBA98u);
    out[2 * k + 1] = tc_prmt(b * 0x80402010u, 0u, 0xBA98u);
  }
}
__device__ __forceinline__ void tc_st32(u32 taddr, const u32* o) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
        "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]),
        "r"(o[16]), "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]), "r"(o[21]), "r"(o[22]), "r"(o[23]),
        "r"(o[24]), "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]), "r"(o[31])
      : "memory");
}

// Thread layout: warp 0 producer, warp 1 MMA, warps 2-5 epilogue, warps 6-9 unpackers (only
// in packed mode, where the A operand is built in shared memory from packed sign bits).
// EPI: 0 = filter pass (lists), 1 = dense keys of every launched row (tests / debug),
//      2 = sample pass: per (CTA, row slot) maximum over the CTA's tiles, one key per slot.
constexpr int EPI_FILTER = 0, EPI_DENSE = 1, EPI_COLMAX = 2;
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_i8_search_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const TcArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  // 1024-byte aligned carve-up: B blocks, A stages, [packed ring], then small state
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up_dev((size_t)tc_smem_raw, 1024));
  unsigned char* sb = base;                                   // [kb][16 KB]
  unsigned char* sa = sb + (size_t)a.kb * TC_TILE_BYTES;      // [stages][16 KB]
  const int row_bytes = a.kb * 16;                            // packed bytes per row
  unsigned char* spk = sa + (a.packed ? 0 : (size_t)a.stages * TC_TILE_BYTES);  // [2][128 * row_bytes] (packed mode)
  int* thr = reinterpret_cast<int*>(spk + (a.packed ? 2 * (size_t)TC_BM * row_bytes : 0));  // [128]
  u32* s_cnt = reinterpret_cast<u32*>(thr + TC_BN);                           // [128]
  u64* bars = reinterpret_cast<u64*>(s_cnt + TC_BN);
  u64* full_a = bars;                          // [stages]
  u64* empty_a = bars + TC_MAX_STAGES;         // [stages]
  u64* b_full = bars + 2 * TC_MAX_STAGES;      // [1]
  u64* tmem_full = b_full + 1;                 // [2]
  u64* tmem_empty = tmem_full + 2;             // [2]
  u64* pk_full = tmem_empty + 2;               // [2]
  u64* pk_empty = pk_full + 2;                 // [2]
  u32* tmem_ptr = reinterpret_cast<u32*>(pk_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.y;           // query block of 128
  const int q0 = qb * TC_BN;
  const u32 stages = (u32)a.stages;

  if (threadIdx.x == 0) {
    for (u32 s = 0; s < stages; ++s) {
      tc_mbar_init(tc_smem(full_a + s), a.packed ? 128 : 1);
      tc_mbar_init(tc_smem(empty_a + s), 1);
    }
    tc_mbar_init(tc_smem(b_full), 1);
    for (int s = 0; s < 2; ++s) {
      tc_mbar_init(tc_smem(tmem_full + s), 1);
      tc_mbar_init(tc_smem(tmem_empty + s), 128 * TC_EPI_GROUPS);
      tc_mbar_init(tc_smem(pk_full + s), 1);
      tc_mbar_init(tc_smem(pk_empty + s), 128 * TC_UNPACK_GROUPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const u32 tmem_cols = a.packed ? 512u : 256u;  // 2 accumulators x 128 columns (+ 256 columns of A ring)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(tmem_ptr)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2 && warp < 6) {
    const int t = threadIdx.x - 64;  // 0..127
    const int qq = q0 + t;
    // bias = -tau_eff: 0 in dense mode (plain scores), far negative for padding columns
    thr[t] = EPI != EPI_FILTER ? 0 : (qq < a.q ? -tc_tau_eff(a.tau[qq]) : -(1 << 30));
    s_cnt[t] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *reinterpret_cast<volatile u32*>(tmem_ptr);

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      tc_mbar_expect_tx(tc_smem(b_full), (u32)(a.kb * TC_TILE_BYTES));
      for (int kb = 0; kb < a.kb; ++kb)
        tc_tma_load_2d(tc_smem(sb + (size_t)kb * TC_TILE_BYTES), &map_b, kb * TC_BK, q0, tc_smem(b_full));
      if (!a.packed) {
        // int8 rows straight from HBM: one 128-row x 128-byte TMA box per K block
        u32 s = 0, ph = 0;
        for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x) {
          const long long row0 = i * a.tile_stride * TC_BM;
          for (int kb = 0; kb < a.kb; ++kb, s = (s + 1 == stages) ? 0 : s + 1, ph ^= (s == 0)) {
            tc_mbar_wait(tc_smem(empty_a + s), ph ^ 1u);
            tc_mbar_expect_tx(tc_smem(full_a + s), TC_TILE_BYTES);
            tc_tma_load_2d(tc_smem(sa + (size_t)s * TC_TILE_BYTES), &map_a, kb * TC_BK, (int)row0, tc_smem(full_a + s));
          }
        }
      } else {
        // packed sign bits: the whole 128-row tile is one contiguous bulk copy
        u32 tcount = 0;
        for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x, ++tcount) {
          const long long row0 = i * a.tile_stride * TC_BM;
          const u32 slot = tcount & 1u;
          tc_mbar_wait_backoff(tc_smem(pk_empty + slot), ((tcount >> 1) & 1u) ^ 1u);
          long long rows = a.n - row0;
          if (rows > TC_BM) rows = TC_BM;
          const u32 bytes = (u32)(rows * row_bytes);
          tc_mbar_expect_tx(tc_smem(pk_full + slot), bytes);
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(tc_smem(spk + (size_t)slot * TC_BM * row_bytes)),
                "l"(a.packed_codes + (size_t)row0 * row_bytes), "r"(bytes), "r"(tc_smem(pk_full + slot)) : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    tc_mbar_wait(tc_smem(b_full), 0);
    tc_fence_after();
    const u64 bdesc0 = tc_smem_desc(tc_smem(sb));
    const u64 adesc0 = tc_smem_desc(tc_smem(sa));
    const u32 full0 = tc_smem(full_a), empty0 = tc_smem(empty_a);
    u32 stage = 0, phase = 0, tcount = 0;
    for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x, ++tcount) {
      const u32 as = tcount & 1u;
      tc_mbar_wait(tc_smem(tmem_empty + as), (tcount >> 1) & 1u);  // read by the epilogue and re-biased
      tc_fence_after();
      const u32 d_tmem = tmem_base + as * TC_BN;
      if (a.packed) {
        // a ring stage holds TC_KBPS K blocks (32 TMEM columns each): one hand-off per 4*TC_KBPS MMAs
        for (int kb = 0; kb < a.kb; kb += TC_KBPS) {
          tc_mbar_wait(full0 + stage * 8, phase);
          tc_fence_after();
          const u32 a_tmem = tmem_base + 2 * TC_BN + stage * (32 * TC_KBPS);  // A ring in tensor memory
          const u64 bd = bdesc0 + (u64)(kb * (TC_TILE_BYTES >> 4));
          if (!(a.debug & 2)) {
#pragma unroll
            for (int j = 0; j < TC_KBPS; ++j) {
              if (kb + j < a.kb) {
                const u64 bdj = bd + (u64)(j * (TC_TILE_BYTES >> 4));
                tc_mma_i8_ts_elect(d_tmem, a_tmem + 32 * j, bdj, TC_IDESC_UA, 1u);
                tc_mma_i8_ts_elect(d_tmem, a_tmem + 32 * j + 8, bdj + 2, TC_IDESC_UA, 1u);
                tc_mma_i8_ts_elect(d_tmem, a_tmem + 32 * j + 16, bdj + 4, TC_IDESC_UA, 1u);
                tc_mma_i8_ts_elect(d_tmem, a_tmem + 32 * j + 24, bdj + 6, TC_IDESC_UA, 1u);
              }
            }
          }
          tc_commit_elect(empty0 + stage * 8);  // frees the A stage when these MMAs have read it
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      } else {
        for (int kb = 0; kb < a.kb; ++kb) {
          tc_mbar_wait(full0 + stage * 8, phase);
          tc_fence_after();
          const u64 ad = adesc0 + (u64)(stage * (TC_TILE_BYTES >> 4));
          const u64 bd = bdesc0 + (u64)(kb * (TC_TILE_BYTES >> 4));
          tc_mma_i8_ss_elect(d_tmem, ad, bd, TC_IDESC, 1u);
          tc_mma_i8_ss_elect(d_tmem, ad + 2, bd + 2, TC_IDESC, 1u);
          tc_mma_i8_ss_elect(d_tmem, ad + 4, bd + 4, TC_IDESC, 1u);
          tc_mma_i8_ss_elect(d_tmem, ad + 6, bd + 6, TC_IDESC, 1u);
          tc_commit_elect(empty0 + stage * 8);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      tc_commit_elect(tc_smem(tmem_full + as));  // accumulator complete
    }
  } else if (warp < 6 || warp >= 6 + 4 * TC_UNPACK_GROUPS) {
    // ===================== epilogue (warps 2..5 and the last four) =====================
    // Two warps share a TMEM lane quarter and split the 128 query columns between them: a
    // single warp's dependent-issue latency is what paces this role.
    // Filter mode keeps the compare out of the instruction stream: the accumulator of query
    // column j starts at -tau_j instead of 0 (the epilogue warps write that bias into tensor
    // memory with tcgen05.st right after they have read a tile, the MMAs always accumulate),
    // so "score >= tau" is the sign bit of the accumulator and a funnel shift per column
    // collects the hit mask.
    // (Measured alternative: no bias stores, the first MMA of a tile overwrites and the epilogue
    // subtracts tau_j from every score - 64 KB of tensor-memory stores per tile less, but 5 % slower.
    // What paces the kernel is the accumulator read-out)
    const int lq = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int half = warp < 6 ? 0 : 1 + (warp - (6 + 4 * TC_UNPACK_GROUPS)) / 4;
    const int c0 = half * TC_EPI_CHUNKS;  // first column chunk of this warp
    // (the biases are re-read from shared memory for every store: this role has slack, and
    // keeping them out of registers lets the kernel run more unpack warps)
    auto store_bias = [&](u32 as) {
#pragma unroll
      for (int cc = 0; cc < TC_EPI_CHUNKS; ++cc) {
        const u32 taddr = tmem_base + ((u32)(lq * 32) << 16) + as * TC_BN + (c0 + cc) * 32;
        u32 o[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int4 tv = reinterpret_cast<const int4*>(thr + (c0 + cc) * 32)[j4];  // thr[] holds the biases
          o[4 * j4 + 0] = (u32)tv.x;
          o[4 * j4 + 1] = (u32)tv.y;
          o[4 * j4 + 2] = (u32)tv.z;
          o[4 * j4 + 3] = (u32)tv.w;
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(taddr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
              "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]),
              "r"(o[16]), "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]), "r"(o[21]), "r"(o[22]), "r"(o[23]),
              "r"(o[24]), "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]), "r"(o[31])
            : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    // both accumulators start biased; each arrival below is "read and re-biased"
    for (u32 as = 0; as < 2; ++as) {
      store_bias(as);
      tc_fence_before();
      tc_mbar_arrive(tc_smem(tmem_empty + as));
    }
    // sample pass: best score seen by this thread's row slot, per query column (a slot sees one
    // row of every tile the CTA takes, so the slots partition the sampled rows)
    int colmax[EPI == EPI_COLMAX ? TC_EPI_CHUNKS * 32 : 1];
#pragma unroll
    for (int j = 0; j < (EPI == EPI_COLMAX ? TC_EPI_CHUNKS * 32 : 1); ++j) colmax[j] = (int)0x80000000;
    u32 tcount = 0;
    for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x, ++tcount) {
      const u32 as = tcount & 1u;
      tc_mbar_wait(tc_smem(tmem_full + as), (tcount >> 1) & 1u);
      tc_fence_after();
      const long long row = i * a.tile_stride * TC_BM + lq * 32 + lane;
      bool valid = row < a.n;
      if (valid && a.tags != nullptr) valid = ((unsigned)a.tags[row] & a.tag_mask) == a.tag_value;
      const long long dense_col = i * TC_BM + lq * 32 + lane;  // position among the launched rows
      const long long dense_ld = a.n_tiles * TC_BM;
      if (!(a.debug & 4)) {
#pragma unroll
        for (int cc = 0; cc < TC_EPI_CHUNKS; ++cc) {
          const int c = c0 + cc;
          u32 v[32];
          const u32 taddr = tmem_base + ((u32)(lq * 32) << 16) + as * TC_BN + c * 32;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (EPI == EPI_DENSE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int qq = q0 + c * 32 + j;
              if (qq < a.q)
                a.dense_keys[(size_t)qq * dense_ld + dense_col] = valid ? ~i32_orderable((int)v[j]) : 0xFFFFFFFFu;
            }
          } else if (EPI == EPI_COLMAX) {
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) colmax[cc * 32 + j] = max(colmax[cc * 32 + j], (int)v[j]);
            }
          } else {
            // four independent funnel-shift chains (one warp runs the epilogue of its lane
            // quarter alone, so dependent-issue latency, not throughput, is what it pays):
            // bit 7-i of m[g] = sign of column 8g+i (1 = below the bound)
            u32 m[4] = {0, 0, 0, 0};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
              for (int g = 0; g < 4; ++g) m[g] = __funnelshift_l(v[8 * g + jj], m[g], 1);
            }
            // byte g of `miss` = m[g]; hit bit 8g + (7-i) <-> column 8g + i
            const u32 miss = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
            u32 hit = valid ? ~miss : 0u;
            while (hit) {  // rare: about k * stride rows per query reach the sampled bound
              const int b = __ffs(hit) - 1;
              hit &= hit - 1;
              const int j = (b & 24) | (7 - (b & 7));
              // 5-level select tree instead of a 32-long dependent chain
              u32 t16[16], t8[8], t4[4], t2[2];
#pragma unroll
              for (int x = 0; x < 16; ++x) t16[x] = (j & 1) ? v[2 * x + 1] : v[2 * x];
#pragma unroll
              for (int x = 0; x < 8; ++x) t8[x] = (j & 2) ? t16[2 * x + 1] : t16[2 * x];
#pragma unroll
              for (int x = 0; x < 4; ++x) t4[x] = (j & 4) ? t8[2 * x + 1] : t8[2 * x];
#pragma unroll
              for (int x = 0; x < 2; ++x) t2[x] = (j & 8) ? t4[2 * x + 1] : t4[2 * x];
              const int s = (int)((j & 16) ? t2[1] : t2[0]);
              const u32 slot = atomicAdd(s_cnt + c * 32 + j, 1u);  // shared-memory counter
              if (slot < (u32)a.cap_cta) {
                const size_t o = ((size_t)(q0 + c * 32 + j) * gridDim.x + blockIdx.x) * a.cap_cta + slot;
                a.list_score[o] = s;  // biased: score - tau_eff (tc_select_lists adds it back)
                a.list_row[o] = (u32)row;
              }
            }
          }
        }
      }
      store_bias(as);
      tc_fence_before();
      tc_mbar_arrive(tc_smem(tmem_empty + as));
    }
    if (EPI == EPI_COLMAX) {
      const size_t ld = (size_t)gridDim.x * TC_BM;
#pragma unroll
      for (int cc = 0; cc < TC_EPI_CHUNKS; ++cc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qq = q0 + (c0 + cc) * 32 + j;
          const int mx = colmax[cc * 32 + j];
          if (qq < a.q)
            a.dense_keys[(size_t)qq * ld + (size_t)blockIdx.x * TC_BM + lq * 32 + lane] =
                mx == (int)0x80000000 ? 0xFFFFFFFFu : ~i32_orderable(mx);
        }
      }
    }
    if (EPI == EPI_FILTER) {
      asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_GROUPS) : "memory");  // the epilogue warps only
      const int t = threadIdx.x - 64;
      if (t < TC_BN && q0 + t < a.q) a.cnt[(size_t)(q0 + t) * gridDim.x + blockIdx.x] = s_cnt[t];
    }
  } else if (a.packed && warp < 6 + 4 * TC_UNPACK_GROUPS) {
    // ===================== unpackers (warps 6..9): packed bits -> +-1 int8 rows in TENSOR MEMORY =====================
    // The A operand never touches shared memory: each thread expands the 128 dims of ITS row
    // for one K block into 32 registers and stores them to its TMEM lane (tcgen05.st), so
    // shared-memory bandwidth is left to the resident B operand.
    const int lq = warp & 3;          // TMEM lane quarter this warp may access
    const int u = lq * 32 + lane;     // row of the tile = TMEM lane
    const u32 group = (u32)(warp - 6) >> 2;  // several warps per quarter take turns on the K blocks,
                                             // so one warp's expand latency hides behind the others
    // The hand-off of K block i (wait for the store, signal the MMA warp) is issued after K
    // block i+1 has been expanded, so the tensor-memory store is in flight during the
    // expansion instead of stalling the warp.
    u32 it = 0, tcount = 0, stage = 0, phase = 0;
    bool pending = false;
    u32 pending_bar = 0;
    for (long long i = blockIdx.x; i < a.n_tiles; i += gridDim.x, ++tcount) {
      const u32 slot = tcount & 1u;
      tc_mbar_wait(tc_smem(pk_full + slot), (tcount >> 1) & 1u);
      const unsigned char* prow = spk + (size_t)slot * TC_BM * row_bytes + (size_t)u * row_bytes;
      for (int kb = 0; kb < a.kb; kb += TC_KBPS, ++it) {  // TC_KBPS K blocks (128 dims each) per ring stage
        const u32 s = stage;
        const u32 ph = phase;
        if (++stage == stages) {
          stage = 0;
          phase ^= 1u;
        }
        if (TC_UNPACK_GROUPS > 1 && it % TC_UNPACK_GROUPS != group) continue;
        if (a.debug & 1) {
          tc_mbar_wait(tc_smem(empty_a + s), ph ^ 1u);
          tc_mbar_arrive(tc_smem(full_a + s));
          continue;
        }
        const uint4 pw = *reinterpret_cast<const uint4*>(prow + kb * 16);  // 128 dims of this row
        uint4 pwx[TC_KBPS];
#pragma unroll
        for (int j = 1; j < TC_KBPS; ++j)
          pwx[j] = (kb + j < a.kb) ? *reinterpret_cast<const uint4*>(prow + (kb + j) * 16) : make_uint4(0u, 0u, 0u, 0u);
        u32 o[32];
        tc_expand32(pw.x, o);
        tc_expand32(pw.y, o + 8);
        tc_expand32(pw.z, o + 16);
        tc_expand32(pw.w, o + 24);
        if (pending) {
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          tc_mbar_arrive(pending_bar);
        }
        tc_mbar_wait(tc_smem(empty_a + s), ph ^ 1u);
        tc_fence_after();
        const u32 taddr = tmem_base + ((u32)(lq * 32) << 16) + 2 * TC_BN + s * (32 * TC_KBPS);
#pragma unroll
        for (int j = 1; j < TC_KBPS; ++j) {
          if (kb + j < a.kb) {
            u32 o1[32];
            tc_expand32(pwx[j].x, o1);
            tc_expand32(pwx[j].y, o1 + 8);
            tc_expand32(pwx[j].z, o1 + 16);
            tc_expand32(pwx[j].w, o1 + 24);
            tc_st32(taddr + 32 * j, o1);
          }
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(taddr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
              "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]),
              "r"(o[16]), "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]), "r"(o[21]), "r"(o[22]), "r"(o[23]),
              "r"(o[24]), "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]), "r"(o[31])
            : "memory");
        pending = true;
        pending_bar = tc_smem(full_a + s);
      }
      tc_mbar_arrive(tc_smem(pk_empty + slot));
    }
    if (pending) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      tc_mbar_arrive(pending_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// packed sign bits (np.packbits order) -> +-1 int8 rows
__global__ void __launch_bounds__(256) unpack_pm1_kernel(const uint8_t* __restrict__ codes, long long n,
                                                         int stride_bytes, int dim, int8_t* __restrict__ out) {
  const int bytes_per_row = dim >> 3;
  const long long total = n * bytes_per_row;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const long long row = i / bytes_per_row;
    const int b = (int)(i % bytes_per_row);
    const unsigned byte = codes[row * stride_bytes + b];
    u32 lo = 0, hi = 0;  // dims 8b..8b+3 and 8b+4..8b+7; dim 8b is bit 7
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lo |= (((byte >> (7 - j)) & 1u) ? 0x01u : 0xFFu) << (8 * j);
      hi |= (((byte >> (3 - j)) & 1u) ? 0x01u : 0xFFu) << (8 * j);
    }
    *reinterpret_cast<uint2*>(out + row * dim + 8 * b) = make_uint2(lo, hi);
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// rows x dim int8 matrix, K-major; box = 128 rows x 128 bytes, SWIZZLE_128B
static int make_map(CUtensorMap* map, const void* ptr, long long rows, int dim) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return RR_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld dim=%d", (int)r, rows, dim);
    return RR_ERR_CUDA;
  }
  return RR_OK;
}

// Pass 0 scores every `stride`-th row tile.  Sampling costs ~ n / stride, the rows that later
// reach the sampled bound (list appends + final select) cost ~ stride * k per query, so the
// optimum grows like sqrt(n / k); the constant is fitted on B200 (profiles/).
static int tc_sample_stride(long long n, int k) {
#ifdef RR_TC_EXPERIMENTS  // never in the shipped library: an environment variable must not change results
  const char* e = getenv("RR_TC_STRIDE");
  if (e && atoi(e) > 0) return atoi(e);
#endif
  const double want = 0.2 * sqrt((double)n / (double)k);
  int s = 4;
  while (s < 128 && (double)s * 1.4142 < want) s <<= 1;
  return s;
}

struct TcPlan {
  long long tiles, sample_tiles, keys_per_q;
  int stride, sample_ctas;
  bool colmax;
  int qblocks, ctas_x, cap_cta;
  size_t off_keys, off_sscore, off_sidx, off_scount, off_tau, off_cnt, off_ls, off_lr, total;
};

static TcPlan tc_plan(long long n, int q, int k) {
  TcPlan p;
  p.tiles = (n + TC_BM - 1) / TC_BM;
  p.stride = tc_sample_stride(n, k);
  p.sample_tiles = (p.tiles + p.stride - 1) / p.stride;
  // the sample must be able to hold k rows; tiny corpora are sampled completely
  if (p.sample_tiles * TC_BM < 4LL * k || p.sample_tiles * TC_BM < 2048) p.sample_tiles = p.tiles;
  p.qblocks = (q + TC_BN - 1) / TC_BN;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  p.ctas_x = sms / p.qblocks;
  if (p.ctas_x < 1) p.ctas_x = 1;
  if (p.ctas_x > p.tiles) p.ctas_x = (int)p.tiles;
  // expected rows that reach the sampled bound: k * (tiles / sample_tiles); CTAs take tiles
  // round-robin, so each sees an even share; 3x headroom + slack, overflow is detected
  const long long expect = (long long)k * (p.tiles / p.sample_tiles + 1);
  long long cap = 3 * (expect / p.ctas_x + 1) + 64;
  const long long rows_per_cta = ((p.tiles + p.ctas_x - 1) / p.ctas_x) * TC_BM;
  if (cap > rows_per_cta) cap = rows_per_cta;
  p.cap_cta = (int)cap;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  // sample pass output: one key per (CTA, row slot) when those are plenty (>= 4k), else the
  // dense keys of every sampled row
  p.sample_ctas = (int)(p.sample_tiles < p.ctas_x ? p.sample_tiles : p.ctas_x);
  p.colmax = (long long)p.sample_ctas * TC_BM >= 4LL * k;
  p.keys_per_q = p.colmax ? (long long)p.sample_ctas * TC_BM : p.sample_tiles * TC_BM;
  p.off_keys = take((size_t)q * p.keys_per_q * 4);
  p.off_sscore = take((size_t)q * k * 4);
  p.off_sidx = take((size_t)q * k * 8);
  p.off_scount = take((size_t)q * 4);
  p.off_tau = take((size_t)q * 4);
  p.off_cnt = take((size_t)q * p.ctas_x * 4);
  p.off_ls = take((size_t)q * p.ctas_x * p.cap_cta * 4);
  p.off_lr = take((size_t)q * p.ctas_x * p.cap_cta * 4);
  p.total = o + 256;
  return p;
}

// optional per-kernel event timing (rr_tc_timing): events 0..4 bracket the four kernels.
// The state is per host thread (the reference drives retrieval from two threads,
// radiant/orchestrator.py:994-998): a thread that enables timing times its own calls only.
static thread_local bool g_tc_timing = false;
static thread_local bool g_tc_timed = false;
static thread_local cudaEvent_t g_tc_ev[5];
static thread_local bool g_tc_ev_ready = false;
static void tc_mark(int i, cudaStream_t st) {
  if (g_tc_timing && g_tc_ev_ready) cudaEventRecord(g_tc_ev[i], st);
}

struct TcSmem {
  int stages;
  size_t bytes;
};

// B resident + A ring (+ packed double buffer) + thresholds / counters / barriers, inside 227 KB
static TcSmem tc_smem_layout(int kb, bool packed) {
  const size_t limit = 232448;  // 227 KB opt-in maximum per block
  TcSmem r;
  if (packed) {
    // A ring lives in tensor memory (256 columns next to the two 128-column accumulators);
    // shared memory holds B, the packed double buffer and the small state only
    r.stages = 8 / TC_KBPS;  // x 32 * TC_KBPS TMEM columns (256 columns of A ring in all)
    r.bytes = 1024 + (size_t)kb * TC_TILE_BYTES + 2 * (size_t)TC_BM * kb * 16 + 2 * TC_BN * 4 + 32 * 8 + 64;
    return r;
  }
  const size_t fixed = 1024 /*alignment slack*/ + (size_t)kb * TC_TILE_BYTES + 2 * TC_BN * 4 + 32 * 8 + 64;
  int stages = (int)((limit - fixed) / TC_TILE_BYTES);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  r.stages = stages;
  r.bytes = fixed + (size_t)stages * TC_TILE_BYTES;
  return r;
}

// mode: 0 = int8 scores (score desc), 1 = Hamming over +-1 rows (dist asc)
static int tc_search(const int8_t* emb, const uint8_t* packed_codes, long long n, int dim, const uint8_t* tags,
                     unsigned tag_mask, unsigned tag_value, const int8_t* queries, int q, int k, long long row_base,
                     int hamming,
                     int* out_a, long long* out_idx, unsigned* overflow_out, unsigned char* overflow_flags,
                     void* ws, size_t ws_bytes, cudaStream_t st) {
  const TcPlan p = tc_plan(n, q, k);
  if (!ws || ws_bytes < p.total) {
    set_error("tensor-core search: workspace %zu < %zu", ws_bytes, p.total);
    return RR_ERR_WORKSPACE;
  }
  char* w = (char*)ws;
  const bool packed = packed_codes != nullptr;
  CUtensorMap map_a, map_b;
  int rc = make_map(&map_b, queries, q, dim);
  if (rc != RR_OK) return rc;
  if (packed) {
    map_a = map_b;  // unused in packed mode
  } else {
    rc = make_map(&map_a, emb, n, dim);
    if (rc != RR_OK) return rc;
  }
  const int kb = dim / TC_BK;
  const TcSmem lay = tc_smem_layout(kb, packed);
  RR_CHECK_ARG(lay.stages >= 2, "not enough shared memory for the operand ring");
  const size_t smem = lay.bytes;
  RR_CUDA(cudaFuncSetAttribute(tc_i8_search_kernel<EPI_COLMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  RR_CUDA(cudaFuncSetAttribute(tc_i8_search_kernel<EPI_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  RR_CUDA(cudaFuncSetAttribute(tc_i8_search_kernel<EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  const int qblocks = p.qblocks;
  const int ctas_x = p.ctas_x;

  TcArgs a;
  a.n = n;
  a.q = q;
  a.kb = kb;
  a.tags = tags;
  a.tag_mask = tag_mask;
  a.tag_value = tag_value;
  a.cnt = (u32*)(w + p.off_cnt);
  a.list_score = (int*)(w + p.off_ls);
  a.list_row = (u32*)(w + p.off_lr);
  a.cap_cta = p.cap_cta;
  a.dense_keys = (u32*)(w + p.off_keys);
  a.stages = lay.stages;
  a.packed = packed ? 1 : 0;
  a.packed_codes = packed_codes;
  a.debug = 0;
#ifdef RR_TC_EXPERIMENTS  // isolation runs (skip a role; results are garbage): experiment builds only
  {
    const char* e = getenv("RR_TC_DEBUG");
    a.debug = e ? atoi(e) : 0;
  }
#endif

  // ---- pass 0: dense scores of a strided sample of row tiles -> tau
  const bool full_sample = p.sample_tiles == p.tiles;
  a.tile_stride = full_sample ? 1 : p.stride;
  a.n_tiles = p.sample_tiles;
  a.tau = nullptr;
  a.dense = 1;
  tc_mark(0, st);
  {
    dim3 grid((unsigned)p.sample_ctas, qblocks);
    if (p.colmax)
      tc_i8_search_kernel<EPI_COLMAX><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, a);
    else
      tc_i8_search_kernel<EPI_DENSE><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, a);
    RR_LAUNCH_CHECK();
  }
  const int kcap = merge_cap(k);
  tc_mark(1, st);
  tau_keys_kernel<false><<<q, TAU_THREADS, 0, st>>>(a.dense_keys, p.keys_per_q, k, (void*)(w + p.off_tau), 1);
  RR_LAUNCH_CHECK();
  tc_mark(2, st);

  // ---- pass 1: filter pass over all rows (every CTA of the grid writes its cnt entries)
  a.tile_stride = 1;
  a.n_tiles = p.tiles;
  a.tau = (const int*)(w + p.off_tau);
  a.dense = 0;
  {
    dim3 grid((unsigned)ctas_x, qblocks);
    tc_i8_search_kernel<EPI_FILTER><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, a);
    RR_LAUNCH_CHECK();
  }
  tc_mark(3, st);
  // ---- pass 2: exact top-k of each query's list segments
  const size_t list_smem = (size_t)kcap * 12 + (size_t)LIST_STAGE_CAP * 8;
  RR_CUDA(cudaFuncSetAttribute(tc_select_lists_kernel<MERGE_HAMMING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)list_smem));
  RR_CUDA(cudaFuncSetAttribute(tc_select_lists_kernel<MERGE_I32_DESC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)list_smem));
  if (hamming)
    tc_select_lists_kernel<MERGE_HAMMING><<<q, LIST_THREADS, list_smem, st>>>(
        a.cnt, a.list_score, a.list_row, ctas_x, p.cap_cta, k, kcap, dim, queries, (const int*)(w + p.off_tau), row_base, out_a,
        out_idx, nullptr, overflow_out, overflow_flags, TC_PACKED_SCALE);
  else
    tc_select_lists_kernel<MERGE_I32_DESC><<<q, LIST_THREADS, list_smem, st>>>(
        a.cnt, a.list_score, a.list_row, ctas_x, p.cap_cta, k, kcap, dim, queries, (const int*)(w + p.off_tau), row_base, out_a,
        out_idx, nullptr, overflow_out, overflow_flags, TC_PACKED_SCALE);
  RR_LAUNCH_CHECK();
  tc_mark(4, st);
  g_tc_timed = g_tc_timing && g_tc_ev_ready;
  return RR_OK;
}

}  // namespace rr

using namespace rr;

extern "C" int rr_unpack_codes_pm1(const uint8_t* codes, int64_t n, int32_t code_stride, int32_t dim,
                                   int8_t* out, void* stream) {
  RR_CHECK_ARG(n >= 0 && dim > 0 && dim % 8 == 0 && code_stride * 8 >= dim, "bad size (dim must be a multiple of 8)");
  if (n == 0) return RR_OK;
  RR_CHECK_ARG(codes && out, "null pointer");
  const long long total = (long long)n * (dim / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  unpack_pm1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(codes, n, code_stride, dim, out);
  RR_LAUNCH_CHECK();
  return RR_OK;
}

extern "C" int rr_tc_timing(int32_t enable) {
  if (enable && !g_tc_ev_ready) {
    for (int i = 0; i < 5; ++i) RR_CUDA(cudaEventCreate(&g_tc_ev[i]));
    g_tc_ev_ready = true;
  }
  g_tc_timing = enable != 0;
  if (!g_tc_timing) g_tc_timed = false;
  return RR_OK;
}

extern "C" int rr_tc_last_timing_ms(float* out_ms) {
  RR_CHECK_ARG(out_ms != nullptr, "null pointer");
  if (!g_tc_timed) {
    set_error("rr_tc_last_timing_ms: no timed tensor-core call (rr_tc_timing(1) first)");
    return RR_ERR_INVALID;
  }
  RR_CUDA(cudaEventSynchronize(g_tc_ev[4]));
  for (int i = 0; i < 4; ++i) RR_CUDA(cudaEventElapsedTime(&out_ms[i], g_tc_ev[i], g_tc_ev[i + 1]));
  return RR_OK;
}

extern "C" size_t rr_tc_search_workspace_bytes(int64_t n, int32_t q, int32_t k) {
  if (n <= 0 || q <= 0 || k <= 0) return 256;
  return tc_plan(n, q, k).total;
}

static int tc_check(int64_t n, int32_t dim, int32_t q, int32_t k) {
  RR_CHECK_ARG(n >= 1 && q >= 1, "bad size");
  RR_CHECK_ARG(dim % TC_BK == 0 && dim >= TC_BK && dim <= TC_BK * TC_MAX_KB,
               "tensor-core path needs dim to be a multiple of 128 in [128, 1024]");
  RR_CHECK_ARG(k >= 1 && k <= RR_MAX_K, "k out of range");
  RR_CHECK_ARG(n < (1LL << 31), "shard larger than 2^31 rows");
  return RR_OK;
}

extern "C" int rr_hamming_topk_tc(const uint32_t* codes, int64_t n, int32_t words, const uint8_t* tags,
                                  uint8_t tag_mask, uint8_t tag_value, const int8_t* q_pm1, int32_t q,
                                  int32_t k, int64_t row_base, int32_t* out_dist, int64_t* out_idx,
                                  uint32_t* overflow, uint8_t* overflow_flags, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  RR_CHECK_ARG(words >= 4 && words <= RR_MAX_WORDS && words % 4 == 0, "words must be a multiple of 4 in [4, 32]");
  const int dim = words * 32;  // padded width; padding bits are 0 in rows and queries alike
  int rc = tc_check(n, dim, q, k);
  if (rc != RR_OK) return rc;
  RR_CHECK_ARG(codes && q_pm1 && out_dist && out_idx && overflow, "null pointer");
  return tc_search(nullptr, (const uint8_t*)codes, n, dim, tags, tag_mask, tag_value, q_pm1, q, k, row_base, 1,
                   out_dist, (long long*)out_idx, overflow, overflow_flags, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

extern "C" int rr_int8_search_topk_tc(const int8_t* emb, int64_t n, int32_t dim, const uint8_t* tags,
                                      uint8_t tag_mask, uint8_t tag_value, const int8_t* queries_i8, int32_t q,
                                      int32_t top_k, int64_t row_base, int32_t* out_score, int64_t* out_idx,
                                      uint32_t* overflow, uint8_t* overflow_flags, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  int rc = tc_check(n, dim, q, top_k);
  if (rc != RR_OK) return rc;
  RR_CHECK_ARG(emb && queries_i8 && out_score && out_idx && overflow, "null pointer");
  return tc_search(emb, nullptr, n, dim, tags, tag_mask, tag_value, queries_i8, q, top_k, row_base, 0, out_score,
                   (long long*)out_idx, overflow, overflow_flags, workspace, workspace_bytes, (cudaStream_t)stream);
}

// Debug / test entry: raw tensor-core scores of every row as order-preserving keys
// (key = ~(score ^ 0x80000000); 0xFFFFFFFF marks padded or filtered rows), u32 [q][ld] with
// ld = ceil(n / 128) * 128.  Lets the parity tests check the MMA path in isolation.
extern "C" int rr_tc_dense_keys(const int8_t* emb, int64_t n, int32_t dim, const int8_t* queries_i8, int32_t q,
                                uint32_t* out_keys, void* stream) {
  int rc = tc_check(n, dim, q, 1);
  if (rc != RR_OK) return rc;
  RR_CHECK_ARG(emb && queries_i8 && out_keys, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap map_a, map_b;
  rc = make_map(&map_a, emb, n, dim);
  if (rc != RR_OK) return rc;
  rc = make_map(&map_b, queries_i8, q, dim);
  if (rc != RR_OK) return rc;
  const int kb = dim / TC_BK;
  const TcSmem lay = tc_smem_layout(kb, false);
  const size_t smem = lay.bytes;
  RR_CUDA(cudaFuncSetAttribute(tc_i8_search_kernel<EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  const int qblocks = (q + TC_BN - 1) / TC_BN;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int ctas_x = sms / qblocks;
  if (ctas_x < 1) ctas_x = 1;
  TcArgs a;
  a.n = n;
  a.q = q;
  a.kb = kb;
  a.tile_stride = 1;
  a.n_tiles = (n + TC_BM - 1) / TC_BM;
  a.tags = nullptr;
  a.tag_mask = a.tag_value = 0;
  a.tau = nullptr;
  a.dense = 1;
  a.dense_keys = out_keys;
  a.cnt = nullptr;
  a.list_score = nullptr;
  a.list_row = nullptr;
  a.cap_cta = 0;
  a.stages = lay.stages;
  a.packed = 0;
  a.packed_codes = nullptr;
  a.debug = 0;
  dim3 grid((unsigned)(a.n_tiles < ctas_x ? a.n_tiles : ctas_x), qblocks);
  tc_i8_search_kernel<EPI_DENSE><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, a);
  RR_LAUNCH_CHECK();
  return RR_OK;
}
