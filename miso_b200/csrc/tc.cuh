// tcgen05 / TMEM helpers for the 64x64 decoder layers (sm_100a).
//
// Shape: one tile = 128 points (UMMA M = 128, one TMEM lane per point), N = 64 hidden units,
// K = 64 in steps of 8 (kind::tf32).  A (activations) is written into TMEM by the thread that owns
// the point (tcgen05.st, 32x32b: thread i of warp w <-> lane 32*(w%4)+i) and consumed straight from
// TMEM (".ts" form, A K-major); B (weights) lives in shared memory in the canonical K-major
// no-swizzle layout (8-row x 16-byte core matrices); D accumulates in TMEM in fp32.
//
// FP32-class accuracy comes from the 3xTF32 split: a = a_hi + a_lo with a_hi = rn_tf32(a) and
// a_lo = rn_tf32(a - a_hi), so  D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  drops only the lo*lo term and the
// rounding of lo (~2^-22 relative each): well inside the 1e-5 forward tolerance, which plain TF32/BF16
// (2^-11 / 2^-8) would miss by orders of magnitude.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace miso {
namespace tc {

constexpr int kTileM = 128;
constexpr int kN = 64;
constexpr int kK = 64;
constexpr uint32_t kLBO = 128;           // bytes between core matrices adjacent in K
constexpr uint32_t kSBO = (kK / 4) * 128;  // bytes between 8-row groups in N (16 core matrices per group)
constexpr int kWeightBytes = kN * kK * 4;  // one canonical 64x64 fp32 operand = 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element B[n][k] inside a canonical K-major no-swizzle operand
__device__ __forceinline__ uint32_t b_offset(int n, int k) {
  return (uint32_t)((n >> 3) * kSBO + (k >> 2) * kLBO + (n & 7) * 16 + (k & 3) * 4);
}

// round-to-nearest tf32 (10 explicit mantissa bits); the low 13 bits of the result are zero
__device__ __forceinline__ float tf32_rn(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return __uint_as_float(r);
}
// a = hi + lo (+ ~2^-22 |a|): hi = rn_tf32(a), lo = rn_tf32(a - hi)  (a - hi is exact in fp32)
__device__ __forceinline__ void tf32_split(float a, float& hi, float& lo) {
  hi = tf32_rn(a);
  lo = tf32_rn(a - hi);
}
// 2-instruction split for the per-point activations: hi = truncate(a) (LOP3), lo = a - hi (exact; the
// tensor core ignores its low 13 bits).  |error| <= 2^-21 |a|; weights keep the round-to-nearest split.
__device__ __forceinline__ void tf32_split_fast(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
  lo = a - hi;
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address, LBO, SBO in
// 16-byte units, version 1, SWIZZLE_NONE
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((kLBO >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((kSBO >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// same, for an operand whose rows hold `k_elems` K-elements (SBO = k_elems/4 core matrices)
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr, int k_elems) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((kLBO >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((((uint32_t)k_elems / 4) * kLBO >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t b_offset_k(int n, int k, int k_elems) {
  return (uint32_t)((n >> 3) * (k_elems / 4) * kLBO + (k >> 2) * kLBO + (n & 7) * 16 + (k & 3) * 4);
}

// instruction descriptor (InstrDescriptor): D=f32, A=B=tf32, both K-major, N=64, M=128
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// D[tmem_d] (+)= A[tmem_a] * B[desc]   (one K=8 step); issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem_d] (+)= A[smem desc] * B[smem desc]   (one K=8 step); issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Full 3xTF32 product of one tile: D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, 24 MMAs, one issuing thread.
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo,
                                                  uint32_t smem_b_hi, uint32_t smem_b_lo) {
  const uint32_t idesc = make_idesc();
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = (pass == 1) ? tmem_a_lo : tmem_a_hi;
    const uint32_t b = (pass == 2) ? smem_b_lo : smem_b_hi;
#pragma unroll
    for (int ks = 0; ks < kK / 8; ++ks)
      mma_tf32_ts(tmem_d, a + ks * 8, make_b_desc(b + ks * 2 * kLBO), idesc, (pass | ks) ? 1u : 0u);
  }
}

// 3xTF32 product with A_hi in TMEM and A_lo in shared memory (canonical K-major layout, 128 rows):
// halves the TMEM footprint of a tile (A_hi 64 + D 64 columns) so four tiles fit in the 512 columns.
__device__ __forceinline__ void issue_gemm_3xtf32_mixed(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t smem_a_lo,
                                                        uint32_t smem_b_hi, uint32_t smem_b_lo) {
  const uint32_t idesc = make_idesc();
#pragma unroll
  for (int ks = 0; ks < kK / 8; ++ks)
    mma_tf32_ts(tmem_d, tmem_a_hi + ks * 8, make_b_desc(smem_b_hi + ks * 2 * kLBO), idesc, ks ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < kK / 8; ++ks)
    mma_tf32_ss(tmem_d, make_b_desc(smem_a_lo + ks * 2 * kLBO), make_b_desc(smem_b_hi + ks * 2 * kLBO), idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < kK / 8; ++ks)
    mma_tf32_ts(tmem_d, tmem_a_hi + ks * 8, make_b_desc(smem_b_lo + ks * 2 * kLBO), idesc, 1u);
}

// D = A * (B_hi + B_lo) for an A that is exactly representable in tf32 (0/1 masks): 16 MMAs, A in TMEM only
__device__ __forceinline__ void issue_gemm_exactA(uint32_t tmem_d, uint32_t tmem_a, uint32_t smem_b_hi,
                                                  uint32_t smem_b_lo) {
  const uint32_t idesc = make_idesc();
#pragma unroll
  for (int ks = 0; ks < kK / 8; ++ks)
    mma_tf32_ts(tmem_d, tmem_a + ks * 8, make_b_desc(smem_b_hi + ks * 2 * kLBO), idesc, ks ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < kK / 8; ++ks)
    mma_tf32_ts(tmem_d, tmem_a + ks * 8, make_b_desc(smem_b_lo + ks * 2 * kLBO), idesc, 1u);
}

// byte offset of row m's first 16-byte chunk inside a canonical K-major A operand (128 rows x 64)
__device__ __forceinline__ uint32_t a_row_offset(int m) { return (uint32_t)((m >> 3) * kSBO + (m & 7) * 16); }

// 3xTF32 product with a short K (the first decoder layer: K = F inputs + 1 bias column, padded to 8):
// A_hi / A_lo both in TMEM, B = canonical [64][KP] operands
template <int KP>
__device__ __forceinline__ void issue_gemm_3xtf32_smallk(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo,
                                                         uint32_t smem_b_hi, uint32_t smem_b_lo) {
  const uint32_t idesc = make_idesc();
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = (pass == 1) ? tmem_a_lo : tmem_a_hi;
    const uint32_t b = (pass == 2) ? smem_b_lo : smem_b_hi;
#pragma unroll
    for (int ks = 0; ks < KP / 8; ++ks)
      mma_tf32_ts(tmem_d, a + ks * 8, make_desc_k(b + ks * 2 * kLBO, KP), idesc, (pass | ks) ? 1u : 0u);
  }
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ---- TMEM <-> registers, 32 lanes x 32 columns per call (this warp's lane quarter) ---------------
#define MISO_R32(v)                                                                                              \
  v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15], v[16], \
      v[17], v[18], v[19], v[20], v[21], v[22], v[23], v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// stage a row-major [n][k] 64x64 fp32 matrix (global) into canonical hi / lo operands in shared memory.
// transpose=true stages its transpose (B[n][k] = W[k][n]).
__device__ __forceinline__ void stage_weights(const float* __restrict__ W, bool transpose, unsigned char* b_hi,
                                              unsigned char* b_lo, int tid, int nthreads) {
  for (int i = tid; i < kN * kK; i += nthreads) {
    const int n = i / kK, k = i % kK;
    const float w = transpose ? W[k * kN + n] : W[n * kK + k];
    float hi, lo;
    tf32_split(w, hi, lo);
    const uint32_t off = b_offset(n, k);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
}

}  // namespace tc
}  // namespace miso
