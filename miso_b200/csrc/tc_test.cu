// Self-test of the tensor-core building block used by the fused decoder: D = A * W^T (or A * W) for a
// 64x64 fp32 matrix with the 3xTF32 split, A fed from TMEM, one 128-row tile per CTA.
#include "common.cuh"
#include "tc.cuh"

namespace miso {

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                          int transpose, float* __restrict__ D, int64_t M) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* b_hi = smem;
  unsigned char* b_lo = smem + tc::kWeightBytes;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  tc::stage_weights(W, transpose != 0, b_hi, b_lo, tid, blockDim.x);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t col_hi = 0, col_lo = 64, col_d = 128;
  uint32_t parity = 0;
  for (int64_t tile = blockIdx.x; tile * 128 < M; tile += gridDim.x) {
    const int64_t row = tile * 128 + tid;
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float a = row < M ? A[row * 64 + half * 32 + i] : 0.f;
        float h, l;
        tc::tf32_split(a, h, l);
        hi[i] = __float_as_uint(h);
        lo[i] = __float_as_uint(l);
      }
      tc::tmem_st32(lane_base + col_hi + half * 32, hi);
      tc::tmem_st32(lane_base + col_lo + half * 32, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::issue_gemm_3xtf32(tbase + col_d, tbase + col_hi, tbase + col_lo, tc::smem_u32(b_hi), tc::smem_u32(b_lo));
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, parity);
    parity ^= 1;
    tc::fence_after_sync();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t d[32];
      tc::tmem_ld32(lane_base + col_d + half * 32, d);
      tc::wait_ld();
      if (row < M) {
#pragma unroll
        for (int i = 0; i < 32; ++i) D[row * 64 + half * 32 + i] = __uint_as_float(d[i]);
      }
    }
    tc::fence_before_sync();
    __syncthreads();  // D and A columns are reused by the next tile
    tc::fence_after_sync();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 256);
}

}  // namespace miso

using namespace miso;

extern "C" int miso_tc_selftest(const float* A, const float* W, int32_t transpose, float* D, int64_t M,
                                miso_stream_t stream) {
  MISO_REQUIRE(A && W && D && M > 0, "tc_selftest: null argument");
  const size_t smem = 2 * tc::kWeightBytes;
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks = grid_for((M + 127) / 128, 1, sm_count());
  tc_selftest_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(A, W, transpose, D, M);
  return check_launch("tc_selftest");
}
