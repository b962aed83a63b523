// Shared device helpers for the MISO B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/miso_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "miso_b200 kernels are written for sm_100a (B200) only"
#endif

namespace miso {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define MISO_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      miso::set_error(__VA_ARGS__);      \
      return MISO_ERR_INVALID_ARG;       \
    }                                    \
  } while (0)

constexpr int kThreads = 256;

// ---- 128-bit helpers -------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Vector reduction to global memory: one 16-byte red instead of four scalar atomics (sm_90+).
__device__ __forceinline__ void red_add_f4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// predicated form: no branch around the reduction (keeps the surrounding code convergent)
__device__ __forceinline__ void red_add_f4_if(unsigned pred, float* addr, float a, float b, float c, float d) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void red_add_f2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) { atomicAdd(addr, a); }
__device__ __forceinline__ void red_add(double* addr, double a) { atomicAdd(addr, a); }

// ---- coordinates -----------------------------------------------------------------------------
// normalize_coordinates (grid_opt/utils/utils.py:49): 2*(x-bmin)/(bmax-bmin) - 1, evaluated with
// the same operation order and without FMA contraction so floor() decisions match the oracle.
__device__ __forceinline__ float normalize_coord(float x, float bmin, float bmax) {
  float len = __fsub_rn(bmax, bmin);
  float t = __fmul_rn(2.0f, __fsub_rn(x, bmin));
  return __fsub_rn(__fdiv_rn(t, len), 1.0f);
}
// ATen grid_sampler_unnormalize, align_corners=False: ((c+1)*size-1)/2 (GridSampler.h), no contraction.
__device__ __forceinline__ float unnormalize_nc(float c, int size) {
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.0f), (float)size), 1.0f), 0.5f);
}

// Per-level trilinear cell: base corner, fractions, validity of the 8 corners.
struct Cell {
  int ix0, iy0, iz0;
  float fx, fy, fz;  // ix - floor(ix)
  unsigned valid;    // bit c (c = dx + 2dy + 4dz) set when corner inside the grid (zeros padding)
  long long base;    // element offset of corner (ix0,iy0,iz0)
};

__device__ __forceinline__ Cell make_cell(float ix, float iy, float iz, const miso_level_t& lv) {
  Cell c;
  float flx = floorf(ix), fly = floorf(iy), flz = floorf(iz);
  // clamp before int conversion so far-out-of-range coordinates cannot overflow
  c.ix0 = (int)fminf(fmaxf(flx, -2.0f), (float)lv.X + 1.0f);
  c.iy0 = (int)fminf(fmaxf(fly, -2.0f), (float)lv.Y + 1.0f);
  c.iz0 = (int)fminf(fmaxf(flz, -2.0f), (float)lv.Z + 1.0f);
  c.fx = ix - flx;
  c.fy = iy - fly;
  c.fz = iz - flz;
  unsigned vx = (c.ix0 >= 0 && c.ix0 < lv.X ? 1u : 0u) | (c.ix0 + 1 >= 0 && c.ix0 + 1 < lv.X ? 2u : 0u);
  unsigned vy = (c.iy0 >= 0 && c.iy0 < lv.Y ? 1u : 0u) | (c.iy0 + 1 >= 0 && c.iy0 + 1 < lv.Y ? 2u : 0u);
  unsigned vz = (c.iz0 >= 0 && c.iz0 < lv.Z ? 1u : 0u) | (c.iz0 + 1 >= 0 && c.iz0 + 1 < lv.Z ? 2u : 0u);
  unsigned v = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    unsigned ok = ((vx >> (k & 1)) & 1u) & ((vy >> ((k >> 1) & 1)) & 1u) & ((vz >> (k >> 2)) & 1u);
    v |= ok << k;
  }
  c.valid = v;
  c.base = (long long)c.iz0 * lv.sZ + (long long)c.iy0 * lv.sY + (long long)c.ix0 * lv.sX;
  return c;
}

// corner weights in the reference's evaluation order (gridsample_cuda.cu:336-343): (wx*wy)*wz
__device__ __forceinline__ void axis_w(const Cell& c, int k, float& wx, float& wy, float& wz) {
  wx = (k & 1) ? c.fx : 1.0f - c.fx;
  wy = (k & 2) ? c.fy : 1.0f - c.fy;
  wz = (k & 4) ? c.fz : 1.0f - c.fz;
}
__device__ __forceinline__ long long corner_off(const miso_level_t& lv, const Cell& c, int k) {
  return c.base + ((k & 1) ? lv.sX : 0) + ((k & 2) ? lv.sY : 0) + ((k & 4) ? lv.sZ : 0);
}

// block-wide sum of `v` (blockDim.x == kThreads); result valid in thread 0
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  T r = 0;
  if (w == 0) {
    r = (l < (int)(blockDim.x >> 5)) ? smem[l] : (T)0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;
}

inline int grid_for(int64_t n, int threads, int max_blocks) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

int sm_count();

}  // namespace miso
