// Probe: can the TMA engine (cp.reduce.async.bulk.tensor ... .add) take the grid-gradient scatter off the LSU pipe?
// One trilinear cell of a channels-last (Z,Y,X,C=4) fp32 grid is a 2x2x2 box of 16-byte voxels; as a 3-D tensor
// {X*4, Y, Z} with box {8,2,2} it is ONE TMA tile, out-of-range corners are clipped by the hardware (== zeros padding).
// Compared against 8 x red.global.add.v4.f32 per point (the round-1 scatter) on the same cells.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_reduce_probe tma_reduce_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Dim { int X, Y, Z; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ float u01(uint32_t h) { return (h >> 8) * (1.0f / 16777216.0f); }

// mode 0: uniform random cells; mode 1: ray-ordered (27 consecutive samples walk along a line through the grid)
__device__ __forceinline__ void cell_of(int n, int mode, Dim d, int& ix, int& iy, int& iz) {
  if (mode == 0) {
    ix = (int)(u01(hash32(3u * n + 1)) * (d.X + 1)) - 1;
    iy = (int)(u01(hash32(3u * n + 2)) * (d.Y + 1)) - 1;
    iz = (int)(u01(hash32(3u * n + 3)) * (d.Z + 1)) - 1;
  } else {
    const int ray = n / 27, k = n % 27;
    const float ox = u01(hash32(7u * ray + 1)) * d.X, oy = u01(hash32(7u * ray + 2)) * d.Y, oz = u01(hash32(7u * ray + 3)) * d.Z;
    float dx = u01(hash32(7u * ray + 4)) - 0.5f, dy = (u01(hash32(7u * ray + 5)) - 0.5f) * 0.3f, dz = u01(hash32(7u * ray + 6)) - 0.5f;
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz + 1e-6f);
    const float len = 0.25f * d.X;   // ray length: a quarter of the grid's x extent
    const float s = (k / 26.0f) * len * inv;
    ix = (int)floorf(ox + dx * s) ; iy = (int)floorf(oy + dy * s); iz = (int)floorf(oz + dz * s);
    ix = max(-1, min(d.X - 1, ix)); iy = max(-1, min(d.Y - 1, iy)); iz = max(-1, min(d.Z - 1, iz));
  }
}

__device__ __forceinline__ void red_add_f4_if(bool ok, float* addr, float a, float b, float c, float d) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
               ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"((unsigned)ok) : "memory");
}

__global__ void __launch_bounds__(256) k_red(float* grid, Dim d, int N, int mode, int levels_red, int clamp_in = 0) {
  for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256) {
    int ix, iy, iz;
    cell_of(n, mode, d, ix, iy, iz);
    if (clamp_in) { ix = max(0, min(d.X - 2, ix)); iy = max(0, min(d.Y - 2, iy)); iz = max(0, min(d.Z - 2, iz)); }
    const float v = 1.0f + (n & 7);
    for (int rep = 0; rep < levels_red; ++rep) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = ix + (k & 1), y = iy + ((k >> 1) & 1), z = iz + (k >> 2);
        const bool ok = x >= 0 && x < d.X && y >= 0 && y < d.Y && z >= 0 && z < d.Z;
        float* p = grid + (ok ? (((size_t)z * d.Y + y) * d.X + x) * 4 : 0);
        red_add_f4_if(ok, p, v * (k + 1), v, -v, 0.5f * v);
      }
    }
  }
}

template <int NSLOT>
__global__ void __launch_bounds__(256) k_tma(const __grid_constant__ CUtensorMap tmap, float* grid, Dim d, int N, int mode,
                                             int levels_tma, int levels_red) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* slots = reinterpret_cast<float4*>(smem) + (size_t)threadIdx.x * NSLOT * 8;
  int it = 0;
  for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256) {
    int ix, iy, iz;
    cell_of(n, mode, d, ix, iy, iz);
    const float v = 1.0f + (n & 7);
    for (int rep = 0; rep < levels_tma; ++rep, ++it) {
      float4* s = slots + (it % NSLOT) * 8;
      if (it >= NSLOT) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = make_float4(v * (k + 1), v, -v, 0.5f * v);   // box order: z, y, then x*4+c
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s);
      asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                   ::"l"(&tmap), "r"(ix * 4), "r"(iy), "r"(iz), "r"(saddr) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    for (int rep = 0; rep < levels_red; ++rep) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = ix + (k & 1), y = iy + ((k >> 1) & 1), z = iz + (k >> 2);
        const bool ok = x >= 0 && x < d.X && y >= 0 && y < d.Y && z >= 0 && z < d.Z;
        float* p = grid + (ok ? (((size_t)z * d.Y + y) * d.X + x) * 4 : 0);
        red_add_f4_if(ok, p, v * (k + 1), v, -v, 0.5f * v);
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// non-tensor bulk reduce: 4 x (32-byte x-pair) per cell, in-range cells only (no clipping available)
template <int NSLOT>
__global__ void __launch_bounds__(256) k_bulk(float* grid, Dim d, int N, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* slots = reinterpret_cast<float4*>(smem) + (size_t)threadIdx.x * NSLOT * 8;
  int it = 0;
  for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256, ++it) {
    int ix, iy, iz;
    cell_of(n, mode, d, ix, iy, iz);
    ix = max(0, min(d.X - 2, ix)); iy = max(0, min(d.Y - 2, iy)); iz = max(0, min(d.Z - 2, iz));
    const float v = 1.0f + (n & 7);
    float4* s = slots + (it % NSLOT) * 8;
    if (it >= NSLOT) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = make_float4(v * (k + 1), v, -v, 0.5f * v);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
      float* p = grid + (((size_t)(iz + (yz >> 1)) * d.Y + iy + (yz & 1)) * d.X + ix) * 4;
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s + 2 * yz);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 32;" ::"l"(p), "r"(saddr) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(float* base, Dim d) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)d.X * 4, (cuuint64_t)d.Y, (cuuint64_t)d.Z};
  cuuint64_t strides[2] = {(cuuint64_t)d.X * 16, (cuuint64_t)d.X * d.Y * 16};
  cuuint32_t box[3] = {8, 2, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

// elected-lane variant: one lane per warp issues the TMA ops of all 32 lanes (explicit uniform issue)
template <int NSLOT>
__global__ void __launch_bounds__(256) k_tma_elect(const __grid_constant__ CUtensorMap tmap, Dim d, int N, int mode, int clamp_in,
                                                   float* red_grid = nullptr, int tma_levels = 1) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* slots = reinterpret_cast<float4*>(smem) + (size_t)threadIdx.x * NSLOT * 8;
  const int lane = threadIdx.x & 31;
  int it = 0;
  for (int n0 = blockIdx.x * 256 + (threadIdx.x & ~31); n0 < N; n0 += gridDim.x * 256, ++it) {
    const int n = n0 + lane;
    int ix, iy, iz;
    cell_of(n, mode, d, ix, iy, iz);
    if (clamp_in) { ix = max(0, min(d.X - 2, ix)); iy = max(0, min(d.Y - 2, iy)); iz = max(0, min(d.Z - 2, iz)); }
    const float v = 1.0f + (n & 7);
    float4* s = slots + (it % NSLOT) * 8;
    if (it >= NSLOT && lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = make_float4(v * (k + 1), v, -v, 0.5f * v);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s);
    for (int l = 0; l < 32; ++l) {
      const int cx = __shfl_sync(0xffffffffu, ix, l), cy = __shfl_sync(0xffffffffu, iy, l), cz = __shfl_sync(0xffffffffu, iz, l);
      const uint32_t sa = __shfl_sync(0xffffffffu, saddr, l);
      if (lane == 0 && n0 + l < N)
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                     ::"l"(&tmap), "r"(cx * 4), "r"(cy), "r"(cz), "r"(sa) : "memory");
    }
    if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (red_grid != nullptr && n < N) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = ix + (k & 1), y = iy + ((k >> 1) & 1), z = iz + (k >> 2);
        const bool ok = x >= 0 && x < d.X && y >= 0 && y < d.Y && z >= 0 && z < d.Z;
        float* p = red_grid + (ok ? (((size_t)z * d.Y + y) * d.X + x) * 4 : 0);
        red_add_f4_if(ok, p, v * (k + 1), v, -v, 0.5f * v);
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void k_one(const __grid_constant__ CUtensorMap tmap, int ix, int iy, int iz) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* s = reinterpret_cast<float4*>(smem);
  if (threadIdx.x < 8) s[threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(&tmap), "r"(ix * 4), "r"(iy), "r"(iz), "r"(sa) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

static void report(const char* what, cudaError_t e) { printf("{\"step\": \"%s\", \"status\": \"%s\"}\n", what, cudaGetErrorString(e)); fflush(stdout); }

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "all";
  const int N = 1 << 20;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  constexpr int NS = 2;
  const size_t smem = 256 * NS * 128;
  CK(cudaFuncSetAttribute(k_tma<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_tma_elect<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_bulk<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (strcmp(which, "all") != 0) {
    // diagnostics: one kernel per process so a sticky error names its culprit
    Dim d = {200, 100, 200};
    const size_t n = (size_t)d.X * d.Y * d.Z * 4;
    float* g;
    CK(cudaMalloc(&g, n * 4)); CK(cudaMemset(g, 0, n * 4));
    CUtensorMap tm = make_map(g, d);
    if (!strcmp(which, "red")) k_red<<<sms * 4, 256>>>(g, d, N, 0, 1);
    else if (!strcmp(which, "tma")) k_tma<NS><<<sms, 256, smem>>>(tm, g, d, N, 0, 1, 0);
    else if (!strcmp(which, "tma_elect")) k_tma_elect<NS><<<sms, 256, smem>>>(tm, d, N, 0, 0);
    else if (!strcmp(which, "tma_elect_in")) k_tma_elect<NS><<<sms, 256, smem>>>(tm, d, N, 0, 1);
    else if (!strcmp(which, "bulk")) k_bulk<NS><<<sms, 256, smem>>>(g, d, N, 0);
    else if (!strncmp(which, "oob", 3)) {
      // one op with a hand-picked coordinate: oob_x-  oob_x+  oob_y-  oob_y+  oob_z-  oob_z+
      int c[3] = {5, 5, 5};
      const int ax = which[4] - 'x';
      const int ext[3] = {d.X, d.Y, d.Z};
      c[ax] = which[5] == '-' ? -1 : ext[ax] - 1;
      k_one<<<1, 32, 128>>>(tm, c[0], c[1], c[2]);
    }
    report(which, cudaDeviceSynchronize());
    return 0;
  }
  printf("{\"sms\": %d, \"N\": %d, \"results\": [\n", sms, N);
  Dim dims[2] = {{200, 100, 200}, {40, 20, 40}};
  bool first = true;
  for (int gi = 0; gi < 2; ++gi) {
    Dim d = dims[gi];
    const size_t n = (size_t)d.X * d.Y * d.Z * 4;
    float *g0, *g1;
    CK(cudaMalloc(&g0, n * 4)); CK(cudaMalloc(&g1, n * 4));
    CUtensorMap tm = make_map(g1, d);
    for (int mode = 0; mode < 2; ++mode) {
      // correctness: one pass each, compare
      CK(cudaMemset(g0, 0, n * 4)); CK(cudaMemset(g1, 0, n * 4));
      k_red<<<sms * 4, 256>>>(g0, d, N, mode, 1, 1);
      k_tma_elect<NS><<<sms * 3, 256, smem>>>(tm, d, N, mode, 1);
      CK(cudaDeviceSynchronize());
      std::vector<float> h0(n), h1(n);
      CK(cudaMemcpy(h0.data(), g0, n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(h1.data(), g1, n * 4, cudaMemcpyDeviceToHost));
      double maxd = 0, sum = 0;
      for (size_t i = 0; i < n; ++i) { maxd = fmax(maxd, fabs((double)h0[i] - h1[i])); sum += fabs(h0[i]); }
      for (int blocks_per_sm = 1; blocks_per_sm <= 3; blocks_per_sm += 2) {
        const int nb = sms * blocks_per_sm;
        float t_red1 = time_ms([&] { k_red<<<sms * 4, 256>>>(g0, d, N, mode, 1, 1); }, 20);
        float t_red2 = time_ms([&] { k_red<<<sms * 4, 256>>>(g0, d, N, mode, 2, 1); }, 20);
        float t_tma1 = time_ms([&] { k_tma_elect<NS><<<nb, 256, smem>>>(tm, d, N, mode, 1); }, 20);
        float t_mix = time_ms([&] { k_tma_elect<NS><<<nb, 256, smem>>>(tm, d, N, mode, 1, g0); }, 20);
        float t_bulk = time_ms([&] { k_bulk<NS><<<nb, 256, smem>>>(g1, d, N, mode); }, 20);
        printf("%s{\"grid\": [%d,%d,%d], \"mode\": \"%s\", \"ctas_per_sm\": %d, \"max_abs_diff\": %.3g, \"sum_abs\": %.6g, "
               "\"ms_red_1level\": %.4f, \"ms_red_2level\": %.4f, \"ms_tma_elect_1level\": %.4f, \"ms_mixed_1tma_1red\": %.4f, \"ms_bulk32x4_1level\": %.4f}",
               first ? "" : ",\n", d.X, d.Y, d.Z, mode == 0 ? "uniform" : "rays", blocks_per_sm, maxd, sum, t_red1, t_red2, t_tma1, t_mix, t_bulk);
        first = false;
      }
    }
    CK(cudaFree(g0)); CK(cudaFree(g1));
  }
  printf("\n]}\n");
  return 0;
}
