// Probe: accumulate the COARSE level's grid gradient (40x20x40 voxels x 4 floats = 512 KB) in the distributed shared
// memory of a 4-CTA cluster (128 KB slab per CTA, red.shared::cluster.add.f32 on the owner's slab) and flush each slab
// once with red.global.add.v4.f32 -- against 8 x red.global.add.v4.f32 per point straight to L2 (the shipped scatter).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dsmem_scatter_probe dsmem_scatter_probe.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Dim { int X, Y, Z; };
constexpr int kCluster = 4;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ float u01(uint32_t h) { return (h >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ void cell_of(int n, Dim d, int& ix, int& iy, int& iz) {
  ix = min(d.X - 2, (int)(u01(hash32(3u * n + 1)) * (d.X - 1)));
  iy = min(d.Y - 2, (int)(u01(hash32(3u * n + 2)) * (d.Y - 1)));
  iz = min(d.Z - 2, (int)(u01(hash32(3u * n + 3)) * (d.Z - 1)));
}
__device__ __forceinline__ void red_add_f4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(1024) k_red(float* grid, Dim d, int N) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    int ix, iy, iz;
    cell_of(n, d, ix, iy, iz);
    const float v = 1.0f + (n & 7);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int x = ix + (k & 1), y = iy + ((k >> 1) & 1), z = iz + (k >> 2);
      red_add_f4(grid + (((size_t)z * d.Y + y) * d.X + x) * 4, v * (k + 1), v, -v, 0.5f * v);
    }
  }
}

// kRemote: slabs spread over the cluster (DSMEM); !kRemote: every CTA keeps a private slab of the first 1/4 of the grid
// and folds the voxel index into it (plain ATOMS rate, no remote traffic; result not comparable)
template <bool kRemote>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(1024) k_dsmem(float* grid, Dim d, int N) {
  extern __shared__ __align__(16) float slab[];
  cg::cluster_group cluster = cg::this_cluster();
  const int vox = d.X * d.Y * d.Z, per = (vox + kCluster - 1) / kCluster;
  const unsigned rank = cluster.block_rank();
  for (int i = threadIdx.x; i < per * 4; i += blockDim.x) slab[i] = 0.f;
  cluster.sync();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(slab);
  const int cid = blockIdx.x / kCluster, nclusters = gridDim.x / kCluster;
  // the cluster walks its share of the points with all 4 CTAs
  for (int n = (cid * kCluster + rank) * blockDim.x + threadIdx.x; n < N; n += nclusters * kCluster * blockDim.x) {
    int ix, iy, iz;
    cell_of(n, d, ix, iy, iz);
    const float v = 1.0f + (n & 7);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int x = ix + (k & 1), y = iy + ((k >> 1) & 1), z = iz + (k >> 2);
      const int vi = (z * d.Y + y) * d.X + x;
      const int owner = vi / per, local = vi - owner * per;
      uint32_t addr = base + local * 16;
      if (kRemote) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(addr), "r"(owner));
      const float vals[4] = {v * (k + 1), v, -v, 0.5f * v};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (kRemote) asm volatile("red.shared::cluster.add.f32 [%0], %1;" ::"r"(addr + 4 * c), "f"(vals[c]) : "memory");
        else asm volatile("red.shared::cta.add.f32 [%0], %1;" ::"r"(addr + 4 * c), "f"(vals[c]) : "memory");
      }
    }
  }
  cluster.sync();
  if (kRemote) {
    const int lo = rank * per, cnt = min(per, vox - lo);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float4 a = reinterpret_cast<const float4*>(slab)[i];
      if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_f4(grid + (size_t)(lo + i) * 4, a.x, a.y, a.z, a.w);
    }
  }
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const int N = 1 << 20;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Dim d = {40, 20, 40};
  const int vox = d.X * d.Y * d.Z, per = (vox + kCluster - 1) / kCluster;
  const size_t smem = (size_t)per * 16, n = (size_t)vox * 4;
  CK(cudaFuncSetAttribute(k_dsmem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_dsmem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float *g0, *g1;
  CK(cudaMalloc(&g0, n * 4)); CK(cudaMalloc(&g1, n * 4));
  CK(cudaMemset(g0, 0, n * 4)); CK(cudaMemset(g1, 0, n * 4));
  const int nb = (sms / kCluster) * kCluster;
  k_red<<<sms, 1024>>>(g0, d, N);
  k_dsmem<true><<<nb, 1024, smem>>>(g1, d, N);
  CK(cudaDeviceSynchronize());
  std::vector<float> h0(n), h1(n);
  CK(cudaMemcpy(h0.data(), g0, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h1.data(), g1, n * 4, cudaMemcpyDeviceToHost));
  double maxrel = 0, sum = 0;
  for (size_t i = 0; i < n; ++i) { maxrel = fmax(maxrel, fabs((double)h0[i] - h1[i]) / (1.0 + fabs(h0[i]))); sum += fabs(h0[i]); }
  const float t_red = time_ms([&] { k_red<<<sms, 1024>>>(g0, d, N); }, 20);
  const float t_ds = time_ms([&] { k_dsmem<true><<<nb, 1024, smem>>>(g1, d, N); }, 20);
  const float t_local = time_ms([&] { k_dsmem<false><<<nb, 1024, smem>>>(g1, d, N); }, 20);
  // flush + zero alone: N = 0 points
  const float t_flush = time_ms([&] { k_dsmem<true><<<nb, 1024, smem>>>(g1, d, 0); }, 20);
  printf("{\"sms\": %d, \"N\": %d, \"grid\": [%d,%d,%d], \"cluster\": %d, \"slab_bytes\": %zu, \"max_rel_diff\": %.3g, \"sum_abs\": %.6g, "
         "\"ms_red_global_v4\": %.4f, \"ms_dsmem_cluster\": %.4f, \"ms_smem_local_only\": %.4f, \"ms_dsmem_zero_and_flush_only\": %.4f}\n",
         sms, N, d.X, d.Y, d.Z, kCluster, smem, maxrel, sum, t_red, t_ds, t_local, t_flush);
  return 0;
}
