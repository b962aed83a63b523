// Error plumbing + the small helpers around the hot path: Adam over dense grids (fused with the
// gradient zero-fill), batched frame->world transform, Morton keys for L2-local batch ordering.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace miso {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return MISO_ERR_CUDA;
  }
  return MISO_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// Device-side step state of miso_adam_step_dev: {counter, gate, ticket}.  Every block derives the bias-correction
// scalars of step = *counter + 1 itself (two float64 pow per block); the block that finishes LAST (ticket) publishes the
// new count -- every block has read the old one by then -- so the optimizer step is ONE launch and the launch sequence
// is identical every step (CUDA-graph capturable).  A non-finite *gate (the step's total loss) skips the update like
// the reference's trainer (`if not torch.isnan(total_loss): backward(); step()`, grid_opt/trainer.py:214-217): p / m / v
// and the counter stay, only the (poisoned) gradient is cleared.
struct AdamDev {
  int32_t* counter;
  const float* gate;
  unsigned* ticket;
};

__device__ __forceinline__ bool adam_dev_begin(const AdamDev& d, float lr, float b1, float b2, float& step_size,
                                               float& bc2_sqrt, float* sh /* 3 floats of shared memory */) {
  if (!d.counter) return false;
  if (threadIdx.x == 0) {
    const bool skip = d.gate && !isfinite(*d.gate);
    const int step = *d.counter + 1;
    sh[0] = (float)((double)lr / (1.0 - pow((double)b1, (double)step)));
    sh[1] = (float)sqrt(1.0 - pow((double)b2, (double)step));
    sh[2] = skip ? 1.f : 0.f;
  }
  __syncthreads();
  step_size = sh[0], bc2_sqrt = sh[1];
  return sh[2] != 0.f;
}

__device__ __forceinline__ void adam_dev_end(const AdamDev& d, bool skipped) {
  if (!d.counter) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(d.ticket, 1u) == gridDim.x - 1) {
      if (!skipped) *d.counter = *d.counter + 1;
      *d.ticket = 0u;
    }
  }
}

__device__ __forceinline__ void adam_clear_grad(float* g, int64_t n, int zero) {
  if (zero)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] = 0.f;
}

// torch.optim.Adam._single_tensor_adam (no amsgrad / weight decay / maximize):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(kThreads)
    adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n4,
                int64_t n, float lr, float b1, float b2, float eps, float step_size, float bc2_sqrt, int zero,
                AdamDev dev) {
  __shared__ float sh_dev[3];
  if (adam_dev_begin(dev, lr, b1, b2, step_size, bc2_sqrt, sh_dev)) {
    adam_clear_grad(g, n, zero);
    adam_dev_end(dev, true);
    return;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    // voxels no sample has ever touched (g = m = v = 0) have an exactly-zero Adam update: skip the parameter
    // read and all four stores (12 B instead of 32 B of traffic per parameter, same result bit for bit)
    if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f && gv.w == 0.f && mv.x == 0.f && mv.y == 0.f && mv.z == 0.f &&
        mv.w == 0.f && vv.x == 0.f && vv.y == 0.f && vv.z == 0.f && vv.w == 0.f)
      continue;
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w},
          va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ma[e] = ma[e] + (ga[e] - ma[e]) * (1.f - b1);  // torch: exp_avg.lerp_(grad, 1-beta1)
      va[e] = va[e] * b2 + (1.f - b2) * ga[e] * ga[e];
      float denom = sqrtf(va[e]) / bc2_sqrt + eps;
      pa[e] = pa[e] - step_size * (ma[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (zero) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // scalar tail
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i];
    float mi = m[i] + (gi - m[i]) * (1.f - b1);
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
    if (zero) g[i] = 0.f;
  }
  adam_dev_end(dev, false);
}

// Same update with an "ever touched" bitmap (one bit per 4-float voxel): a voxel whose bit is clear has
// g = m = v = 0 by construction, so when its gradient is zero again NOTHING but the 16-byte gradient (L2-resident right
// after the scatter) is read -- the plain kernel above still streams m and v (2 x 16 B per voxel) from HBM to find that
// out.  Thread i <-> voxel i; the 32 voxels of a warp share one bitmap word (broadcast load, ballot, one store).
__global__ void __launch_bounds__(kThreads)
    adam_tracked_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                        uint32_t* __restrict__ touched, int64_t n4, float lr, float b1, float b2, float eps,
                        float step_size, float bc2_sqrt, int zero, AdamDev dev) {
  __shared__ float sh_dev[3];
  if (adam_dev_begin(dev, lr, b1, b2, step_size, bc2_sqrt, sh_dev)) {
    adam_clear_grad(g, n4 * 4, zero);
    adam_dev_end(dev, true);
    return;
  }
  // A warp takes kU consecutive bitmap words = kU x 32 voxels per trip and issues all of their loads before it touches
  // any of them: with one voxel per thread per trip the sweep over a mostly-untouched level (one 16-byte gradient load
  // and nothing else per voxel) ran at ~1.2 TB/s, latency-bound; kU independent loads per thread bring it to the HBM rate.
#ifndef MISO_ADAM_U
#define MISO_ADAM_U 4
#endif
  constexpr int kU = MISO_ADAM_U;
  const int lane = threadIdx.x & 31;
  const int64_t nwords = (n4 + 31) >> 5;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w0 = warp0 * kU; w0 < nwords; w0 += nwarps * kU) {
    uint32_t word[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) word[u] = (w0 + u < nwords) ? touched[w0 + u] : 0u;
    float4 gv[kU], mv[kU], vv[kU], pv[kU];
    bool live[kU], was[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = (w0 + u) * 32 + lane;
      live[u] = (w0 + u < nwords) && i < n4;
      was[u] = (word[u] >> lane) & 1u;
      gv[u] = mv[u] = vv[u] = pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live[u]) gv[u] = reinterpret_cast<float4*>(g)[i];
      // a voxel that was touched before needs p, m, v whatever its gradient is now: issue those loads together with
      // the gradient's instead of after it (one memory round trip per trip instead of two)
      if (live[u] && was[u]) {
        mv[u] = reinterpret_cast<float4*>(m)[i];
        vv[u] = reinterpret_cast<float4*>(v)[i];
        pv[u] = reinterpret_cast<float4*>(p)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (w0 + u >= nwords) continue;          // warp-uniform
      const int64_t i = (w0 + u) * 32 + lane;
      const bool gnz = gv[u].x != 0.f || gv[u].y != 0.f || gv[u].z != 0.f || gv[u].w != 0.f;
      const uint32_t now = __ballot_sync(0xffffffffu, was[u] || gnz);
      if (lane == 0 && now != word[u]) touched[w0 + u] = now;
      if (!live[u] || !(was[u] || gnz)) continue;
      if (!was[u]) pv[u] = reinterpret_cast<float4*>(p)[i];   // first touch: m = v = 0 by construction
      float pa[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w}, ga[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w},
            ma[4] = {mv[u].x, mv[u].y, mv[u].z, mv[u].w}, va[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ma[e] = ma[e] + (ga[e] - ma[e]) * (1.f - b1);
        va[e] = va[e] * b2 + (1.f - b2) * ga[e] * ga[e];
        float denom = sqrtf(va[e]) / bc2_sqrt + eps;
        pa[e] = pa[e] - step_size * (ma[e] / denom);
      }
      reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
      reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
      reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
      if (zero && gnz) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  adam_dev_end(dev, false);
}

__global__ void __launch_bounds__(kThreads)
    transform_kernel(const float* __restrict__ x, const int64_t* __restrict__ ids, const float* __restrict__ R,
                     const float* __restrict__ t, int num_frames, int64_t N, float* __restrict__ y) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float a = x[3 * n], b = x[3 * n + 1], c = x[3 * n + 2];
    int64_t id = ids[n];
    if (id < 0 || id >= num_frames) {
      y[3 * n] = a, y[3 * n + 1] = b, y[3 * n + 2] = c;
      continue;
    }
    const float* Rm = R + id * 9;
    const float* tv = t + id * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) y[3 * n + j] = fmaf(c, Rm[3 * j + 2], fmaf(b, Rm[3 * j + 1], a * Rm[3 * j])) + tv[j];
  }
}

// Slab selection of a replicated batch (domain-decomposed multi-GPU fit, miso_b200/sharded_fit.py): every rank reads
// the whole batch and keeps the samples whose trilinear cell of `level` starts in its range of z-planes
// [z_begin, z_end) (z = the slowest axis of the channels-last level, so a range of planes is one contiguous piece of the
// level, of its gradient and of the Adam moments).  The z index is computed with the fused kernels' own arithmetic
// (frame -> world, normalize_coord, unnormalize_nc, floor), so a sample is owned by exactly one rank and every corner
// it touches lies in planes [z_begin, z_end].  Compaction keeps the order of the samples inside a 1024-sample chunk
// (consecutive ray samples stay adjacent for the merged reductions of the step kernel); the number kept lands in *count.
constexpr int kSelPer = 4;   // consecutive samples per thread: a block trip covers 1024 samples with two barriers

__global__ void __launch_bounds__(kThreads)
    slab_select_kernel(const float* __restrict__ x, const int64_t* __restrict__ ids, const float* __restrict__ R,
                       const float* __restrict__ t, int num_frames, int64_t N, float zmin, float zmax, int Z, int axis,
                       int z_begin, int z_end, const float* __restrict__ sdf, const uint8_t* __restrict__ valid,
                       const float* __restrict__ sign, const float* __restrict__ weights, float* __restrict__ x_out,
                       int64_t* __restrict__ ids_out, float* __restrict__ sdf_out, uint8_t* __restrict__ valid_out,
                       float* __restrict__ sign_out, float* __restrict__ weights_out, int32_t* __restrict__ count) {
  __shared__ int warp_tot[kThreads / 32];
  __shared__ int chunk_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int kChunk = kThreads * kSelPer;
  const int64_t chunks = (N + kChunk - 1) / kChunk;
  for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    const int64_t n0 = c * kChunk + (int64_t)threadIdx.x * kSelPer;
    float px[kSelPer][3];
    int64_t id[kSelPer];
    unsigned keep = 0;
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) {
      const int64_t n = n0 + k;
      id[k] = 0;
      px[k][0] = px[k][1] = px[k][2] = 0.f;
      if (n < N) {
        px[k][0] = x[3 * n], px[k][1] = x[3 * n + 1], px[k][2] = x[3 * n + 2];
        if (ids) id[k] = ids[n];
      }
    }
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) {
      if (n0 + k >= N) continue;
      float zw = axis == 0 ? px[k][0] : (axis == 1 ? px[k][1] : px[k][2]);   // world coordinate along the slab axis
      if (ids) {
        const int64_t q = (id[k] < 0 || id[k] >= num_frames) ? 0 : id[k];
        const float* Rr = R + q * 9 + 3 * axis;
        zw = fmaf(px[k][2], Rr[2], fmaf(px[k][1], Rr[1], px[k][0] * Rr[0])) + t[q * 3 + axis];
      }
      const float iz = unnormalize_nc(normalize_coord(zw, zmin, zmax), Z);
      // NaN (a keyframe without a pose) and far-away samples go to the edge planes: some rank must own them
      float fz = floorf(iz);
      fz = fz == fz ? fminf(fmaxf(fz, 0.f), (float)(Z - 1)) : 0.f;
      const int plane = (int)fz;
      keep |= (plane >= z_begin && plane < z_end) ? (1u << k) : 0u;
    }
    // the kept samples' targets are fetched now, so that their latency overlaps the prefix sum and the chunk's atomic
    float g_sdf[kSelPer], g_sign[kSelPer], g_w[kSelPer];
    uint8_t g_valid[kSelPer];
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) {
      g_sdf[k] = g_sign[k] = 0.f, g_w[k] = 1.f, g_valid[k] = 0;
      if ((keep >> k) & 1u) {
        const int64_t n = n0 + k;
        g_sdf[k] = sdf[n], g_valid[k] = valid[n], g_sign[k] = sign[n];
        if (weights) g_w[k] = weights[n];
      }
    }
    // exclusive prefix of the per-thread keep counts: inside the warp by shuffles, across warps through shared memory
    const int mine = __popc(keep);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int k = 0; k < kThreads / 32; ++k) {
        const int v = warp_tot[k];
        warp_tot[k] = tot;
        tot += v;
      }
      chunk_base = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    int64_t o = (int64_t)chunk_base + warp_tot[w] + (incl - mine);
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) {
      if (!((keep >> k) & 1u)) continue;
      x_out[3 * o] = px[k][0], x_out[3 * o + 1] = px[k][1], x_out[3 * o + 2] = px[k][2];
      if (ids_out) ids_out[o] = id[k];
      sdf_out[o] = g_sdf[k];
      valid_out[o] = g_valid[k];
      sign_out[o] = g_sign[k];
      if (weights_out) weights_out[o] = g_w[k];
      ++o;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(kThreads)
    morton_kernel(const float* __restrict__ x, int64_t N, float bx0, float bx1, float by0, float by1, float bz0,
                  float bz1, uint32_t* __restrict__ keys) {
  const float sx = 1024.f / (bx1 - bx0), sy = 1024.f / (by1 - by0), sz = 1024.f / (bz1 - bz0);
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float fx = fminf(fmaxf((x[3 * n] - bx0) * sx, 0.f), 1023.f);
    float fy = fminf(fmaxf((x[3 * n + 1] - by0) * sy, 0.f), 1023.f);
    float fz = fminf(fmaxf((x[3 * n + 2] - bz0) * sz, 0.f), 1023.f);
    keys[n] = spread10((uint32_t)fx) | (spread10((uint32_t)fy) << 1) | (spread10((uint32_t)fz) << 2);
  }
}

}  // namespace miso

namespace miso {
// compact host batch -> the tensors the mapping step consumes (see miso_expand_batch in the header)
__global__ void expand_batch_kernel(const int16_t* __restrict__ ids16, const float* __restrict__ sdf, float trunc,
                                    int64_t N, int64_t* __restrict__ ids64, uint8_t* __restrict__ valid,
                                    float* __restrict__ sign) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    if (ids64) ids64[n] = (int64_t)ids16[n];
    const float s = sdf[n];
    valid[n] = fabsf(s) < trunc ? 1 : 0;                      // sdf_rgbd.py:452
    sign[n] = s > trunc ? 1.0f : (s < -trunc ? -1.0f : 0.0f);  // sdf_rgbd.py:453-455
  }
}
}  // namespace miso

using namespace miso;

extern "C" int miso_expand_batch(const int16_t* ids16, const float* sdf, float trunc_dist, int64_t N, int64_t* ids64,
                                 uint8_t* valid, float* sign, miso_stream_t stream) {
  MISO_REQUIRE(N >= 0 && (N == 0 || (sdf && valid && sign)), "expand_batch: null argument");
  MISO_REQUIRE(!ids64 || ids16, "expand_batch: ids64 requested without ids16");
  MISO_REQUIRE(trunc_dist > 0.f, "expand_batch: trunc_dist must be positive");
  if (N == 0) return MISO_OK;
  expand_batch_kernel<<<grid_for(N, kThreads, sm_count() * 8), kThreads, 0, (cudaStream_t)stream>>>(
      ids16, sdf, trunc_dist, N, ids64, valid, sign);
  return check_launch("expand_batch");
}

extern "C" const char* miso_last_error_string(void) { return g_err; }
extern "C" int miso_abi_version(void) { return MISO_ABI_VERSION; }
extern "C" int miso_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("device_sm_count: no CUDA device");
    return MISO_ERR_CUDA;
  }
  return n;
}

extern "C" int miso_adam_step(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                              float eps, int32_t step, int32_t zero_grad, miso_stream_t stream) {
  MISO_REQUIRE(p && g && m && v, "adam_step: null tensor");
  MISO_REQUIRE(n >= 0 && step >= 1, "adam_step: n >= 0 and step >= 1 required");
  if (n == 0) return MISO_OK;
  const bool aligned = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const int blocks = grid_for(n4 > 0 ? n4 : n, kThreads, sm_count() * 8);
  adam_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n4, n, lr, beta1, beta2, eps, step_size,
                                                             bc2_sqrt, zero_grad, AdamDev{nullptr, nullptr, nullptr});
  return check_launch("adam_step");
}

// Adam on ONE boundary plane of a slab-sharded level (miso_b200/sharded_fit.py), fused with both halo exchanges over
// NVLink peer memory: the plane's gradient is g + g_peer, where g_peer is the SAME plane in the lower neighbour's
// gradient buffer (its samples' upper corners land there) read with peer loads and cleared in place; the updated
// parameters are stored locally and into the neighbour's copy of the plane (p_peer, peer stores), which its next step
// reads.  Replaces isend/irecv of the gradient plane + add + Adam + isend/irecv of the parameter plane.  Ordering
// between the ranks comes from the collectives around it (see SlabShardedFit._exchange_and_update).
__global__ void __launch_bounds__(kThreads)
    adam_halo_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     float* __restrict__ g_peer, float* __restrict__ p_peer, int64_t n4, float lr, float b1, float b2,
                     float eps, AdamDev dev) {
  __shared__ float sh_dev[3];
  float step_size = 0.f, bc2_sqrt = 1.f;
  adam_dev_begin(dev, lr, b1, b2, step_size, bc2_sqrt, sh_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 gv = reinterpret_cast<float4*>(g)[i];
    if (g_peer) {
      const float4 gp = reinterpret_cast<const float4*>(g_peer)[i];
      gv.x += gp.x, gv.y += gp.y, gv.z += gp.z, gv.w += gp.w;
      if (gp.x != 0.f || gp.y != 0.f || gp.z != 0.f || gp.w != 0.f) reinterpret_cast<float4*>(g_peer)[i] = zero4;
    }
    const float4 mv = reinterpret_cast<float4*>(m)[i];
    const float4 vv = reinterpret_cast<float4*>(v)[i];
    if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f && gv.w == 0.f && mv.x == 0.f && mv.y == 0.f && mv.z == 0.f &&
        mv.w == 0.f && vv.x == 0.f && vv.y == 0.f && vv.z == 0.f && vv.w == 0.f)
      continue;   // never touched: exactly-zero update, the neighbour's copy is already equal
    const float4 pv = reinterpret_cast<float4*>(p)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w},
          va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ma[e] = ma[e] + (ga[e] - ma[e]) * (1.f - b1);
      va[e] = va[e] * b2 + (1.f - b2) * ga[e] * ga[e];
      const float denom = sqrtf(va[e]) / bc2_sqrt + eps;
      pa[e] = pa[e] - step_size * (ma[e] / denom);
    }
    const float4 pn = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(p)[i] = pn;
    if (p_peer) reinterpret_cast<float4*>(p_peer)[i] = pn;
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    reinterpret_cast<float4*>(g)[i] = zero4;
  }
  adam_dev_end(dev, false);
}

extern "C" int miso_adam_step_halo(float* p, float* g, float* m, float* v, int64_t n, float* g_peer, float* p_peer,
                                   float lr, float beta1, float beta2, float eps, int32_t* step_counter, float* scalars,
                                   miso_stream_t stream) {
  MISO_REQUIRE(p && g && m && v && step_counter && scalars, "adam_step_halo: null tensor");
  MISO_REQUIRE(n > 0 && n % 4 == 0, "adam_step_halo: n must be a positive multiple of 4");
  MISO_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g_peer | (uintptr_t)p_peer) % 16) == 0,
               "adam_step_halo: tensors not 16-byte aligned");
  const AdamDev dev{step_counter, nullptr, reinterpret_cast<unsigned*>(scalars)};
  adam_halo_kernel<<<grid_for(n / 4, kThreads, sm_count() * 4), kThreads, 0, (cudaStream_t)stream>>>(
      p, g, m, v, g_peer, p_peer, n / 4, lr, beta1, beta2, eps, dev);
  return check_launch("adam_step_halo");
}

// Neighbour-to-neighbour ordering without a collective (slab-sharded fit): after its boundary-plane Adam a rank bumps a
// counter in the LOWER neighbour's memory; before its next step kernel that neighbour waits until the counter has reached
// the number of steps it has taken.  sync = {flag (written by the peer), expected, error, -}.  The wait gives up after
// ~30 s (error word set, checked by the host) instead of hanging the GPU.
__global__ void peer_signal_kernel(unsigned* flag_peer) {
  __threadfence_system();
  atomicAdd_system(flag_peer, 1u);
}
__global__ void peer_wait_kernel(unsigned* sync) {
  const unsigned want = sync[1];
  sync[1] = want + 1u;
  const long long t0 = clock64();
  while (*reinterpret_cast<volatile unsigned*>(sync) < want) {
    if (clock64() - t0 > 60000000000LL) {
      sync[2] = 1u;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}
extern "C" int miso_peer_signal(uint32_t* flag_peer, miso_stream_t stream) {
  MISO_REQUIRE(flag_peer, "peer_signal: null flag");
  peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_peer);
  return check_launch("peer_signal");
}
extern "C" int miso_peer_wait(uint32_t* sync, miso_stream_t stream) {
  MISO_REQUIRE(sync, "peer_wait: null sync words");
  peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sync);
  return check_launch("peer_wait");
}

// Same-node peer mapping of a device allocation (CUDA IPC): export on the owner, import in the neighbour's process.
// The handle names the whole cudaMalloc allocation `ptr` lives in; `offset` is ptr's distance from its base.
extern "C" int miso_ipc_export(const void* ptr, unsigned char handle[64], int64_t* offset) {
  MISO_REQUIRE(ptr && handle && offset, "ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn get_range = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (range_fn)f;
  }();
  MISO_REQUIRE(get_range, "ipc_export: cuMemGetAddressRange not available");
  unsigned long long base = 0;
  size_t size = 0;
  MISO_REQUIRE(get_range(&base, &size, (unsigned long long)(uintptr_t)ptr) == 0, "ipc_export: not a device allocation");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base);
  MISO_REQUIRE(e == cudaSuccess, "ipc_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle, &h, 64);
  *offset = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
  return MISO_OK;
}

extern "C" int miso_ipc_import(const unsigned char handle[64], int64_t offset, void** ptr) {
  MISO_REQUIRE(handle && ptr && offset >= 0, "ipc_import: bad argument");
  // an allocation can be opened once per process and device: cache the mappings
  struct Entry { unsigned char h[64]; int dev; void* base; };
  static std::vector<Entry> cache;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (const Entry& en : cache)
    if (en.dev == dev && !memcmp(en.h, handle, 64)) {
      *ptr = (char*)en.base + offset;
      return MISO_OK;
    }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  MISO_REQUIRE(e == cudaSuccess, "ipc_import: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  Entry en;
  memcpy(en.h, handle, 64);
  en.dev = dev, en.base = base;
  cache.push_back(en);
  *ptr = (char*)base + offset;
  return MISO_OK;
}

// step = ++(*counter); scalars = {lr / (1 - b1^step), sqrt(1 - b2^step)} in float64, exactly the host formula above
extern "C" int miso_adam_step_dev(float* p, float* g, float* m, float* v, uint32_t* touched, int64_t n, float lr,
                                  float beta1, float beta2, float eps, int32_t* step_counter, float* scalars,
                                  const float* gate, int32_t zero_grad, miso_stream_t stream) {
  MISO_REQUIRE(p && g && m && v && step_counter && scalars, "adam_step_dev: null tensor");
  MISO_REQUIRE(n > 0, "adam_step_dev: n > 0 required");
  cudaStream_t s = (cudaStream_t)stream;
  const AdamDev dev{step_counter, gate, reinterpret_cast<unsigned*>(scalars)};   // scalars[0]: the block ticket (zeroed by the caller once)
  const bool aligned = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  if (touched) {
    MISO_REQUIRE(aligned && n % 4 == 0, "adam_step_dev: the tracked variant needs 16-byte aligned tensors and n % 4 == 0");
    const int64_t n4 = n / 4;
    adam_tracked_kernel<<<grid_for(n4, kThreads, sm_count() * 8), kThreads, 0, s>>>(p, g, m, v, touched, n4, lr, beta1, beta2,
                                                                                   eps, 0.f, 1.f, zero_grad, dev);
  } else {
    const int64_t n4 = aligned ? n / 4 : 0;
    adam_kernel<<<grid_for(n4 > 0 ? n4 : n, kThreads, sm_count() * 8), kThreads, 0, s>>>(p, g, m, v, n4, n, lr, beta1, beta2,
                                                                                         eps, 0.f, 1.f, zero_grad, dev);
  }
  return check_launch("adam_step_dev");
}

extern "C" int miso_adam_step_tracked(float* p, float* g, float* m, float* v, uint32_t* touched, int64_t n, float lr,
                                      float beta1, float beta2, float eps, int32_t step, int32_t zero_grad,
                                      miso_stream_t stream) {
  MISO_REQUIRE(p && g && m && v && touched, "adam_step_tracked: null tensor");
  MISO_REQUIRE(n >= 0 && n % 4 == 0 && step >= 1, "adam_step_tracked: n must be a non-negative multiple of 4, step >= 1");
  MISO_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0, "adam_step_tracked: tensors not 16-byte aligned");
  if (n == 0) return MISO_OK;
  const int64_t n4 = n / 4;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int blocks = grid_for(n4, kThreads, sm_count() * 8);
  adam_tracked_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, touched, n4, lr, beta1, beta2, eps,
                                                                     (float)((double)lr / bc1), (float)sqrt(bc2), zero_grad,
                                                                     AdamDev{nullptr, nullptr, nullptr});
  return check_launch("adam_step_tracked");
}

extern "C" int miso_transform_points(const float* x, const int64_t* ids, const float* R, const float* t,
                                     int32_t num_frames, int64_t N, float* y, miso_stream_t stream) {
  MISO_REQUIRE(N >= 0 && (N == 0 || (x && ids && R && t && y)), "transform_points: null argument");
  MISO_REQUIRE(num_frames > 0, "transform_points: num_frames must be positive");
  if (N == 0) return MISO_OK;
  transform_kernel<<<grid_for(N, kThreads, sm_count() * 8), kThreads, 0, (cudaStream_t)stream>>>(x, ids, R, t, num_frames, N, y);
  return check_launch("transform_points");
}

extern "C" int miso_slab_select(const miso_frames_t* frames, const float* x, int64_t N, float zmin, float zmax,
                                int32_t Z, int32_t axis, int32_t z_begin, int32_t z_end, const float* gt_sdf, const uint8_t* gt_valid,
                                const float* gt_sign, const float* weights, float* x_out, int64_t* ids_out,
                                float* sdf_out, uint8_t* valid_out, float* sign_out, float* weights_out,
                                int32_t* count, miso_stream_t stream) {
  MISO_REQUIRE(count && N >= 0 && (N == 0 || (x && gt_sdf && gt_valid && gt_sign && x_out && sdf_out && valid_out && sign_out)),
               "slab_select: null argument");
  MISO_REQUIRE(Z > 0 && zmax > zmin && z_begin >= 0 && z_end <= Z && z_begin <= z_end, "slab_select: bad slab [%d,%d) of %d", z_begin, z_end, Z);
  MISO_REQUIRE(axis >= 0 && axis <= 2, "slab_select: axis must be 0 (x), 1 (y) or 2 (z)");
  const bool have_frames = frames && frames->ids;
  MISO_REQUIRE(!have_frames || (frames->R && frames->t && frames->num_frames > 0 && ids_out), "slab_select: frames without poses / ids_out");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(count, 0, sizeof(int32_t), s);
  if (N == 0) return check_launch("slab_select(memset)");
  MISO_REQUIRE(N < ((int64_t)1 << 31), "slab_select: N must fit int32");
  // one full wave of resident blocks (the kernel is latency-bound: a partial second wave costs a whole extra trip)
  static int per_sm = 0;
  if (!per_sm) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, slab_select_kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
  }
  const int blocks = grid_for((N + kThreads * kSelPer - 1) / (kThreads * kSelPer), 1, sm_count() * per_sm);
  slab_select_kernel<<<blocks, kThreads, 0, s>>>(x, have_frames ? frames->ids : nullptr, have_frames ? frames->R : nullptr,
                                                 have_frames ? frames->t : nullptr, have_frames ? frames->num_frames : 0, N,
                                                 zmin, zmax, Z, axis, z_begin, z_end, gt_sdf, gt_valid, gt_sign, weights, x_out,
                                                 ids_out, sdf_out, valid_out, sign_out, weights_out, count);
  return check_launch("slab_select");
}

extern "C" int miso_morton_keys(const float* x, int64_t N, const float bound[6], uint32_t* keys, miso_stream_t stream) {
  MISO_REQUIRE(N >= 0 && (N == 0 || (x && keys)) && bound, "morton_keys: null argument");
  if (N == 0) return MISO_OK;
  morton_kernel<<<grid_for(N, kThreads, sm_count() * 8), kThreads, 0, (cudaStream_t)stream>>>(
      x, N, bound[0], bound[1], bound[2], bound[3], bound[4], bound[5], keys);
  return check_launch("morton_keys");
}
