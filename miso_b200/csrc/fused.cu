// Fused multiresolution-grid + MLP-decoder kernels: per-point features never touch HBM.
//
// Replaces the chain
//   normalize_coordinates (utils.py:22-51) -> FeatureGrid.interpolate per level (grid_modules.py:72-95)
//   -> torch.cat (utils.py:164) -> MLPNet.forward (modules.py:31-32) -> gradient3d (diff.py:14-38)
//   -> miso_loss_regression / miso_loss_free_space / miso_loss_eikonal (loss.py:594-700)
//   -> autograd backward through aten::grid_sampler_3d_backward and grid_sampler_3d_grad2_kernel.
//
// Math: with J = d sdf/d feat = W3 D2 W2 D1 W1 (ReLU masks D), and because
// d2relu = 0 a.e., both the first-order loss gradient and the eikonal double-backward reduce to ONE
// scatter per corner:   grad_l[corner] += (a * w_c + v . dw_c/dx) * J_l
// with a = dL/dsdf and v = dL/d(grad_x sdf).
//
// SIMT design for sm_100a: one thread per point, decoder weights staged once per CTA in shared
// memory and read as warp-uniform 128-bit broadcasts, all MLP math as packed FFMA2 (fma.rn.f32x2),
// corner fetches as 128-bit read-only loads from the channels-last grid, scatter as 128-bit
// red.global.add.v4.f32.  Persistent CTAs (grid = SMs x occupancy) loop over 256-point tiles.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace miso {

#ifndef MISO_MAP_MIN_BLOCKS
#define MISO_MAP_MIN_BLOCKS 2   // 2 CTAs/SM (<=128 regs, ~170 B spill) measured faster than 1 CTA at 175 regs
#endif

constexpr int H = 64;  // decoder hidden_dim (configs/rgbd/scannet.yaml:12, configs/lidar/ncd_quad.yaml:11)

template <int F>
struct DecoderSmem {
  float W1[H * F];  // [k][i]
  float W2[H * H];  // [j][k]
  float b1[H];
  float b2[H];
  float W3[H];
  float b3[4];
};

template <int F>
__device__ __forceinline__ void load_decoder(DecoderSmem<F>* s, const miso_decoder_t& d) {
  for (int i = threadIdx.x; i < H * F; i += blockDim.x) s->W1[i] = d.W1[i];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) s->W2[i] = d.W2[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    s->b1[i] = d.b1[i];
    s->b2[i] = d.b2[i];
    s->W3[i] = d.W3[i];
  }
  if (threadIdx.x == 0) s->b3[0] = d.b3[0];
  __syncthreads();
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// MLP forward + Jacobian wrt input.  f: F inputs.  Returns sdf; J[F] = d sdf / d f.
template <int F, bool kJac>
__device__ __forceinline__ float mlp_eval(const DecoderSmem<F>* s, const float (&f)[F], float (&J)[F]) {
  static_assert(F % 4 == 0, "F must be a multiple of 4");
  float2 fp[F / 2];
#pragma unroll
  for (int i = 0; i < F / 2; ++i) fp[i] = make_float2(f[2 * i], f[2 * i + 1]);

  // ---- layer 1: h1 = relu(W1 f + b1), kept as 32 packed pairs --------------------------------
  float2 h1[H / 2];
  unsigned m1lo = 0, m1hi = 0;
#pragma unroll
  for (int kp = 0; kp < H / 2; ++kp) {
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
    const float4* r0 = reinterpret_cast<const float4*>(s->W1 + (2 * kp) * F);
    const float4* r1 = reinterpret_cast<const float4*>(s->W1 + (2 * kp + 1) * F);
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      float4 w0 = r0[q], w1 = r1[q];
      a0 = ffma2(make_float2(w0.x, w0.y), fp[2 * q], a0);
      a0 = ffma2(make_float2(w0.z, w0.w), fp[2 * q + 1], a0);
      a1 = ffma2(make_float2(w1.x, w1.y), fp[2 * q], a1);
      a1 = ffma2(make_float2(w1.z, w1.w), fp[2 * q + 1], a1);
    }
    float2 bb = *reinterpret_cast<const float2*>(s->b1 + 2 * kp);
    float x0 = (a0.x + a0.y) + bb.x, x1 = (a1.x + a1.y) + bb.y;
    unsigned p0 = x0 > 0.f, p1 = x1 > 0.f;
    if (kp < 16) m1lo |= (p0 << (2 * kp)) | (p1 << (2 * kp + 1));
    else m1hi |= (p0 << (2 * kp - 32)) | (p1 << (2 * kp - 31));
    h1[kp] = make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  }

  // ---- layer 2 + 3 forward: sdf = W3 relu(W2 h1 + b2) + b3 ------------------------------------
  float sdf = s->b3[0];
  unsigned m2lo = 0, m2hi = 0;
#pragma unroll 1
  for (int j = 0; j < H; j += 4) {
    float2 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float4 w = reinterpret_cast<const float4*>(s->W2 + (j + r) * H)[q];
        acc[r] = ffma2(make_float2(w.x, w.y), h1[2 * q], acc[r]);
        acc[r] = ffma2(make_float2(w.z, w.w), h1[2 * q + 1], acc[r]);
      }
    }
    float4 b2v = *reinterpret_cast<const float4*>(s->b2 + j);
    float4 w3v = *reinterpret_cast<const float4*>(s->W3 + j);
    float hb[4] = {b2v.x, b2v.y, b2v.z, b2v.w};
    float hw[4] = {w3v.x, w3v.y, w3v.z, w3v.w};
    unsigned bits = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float h2 = (acc[r].x + acc[r].y) + hb[r];
      bits |= (h2 > 0.f ? 1u : 0u) << r;
      sdf = fmaf(hw[r], fmaxf(h2, 0.f), sdf);
    }
    if (j < 32) m2lo |= bits << j;
    else m2hi |= bits << (j - 32);
  }
  if constexpr (!kJac) return sdf;

  // ---- backward for J: g1 = D1 W2^T D2 W3^T ; J = W1^T g1 -------------------------------------
  float2 g1[H / 2];
#pragma unroll
  for (int kp = 0; kp < H / 2; ++kp) g1[kp] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int j = 0; j < H; j += 2) {
    unsigned word = (j < 32) ? m2lo : m2hi;
    unsigned sh = j & 31;
    float2 w3 = *reinterpret_cast<const float2*>(s->W3 + j);
    float t0 = ((word >> sh) & 1u) ? w3.x : 0.f;
    float t1 = ((word >> (sh + 1)) & 1u) ? w3.y : 0.f;
    float2 tt0 = make_float2(t0, t0), tt1 = make_float2(t1, t1);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      float4 w0 = reinterpret_cast<const float4*>(s->W2 + j * H)[q];
      float4 w1 = reinterpret_cast<const float4*>(s->W2 + (j + 1) * H)[q];
      g1[2 * q] = ffma2(make_float2(w0.x, w0.y), tt0, g1[2 * q]);
      g1[2 * q + 1] = ffma2(make_float2(w0.z, w0.w), tt0, g1[2 * q + 1]);
      g1[2 * q] = ffma2(make_float2(w1.x, w1.y), tt1, g1[2 * q]);
      g1[2 * q + 1] = ffma2(make_float2(w1.z, w1.w), tt1, g1[2 * q + 1]);
    }
  }
  // compiler barrier: without it nvcc CSEs these W1 reads with layer 1's and keeps 64*F weights
  // alive across the whole MLP (2.5 KB of local-memory spills per thread)
  asm volatile("" ::: "memory");
  float2 Jp[F / 2];
#pragma unroll
  for (int i = 0; i < F / 2; ++i) Jp[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int kp = 0; kp < H / 2; ++kp) {
    unsigned word = (kp < 16) ? m1lo : m1hi;
    unsigned sh = (2 * kp) & 31;
    float e0 = ((word >> sh) & 1u) ? g1[kp].x : 0.f;
    float e1 = ((word >> (sh + 1)) & 1u) ? g1[kp].y : 0.f;
    float2 ee0 = make_float2(e0, e0), ee1 = make_float2(e1, e1);
    const float4* r0 = reinterpret_cast<const float4*>(s->W1 + (2 * kp) * F);
    const float4* r1 = reinterpret_cast<const float4*>(s->W1 + (2 * kp + 1) * F);
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      float4 w0 = r0[q], w1 = r1[q];
      Jp[2 * q] = ffma2(make_float2(w0.x, w0.y), ee0, Jp[2 * q]);
      Jp[2 * q + 1] = ffma2(make_float2(w0.z, w0.w), ee0, Jp[2 * q + 1]);
      Jp[2 * q] = ffma2(make_float2(w1.x, w1.y), ee1, Jp[2 * q]);
      Jp[2 * q + 1] = ffma2(make_float2(w1.z, w1.w), ee1, Jp[2 * q + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < F / 2; ++i) {
    J[2 * i] = Jp[i].x;
    J[2 * i + 1] = Jp[i].y;
  }
  return sdf;
}

// ---- per-point geometry -------------------------------------------------------------------------
struct FieldGeom {
  float bmin[3], bmax[3];
  float inv_len[3];  // 1/(bmax-bmin)
};

__device__ __forceinline__ FieldGeom field_geom(const miso_field_t& fl) {
  FieldGeom g;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    g.bmin[d] = fl.bound[2 * d];
    g.bmax[d] = fl.bound[2 * d + 1];
    g.inv_len[d] = 1.0f / (g.bmax[d] - g.bmin[d]);
  }
  return g;
}

// A sample whose keyframe id has no pose (outside [0, num_frames), or a table row the host marked with a NaN
// translation because no 'KF<id>' key was registered) must not train silently with somebody else's pose -- the
// reference asserts "Key KF.. not found" (grid_net.py:243).  Such a sample gets NaN coordinates (it then interpolates
// zeros, scatters to no voxel) and, in the mapping step, raises the poison word that turns the step's loss into NaN.
__device__ __forceinline__ void flag_bad_frame(float* poison) {
  if (poison) *poison = __int_as_float(0x7fc00000);
}

__device__ __forceinline__ void load_point(const float* __restrict__ x, const miso_frames_t& fr, int64_t n,
                                           float (&p)[3], float* poison = nullptr) {
  float a = x[3 * n], b = x[3 * n + 1], c = x[3 * n + 2];
  if (fr.ids) {
    // transform_points_to (utils_geometry.py:214-225): x R^T + t^T, pose picked per sample (loss.py:764-774)
    int64_t id = fr.ids[n];
    const bool bad = id < 0 || id >= fr.num_frames;
    if (bad) id = 0;
    const float* R = fr.R + id * 9;
    const float* t = fr.t + id * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) p[j] = fmaf(c, R[3 * j + 2], fmaf(b, R[3 * j + 1], a * R[3 * j])) + t[j];
    if (bad) p[0] = p[1] = p[2] = __int_as_float(0x7fc00000);
    if (p[0] != p[0]) flag_bad_frame(poison);
  } else {
    p[0] = a, p[1] = b, p[2] = c;
  }
}

__device__ __forceinline__ Cell level_cell(const miso_level_t& lv, const float (&xn)[3]) {
  return make_cell(unnormalize_nc(xn[0], lv.X), unnormalize_nc(xn[1], lv.Y), unnormalize_nc(xn[2], lv.Z), lv);
}

// Gather C channels of one level: features f[C] and index-space derivatives d[3][C].
// Separable evaluation (lerp along x, then y, then z) of the trilinear form and of its three partial
// derivatives: 22 flops per channel instead of 32 FMAs + 32 corner-weight products.  Out-of-range corners
// contribute the value 0 (zeros padding, gridsample_cuda.cu:385-441), exactly like masking their weights.
template <int C, bool kDeriv>
__device__ __forceinline__ void gather_level(const miso_level_t& lv, const Cell& c, float* __restrict__ f,
                                             float* __restrict__ dfx, float* __restrict__ dfy,
                                             float* __restrict__ dfz) {
  long long off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = ((c.valid >> k) & 1u) ? corner_off(lv, c, k) : -1;
#pragma unroll
  for (int ch = 0; ch < C; ch += 4) {
    float v[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (off[k] >= 0) t = ldg_f4(lv.feat + off[k] + ch);
      v[k][0] = t.x, v[k][1] = t.y, v[k][2] = t.z, v[k][3] = t.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a[4], d[4];
#pragma unroll
      for (int yz = 0; yz < 4; ++yz) {
        d[yz] = v[2 * yz + 1][e] - v[2 * yz][e];
        a[yz] = fmaf(c.fx, d[yz], v[2 * yz][e]);
      }
      float ay[2], ey[2], dxy[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        ey[z] = a[2 * z + 1] - a[2 * z];
        ay[z] = fmaf(c.fy, ey[z], a[2 * z]);
        if constexpr (kDeriv) dxy[z] = fmaf(c.fy, d[2 * z + 1] - d[2 * z], d[2 * z]);
      }
      const float ez = ay[1] - ay[0];
      f[ch + e] = fmaf(c.fz, ez, ay[0]);
      if constexpr (kDeriv) {
        dfz[ch + e] = ez;
        dfy[ch + e] = fmaf(c.fz, ey[1] - ey[0], ey[0]);
        dfx[ch + e] = fmaf(c.fz, dxy[1] - dxy[0], dxy[0]);
      }
    }
  }
}

__device__ __forceinline__ float ldg_early_f32(const float* p) {
  float r;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned ldg_early_u8(const uint8_t* p) {
  unsigned r;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// Scatter (a*w_c + vi . dw_c/di) * J into one level's gradient buffer; vi is in index space.
// The per-corner coefficient is built separably:  coef = wz*(wy*(a*wx + vix*sx) + viy*sy*wx) + viz*sz*wx*wy.
template <int C>
__device__ __forceinline__ void scatter_level(const miso_level_t& lv, const Cell& c, float a, float vix, float viy,
                                              float viz, const float* __restrict__ J) {
  const float wx[2] = {1.0f - c.fx, c.fx}, wy[2] = {1.0f - c.fy, c.fy}, wz[2] = {1.0f - c.fz, c.fz};
  float px[2] = {fmaf(a, wx[0], -vix), fmaf(a, wx[1], vix)};
  float r[4], q[4];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float sy = dy ? viy : -viy;
      r[2 * dy + dx] = fmaf(wy[dy], px[dx], sy * wx[dx]);
      q[2 * dy + dx] = wx[dx] * wy[dy];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (!((c.valid >> k) & 1u)) continue;
    const int dz = k >> 2;
    const float sz = dz ? viz : -viz;
    const float coef = fmaf(wz[dz], r[k & 3], sz * q[k & 3]);
    float* dst = lv.grad + corner_off(lv, c, k);
#pragma unroll
    for (int ch = 0; ch < C; ch += 4)
      red_add_f4(dst + ch, coef * J[ch], coef * J[ch + 1], coef * J[ch + 2], coef * J[ch + 3]);
  }
}

// ---- lean per-level cell for the tensor-core kernels: 32-bit element index (host guarantees every level has
// < 2^31 elements on this path), kept in registers from the gather to the scatter so the index math, the
// validity tests and the 64-bit address arithmetic are done once per point and level -----------------------
struct CellLite {
  int base;        // element index of corner (ix0,iy0,iz0); may be "virtual" (negative) when that corner is outside
  float fx, fy, fz;
  unsigned valid;  // bit k set when corner k lies inside the grid
};

__device__ __forceinline__ CellLite make_cell_lite(const miso_level_t& lv, const float (&xn)[3]) {
  const float ix = unnormalize_nc(xn[0], lv.X), iy = unnormalize_nc(xn[1], lv.Y), iz = unnormalize_nc(xn[2], lv.Z);
  const float flx = floorf(ix), fly = floorf(iy), flz = floorf(iz);
  const int x0 = (int)fminf(fmaxf(flx, -2.0f), (float)lv.X + 1.0f);
  const int y0 = (int)fminf(fmaxf(fly, -2.0f), (float)lv.Y + 1.0f);
  const int z0 = (int)fminf(fmaxf(flz, -2.0f), (float)lv.Z + 1.0f);
  CellLite c;
  c.fx = ix - flx, c.fy = iy - fly, c.fz = iz - flz;
  c.base = z0 * (int)lv.sZ + y0 * (int)lv.sY + x0 * (int)lv.sX;
  if (x0 >= 0 && x0 + 1 < lv.X && y0 >= 0 && y0 + 1 < lv.Y && z0 >= 0 && z0 + 1 < lv.Z) {
    c.valid = 0xffu;  // interior: the overwhelmingly common case, no per-corner tests
  } else {
    const unsigned vx = ((unsigned)x0 < (unsigned)lv.X ? 1u : 0u) | ((unsigned)(x0 + 1) < (unsigned)lv.X ? 2u : 0u);
    const unsigned vy = ((unsigned)y0 < (unsigned)lv.Y ? 1u : 0u) | ((unsigned)(y0 + 1) < (unsigned)lv.Y ? 2u : 0u);
    const unsigned vz = ((unsigned)z0 < (unsigned)lv.Z ? 1u : 0u) | ((unsigned)(z0 + 1) < (unsigned)lv.Z ? 2u : 0u);
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v |= (((vx >> (k & 1)) & (vy >> ((k >> 1) & 1)) & (vz >> (k >> 2))) & 1u) << k;
    c.valid = v;
  }
  return c;
}

__device__ __forceinline__ int corner_delta(const miso_level_t& lv, int k) {
  return ((k & 1) ? (int)lv.sX : 0) + ((k & 2) ? (int)lv.sY : 0) + ((k & 4) ? (int)lv.sZ : 0);
}

template <bool kDeriv>
__device__ __forceinline__ void lerp_corners4(const float4 (&v)[8], const CellLite& c, float* __restrict__ f,
                                              float* __restrict__ dfx, float* __restrict__ dfy,
                                              float* __restrict__ dfz) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float a[4], d[4];
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
      const float v0 = reinterpret_cast<const float*>(&v[2 * yz])[e];
      const float v1 = reinterpret_cast<const float*>(&v[2 * yz + 1])[e];
      d[yz] = v1 - v0;
      a[yz] = fmaf(c.fx, d[yz], v0);
    }
    float ay[2], ey[2], dxy[2];
#pragma unroll
    for (int z = 0; z < 2; ++z) {
      ey[z] = a[2 * z + 1] - a[2 * z];
      ay[z] = fmaf(c.fy, ey[z], a[2 * z]);
      if constexpr (kDeriv) dxy[z] = fmaf(c.fy, d[2 * z + 1] - d[2 * z], d[2 * z]);
    }
    const float ez = ay[1] - ay[0];
    f[e] = fmaf(c.fz, ez, ay[0]);
    if constexpr (kDeriv) {
      dfz[e] = ez;
      dfy[e] = fmaf(c.fz, ey[1] - ey[0], ey[0]);
      dfx[e] = fmaf(c.fz, dxy[1] - dxy[0], dxy[0]);
    }
  }
}

template <int C, bool kDeriv>
__device__ __forceinline__ void gather_level_lite(const miso_level_t& lv, const CellLite& c, float* __restrict__ f,
                                                  float* __restrict__ dfx, float* __restrict__ dfy,
                                                  float* __restrict__ dfz) {
#pragma unroll
  for (int ch = 0; ch < C; ch += 4) {
    float4 v[8];
    if (c.valid == 0xffu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ldg_f4(lv.feat + (c.base + corner_delta(lv, k) + ch));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((c.valid >> k) & 1u) v[k] = ldg_f4(lv.feat + (c.base + corner_delta(lv, k) + ch));
      }
    }
    lerp_corners4<kDeriv>(v, c, f + ch, dfx + ch, dfy + ch, dfz + ch);
  }
}

template <int C>
__device__ __forceinline__ void scatter_level_lite(const miso_level_t& lv, const CellLite& c, unsigned on, float a,
                                                   float vix, float viy, float viz, const float* __restrict__ J) {
  const float wx[2] = {1.0f - c.fx, c.fx}, wy[2] = {1.0f - c.fy, c.fy}, wz[2] = {1.0f - c.fz, c.fz};
  const float px[2] = {fmaf(a, wx[0], -vix), fmaf(a, wx[1], vix)};
  float r[4], q[4];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float sy = dy ? viy : -viy;
      r[2 * dy + dx] = fmaf(wy[dy], px[dx], sy * wx[dx]);
      q[2 * dy + dx] = wx[dx] * wy[dy];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dz = k >> 2;
    const float sz = dz ? viz : -viz;
    const float coef = fmaf(wz[dz], r[k & 3], sz * q[k & 3]);
    // out-of-grid corners have a "virtual" index: clamp the address, the predicate suppresses the access
    const unsigned ok = on & (c.valid >> k) & 1u;
    float* dst = lv.grad + (ok ? c.base + corner_delta(lv, k) : 0);
#pragma unroll
    for (int ch = 0; ch < C; ch += 4)
      red_add_f4_if(ok, dst + ch, coef * J[ch], coef * J[ch + 1], coef * J[ch + 2], coef * J[ch + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// kernel: features only (grid_interp_regular)
// ---------------------------------------------------------------------------------------------
template <int L, int C>
__global__ void __launch_bounds__(kThreads) field_features_kernel(const __grid_constant__ miso_field_t fl,
                                                                  const float* __restrict__ x, int64_t N,
                                                                  float* __restrict__ feats) {
  const FieldGeom g = field_geom(fl);
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float xn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(x[3 * n + d], g.bmin[d], g.bmax[d]);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      float f[C];
      if ((fl.ignore_mask >> l) & 1u) {
#pragma unroll
        for (int i = 0; i < C; ++i) f[i] = 0.f;
      } else {
        Cell c = level_cell(fl.level[l], xn);
        gather_level<C, false>(fl.level[l], c, f, nullptr, nullptr, nullptr);
      }
#pragma unroll
      for (int ch = 0; ch < C; ch += 4)
        *reinterpret_cast<float4*>(feats + n * (L * C) + l * C + ch) = make_float4(f[ch], f[ch + 1], f[ch + 2], f[ch + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kernel: fused forward (sdf, jac, gradx)
// ---------------------------------------------------------------------------------------------
template <int L, int C, bool kJac>
__global__ void __launch_bounds__(kThreads)
    sdf_forward_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                       const __grid_constant__ miso_frames_t fr, const float* __restrict__ x, int64_t N,
                       float* __restrict__ sdf, float* __restrict__ jac, float* __restrict__ gradx,
                       float* __restrict__ xw) {
  constexpr int F = L * C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecoderSmem<F>* s = reinterpret_cast<DecoderSmem<F>*>(smem_raw);
  load_decoder<F>(s, dec);
  const FieldGeom g = field_geom(fl);
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float p[3], xn[3];
    load_point(x, fr, n, p);
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(p[d], g.bmin[d], g.bmax[d]);
    float f[F], dfx[F], dfy[F], dfz[F];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if ((fl.ignore_mask >> l) & 1u) {
#pragma unroll
        for (int i = 0; i < C; ++i) f[l * C + i] = dfx[l * C + i] = dfy[l * C + i] = dfz[l * C + i] = 0.f;
      } else {
        Cell c = level_cell(fl.level[l], xn);
        gather_level<C, kJac>(fl.level[l], c, f + l * C, dfx + l * C, dfy + l * C, dfz + l * C);
        if constexpr (kJac) {
          // chain d(index)/dx = S/len (ATen gi*_mult = S/2 times normalize's 2/len)
          float kx = (float)fl.level[l].X * g.inv_len[0], ky = (float)fl.level[l].Y * g.inv_len[1],
                kz = (float)fl.level[l].Z * g.inv_len[2];
#pragma unroll
          for (int i = 0; i < C; ++i) dfx[l * C + i] *= kx, dfy[l * C + i] *= ky, dfz[l * C + i] *= kz;
        }
      }
    }
    float J[F];
    float val = mlp_eval<F, kJac>(s, f, J);
    sdf[n] = val;
    if (xw) xw[3 * n] = p[0], xw[3 * n + 1] = p[1], xw[3 * n + 2] = p[2];
    if constexpr (kJac) {
      if (jac) {
#pragma unroll
        for (int i = 0; i < F; i += 4) *reinterpret_cast<float4*>(jac + n * F + i) = make_float4(J[i], J[i + 1], J[i + 2], J[i + 3]);
      }
      if (gradx) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int i = 0; i < F; ++i) gx = fmaf(J[i], dfx[i], gx), gy = fmaf(J[i], dfy[i], gy), gz = fmaf(J[i], dfz[i], gz);
        gradx[3 * n] = gx, gradx[3 * n + 1] = gy, gradx[3 * n + 2] = gz;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kernel: Gauss-Newton / LM normal equations of one keyframe (Tracker.lm_step, grid_opt/slam/tracker.py:148-212)
//   x_w = R x + t ; r = sdf(x_w) - gt ; g = grad_x sdf(x_w) ; J = [ ((R x) x g)^T R , g^T ] (N,6)
//   w = 1 (L2) or c / (c + r^2)^2 (Geman-McClure, :139-146) ; H = J^T W J ; b = J^T W r
// in ONE launch: transform, interpolation of every level, decoder + analytic gradient, residual weight, the 6-vector
// J and the 21 + 6 reductions (per-thread partials -> warp shuffle -> float64 atomics), plus the in-bound count the
// reference reports as fov_overlap (:176) and the number of samples that passed the |gt| < trunc filter (:158-164).
// out (double[45]): H row-major [0,36) (both triangles), b [36,42), in-bound count [42], used count [43], sum w r^2 [44]
// ---------------------------------------------------------------------------------------------
template <int L, int C>
__global__ void __launch_bounds__(kThreads)
    track_normal_equations_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                                  const float* __restrict__ x, const float* __restrict__ gt_sdf, int64_t N,
                                  const float* __restrict__ Rt, int loss_type, float gm_scale, float trunc_dist,
                                  double* __restrict__ out) {
  constexpr int F = L * C;
  constexpr int kAcc = 21 + 6 + 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecoderSmem<F>* s = reinterpret_cast<DecoderSmem<F>*>(smem_raw);
  __shared__ float s_part[kThreads / 32][kAcc];
  load_decoder<F>(s, dec);
  const FieldGeom g = field_geom(fl);
  float R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = Rt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = Rt[9 + i];
  float acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const float gtv = gt_sdf[n];
    if (trunc_dist >= 0.f && !(fabsf(gtv) < trunc_dist)) continue;
    const float a = x[3 * n], b = x[3 * n + 1], c = x[3 * n + 2];
    float xr[3], p[3], xn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      xr[j] = fmaf(c, R[3 * j + 2], fmaf(b, R[3 * j + 1], a * R[3 * j]));   // x R^T (transform_points_to, t = 0)
      p[j] = xr[j] + t[j];
      xn[j] = normalize_coord(p[j], g.bmin[j], g.bmax[j]);
    }
    const bool inb = p[0] >= g.bmin[0] && p[0] <= g.bmax[0] && p[1] >= g.bmin[1] && p[1] <= g.bmax[1] &&
                     p[2] >= g.bmin[2] && p[2] <= g.bmax[2];
    float f[F], dfx[F], dfy[F], dfz[F];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if ((fl.ignore_mask >> l) & 1u) {
#pragma unroll
        for (int i = 0; i < C; ++i) f[l * C + i] = dfx[l * C + i] = dfy[l * C + i] = dfz[l * C + i] = 0.f;
      } else {
        Cell cl = level_cell(fl.level[l], xn);
        gather_level<C, true>(fl.level[l], cl, f + l * C, dfx + l * C, dfy + l * C, dfz + l * C);
        const float kx = (float)fl.level[l].X * g.inv_len[0], ky = (float)fl.level[l].Y * g.inv_len[1],
                    kz = (float)fl.level[l].Z * g.inv_len[2];
#pragma unroll
        for (int i = 0; i < C; ++i) dfx[l * C + i] *= kx, dfy[l * C + i] *= ky, dfz[l * C + i] *= kz;
      }
    }
    float Jf[F];
    const float sdf = mlp_eval<F, true>(s, f, Jf);
    float gw[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < F; ++i) gw[0] = fmaf(Jf[i], dfx[i], gw[0]), gw[1] = fmaf(Jf[i], dfy[i], gw[1]), gw[2] = fmaf(Jf[i], dfz[i], gw[2]);
    // hat(R x) g = (R x) x g ; J_R = (that)^T R ; J_t = g
    const float cx = xr[1] * gw[2] - xr[2] * gw[1], cy = xr[2] * gw[0] - xr[0] * gw[2], cz = xr[0] * gw[1] - xr[1] * gw[0];
    float J[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) J[j] = cx * R[j] + cy * R[3 + j] + cz * R[6 + j];
    J[3] = gw[0], J[4] = gw[1], J[5] = gw[2];
    const float r = sdf - gtv;
    const float w = loss_type == 0 ? 1.0f : gm_scale / ((gm_scale + r * r) * (gm_scale + r * r));
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) acc[k++] += w * J[i] * J[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[21 + i] += w * J[i] * r;
    acc[27] += inb ? 1.f : 0.f;
    acc[28] += 1.f;
    acc[29] += w * r * r;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double v = 0.0;
    for (int w2 = 0; w2 < kThreads / 32; ++w2) v += (double)s_part[w2][threadIdx.x];
    if (v != 0.0) {
      const int i = threadIdx.x;
      if (i < 21) {
        int a2 = 0, tt = i;
        while (tt >= 6 - a2) tt -= 6 - a2, ++a2;
        const int b2 = a2 + tt;
        atomicAdd(out + a2 * 6 + b2, v);
        if (a2 != b2) atomicAdd(out + b2 * 6 + a2, v);
      } else {
        atomicAdd(out + 36 + (i - 21), v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kernel: scatter backward given per-point J, a, v  (+ optional Hessian-vector product wrt x)
// ---------------------------------------------------------------------------------------------
template <int L, int C>
__global__ void __launch_bounds__(kThreads)
    sdf_backward_kernel(const __grid_constant__ miso_field_t fl, const float* __restrict__ xw, int64_t N,
                        const float* __restrict__ jac, const float* __restrict__ a, const float* __restrict__ v,
                        float* __restrict__ hv) {
  constexpr int F = L * C;
  const FieldGeom g = field_geom(fl);
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float av = a ? a[n] : 0.f;
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (v) vx = v[3 * n], vy = v[3 * n + 1], vz = v[3 * n + 2];
    const bool nothing = av == 0.f && vx == 0.f && vy == 0.f && vz == 0.f;
    if (nothing) {
      if (hv) hv[3 * n] = 0.f, hv[3 * n + 1] = 0.f, hv[3 * n + 2] = 0.f;
      continue;
    }
    float xn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(xw[3 * n + d], g.bmin[d], g.bmax[d]);
    float J[F];
#pragma unroll
    for (int i = 0; i < F; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(jac + n * F + i);
      J[i] = t.x, J[i + 1] = t.y, J[i + 2] = t.z, J[i + 3] = t.w;
    }
    float hx = 0.f, hy = 0.f, hz = 0.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if ((fl.ignore_mask >> l) & 1u) continue;
      const miso_level_t& lv = fl.level[l];
      Cell c = level_cell(lv, xn);
      float kx = (float)lv.X * g.inv_len[0], ky = (float)lv.Y * g.inv_len[1], kz = (float)lv.Z * g.inv_len[2];
      const float ux = vx * kx, uy = vy * ky, uz = vz * kz;
      if (lv.grad) scatter_level<C>(lv, c, av, ux, uy, uz, J + l * C);
      if (hv) {
        // d/dx of (v . grad_x sdf) with J held fixed: mixed second derivatives of the trilinear
        // weights (dxy, dxz, dyz of gridsample_cuda.cu:484-526); the diagonal is zero.
        float dxy = 0.f, dxz = 0.f, dyz = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!((c.valid >> k) & 1u)) continue;
          float wx, wy, wz;
          axis_w(c, k, wx, wy, wz);
          float sx = (k & 1) ? 1.f : -1.f, sy = (k & 2) ? 1.f : -1.f, sz = (k & 4) ? 1.f : -1.f;
          float sc = 0.f;
#pragma unroll
          for (int ch = 0; ch < C; ch += 4) {
            float4 val = ldg_f4(lv.feat + corner_off(lv, c, k) + ch);
            sc = fmaf(val.x, J[l * C + ch], sc);
            sc = fmaf(val.y, J[l * C + ch + 1], sc);
            sc = fmaf(val.z, J[l * C + ch + 2], sc);
            sc = fmaf(val.w, J[l * C + ch + 3], sc);
          }
          dxy = fmaf(sc, sx * sy * wz, dxy);
          dxz = fmaf(sc, sx * wy * sz, dxz);
          dyz = fmaf(sc, wx * sy * sz, dyz);
        }
        hx = fmaf(kx, uy * dxy + uz * dxz, hx);
        hy = fmaf(ky, ux * dxy + uz * dyz, hy);
        hz = fmaf(kz, ux * dxz + uy * dyz, hz);
      }
    }
    if (hv) hv[3 * n] = hx, hv[3 * n + 1] = hy, hv[3 * n + 2] = hz;
  }
}

// ---------------------------------------------------------------------------------------------
// kernel: whole mapping step (forward + losses + backward scatter) -- the headline kernel
// ---------------------------------------------------------------------------------------------
struct MapArgs {
  const float* x;
  int64_t N;
  const float* gt_sdf;
  const uint8_t* gt_valid;
  const float* gt_sign;
  const float* weights;
  miso_mapping_cfg_t cfg;
  const int32_t* eik_count;
  float* partials;
  float* poison;   // one float past the per-block partials: set to NaN by a sample without a pose, consumed by finalize
  float* sdf_out;
  float* jac;      // forward modes of the two-thread kernel: (N,F) Jacobian, (N,3) grad_x sdf, (N,3) world coordinates
  float* gradx;
  float* xw;
  const float* a_ext;   // two-thread kernel, step mode: per-point d total / d sdf given by the caller (no loss terms)
  int fd_n;             // > 0: the launch runs over 6 * fd_n virtual points displaced by +-fd_eps (finite differences)
  float fd_eps;
  float inv_len[3];                       // 1/(bmax-bmin), computed on the host with the device's fp32 ops
  float lvl_scale[MISO_MAX_LEVELS][3];    // (float)dim * inv_len: index-space -> world-space derivative scale
  int dbg;   // MISO_DBG ablation bits (profiling only): 1 = no reductions, 2 = no corner loads, 4 = no g1 TMEM loads
};

template <int L, int C>
__global__ void __launch_bounds__(kThreads, MISO_MAP_MIN_BLOCKS)
    mapping_step_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                        const __grid_constant__ miso_frames_t fr, const __grid_constant__ MapArgs m) {
  constexpr int F = L * C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecoderSmem<F>* s = reinterpret_cast<DecoderSmem<F>*>(smem_raw);
  __shared__ float red[32];
  load_decoder<F>(s, dec);
  const FieldGeom g = field_geom(fl);
  const bool eik_on = m.cfg.eik_mode != 0 && m.cfg.weight_eik != 0.f;
  const bool eik_filter = m.cfg.eik_trunc_dist >= 0.f;
  const float n_den = (float)(m.cfg.n_total > 0 ? m.cfg.n_total : m.N);
  const float invN = 1.0f / n_den;
  float n_eik = n_den;
  if (eik_on && eik_filter) n_eik = (float)(*m.eik_count);
  const float inv_neik = 1.0f / n_eik;

  float acc_sdf = 0.f, acc_fs = 0.f, acc_eik = 0.f;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < m.N; n += (int64_t)gridDim.x * blockDim.x) {
    float p[3], xn[3];
    load_point(m.x, fr, n, p, m.poison);
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(p[d], g.bmin[d], g.bmax[d]);
    float f[F], dfx[F], dfy[F], dfz[F];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if ((fl.ignore_mask >> l) & 1u) {
#pragma unroll
        for (int i = 0; i < C; ++i) f[l * C + i] = dfx[l * C + i] = dfy[l * C + i] = dfz[l * C + i] = 0.f;
      } else {
        Cell c = level_cell(fl.level[l], xn);
        gather_level<C, true>(fl.level[l], c, f + l * C, dfx + l * C, dfy + l * C, dfz + l * C);
      }
    }
    float J[F];
    const float pred = mlp_eval<F, true>(s, f, J);
    if (m.sdf_out) m.sdf_out[n] = pred;

    const float gt = m.gt_sdf[n];
    // --- regression term (loss.py:594-635): where(valid==1, rho(pred-gt), 0) * weight, mean over N
    float a = 0.f;
    if (m.gt_valid[n]) {
      const float w = m.weights ? m.weights[n] : 1.f;
      const float e = pred - gt;
      if (m.cfg.loss_type == 0) {
        acc_sdf += w * fabsf(e);
        a += m.cfg.weight_sdf * w * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
      } else {
        acc_sdf += w * e * e;
        a += m.cfg.weight_sdf * w * 2.f * e;
      }
    }
    // --- free-space term (loss.py:668-700): max(relu(pred-gt), relu(trunc-pred)) where sign==1
    if (m.cfg.weight_fs != 0.f && m.gt_sign[n] == 1.f) {
      const float up = fmaxf(pred - gt, 0.f), lo = fmaxf(m.cfg.trunc_dist - pred, 0.f);
      acc_fs += fmaxf(up, lo);
      if (up > lo) a += m.cfg.weight_fs;
      else if (lo > up) a -= m.cfg.weight_fs;
    }
    a *= invN * m.cfg.grad_scale;
    // --- eikonal term (loss.py:638-665): mean((|grad_x sdf| - 1)^2) over the filtered samples
    float v[3] = {0.f, 0.f, 0.f};
    if (eik_on && (!eik_filter || fabsf(gt) < m.cfg.eik_trunc_dist)) {
      float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int i = 0; i < C; ++i) {
          sx = fmaf(J[l * C + i], dfx[l * C + i], sx);
          sy = fmaf(J[l * C + i], dfy[l * C + i], sy);
          sz = fmaf(J[l * C + i], dfz[l * C + i], sz);
        }
        gx = fmaf(sx, (float)fl.level[l].X * g.inv_len[0], gx);
        gy = fmaf(sy, (float)fl.level[l].Y * g.inv_len[1], gy);
        gz = fmaf(sz, (float)fl.level[l].Z * g.inv_len[2], gz);
      }
      const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
      const float e = nrm - 1.f;
      acc_eik += e * e;
      if (nrm > 0.f) {
        const float k = m.cfg.weight_eik * m.cfg.grad_scale * 2.f * e * inv_neik / nrm;
        v[0] = k * gx, v[1] = k * gy, v[2] = k * gz;
      }
    }
    if (a != 0.f || v[0] != 0.f || v[1] != 0.f || v[2] != 0.f) {
#pragma unroll
      for (int l = 0; l < L; ++l) {
        if (((fl.ignore_mask >> l) & 1u) || !fl.level[l].grad) continue;
        const miso_level_t& lv = fl.level[l];
        Cell c = level_cell(lv, xn);
        float kx = (float)lv.X * g.inv_len[0], ky = (float)lv.Y * g.inv_len[1], kz = (float)lv.Z * g.inv_len[2];
        scatter_level<C>(lv, c, a, v[0] * kx, v[1] * ky, v[2] * kz, J + l * C);
      }
    }
  }
  float s0 = block_sum(acc_sdf, red);
  float s1 = block_sum(acc_fs, red);
  float s2 = block_sum(acc_eik, red);
  if (threadIdx.x == 0) {
    m.partials[blockIdx.x * 4 + 0] = s0;
    m.partials[blockIdx.x * 4 + 1] = s1;
    m.partials[blockIdx.x * 4 + 2] = s2;
    m.partials[blockIdx.x * 4 + 3] = 0.f;
  }
}

#include "fused_wgrad.cuh"

// ---------------------------------------------------------------------------------------------
// Decoder on the tensor cores (tcgen05, 3xTF32).
//
// CTA = 4 warpgroups x 128 threads; a warpgroup owns one 128-point tile at a time (thread <-> point <->
// TMEM lane).  Per tile: gather + layer 1 on SIMT -> h1 split: hi into TMEM (64 columns), lo into
// shared memory (canonical K-major operand) -> 24 x tcgen05.mma against W2 -> each thread reads its row
// of h2 from TMEM: bias, ReLU, W3 dot, mask -> t = D2 W3^T split the same way -> 24 x tcgen05.mma
// against W2^T -> thread reads g1, applies D1, J = W1^T g1 on SIMT.  TMEM per tile = 64 (A_hi) + 64 (D)
// columns, so four tiles are in flight per SM (16 warps) and the warpgroups hide each other's MMA,
// gather and scatter latency.
// ---------------------------------------------------------------------------------------------
constexpr int kTcWgs = 4;
constexpr int kTcThreads = kTcWgs * 128;
constexpr int kTcABytes = tc::kTileM * tc::kK * 4;  // one A_lo operand: 32 KB
constexpr int kTcMaxSmemPoses = 256;                // keyframe pose table kept in shared memory (12 KB)

// load_point with the pose table in shared memory (same arithmetic as load_point)
__device__ __forceinline__ void load_point_smem(const float* __restrict__ x, const miso_frames_t& fr, int64_t n,
                                                const float* __restrict__ spose, float (&p)[3], float* poison = nullptr) {
  if (!fr.ids || !spose) {
    load_point(x, fr, n, p, poison);
    return;
  }
  const float a = x[3 * n], b = x[3 * n + 1], c = x[3 * n + 2];
  int64_t id = fr.ids[n];
  const bool bad = id < 0 || id >= fr.num_frames;
  if (bad) id = 0;
  const float4* P = reinterpret_cast<const float4*>(spose + id * 12);
  const float4 q0 = P[0], q1 = P[1], q2 = P[2];
  const float R[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
  const float tt[3] = {q2.y, q2.z, q2.w};
#pragma unroll
  for (int j = 0; j < 3; ++j) p[j] = fmaf(c, R[3 * j + 2], fmaf(b, R[3 * j + 1], a * R[3 * j])) + tt[j];
  if (bad) p[0] = p[1] = p[2] = __int_as_float(0x7fc00000);
  if (p[0] != p[0]) flag_bad_frame(poison);
}

template <int F>
struct TcSmem {
  alignas(128) unsigned char w2_hi[tc::kWeightBytes];
  alignas(128) unsigned char w2_lo[tc::kWeightBytes];
  alignas(128) unsigned char w2t_hi[tc::kWeightBytes];
  alignas(128) unsigned char w2t_lo[tc::kWeightBytes];
  alignas(128) unsigned char a_lo[kTcWgs][kTcABytes];
  alignas(128) unsigned char w1e_hi[H * (F + 8) * 4];  // layer 1: canonical [64][KP], row = {W1[n][0..F), b1[n], 0..}
  alignas(128) unsigned char w1e_lo[H * (F + 8) * 4];
  alignas(16) float W1T[F * H];   // [i][k]  (Jacobian product, pairs over k)
  alignas(16) float2 ep[H];       // {b2, W3} per hidden unit
  float b3[4];
  uint64_t bar[kTcWgs];
  uint32_t tmem_base;
  alignas(16) float poses[kTcMaxSmemPoses * 12];   // R (9) + t (3) per keyframe, staged when they fit
};

template <int F>
struct TcTile {
  TcSmem<F>* s;
  uint32_t a_tmem;      // TMEM address of A_hi (lane 0 of this warpgroup's slice)
  uint32_t d_tmem;      // TMEM address of D
  uint32_t lane_bits;   // this warp's lane quarter, already shifted
  uint32_t a_lo_smem;   // shared-memory address of this warpgroup's A_lo operand
  unsigned char* a_lo_row;  // this thread's row inside it
  uint64_t* bar;
  uint32_t parity;
  int wg, wtid;
};

__device__ __forceinline__ void wg_barrier(int wg) {
  asm volatile("bar.sync %0, %1;" ::"r"(wg + 1), "r"(128) : "memory");
}

// one-time CTA setup: TMEM allocation, barriers, weights in canonical layout.  Ends with a CTA barrier.
template <int F>
__device__ __forceinline__ TcTile<F> tc_setup(unsigned char* smem_raw, const miso_decoder_t& dec) {
  TcSmem<F>* s = reinterpret_cast<TcSmem<F>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&s->tmem_base, 512);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kTcWgs; ++i) tc::mbar_init(&s->bar[i], 1);
    tc::fence_mbar_init();
  }
  tc::stage_weights(dec.W2, false, s->w2_hi, s->w2_lo, tid, kTcThreads);   // B[n=j][k] = W2[j][k]
  // backward operand with W3 folded in: B'[n=k][j] = W3[j] * W2[j][k], so A is the exact 0/1 ReLU mask
  for (int i = tid; i < H * H; i += kTcThreads) {
    const int k = i / H, j = i % H;
    float hi, lo;
    tc::tf32_split(dec.W3[j] * dec.W2[j * H + k], hi, lo);
    const uint32_t off = tc::b_offset(k, j);
    *reinterpret_cast<float*>(s->w2t_hi + off) = hi;
    *reinterpret_cast<float*>(s->w2t_lo + off) = lo;
  }
  constexpr int KP = ((F + 1 + 7) / 8) * 8;   // F inputs + the bias column, padded to the tf32 K step
  for (int i = tid; i < H * KP; i += kTcThreads) {
    const int n = i / KP, k = i % KP;
    const float w = k < F ? dec.W1[n * F + k] : (k == F ? dec.b1[n] : 0.f);
    float hi, lo;
    tc::tf32_split(w, hi, lo);
    const uint32_t off = tc::b_offset_k(n, k, KP);
    *reinterpret_cast<float*>(s->w1e_hi + off) = hi;
    *reinterpret_cast<float*>(s->w1e_lo + off) = lo;
  }
  for (int i = tid; i < H * F; i += kTcThreads) s->W1T[(i % F) * H + i / F] = dec.W1[i];
  for (int i = tid; i < H; i += kTcThreads) s->ep[i] = make_float2(dec.b2[i], dec.W3[i]);
  if (tid == 0) s->b3[0] = dec.b3[0];
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  TcTile<F> t;
  t.s = s;
  t.wg = tid >> 7;
  t.wtid = tid & 127;
  const uint32_t tbase = s->tmem_base + (uint32_t)(t.wg * 128);
  t.a_tmem = tbase;
  t.d_tmem = tbase + 64;
  t.lane_bits = (uint32_t)((warp & 3) * 32) << 16;
  t.a_lo_smem = tc::smem_u32(s->a_lo[t.wg]);
  t.a_lo_row = s->a_lo[t.wg] + tc::a_row_offset(t.wtid);
  t.bar = &s->bar[t.wg];
  t.parity = 0;
  return t;
}

template <int F>
__device__ __forceinline__ void tc_teardown(const TcTile<F>& t) {
  tc::fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(t.s->tmem_base, 512);
}

// store 16 hidden values of this thread's row: hi -> TMEM columns [16q,16q+16), lo -> shared memory
__device__ __forceinline__ void tc_store_chunk(uint32_t a_tmem_lane, unsigned char* a_lo_row, int q,
                                               const uint32_t (&hi)[16], const float (&lo)[16]) {
  tc::tmem_st16(a_tmem_lane + q * 16, hi);
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<float4*>(a_lo_row + (q * 4 + c) * tc::kLBO) =
        make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
}

// whole-warpgroup collective: publish A (TMEM + smem), run the 24 MMAs against (b_hi, b_lo), wait for D
template <int F, bool kExactA>
__device__ __forceinline__ void tc_gemm(TcTile<F>& t, uint32_t b_hi, uint32_t b_lo) {
  tc::wait_st();
  if constexpr (!kExactA) tc::fence_proxy_async();  // generic-proxy STS of A_lo -> visible to the async proxy
  tc::fence_before_sync();
  wg_barrier(t.wg);
  if (t.wtid == 0) {
    tc::fence_after_sync();
    if constexpr (kExactA) tc::issue_gemm_exactA(t.d_tmem, t.a_tmem, b_hi, b_lo);
    else tc::issue_gemm_3xtf32_mixed(t.d_tmem, t.a_tmem, t.a_lo_smem, b_hi, b_lo);
    tc::mma_commit(t.bar);
  }
  tc::mbar_wait(t.bar, t.parity);
  t.parity ^= 1;
  tc::fence_after_sync();
}

// MLP forward (+ Jacobian wrt the F inputs) for this thread's point.  Must be called by all 128 threads
// of the warpgroup (inactive points pass zeros).
template <int F, bool kJac>
__device__ __forceinline__ float decoder_tc(TcTile<F>& t, const float (&f)[F], float (&J)[F]) {
  TcSmem<F>* s = t.s;
  const uint32_t a_lane = t.a_tmem + t.lane_bits, d_lane = t.d_tmem + t.lane_bits;
  float2 fp[F / 2];
#pragma unroll
  for (int i = 0; i < F / 2; ++i) fp[i] = make_float2(f[2 * i], f[2 * i + 1]);
  // ---- layer 1 on the tensor core: [f, 1] (K = F+1 padded to 8) x [W1 | b1]^T, 3xTF32 ---------------
  constexpr int KP = ((F + 1 + 7) / 8) * 8;
  {
    uint32_t fh[KP], flo[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      float h = 0.f, l = 0.f;
      if (i < F) tc::tf32_split_fast(f[i], h, l);
      else if (i == F) h = 1.0f;
      fh[i] = __float_as_uint(h);
      flo[i] = __float_as_uint(l);
    }
#pragma unroll
    for (int c = 0; c < KP / 8; ++c) {
      uint32_t a8[8], b8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a8[i] = fh[8 * c + i], b8[i] = flo[8 * c + i];
      tc::tmem_st8(a_lane + 8 * c, a8);
      tc::tmem_st8(a_lane + KP + 8 * c, b8);
    }
    tc::wait_st();
    tc::fence_before_sync();
    wg_barrier(t.wg);
    if (t.wtid == 0) {
      tc::fence_after_sync();
      tc::issue_gemm_3xtf32_smallk<KP>(t.d_tmem, t.a_tmem, t.a_tmem + KP, tc::smem_u32(s->w1e_hi), tc::smem_u32(s->w1e_lo));
      tc::mma_commit(t.bar);
    }
    tc::mbar_wait(t.bar, t.parity);
    t.parity ^= 1;
    tc::fence_after_sync();
  }
  // epilogue 1: h1 = relu(D) -> hi into TMEM, lo into shared memory; sign bits kept for the Jacobian pass
  unsigned m1[2] = {0u, 0u};   // bit (31 - (k & 31)) of word k>>5 = sign of pre-activation k
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t d[16];
    tc::tmem_ld16(d_lane + q * 16, d);
    tc::wait_ld();
    uint32_t hi[16];
    float lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      m1[q >> 1] = __funnelshift_l(d[i], m1[q >> 1], 1);
      float h;
      tc::tf32_split_fast(fmaxf(__uint_as_float(d[i]), 0.f), h, lo[i]);
      hi[i] = __float_as_uint(h);
    }
    tc_store_chunk(a_lane, t.a_lo_row, q, hi, lo);
  }
  tc_gemm<F, false>(t, tc::smem_u32(s->w2_hi), tc::smem_u32(s->w2_lo));
  // ---- layer 2 epilogue + layer 3: sdf = W3 relu(h2 + b2) + b3 ; A <- ReLU mask (exact in tf32) ------
  float pred = s->b3[0];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t d[16];
    tc::tmem_ld16(d_lane + q * 16, d);
    tc::wait_ld();
    uint32_t mk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 e = s->ep[q * 16 + i];
      const float h2 = __uint_as_float(d[i]) + e.x;
      pred = fmaf(e.y, fmaxf(h2, 0.f), pred);
      mk[i] = h2 > 0.f ? 0x3f800000u : 0u;
    }
    if constexpr (kJac) tc::tmem_st16(a_lane + q * 16, mk);
  }
  if constexpr (!kJac) {
    tc::fence_before_sync();  // D is overwritten by the next tile's MMA only after the next barrier
    return pred;
  }
  tc_gemm<F, true>(t, tc::smem_u32(s->w2t_hi), tc::smem_u32(s->w2t_lo));
  // ---- g1 = D1 (W2^T D2 W3^T) ; J = W1^T g1 (SIMT, packed over pairs of hidden units) ----------------
  float2 Jp[F];
#pragma unroll
  for (int i = 0; i < F; ++i) Jp[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t d[16];
    tc::tmem_ld16(d_lane + q * 16, d);
    tc::wait_ld();
    float2 e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // sign bit of hidden unit k sits at bit 31 - (k & 31); set = pre-activation negative = ReLU off
      const int k0 = (q & 1) * 16 + 2 * i;
      const unsigned w = m1[q >> 1];
      e[i] = make_float2(((w >> (31 - k0)) & 1u) ? 0.f : __uint_as_float(d[2 * i]),
                         ((w >> (30 - k0)) & 1u) ? 0.f : __uint_as_float(d[2 * i + 1]));
    }
#pragma unroll
    for (int i = 0; i < F; ++i) {
      const float4* r0 = reinterpret_cast<const float4*>(s->W1T + i * H + q * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 w0 = r0[c];
        Jp[i] = ffma2(make_float2(w0.x, w0.y), e[2 * c], Jp[i]);
        Jp[i] = ffma2(make_float2(w0.z, w0.w), e[2 * c + 1], Jp[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < F; ++i) J[i] = Jp[i].x + Jp[i].y;
  return pred;
}

template <int L, int C>
__global__ void __launch_bounds__(kTcThreads, 1)
    mapping_step_tc_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                           const __grid_constant__ miso_frames_t fr, const __grid_constant__ MapArgs m) {
  constexpr int F = L * C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float red[32];
  TcTile<F> t = tc_setup<F>(smem_raw, dec);

  const FieldGeom g = field_geom(fl);
  const bool eik_on = m.cfg.eik_mode != 0 && m.cfg.weight_eik != 0.f;
  const bool eik_filter = m.cfg.eik_trunc_dist >= 0.f;
  const float n_den = (float)(m.cfg.n_total > 0 ? m.cfg.n_total : m.N);
  const float invN = 1.0f / n_den;
  float n_eik = n_den;
  if (eik_on && eik_filter) n_eik = (float)(*m.eik_count);
  const float inv_neik = 1.0f / n_eik;

  // keyframe pose table -> shared memory (one dependent global round trip less per point)
  const bool poses_in_smem = fr.ids != nullptr && fr.num_frames <= kTcMaxSmemPoses;
  if (poses_in_smem) {
    for (int i = threadIdx.x; i < fr.num_frames * 9; i += kTcThreads) t.s->poses[(i / 9) * 12 + i % 9] = fr.R[i];
    for (int i = threadIdx.x; i < fr.num_frames * 3; i += kTcThreads) t.s->poses[(i / 3) * 12 + 9 + i % 3] = fr.t[i];
  }
  __syncthreads();

  float acc_sdf = 0.f, acc_fs = 0.f, acc_eik = 0.f;
  const int64_t num_tiles = (m.N + 127) / 128;
  for (int64_t tile = (int64_t)blockIdx.x * kTcWgs + t.wg; tile < num_tiles; tile += (int64_t)gridDim.x * kTcWgs) {
    const int64_t n = tile * 128 + t.wtid;
    const bool active = n < m.N;
    float p[3] = {0.f, 0.f, 0.f}, xn[3];
    // loss inputs are fetched up front so their latency hides behind the gather and the decoder
    float gt = 0.f, wgt = 1.f, sgn = 0.f;
    unsigned char vld = 0;
    if (active) {
      load_point_smem(m.x, fr, n, poses_in_smem ? t.s->poses : nullptr, p, m.poison);
      // asm volatile: keeps these loads up here (the compiler would otherwise sink them to their first use
      // after the decoder, exposing a full memory latency in the loss epilogue)
      gt = ldg_early_f32(m.gt_sdf + n);
      vld = (unsigned char)ldg_early_u8(m.gt_valid + n);
      sgn = ldg_early_f32(m.gt_sign + n);
      if (m.weights) wgt = ldg_early_f32(m.weights + n);
      // pull the NEXT tile's per-point inputs towards L1 while this tile is processed
      const int64_t n2 = n + (int64_t)gridDim.x * kTcWgs * 128;
      if (n2 < m.N) {
        prefetch_l1(m.x + 3 * n2);
        prefetch_l1(m.x + 3 * n2 + 2);
        if (fr.ids) prefetch_l1(fr.ids + n2);
        prefetch_l1(m.gt_sdf + n2);
        prefetch_l1(m.gt_sign + n2);
        prefetch_l1(m.gt_valid + n2);
        if (m.weights) prefetch_l1(m.weights + n2);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(p[d], g.bmin[d], g.bmax[d]);
    float f[F], dfx[F], dfy[F], dfz[F];
    CellLite cells[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      cells[l] = make_cell_lite(fl.level[l], xn);
      if (!active || ((fl.ignore_mask >> l) & 1u)) cells[l].valid = 0u;   // contributes zeros, scatters nothing
      CellLite cg = cells[l];
      if (m.dbg & 2) cg.valid = 0u;
      gather_level_lite<C, true>(fl.level[l], cg, f + l * C, dfx + l * C, dfy + l * C, dfz + l * C);
    }
    float J[F];
    const float pred = decoder_tc<F, true>(t, f, J);

    // ---- loss terms + scatter, written without divergent regions (inactive lanes carry zeros) ----------
    if (active && m.sdf_out) m.sdf_out[n] = pred;
    float a = 0.f;
    if (vld) {
      const float e = pred - gt;
      if (m.cfg.loss_type == 0) {
        acc_sdf += wgt * fabsf(e);
        a = m.cfg.weight_sdf * wgt * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
      } else {
        acc_sdf += wgt * e * e;
        a = m.cfg.weight_sdf * wgt * 2.f * e;
      }
    }
    if (m.cfg.weight_fs != 0.f && sgn == 1.f) {
      const float up = fmaxf(pred - gt, 0.f), lo = fmaxf(m.cfg.trunc_dist - pred, 0.f);
      acc_fs += fmaxf(up, lo);
      a += up > lo ? m.cfg.weight_fs : (lo > up ? -m.cfg.weight_fs : 0.f);
    }
    a *= invN * m.cfg.grad_scale;
    float v[3] = {0.f, 0.f, 0.f};
    if (eik_on) {
      float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int i = 0; i < C; ++i) {
          sx = fmaf(J[l * C + i], dfx[l * C + i], sx);
          sy = fmaf(J[l * C + i], dfy[l * C + i], sy);
          sz = fmaf(J[l * C + i], dfz[l * C + i], sz);
        }
        gx = fmaf(sx, (float)fl.level[l].X * g.inv_len[0], gx);
        gy = fmaf(sy, (float)fl.level[l].Y * g.inv_len[1], gy);
        gz = fmaf(sz, (float)fl.level[l].Z * g.inv_len[2], gz);
      }
      const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
      const float e = nrm - 1.f;
      const bool use = active && (!eik_filter || fabsf(gt) < m.cfg.eik_trunc_dist);
      acc_eik += use ? e * e : 0.f;
      const float k = (use && nrm > 0.f) ? m.cfg.weight_eik * m.cfg.grad_scale * 2.f * e * inv_neik / nrm : 0.f;
      v[0] = k * gx, v[1] = k * gy, v[2] = k * gz;
    }
    const unsigned nz = ((a != 0.f || v[0] != 0.f || v[1] != 0.f || v[2] != 0.f) && !(m.dbg & 1)) ? 1u : 0u;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const miso_level_t& lv = fl.level[l];
      float kx = (float)lv.X * g.inv_len[0], ky = (float)lv.Y * g.inv_len[1], kz = (float)lv.Z * g.inv_len[2];
      scatter_level_lite<C>(lv, cells[l], (nz && lv.grad) ? 1u : 0u, a, v[0] * kx, v[1] * ky, v[2] * kz, J + l * C);
    }
  }
  float s0 = block_sum(acc_sdf, red);
  float s1 = block_sum(acc_fs, red);
  float s2 = block_sum(acc_eik, red);
  if (threadIdx.x == 0) {
    m.partials[blockIdx.x * 4 + 0] = s0;
    m.partials[blockIdx.x * 4 + 1] = s1;
    m.partials[blockIdx.x * 4 + 2] = s2;
    m.partials[blockIdx.x * 4 + 3] = 0.f;
  }
  tc_teardown<F>(t);
}

// fused forward (sdf, jac, gradx) with the tensor-core decoder
template <int L, int C, bool kJac>
__global__ void __launch_bounds__(kTcThreads, 1)
    sdf_forward_tc_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                          const __grid_constant__ miso_frames_t fr, const float* __restrict__ x, int64_t N,
                          float* __restrict__ sdf, float* __restrict__ jac, float* __restrict__ gradx,
                          float* __restrict__ xw) {
  constexpr int F = L * C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcTile<F> t = tc_setup<F>(smem_raw, dec);
  const FieldGeom g = field_geom(fl);
  const int64_t num_tiles = (N + 127) / 128;
  for (int64_t tile = (int64_t)blockIdx.x * kTcWgs + t.wg; tile < num_tiles; tile += (int64_t)gridDim.x * kTcWgs) {
    const int64_t n = tile * 128 + t.wtid;
    const bool active = n < N;
    float p[3] = {0.f, 0.f, 0.f}, xn[3];
    if (active) load_point(x, fr, n, p);
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(p[d], g.bmin[d], g.bmax[d]);
    float f[F], dfx[F], dfy[F], dfz[F];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if (!active || ((fl.ignore_mask >> l) & 1u)) {
#pragma unroll
        for (int i = 0; i < C; ++i) f[l * C + i] = dfx[l * C + i] = dfy[l * C + i] = dfz[l * C + i] = 0.f;
      } else {
        Cell c = level_cell(fl.level[l], xn);
        gather_level<C, kJac>(fl.level[l], c, f + l * C, dfx + l * C, dfy + l * C, dfz + l * C);
        if constexpr (kJac) {
          float kx = (float)fl.level[l].X * g.inv_len[0], ky = (float)fl.level[l].Y * g.inv_len[1],
                kz = (float)fl.level[l].Z * g.inv_len[2];
#pragma unroll
          for (int i = 0; i < C; ++i) dfx[l * C + i] *= kx, dfy[l * C + i] *= ky, dfz[l * C + i] *= kz;
        }
      }
    }
    float J[F];
    const float val = decoder_tc<F, kJac>(t, f, J);
    if (!active) continue;
    sdf[n] = val;
    if (xw) xw[3 * n] = p[0], xw[3 * n + 1] = p[1], xw[3 * n + 2] = p[2];
    if constexpr (kJac) {
      if (jac) {
#pragma unroll
        for (int i = 0; i < F; i += 4) *reinterpret_cast<float4*>(jac + n * F + i) = make_float4(J[i], J[i + 1], J[i + 2], J[i + 3]);
      }
      if (gradx) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int i = 0; i < F; ++i) gx = fmaf(J[i], dfx[i], gx), gy = fmaf(J[i], dfy[i], gy), gz = fmaf(J[i], dfz[i], gz);
        gradx[3 * n] = gx, gradx[3 * n + 1] = gy, gradx[3 * n + 2] = gz;
      }
    }
  }
  tc_teardown<F>(t);
}

}  // namespace miso
#include "fused_tc2.cuh"
namespace miso {

__global__ void mapping_finalize_kernel(const float* __restrict__ partials, int nblocks, int64_t N,
                                        miso_mapping_cfg_t cfg, const int32_t* eik_count, float* __restrict__ out,
                                        float* __restrict__ poison) {
  // fixed-order reduction => deterministic loss values
  __shared__ double red[32];
  double s[3] = {0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
    for (int k = 0; k < 3; ++k) s[k] += (double)partials[b * 4 + k];
  double t[3];
  for (int k = 0; k < 3; ++k) t[k] = block_sum(s[k], red);
  if (threadIdx.x == 0) {
    const bool eik_on = cfg.eik_mode != 0 && cfg.weight_eik != 0.f;
    const double nden = (double)(cfg.n_total > 0 ? cfg.n_total : N);
    double neik = nden;
    if (eik_on && cfg.eik_trunc_dist >= 0.f) neik = (double)(*eik_count);
    float l_sdf = (float)(t[0] / nden);
    float l_fs = cfg.weight_fs != 0.f ? (float)(t[1] / nden) : 0.f;
    float l_eik = eik_on ? (float)(t[2] / neik) : 0.f;
    out[0] = l_sdf;
    out[1] = l_fs;
    out[2] = l_eik;
    out[3] = cfg.weight_sdf * l_sdf + cfg.weight_fs * l_fs + (eik_on ? cfg.weight_eik * l_eik : 0.f);
    if (*poison != *poison) {   // some sample had no keyframe pose: the step is invalid, say so loudly
      out[0] = out[1] = out[2] = out[3] = *poison;
      *poison = 0.f;
    }
  }
}

__global__ void mapping_count_kernel(const float* __restrict__ gt, int64_t N, float trunc, int32_t* out) {
  int cnt = 0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x)
    cnt += fabsf(gt[n]) < trunc ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
static int validate_field(const miso_field_t* f, bool need_grad_layout) {
  MISO_REQUIRE(f, "field: null");
  MISO_REQUIRE(f->num_levels >= 1 && f->num_levels <= MISO_MAX_LEVELS, "field: num_levels %d not in [1,%d]",
               f->num_levels, MISO_MAX_LEVELS);
  const int C = f->level[0].C;
  for (int l = 0; l < f->num_levels; ++l) {
    const miso_level_t& lv = f->level[l];
    MISO_REQUIRE(lv.feat, "field: level %d has null features", l);
    MISO_REQUIRE(lv.C == C, "field: fused kernels need the same channel count on every level");
    MISO_REQUIRE(lv.C % 4 == 0 && lv.C >= 4 && lv.C <= 16, "field: fused kernels need C in {4,8,12,16}, got %d", lv.C);
    MISO_REQUIRE(lv.sC == 1, "field: level %d is not channels-last (stride_C=%lld); convert with channels_last_3d", l,
                 (long long)lv.sC);
    MISO_REQUIRE(lv.sX % 4 == 0 && lv.sY % 4 == 0 && lv.sZ % 4 == 0 && ((uintptr_t)lv.feat) % 16 == 0,
                 "field: level %d not 16-byte aligned", l);
    MISO_REQUIRE(!lv.grad || ((uintptr_t)lv.grad) % 16 == 0, "field: level %d grad not 16-byte aligned", l);
    MISO_REQUIRE(lv.X > 0 && lv.Y > 0 && lv.Z > 0, "field: level %d has an empty dimension", l);
  }
  for (int d = 0; d < 3; ++d)
    MISO_REQUIRE(f->bound[2 * d + 1] > f->bound[2 * d], "field: empty bound on axis %d", d);
  (void)need_grad_layout;
  return MISO_OK;
}

static int validate_decoder(const miso_decoder_t* d, int F) {
  MISO_REQUIRE(d && d->W1 && d->b1 && d->W2 && d->b2 && d->W3 && d->b3, "decoder: null weights (bias=True required)");
  MISO_REQUIRE(d->hidden_dim == H, "decoder: fused path supports hidden_dim=64, got %d", d->hidden_dim);
  MISO_REQUIRE(d->in_dim == F, "decoder: in_dim %d != levels*channels %d", d->in_dim, F);
  return MISO_OK;
}

template <typename K>
static int blocks_for(K kernel, size_t smem, int64_t N) {
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem);
  if (occ < 1) occ = 1;
  return grid_for(N, kThreads, sm_count() * occ);
}

#define MISO_DISPATCH_LC(L_, C_, ...)                                              \
  do {                                                                             \
    const int key_ = (L_)*100 + (C_);                                              \
    switch (key_) {                                                                \
      case 104: { constexpr int L = 1, C = 4; __VA_ARGS__; } break;                \
      case 204: { constexpr int L = 2, C = 4; __VA_ARGS__; } break;                \
      case 304: { constexpr int L = 3, C = 4; __VA_ARGS__; } break;                \
      case 404: { constexpr int L = 4, C = 4; __VA_ARGS__; } break;                \
      case 108: { constexpr int L = 1, C = 8; __VA_ARGS__; } break;                \
      case 208: { constexpr int L = 2, C = 8; __VA_ARGS__; } break;                \
      case 116: { constexpr int L = 1, C = 16; __VA_ARGS__; } break;               \
      default:                                                                     \
        miso::set_error("fused path: unsupported (levels=%d, channels=%d)", (L_), (C_)); \
        return MISO_ERR_UNSUPPORTED;                                               \
    }                                                                              \
  } while (0)

// feature-only kernel has no decoder, allow more shapes
#define MISO_DISPATCH_LC_FEAT(L_, C_, ...)                                         \
  do {                                                                             \
    const int key_ = (L_)*100 + (C_);                                              \
    switch (key_) {                                                                \
      case 104: { constexpr int L = 1, C = 4; __VA_ARGS__; } break;                \
      case 204: { constexpr int L = 2, C = 4; __VA_ARGS__; } break;                \
      case 304: { constexpr int L = 3, C = 4; __VA_ARGS__; } break;                \
      case 404: { constexpr int L = 4, C = 4; __VA_ARGS__; } break;                \
      case 108: { constexpr int L = 1, C = 8; __VA_ARGS__; } break;                \
      case 208: { constexpr int L = 2, C = 8; __VA_ARGS__; } break;                \
      case 308: { constexpr int L = 3, C = 8; __VA_ARGS__; } break;                \
      case 408: { constexpr int L = 4, C = 8; __VA_ARGS__; } break;                \
      case 116: { constexpr int L = 1, C = 16; __VA_ARGS__; } break;               \
      case 216: { constexpr int L = 2, C = 16; __VA_ARGS__; } break;               \
      case 316: { constexpr int L = 3, C = 16; __VA_ARGS__; } break;               \
      case 416: { constexpr int L = 4, C = 16; __VA_ARGS__; } break;               \
      default:                                                                     \
        miso::set_error("features: unsupported (levels=%d, channels=%d)", (L_), (C_)); \
        return MISO_ERR_UNSUPPORTED;                                               \
    }                                                                              \
  } while (0)

// Kernel-variant switches.  Defaults come from the environment once (MISO_MLP=simt, MISO_TC=1, MISO_TC2_GROUPS=3,
// MISO_PAIR=0, MISO_FWD_TC2=1, MISO_DBG=<bits>); miso_set_tuning() overrides them at run time so the parity tests can
// walk every variant inside one process.
struct Tuning {
  int mlp_tc;       // 1: tcgen05 decoder (default), 0: FP32 SIMT decoder
  int tc2_groups;   // 0: one-thread-per-point tensor-core kernel, 3 | 4: two-threads-per-point kernel, tiles in flight
  int pair;         // 1: lane-paired gather / scatter (default)
  int fwd_tc2;      // 1: two-thread kernel also for value-only forwards
  int dbg;          // ablation bits (profiling only): 1 = no reductions, 2 = no corner loads
  int force_int64;  // 1: treat every field as not 32-bit addressable (exercises the SIMT / 64-bit-offset route)
};
static Tuning& tuning() {
  static Tuning t = [] {
    Tuning d;
    const char* e = getenv("MISO_MLP");
    d.mlp_tc = (e && (e[0] == 's' || e[0] == 'S')) ? 0 : 1;
    e = getenv("MISO_TC");
    const char* g = getenv("MISO_TC2_GROUPS");
    d.tc2_groups = (e && e[0] == '1') ? 0 : ((g && g[0] == '3') ? 3 : 4);
    e = getenv("MISO_PAIR");
    d.pair = (e && e[0] == '0') ? 0 : 1;
    e = getenv("MISO_FWD_TC2");
    d.fwd_tc2 = (e && e[0] == '1') ? 1 : 0;
    e = getenv("MISO_DBG");
    d.dbg = e ? atoi(e) : 0;
    d.force_int64 = 0;
    return d;
  }();
  return t;
}
// MISO_MLP=simt forces the FP32 SIMT decoder; default is the tcgen05 (3xTF32) decoder
static bool fits_int32(const miso_field_t* f) {
  if (tuning().force_int64) return false;
  for (int l = 0; l < f->num_levels; ++l) {
    const miso_level_t& lv = f->level[l];
    const int64_t span = (int64_t)(lv.Z + 2) * lv.sZ + (int64_t)(lv.Y + 2) * lv.sY + (int64_t)(lv.X + 2) * lv.sX;
    if (span >= (int64_t)0x7fffffff) return false;
  }
  return true;
}

static bool use_tensor_cores() { return tuning().mlp_tc == 1; }
static int tc2_groups() { return tuning().tc2_groups; }

#define MISO_DISPATCH_LC_TC2(L_, C_, ...)                                          \
  do {                                                                             \
    const int key_ = (L_)*100 + (C_);                                              \
    switch (key_) {                                                                \
      case 204: { constexpr int L = 2, C = 4; __VA_ARGS__; } break;                \
      case 404: { constexpr int L = 4, C = 4; __VA_ARGS__; } break;                \
      case 108: { constexpr int L = 1, C = 8; __VA_ARGS__; } break;                \
      case 208: { constexpr int L = 2, C = 8; __VA_ARGS__; } break;                \
      case 116: { constexpr int L = 1, C = 16; __VA_ARGS__; } break;               \
      default: break;                                                              \
    }                                                                              \
  } while (0)

static bool tc2_paired() { return tuning().pair == 1; }

template <int L, int C, int G>
static int launch_tc2(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t& fr, const MapArgs& m,
                      cudaStream_t s) {
  constexpr size_t smem = sizeof(Tc2Smem<L * C, G>) + 128;
  auto k = m.a_ext ? mapping_step_tc2_kernel<L, C, G, true, 3>
                   : (m.cfg.n_device ? mapping_step_tc2_kernel<L, C, G, true, 4>
                                     : (tc2_paired() ? mapping_step_tc2_kernel<L, C, G, true, 0>
                                                     : mapping_step_tc2_kernel<L, C, G, false, 0>));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nblocks = grid_for((m.N + G * 128 - 1) / (G * 128), 1, sm_count());
  k<<<nblocks, G * 256, smem, s>>>(*field, *dec, fr, m);
  return nblocks;
}

static void fill_scales(const miso_field_t* field, MapArgs& m) {
  for (int d = 0; d < 3; ++d) {
    m.inv_len[d] = 1.0f / (field->bound[2 * d + 1] - field->bound[2 * d]);
    for (int l = 0; l < field->num_levels; ++l) {
      const miso_level_t& lv = field->level[l];
      m.lvl_scale[l][d] = (float)(d == 0 ? lv.X : (d == 1 ? lv.Y : lv.Z)) * m.inv_len[d];
    }
  }
}

// forward modes of the two-thread kernel (miso_sdf_forward): kMode 1 = with Jacobian / grad_x, 2 = values only
template <int L, int C, int kMode>
static int launch_tc2_forward(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t& fr,
                              const MapArgs& m, cudaStream_t s) {
  constexpr int G = 4;
  constexpr size_t smem = sizeof(Tc2Smem<L * C, G>) + 128;
  auto k = mapping_step_tc2_kernel<L, C, G, true, kMode>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nblocks = grid_for((m.N + G * 128 - 1) / (G * 128), 1, sm_count());
  k<<<nblocks, G * 256, smem, s>>>(*field, *dec, fr, m);
  return nblocks;
}

static miso_frames_t frames_or_none(const miso_frames_t* fr) {
  miso_frames_t z;
  z.ids = nullptr, z.R = nullptr, z.t = nullptr, z.num_frames = 0;
  return fr ? *fr : z;
}

static int validate_frames(const miso_frames_t* fr) {
  if (!fr || !fr->ids) return MISO_OK;
  MISO_REQUIRE(fr->R && fr->t && fr->num_frames > 0, "frames: ids given without poses");
  return MISO_OK;
}

}  // namespace miso

using namespace miso;

extern "C" int miso_field_features(const miso_field_t* field, const float* x, int64_t N, float* feats,
                                   miso_stream_t stream) {
  if (int e = validate_field(field, false)) return e;
  MISO_REQUIRE(N >= 0 && (N == 0 || (x && feats)), "field_features: null x/feats");
  if (N == 0) return MISO_OK;
  MISO_REQUIRE(((uintptr_t)feats) % 16 == 0, "field_features: feats not 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  MISO_DISPATCH_LC_FEAT(field->num_levels, field->level[0].C, {
    auto k = field_features_kernel<L, C>;
    k<<<blocks_for(k, 0, N), kThreads, 0, s>>>(*field, x, N, feats);
  });
  return check_launch("field_features");
}

extern "C" int miso_sdf_forward(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                                const float* x, int64_t N, float* sdf, float* jac, float* gradx, float* xw,
                                miso_stream_t stream) {
  if (int e = validate_field(field, false)) return e;
  if (int e = validate_decoder(dec, field->num_levels * field->level[0].C)) return e;
  if (int e = validate_frames(frames)) return e;
  MISO_REQUIRE(N >= 0 && (N == 0 || (x && sdf)), "sdf_forward: null x/sdf");
  if (N == 0) return MISO_OK;
  MISO_REQUIRE(!jac || ((uintptr_t)jac) % 16 == 0, "sdf_forward: jac not 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const miso_frames_t fr = frames_or_none(frames);
  const bool want_jac = jac || gradx;
  const int F_ = field->num_levels * field->level[0].C;
  // with Jacobian / grad_x outputs the two-threads-per-point kernel is faster (0.17 vs 0.20 ms per 2^20 points);
  // values only: the one-thread kernel wins on lattice-ordered dense queries (0.73 vs 0.80 ms per 2^23), so mode 2
  // is only taken when MISO_FWD_TC2=1 asks for it
  const bool fwd_tc2 = tuning().fwd_tc2 == 1;
  if (use_tensor_cores() && N >= 4096 && tc2_groups() != 0 && F_ % 8 == 0 && fits_int32(field) &&
      N < ((int64_t)1 << 31) - ((int64_t)1 << 26) && (want_jac || fwd_tc2)) {
    MapArgs m;
    memset(&m, 0, sizeof(m));
    m.x = x, m.N = N, m.sdf_out = sdf, m.jac = jac, m.gradx = gradx, m.xw = xw;
    fill_scales(field, m);
    int nblocks = 0;
    MISO_DISPATCH_LC_TC2(field->num_levels, field->level[0].C, {
      nblocks = want_jac ? launch_tc2_forward<L, C, 1>(field, dec, fr, m, s) : launch_tc2_forward<L, C, 2>(field, dec, fr, m, s);
    });
    MISO_REQUIRE(nblocks > 0, "sdf_forward(tc2): unsupported (levels=%d, channels=%d)", field->num_levels, field->level[0].C);
    return check_launch("sdf_forward(tc2)");
  }
  if (use_tensor_cores() && N >= 4096) {
    MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
      constexpr size_t smem = sizeof(TcSmem<L * C>) + 128;
      const int nb = grid_for((N + kTcThreads - 1) / kTcThreads, 1, sm_count());
      if (want_jac) {
        auto k = sdf_forward_tc_kernel<L, C, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<nb, kTcThreads, smem, s>>>(*field, *dec, fr, x, N, sdf, jac, gradx, xw);
      } else {
        auto k = sdf_forward_tc_kernel<L, C, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<nb, kTcThreads, smem, s>>>(*field, *dec, fr, x, N, sdf, jac, gradx, xw);
      }
    });
    return check_launch("sdf_forward(tc)");
  }
  MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
    constexpr size_t smem = sizeof(DecoderSmem<L * C>);
    if (want_jac) {
      auto k = sdf_forward_kernel<L, C, true>;
      k<<<blocks_for(k, smem, N), kThreads, smem, s>>>(*field, *dec, fr, x, N, sdf, jac, gradx, xw);
    } else {
      auto k = sdf_forward_kernel<L, C, false>;
      k<<<blocks_for(k, smem, N), kThreads, smem, s>>>(*field, *dec, fr, x, N, sdf, jac, gradx, xw);
    }
  });
  return check_launch("sdf_forward");
}

extern "C" int miso_track_normal_equations(const miso_field_t* field, const miso_decoder_t* dec, const float* x_frame,
                                           const float* gt_sdf, int64_t N, const float* Rt, int32_t loss_type,
                                           float gm_scale, float trunc_dist, double* out, miso_stream_t stream) {
  if (int e = validate_field(field, false)) return e;
  if (int e = validate_decoder(dec, field->num_levels * field->level[0].C)) return e;
  MISO_REQUIRE(out && Rt && N >= 0 && (N == 0 || (x_frame && gt_sdf)), "track_normal_equations: null argument");
  MISO_REQUIRE(loss_type == 0 || loss_type == 1, "track_normal_equations: loss_type must be 0 (L2) or 1 (GM)");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double) * 45, s);
  if (N == 0) return check_launch("track_normal_equations(memset)");
  MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
    constexpr size_t smem = sizeof(DecoderSmem<L * C>);
    auto k = track_normal_equations_kernel<L, C>;
    k<<<grid_for(N, kThreads, sm_count()), kThreads, smem, s>>>(*field, *dec, x_frame, gt_sdf, N, Rt, loss_type, gm_scale,
                                                                 trunc_dist, out);
  });
  return check_launch("track_normal_equations");
}

extern "C" int miso_sdf_backward(const miso_field_t* field, const float* xw, int64_t N, const float* jac,
                                 const float* a, const float* v, float* hv, miso_stream_t stream) {
  if (int e = validate_field(field, true)) return e;
  MISO_REQUIRE(N >= 0 && (N == 0 || (xw && jac)), "sdf_backward: null xw/jac");
  if (N == 0 || (!a && !v)) return MISO_OK;
  MISO_REQUIRE(((uintptr_t)jac) % 16 == 0, "sdf_backward: jac not 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
    auto k = sdf_backward_kernel<L, C>;
    k<<<blocks_for(k, 0, N), kThreads, 0, s>>>(*field, xw, N, jac, a, v, hv);
  });
  return check_launch("sdf_backward");
}

extern "C" int miso_set_tuning(const char* key, int32_t value) {
  MISO_REQUIRE(key, "set_tuning: null key");
  Tuning& t = tuning();
  if (!strcmp(key, "mlp_tc")) t.mlp_tc = value ? 1 : 0;
  else if (!strcmp(key, "tc2_groups")) {
    MISO_REQUIRE(value == 0 || value == 3 || value == 4, "set_tuning: tc2_groups must be 0, 3 or 4");
    t.tc2_groups = value;
  } else if (!strcmp(key, "pair")) t.pair = value ? 1 : 0;
  else if (!strcmp(key, "fwd_tc2")) t.fwd_tc2 = value ? 1 : 0;
  else if (!strcmp(key, "dbg")) t.dbg = value;
  else if (!strcmp(key, "force_int64")) t.force_int64 = value ? 1 : 0;
  else MISO_REQUIRE(false, "set_tuning: unknown key '%s'", key);
  return MISO_OK;
}
extern "C" int miso_get_tuning(const char* key) {
  if (!key) return -1;
  const Tuning& t = tuning();
  if (!strcmp(key, "mlp_tc")) return t.mlp_tc;
  if (!strcmp(key, "tc2_groups")) return t.tc2_groups;
  if (!strcmp(key, "pair")) return t.pair;
  if (!strcmp(key, "fwd_tc2")) return t.fwd_tc2;
  if (!strcmp(key, "dbg")) return t.dbg;
  if (!strcmp(key, "force_int64")) return t.force_int64;
  return -1;
}

static int64_t partial_floats() { return (int64_t)sm_count() * 8 * 4; }
extern "C" int64_t miso_mapping_workspace_floats(void) { return partial_floats() + 4; }   // + the poison word

extern "C" int miso_mapping_count(const float* gt_sdf, int64_t N, float eik_trunc_dist, int32_t* eik_count,
                                  miso_stream_t stream) {
  MISO_REQUIRE(eik_count && (N == 0 || gt_sdf), "mapping_count: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(eik_count, 0, sizeof(int32_t), s);
  if (N > 0) mapping_count_kernel<<<grid_for(N, kThreads, sm_count() * 8), kThreads, 0, s>>>(gt_sdf, N, eik_trunc_dist, eik_count);
  return check_launch("mapping_count");
}

// finite-difference eikonal (diff.py:18-26 + loss.py:638-665) from the six displaced values f[k N + i]:
//   g_d = (f[2d] - f[2d+1]) / (2 eps),  term = (|g| - 1)^2 over the samples that pass the |gt| < trunc filter,
//   a_ext[(2d) N + i] = +c g_d / (2 eps), a_ext[(2d+1) N + i] = -c g_d / (2 eps),  c = w_eik * scale * 2 (|g|-1) / (n_eik |g|)
// i.e. d(w_eik * eik)/d f at each displaced evaluation -- the cotangents of the six first-order backward passes.
__global__ void __launch_bounds__(kThreads)
    fd_eikonal_kernel(const float* __restrict__ f, const float* __restrict__ gt_sdf, int64_t N, float eps,
                      miso_mapping_cfg_t cfg, const int32_t* eik_count, float* __restrict__ a_ext,
                      float* __restrict__ partials) {
  __shared__ float red[32];
  const bool filter = cfg.eik_trunc_dist >= 0.f;
  const float n_den = (float)(cfg.n_total > 0 ? cfg.n_total : N);
  const float n_eik = filter ? (float)(*eik_count) : n_den;
  const float scale = cfg.weight_eik * cfg.grad_scale * 2.f * (1.0f / n_eik);
  const float inv2e = 1.0f / (eps * 2.0f);
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] = (f[(2 * d) * N + i] - f[(2 * d + 1) * N + i]) / (eps * 2.0f);
    const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float e = nrm - 1.f;
    const bool use = !filter || fabsf(gt_sdf[i]) < cfg.eik_trunc_dist;
    acc += use ? e * e : 0.f;
    const float c = (use && nrm > 0.f) ? scale * e / nrm : 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float a = c * g[d] * inv2e;
      a_ext[(2 * d) * N + i] = a;
      a_ext[(2 * d + 1) * N + i] = -a;
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void fd_eikonal_finalize_kernel(const float* __restrict__ partials, int nblocks, int64_t N,
                                           miso_mapping_cfg_t cfg, const int32_t* eik_count, float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += (double)partials[b];
  const double t = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double nden = (double)(cfg.n_total > 0 ? cfg.n_total : N);
    const double neik = cfg.eik_trunc_dist >= 0.f ? (double)(*eik_count) : nden;
    const float l_eik = (float)(t / neik);
    out[2] = l_eik;                         // NaN-poisoned steps stay NaN: NaN + x = NaN
    out[3] = out[3] + cfg.weight_eik * l_eik;
  }
}

static int mapping_step_impl(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                             const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                             const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                             const int32_t* eik_count, float* partials, float* loss_out, float* sdf_out,
                             miso_stream_t stream, const float* a_ext, int fd_n, float fd_eps);

extern "C" int miso_mapping_step(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                                 const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                                 const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                                 const int32_t* eik_count, float* partials, float* loss_out, float* sdf_out,
                                 miso_stream_t stream) {
  return mapping_step_impl(field, dec, frames, x, N, gt_sdf, gt_valid, gt_sign, weights, cfg, eik_count, partials,
                           loss_out, sdf_out, stream, nullptr, 0, 0.f);
}

// Decoder-parameter gradients of the same step (decoder.fix: False): second pass over the batch, see fused_wgrad.cuh.
static int wgrad_blocks() { return sm_count() * 2; }
extern "C" int64_t miso_mapping_wgrad_workspace_floats(void) {
  return (int64_t)wgrad_blocks() * wgrad::row_stride<16>();   // widest supported decoder input
}
extern "C" int miso_mapping_step_wgrad(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                                       const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                                       const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                                       const int32_t* eik_count, const miso_decoder_grad_t* grad, float* workspace,
                                       miso_stream_t stream) {
  if (int e = validate_field(field, false)) return e;
  if (int e = validate_decoder(dec, field->num_levels * field->level[0].C)) return e;
  if (int e = validate_frames(frames)) return e;
  MISO_REQUIRE(cfg && grad && workspace, "mapping_step_wgrad: null cfg/grad/workspace");
  MISO_REQUIRE(N > 0 && x && gt_sdf && gt_valid && gt_sign, "mapping_step_wgrad: null inputs or N == 0");
  MISO_REQUIRE(cfg->loss_type == 0 || cfg->loss_type == 1, "mapping_step_wgrad: loss_type must be 0 (L1) or 1 (L2)");
  MISO_REQUIRE(cfg->eik_mode == 0 || cfg->eik_mode == 1, "mapping_step_wgrad: eik_mode must be 0 or 1");
  MISO_REQUIRE(!cfg->n_device, "mapping_step_wgrad: device-side sample counts are not supported");
  const bool eik_on = cfg->eik_mode != 0 && cfg->weight_eik != 0.f;
  MISO_REQUIRE(!(eik_on && cfg->eik_trunc_dist >= 0.f) || eik_count, "mapping_step_wgrad: eik filter needs eik_count");
  cudaStream_t s = (cudaStream_t)stream;
  const miso_frames_t fr = frames_or_none(frames);
  MapArgs m;
  memset(&m, 0, sizeof(m));
  m.x = x, m.N = N, m.gt_sdf = gt_sdf, m.gt_valid = gt_valid, m.gt_sign = gt_sign, m.weights = weights;
  m.cfg = *cfg, m.eik_count = eik_count;
  const int nblocks = (int)std::min<int64_t>(wgrad_blocks(), (N + wgrad::kT - 1) / wgrad::kT);
  MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
    constexpr size_t smem = sizeof(wgrad::Smem<L * C>);
    auto k = mapping_wgrad_kernel<L, C>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<nblocks, kThreads, smem, s>>>(*field, *dec, fr, m, workspace);
    if (int e = check_launch("mapping_step_wgrad")) return e;
    constexpr int np = wgrad::param_count<L * C>();
    mapping_wgrad_finalize_kernel<L * C><<<(np + kThreads - 1) / kThreads, kThreads, 0, s>>>(workspace, nblocks, *grad);
  });
  return check_launch("mapping_step_wgrad(finalize)");
}

extern "C" int miso_mapping_step_fd(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                                    const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                                    const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                                    const int32_t* eik_count, float* partials, float* loss_out, float* sdf_out,
                                    float finite_diff_eps, float* fd_workspace, miso_stream_t stream) {
  if (int e = validate_field(field, true)) return e;
  if (int e = validate_decoder(dec, field->num_levels * field->level[0].C)) return e;
  MISO_REQUIRE(cfg && fd_workspace && finite_diff_eps > 0.f, "mapping_step_fd: null cfg/workspace or eps <= 0");
  MISO_REQUIRE(cfg->weight_eik != 0.f, "mapping_step_fd: weight_eik == 0, call miso_mapping_step");
  MISO_REQUIRE(!(cfg->eik_trunc_dist >= 0.f) || eik_count, "mapping_step_fd: eik filter needs eik_count");
  const int F_ = field->num_levels * field->level[0].C;
  const bool tc2_ok = use_tensor_cores() && fits_int32(field) && tc2_groups() != 0 && F_ % 8 == 0 &&
                      6 * N < ((int64_t)1 << 31) - ((int64_t)1 << 26);
  if (!tc2_ok) {
    set_error("mapping_step_fd: needs the two-threads-per-point tensor-core kernel (levels*channels %% 8 == 0, "
              "32-bit addressable grids, 6 N < 2^31)");
    return MISO_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* f6 = fd_workspace;            // (6, N) displaced values
  float* a6 = fd_workspace + 6 * N;    // (6, N) their cotangents
  // (1) sdf + free-space terms and their scatter on the N samples, eikonal off
  miso_mapping_cfg_t base = *cfg;
  base.eik_mode = 0;
  if (int e = mapping_step_impl(field, dec, frames, x, N, gt_sdf, gt_valid, gt_sign, weights, &base, eik_count, partials,
                                loss_out, sdf_out, stream, nullptr, 0, 0.f))
    return e;
  // (2) six displaced forward evaluations in one launch (values only)
  const miso_frames_t fr = frames_or_none(frames);
  {
    MapArgs m;
    memset(&m, 0, sizeof(m));
    m.x = x, m.N = 6 * N, m.sdf_out = f6, m.fd_n = (int)N, m.fd_eps = finite_diff_eps;
    fill_scales(field, m);
    int nb = 0;
    MISO_DISPATCH_LC_TC2(field->num_levels, field->level[0].C, { nb = launch_tc2_forward<L, C, 2>(field, dec, fr, m, s); });
    MISO_REQUIRE(nb > 0, "mapping_step_fd: unsupported (levels=%d, channels=%d)", field->num_levels, field->level[0].C);
    if (int e = check_launch("mapping_step_fd(forward)")) return e;
  }
  // (3) eikonal term + cotangents of the six evaluations
  const int nbk = grid_for(N, kThreads, (int)(partial_floats() / 4));
  fd_eikonal_kernel<<<nbk, kThreads, 0, s>>>(f6, gt_sdf, N, finite_diff_eps, *cfg, eik_count, a6, partials);
  fd_eikonal_finalize_kernel<<<1, kThreads, 0, s>>>(partials, nbk, N, *cfg, eik_count, loss_out);
  if (int e = check_launch("mapping_step_fd(eikonal)")) return e;
  // (4) six first-order backward passes in one launch: scatter a_ext * w_c * J at the displaced points
  {
    float scratch_loss_unused = 0.f;
    (void)scratch_loss_unused;
    if (int e = mapping_step_impl(field, dec, frames, x, 6 * N, nullptr, nullptr, nullptr, nullptr, &base, eik_count,
                                  partials, nullptr, nullptr, stream, a6, (int)N, finite_diff_eps))
      return e;
  }
  return MISO_OK;
}

static int mapping_step_impl(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                             const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                             const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                             const int32_t* eik_count, float* partials, float* loss_out, float* sdf_out,
                             miso_stream_t stream, const float* a_ext, int fd_n, float fd_eps) {
  if (int e = validate_field(field, true)) return e;
  if (int e = validate_decoder(dec, field->num_levels * field->level[0].C)) return e;
  if (int e = validate_frames(frames)) return e;
  MISO_REQUIRE(cfg && partials && (loss_out || a_ext), "mapping_step: null cfg/partials/loss_out");
  MISO_REQUIRE(N > 0 && x && (a_ext || (gt_sdf && gt_valid && gt_sign)), "mapping_step: null inputs or N == 0");
  MISO_REQUIRE(cfg->loss_type == 0 || cfg->loss_type == 1, "mapping_step: loss_type must be 0 (L1) or 1 (L2)");
  MISO_REQUIRE(cfg->eik_mode == 0 || cfg->eik_mode == 1, "mapping_step: eik_mode must be 0 or 1");
  const bool eik_on = cfg->eik_mode != 0 && cfg->weight_eik != 0.f;
  MISO_REQUIRE(!(eik_on && cfg->eik_trunc_dist >= 0.f) || eik_count, "mapping_step: eik filter needs eik_count");
  cudaStream_t s = (cudaStream_t)stream;
  const miso_frames_t fr = frames_or_none(frames);
  MapArgs m;
  m.x = x, m.N = N, m.gt_sdf = gt_sdf, m.gt_valid = gt_valid, m.gt_sign = gt_sign, m.weights = weights;
  m.cfg = *cfg, m.eik_count = eik_count, m.partials = partials, m.sdf_out = sdf_out;
  m.poison = a_ext ? nullptr : partials + partial_floats();   // the step's own pass has already judged the keyframe ids
  m.jac = nullptr, m.gradx = nullptr, m.xw = nullptr;
  m.a_ext = a_ext, m.fd_n = fd_n, m.fd_eps = fd_eps;
  fill_scales(field, m);
  m.dbg = tuning().dbg;
  int nblocks = 0;
  const int F_ = field->num_levels * field->level[0].C;
  const bool tc2_route = use_tensor_cores() && fits_int32(field) && tc2_groups() != 0 && F_ % 8 == 0 &&
                         N < ((int64_t)1 << 31) - ((int64_t)1 << 26);
  MISO_REQUIRE(!cfg->n_device || (tc2_route && cfg->n_total > 0),
               "mapping_step: n_device needs the two-threads-per-point kernel and an explicit n_total");
  if (tc2_route) {
    const int G_ = tc2_groups();
    MISO_DISPATCH_LC_TC2(field->num_levels, field->level[0].C, {
      nblocks = G_ == 3 ? launch_tc2<L, C, 3>(field, dec, fr, m, s) : launch_tc2<L, C, 4>(field, dec, fr, m, s);
    });
    MISO_REQUIRE(nblocks > 0, "mapping_step(tc2): unsupported (levels=%d, channels=%d)", field->num_levels, field->level[0].C);
  } else if (use_tensor_cores() && fits_int32(field)) {
    MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
      constexpr size_t smem = sizeof(TcSmem<L * C>) + 128;
      auto k = mapping_step_tc_kernel<L, C>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      nblocks = grid_for((N + kTcThreads - 1) / kTcThreads, 1, sm_count());
      k<<<nblocks, kTcThreads, smem, s>>>(*field, *dec, fr, m);
    });
  } else {
    MISO_DISPATCH_LC(field->num_levels, field->level[0].C, {
      constexpr size_t smem = sizeof(DecoderSmem<L * C>);
      auto k = mapping_step_kernel<L, C>;
      nblocks = blocks_for(k, smem, N);
      if ((int64_t)nblocks * 4 > partial_floats()) nblocks = (int)(partial_floats() / 4);
      k<<<nblocks, kThreads, smem, s>>>(*field, *dec, fr, m);
    });
  }
  if (int e = check_launch("mapping_step")) return e;
  if (a_ext) return MISO_OK;   // backward-only pass: no loss terms to finalize
  mapping_finalize_kernel<<<1, kThreads, 0, s>>>(partials, nblocks, N, *cfg, eik_count, loss_out, m.poison);
  return check_launch("mapping_finalize");
}
