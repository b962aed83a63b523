// Latent-space submap alignment: transform + in-bound mask + dual interpolation + residual +
// pose-gradient / Gauss-Newton reductions, batched over every submap pair of an iteration.
//
// Replaces the body of pairwise_loss_latent (grid_opt/align/miso.py:116-211) and its autograd
// backward, called per pair per iteration from generic_align_multiple_submaps
// (grid_opt/align/base.py:127-159), plus check_submap_intersection (grid_opt/models/grid_atlas.py:405-420).
//
// The reference materialises coords_world, coords_to, the mask, nonzero indices, two (M,8) feature
// matrices and their autograd graph per pair (>=20 launches and ~1 KB/point of HBM traffic); here a
// point costs 12 B of coordinates + its corner fetches, and everything else stays in registers until
// the block-level reduction (float partials -> float64 atomics, 24 or 51 values per pair).
#include <algorithm>

#include <type_traits>

#include "common.cuh"

namespace miso {

constexpr int kAlignAcc = 24;       // S, count, G0(3), G1(9), G2(9), sum|r|
constexpr int kGnAcc = 27;          // Jtr(6) + upper triangle of JtJ(21)

struct Pose24 {
  float A1[9], b1[3], A2[9], b2[3];
};

__device__ __forceinline__ void load_pose(const float* __restrict__ poses, int pair, Pose24& P) {
  const float* s = poses + (int64_t)pair * 24;
#pragma unroll
  for (int i = 0; i < 9; ++i) P.A1[i] = s[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) P.b1[i] = s[9 + i];
#pragma unroll
  for (int i = 0; i < 9; ++i) P.A2[i] = s[12 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) P.b2[i] = s[21 + i];
}

// y = x A^T + b^T evaluated like torch's (N,3)@(3,3) + (1,3): products summed left to right
__device__ __forceinline__ void xform(const float (&A)[9], const float (&b)[3], const float (&x)[3], float (&y)[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float s = __fmul_rn(x[0], A[3 * j]);
    s = __fmaf_rn(x[1], A[3 * j + 1], s);
    s = __fmaf_rn(x[2], A[3 * j + 2], s);
    y[j] = __fadd_rn(s, b[j]);
  }
}

__device__ __forceinline__ bool in_bound(const float (&q)[3], const float* bound) {
  // coords_in_bound (utils_geometry.py:21-23): inclusive on both sides
  return q[0] >= bound[0] && q[0] <= bound[1] && q[1] >= bound[2] && q[1] <= bound[3] && q[2] >= bound[4] &&
         q[2] <= bound[5];
}

template <int C>
__device__ __forceinline__ void gather4(const miso_level_t& lv, const Cell& c, float* f, float* dx, float* dy,
                                        float* dz, bool deriv) {
  float w[8], wdx[8], wdy[8], wdz[8];
  long long off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float wx, wy, wz;
    axis_w(c, k, wx, wy, wz);
    bool ok = (c.valid >> k) & 1u;
    float sx = (k & 1) ? 1.f : -1.f, sy = (k & 2) ? 1.f : -1.f, sz = (k & 4) ? 1.f : -1.f;
    w[k] = ok ? (wx * wy) * wz : 0.f;
    wdx[k] = ok ? sx * wy * wz : 0.f;
    wdy[k] = ok ? wx * sy * wz : 0.f;
    wdz[k] = ok ? wx * wy * sz : 0.f;
    off[k] = ok ? corner_off(lv, c, k) : 0;
  }
#pragma unroll
  for (int ch = 0; ch < C; ch += 4) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ldg_f4(lv.feat + off[k] + ch);
    float a[4] = {0, 0, 0, 0}, ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0}, az[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float vv[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a[e] = fmaf(vv[e], w[k], a[e]);
        if (deriv) {
          ax[e] = fmaf(vv[e], wdx[k], ax[e]);
          ay[e] = fmaf(vv[e], wdy[k], ay[e]);
          az[e] = fmaf(vv[e], wdz[k], az[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[ch + e] = a[e];
      if (deriv) dx[ch + e] = ax[e], dy[ch + e] = ay[e], dz[ch + e] = az[e];
    }
  }
}

// Separable trilinear evaluation (lerp along x, then y, then z) of the value and its three index-space derivatives:
// the 8 corner vectors are consumed channel group by channel group, no per-corner weight / derivative-weight arrays
// (40 registers in gather4) -- out-of-range corners contribute zeros (zeros padding), exactly like masking their weights.
template <int C, bool kDeriv>
__device__ __forceinline__ void gather_sep(const miso_level_t& lv, const Cell& c, float* __restrict__ f,
                                           float* __restrict__ dx, float* __restrict__ dy, float* __restrict__ dz) {
#pragma unroll
  for (int ch = 0; ch < C; ch += 4) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((c.valid >> k) & 1u) v[k] = ldg_f4(lv.feat + corner_off(lv, c, k) + ch);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a[4], d[4];
#pragma unroll
      for (int yz = 0; yz < 4; ++yz) {
        const float lo = reinterpret_cast<const float*>(&v[2 * yz])[e], hi = reinterpret_cast<const float*>(&v[2 * yz + 1])[e];
        d[yz] = hi - lo;
        a[yz] = fmaf(c.fx, d[yz], lo);
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
        dz[ch + e] = ez;
        dy[ch + e] = fmaf(c.fz, ey[1] - ey[0], ey[0]);
        dx[ch + e] = fmaf(c.fz, dxy[1] - dxy[0], dxy[0]);
      }
    }
  }
}

template <int C>
__device__ __forceinline__ void scatter4(const miso_level_t& lv, const Cell& c, float coef, const float* r) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (!((c.valid >> k) & 1u)) continue;
    float wx, wy, wz;
    axis_w(c, k, wx, wy, wz);
    float w = coef * ((wx * wy) * wz);
    float* dst = lv.grad + corner_off(lv, c, k);
#pragma unroll
    for (int ch = 0; ch < C; ch += 4) red_add_f4(dst + ch, w * r[ch], w * r[ch + 1], w * r[ch + 2], w * r[ch + 3]);
  }
}

// kLoss: 0 = L2 (mean r^2 over points x channels, miso.py:200-201), 1 = "L1" (mean over points of |r|_2, :202-203),
// 2 = cos (mean over points of 1 - cosine_similarity(f_s, f_d), :204-205; ATen clamps each norm at eps = 1e-8).
// acc[0] always holds the sum of the per-point loss values and gamma = d(that sum)/dq, so the host-side
// normalisation is weight / (count * K) for L2 and weight / count for the other two.
#ifndef MISO_ALIGN_MIN_BLOCKS
#define MISO_ALIGN_MIN_BLOCKS 3
#endif
template <int C, bool kGN, int kLoss>
__global__ void __launch_bounds__(kThreads, (kGN || kLoss != 0) ? 1 : MISO_ALIGN_MIN_BLOCKS)
    align_batch_kernel(const miso_field_t* __restrict__ fields, const miso_align_pair_t* __restrict__ pairs,
                       const float* __restrict__ poses, double* __restrict__ out) {
  constexpr int NACC = kAlignAcc + (kGN ? kGnAcc : 0);
  // L1 / cos weigh a sample by 1/|r| resp. 1/(|f_s||f_d|): per-sample magnitudes spread over orders of magnitude and
  // the pose-gradient sums cancel heavily, so these two variants accumulate in float64 (the L2 hot path stays float32)
  using AccT = typename std::conditional<kLoss != 0, double, float>::type;
  __shared__ AccT red[NACC][kThreads / 32];
  const int pi = blockIdx.y;
  const miso_align_pair_t pr = pairs[pi];
  if (pr.enabled && *pr.enabled == 0) return;
  if (pr.M <= 0) return;
  // the two field descriptors and the pair's pose live in shared memory: level pointers / dims / strides and the 24 pose
  // floats are read where they are used (broadcast LDS) instead of being pinned in ~80 registers for the whole loop --
  // the kernel is latency-bound at 8 warps per SM with 231 registers (ncu: warps active 12.5 %, issue 33 %)
  __shared__ miso_field_t s_field[2];
  __shared__ Pose24 s_pose;
  {
    const uint32_t* gs = reinterpret_cast<const uint32_t*>(&fields[pr.src]);
    const uint32_t* gd = reinterpret_cast<const uint32_t*>(&fields[pr.dst]);
    uint32_t* ds = reinterpret_cast<uint32_t*>(&s_field[0]);
    uint32_t* dd = reinterpret_cast<uint32_t*>(&s_field[1]);
    for (int i = threadIdx.x; i < (int)(sizeof(miso_field_t) / 4); i += blockDim.x) ds[i] = gs[i], dd[i] = gd[i];
    float* dp = reinterpret_cast<float*>(&s_pose);
    for (int i = threadIdx.x; i < 24; i += blockDim.x) dp[i] = poses[(int64_t)pi * 24 + i];
  }
  __syncthreads();
  const miso_field_t& src = s_field[0];
  const miso_field_t& dst = s_field[1];
  const Pose24& P = s_pose;
  const float* dbound = dst.bound;
  const float* sbound = src.bound;
  const int LU = pr.levels_used;
  const int K = LU * C;

  AccT acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0;

  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < pr.M; n += (int64_t)gridDim.x * blockDim.x) {
    float p[3] = {pr.p[3 * n], pr.p[3 * n + 1], pr.p[3 * n + 2]};
    float u[3], q[3];
    xform(P.A1, P.b1, p, u);
    xform(P.A2, P.b2, u, q);
    const bool ok = in_bound(q, dbound);
    if (pr.mask_out) pr.mask_out[n] = ok ? 1 : 0;
    if (!ok) continue;
    float pn[3], qn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      pn[d] = normalize_coord(p[d], sbound[2 * d], sbound[2 * d + 1]);
      qn[d] = normalize_coord(q[d], dbound[2 * d], dbound[2 * d + 1]);
    }
    float gam[3] = {0.f, 0.f, 0.f};
    float rr = 0.f;
    if constexpr (kLoss != 0) {
      // the per-point factor of these losses needs the whole feature vector first: keep f_s, f_d and grad_q f_d
      float fsv[MISO_MAX_LEVELS * C], fdv[MISO_MAX_LEVELS * C], gq[MISO_MAX_LEVELS * C][3];
#pragma unroll
      for (int l = 0; l < MISO_MAX_LEVELS; ++l) {
        if (l >= LU) continue;
        float fs[C], fd[C], dx[C], dy[C], dz[C];
        const miso_level_t& sl = src.level[l];
        const miso_level_t& dl = dst.level[l];
        if (pr.fsrc) {
#pragma unroll
          for (int ch = 0; ch < C; ch += 4) {
            float4 t = *reinterpret_cast<const float4*>(pr.fsrc + n * K + l * C + ch);
            fs[ch] = t.x, fs[ch + 1] = t.y, fs[ch + 2] = t.z, fs[ch + 3] = t.w;
          }
        } else if ((src.ignore_mask >> l) & 1u) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) fs[ch] = 0.f;
        } else {
          Cell cs = make_cell(unnormalize_nc(pn[0], sl.X), unnormalize_nc(pn[1], sl.Y), unnormalize_nc(pn[2], sl.Z), sl);
          gather4<C>(sl, cs, fs, nullptr, nullptr, nullptr, false);
        }
        if ((dst.ignore_mask >> l) & 1u) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) fd[ch] = dx[ch] = dy[ch] = dz[ch] = 0.f;
        } else {
          Cell cd = make_cell(unnormalize_nc(qn[0], dl.X), unnormalize_nc(qn[1], dl.Y), unnormalize_nc(qn[2], dl.Z), dl);
          gather4<C>(dl, cd, fd, dx, dy, dz, true);
        }
        const float kx = (float)dl.X / (dbound[1] - dbound[0]), ky = (float)dl.Y / (dbound[3] - dbound[2]),
                    kz = (float)dl.Z / (dbound[5] - dbound[4]);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          fsv[l * C + ch] = fs[ch], fdv[l * C + ch] = fd[ch];
          gq[l * C + ch][0] = dx[ch] * kx, gq[l * C + ch][1] = dy[ch] * ky, gq[l * C + ch][2] = dz[ch] * kz;
        }
      }
      // the per-sample factors cancel heavily where |f_d| is small (1/|f_d| weights against a projection that removes
      // the f_d direction): the few scalars per sample are formed in float64, the interpolation stays float32
      double ss = 0, dd = 0, sd = 0, r2 = 0;
#pragma unroll
      for (int k = 0; k < MISO_MAX_LEVELS * C; ++k) {
        if (k >= K) continue;
        const double a = fsv[k], b = fdv[k];
        r2 += (a - b) * (a - b), ss += a * a, dd += b * b, sd += a * b;
      }
      double a_s = 0, a_d = 0;   // d(loss_i)/d(f_d) = a_s * f_s + a_d * f_d
      if constexpr (kLoss == 1) {
        const double nr = sqrt(r2);
        rr = (float)nr;                            // loss_i = |r|_2 ; d/df_d = -(f_s - f_d)/|r| (0 at r = 0, as autograd)
        if (nr > 0) a_s = -1.0 / nr, a_d = 1.0 / nr;
      } else {
        const double eps = 1e-8;
        const double ns = sqrt(ss), nd = sqrt(dd);
        const double cs_ = fmax(ns, eps), cd_ = fmax(nd, eps);
        rr = (float)(1.0 - sd / (cs_ * cd_));
        // cos = (f_s/cs).(f_d/cd), cd = max(|f_d|, eps):  d cos/d f_d = f_s/(cs cd) - [|f_d| > eps] (f_s.f_d) f_d/(cs cd^2 |f_d|)
        a_s = -1.0 / (cs_ * cd_);
        a_d = nd > eps ? sd / (cs_ * cd_ * cd_ * nd) : 0.0;
      }
      double g0 = 0, g1 = 0, g2 = 0;
#pragma unroll
      for (int k = 0; k < MISO_MAX_LEVELS * C; ++k) {
        if (k >= K) continue;
        const double c = a_s * (double)fsv[k] + a_d * (double)fdv[k];
        g0 += c * (double)gq[k][0], g1 += c * (double)gq[k][1], g2 += c * (double)gq[k][2];
      }
      gam[0] = (float)g0, gam[1] = (float)g1, gam[2] = (float)g2;
      acc[23] += sqrt(r2);
    } else {
    for (int l = 0; l < LU; ++l) {
      float fs[C], fd[C], dx[C], dy[C], dz[C];
      const miso_level_t& sl = src.level[l];
      const miso_level_t& dl = dst.level[l];
      Cell cs;
      const bool src_ignored = (src.ignore_mask >> l) & 1u;
      const bool dst_ignored = (dst.ignore_mask >> l) & 1u;
      if (pr.fsrc) {
#pragma unroll
        for (int ch = 0; ch < C; ch += 4) {
          float4 t = *reinterpret_cast<const float4*>(pr.fsrc + n * K + l * C + ch);
          fs[ch] = t.x, fs[ch + 1] = t.y, fs[ch + 2] = t.z, fs[ch + 3] = t.w;
        }
        if (sl.grad) cs = make_cell(unnormalize_nc(pn[0], sl.X), unnormalize_nc(pn[1], sl.Y), unnormalize_nc(pn[2], sl.Z), sl);
      } else if (src_ignored) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) fs[ch] = 0.f;
      } else {
        cs = make_cell(unnormalize_nc(pn[0], sl.X), unnormalize_nc(pn[1], sl.Y), unnormalize_nc(pn[2], sl.Z), sl);
        gather_sep<C, false>(sl, cs, fs, nullptr, nullptr, nullptr);
      }
      Cell cd = make_cell(unnormalize_nc(qn[0], dl.X), unnormalize_nc(qn[1], dl.Y), unnormalize_nc(qn[2], dl.Z), dl);
      if (dst_ignored) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) fd[ch] = dx[ch] = dy[ch] = dz[ch] = 0.f;
      } else {
        gather_sep<C, true>(dl, cd, fd, dx, dy, dz);
      }
      const float kx = (float)dl.X / (dbound[1] - dbound[0]), ky = (float)dl.Y / (dbound[3] - dbound[2]),
                  kz = (float)dl.Z / (dbound[5] - dbound[4]);
      float r[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        r[ch] = fs[ch] - fd[ch];
        rr = fmaf(r[ch], r[ch], rr);
        const float gx = dx[ch] * kx, gy = dy[ch] * ky, gz = dz[ch] * kz;  // grad_q f_d,ch
        gam[0] = fmaf(-2.f * r[ch], gx, gam[0]);
        gam[1] = fmaf(-2.f * r[ch], gy, gam[1]);
        gam[2] = fmaf(-2.f * r[ch], gz, gam[2]);
        if constexpr (kGN) {
          // J row (1x6) of r_ch wrt a right-multiplied dst twist: [ (q x g)^T , (R_d g)^T ],  R_d = A2^T
          float Jr[6];
          Jr[0] = q[1] * gz - q[2] * gy;
          Jr[1] = q[2] * gx - q[0] * gz;
          Jr[2] = q[0] * gy - q[1] * gx;
          Jr[3] = P.A2[0] * gx + P.A2[3] * gy + P.A2[6] * gz;
          Jr[4] = P.A2[1] * gx + P.A2[4] * gy + P.A2[7] * gz;
          Jr[5] = P.A2[2] * gx + P.A2[5] * gy + P.A2[8] * gz;
          int t = kAlignAcc + 6;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            acc[kAlignAcc + a] = fmaf(Jr[a], r[ch], acc[kAlignAcc + a]);
#pragma unroll
            for (int b = a; b < 6; ++b) {
              acc[t] = fmaf(Jr[a], Jr[b], acc[t]);
              ++t;
            }
          }
        }
      }
      // optional dS/dfeature scatter (the reference back-propagates into both submaps' grids)
      if (sl.grad && !src_ignored && pr.src_grad_scale != 0.f) scatter4<C>(sl, cs, 2.f * pr.src_grad_scale, r);
      if (dl.grad && !dst_ignored && pr.dst_grad_scale != 0.f) scatter4<C>(dl, cd, -2.f * pr.dst_grad_scale, r);
    }
    acc[23] += sqrtf(rr);
    }
    acc[0] += rr;
    acc[1] += 1.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      acc[2 + i] += gam[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if constexpr (kLoss != 0) {
          acc[5 + 3 * i + j] += (double)gam[i] * (double)u[j];
          acc[14 + 3 * i + j] += (double)gam[i] * (double)p[j];
        } else {
          // G1 = sum gamma u^T follows from u = A1 p + b1:  G1 = G2 A1^T + G0 b1^T -- formed once per block below
          // instead of carrying nine more accumulators through the loop (the GN variant keeps the direct sum)
          if constexpr (kGN) acc[5 + 3 * i + j] = fmaf(gam[i], u[j], acc[5 + 3 * i + j]);
          acc[14 + 3 * i + j] = fmaf(gam[i], p[j], acc[14 + 3 * i + j]);
        }
      }
    }
  }
  // block reduction, then one float64 atomic per value per block
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    AccT v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[i][w] = v;
  }
  __syncthreads();
  __shared__ double block_sum_[NACC];
  if (threadIdx.x < NACC) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) s += (double)red[threadIdx.x][k];
    block_sum_[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double s = block_sum_[threadIdx.x];
    double* o = out + (int64_t)pi * MISO_ALIGN_OUT;
    int idx = threadIdx.x;
    if constexpr (kLoss == 0 && !kGN) {
      if (idx >= 5 && idx < 14) {
        const int i = (idx - 5) / 3, j = (idx - 5) % 3;
        s = block_sum_[2 + i] * (double)P.b1[j];
#pragma unroll
        for (int k = 0; k < 3; ++k) s += block_sum_[14 + 3 * i + k] * (double)P.A1[3 * j + k];
      }
    }
    if (idx < kAlignAcc) {
      if (s != 0.0) atomicAdd(o + idx, s);
    } else if (idx < kAlignAcc + 6) {
      if (s != 0.0) atomicAdd(o + idx, s);  // Jtr at [24..29]
    } else {
      // upper-triangle entry t -> (a,b); mirror into the full 6x6 at [30..65]
      int t = idx - (kAlignAcc + 6), a = 0;
      while (t >= 6 - a) {
        t -= 6 - a;
        ++a;
      }
      int b = a + t;
      if (s != 0.0) {
        atomicAdd(o + 30 + a * 6 + b, s);
        if (a != b) atomicAdd(o + 30 + b * 6 + a, s);
      }
    }
  }
}

// check_submap_intersection for every pair, grouped by source submap: a block reads each source vertex ONCE
// and tests it against all destination submaps paired with that source (up to kMaxGroup), instead of
// re-reading the 48 MB vertex list once per pair.  `pairs` must be sorted so that pairs sharing (p, M) are
// contiguous; groups[g] = {first pair, number of pairs}.
constexpr int kMaxGroup = 32;

__global__ void __launch_bounds__(kThreads)
    align_intersection_kernel(const miso_field_t* __restrict__ fields, const miso_align_pair_t* __restrict__ pairs,
                              const int2* __restrict__ groups, const float* __restrict__ poses,
                              unsigned long long* __restrict__ counts) {
  const int2 grp = groups[blockIdx.y];
  const int first = grp.x, np = grp.y;
  const miso_align_pair_t pr0 = pairs[first];
  if (pr0.M <= 0) return;
  __shared__ float s_pose[kMaxGroup][24];
  __shared__ float s_bound[kMaxGroup][6];
  __shared__ int s_slot[kMaxGroup];
  __shared__ float s_eps[kMaxGroup][3];   // lattice path: decision margin per destination and axis
  for (int i = threadIdx.x; i < np * 3; i += blockDim.x) {
    const float* b = fields[pairs[first + i / 3].dst].bound + 2 * (i % 3);
    s_eps[i / 3][i % 3] = 1e-3f + 1e-5f * fmaxf(fabsf(b[0]), fabsf(b[1]));
  }
  for (int i = threadIdx.x; i < np * 24; i += blockDim.x) {
    const int j = i / 24;
    s_pose[j][i % 24] = poses[(int64_t)pairs[first + j].reserved * 24 + i % 24];
  }
  for (int i = threadIdx.x; i < np * 6; i += blockDim.x) s_bound[i / 6][i % 6] = fields[pairs[first + i / 6].dst].bound[i % 6];
  for (int i = threadIdx.x; i < np; i += blockDim.x) s_slot[i] = pairs[first + i].reserved;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned my_cnt = 0;
  float A1[9], b1[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) A1[i] = s_pose[0][i];   // src -> world is shared by the whole group
#pragma unroll
  for (int i = 0; i < 3; ++i) b1[i] = s_pose[0][9 + i];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // levels_used > 0: p is a lattice listed row by row -- X = low 16 bits vertices per row (collinear, equispaced),
  // Y = high 16 bits rows per z-plane; vertex (iz,iy,ix) = (p[3 ix], p[3 iy X + 1], p[3 iz X Y + 2]) exactly, because
  // vertex_positions() is a per-axis outer product (grid_modules.py:111-123, utils.py:294-307)
  const int row_len = pr0.levels_used & 0xffff;
  const int rows_per_plane = pr0.levels_used >> 16;
  if (row_len > 0) {
    // Lattice path.  A thread takes one ROW of the lattice (row_len collinear, equispaced vertices).  Along the row the
    // map vertex index -> q is affine up to fp32 rounding, and the destination bound is convex, so the inside indices
    // form an interval.  For every destination the thread intersects the row with the bound twice in closed form:
    // with the bound SHRUNK by a margin (indices in that interval are surely inside) and GROWN by it (indices outside
    // that interval are surely outside); only the few indices between the two intervals are tested one by one, with
    // the very same arithmetic as the plain path.  The margin (1e-3 + 1e-5 |bound|) is two orders above the rounding
    // error of the two chained transforms and of the linear model, so the count is the exact per-vertex count; a row
    // that runs (nearly) parallel to a face inside the margin band simply has a long uncertain interval and is
    // counted vertex by vertex.  Work per source drops from O(vertices) transforms to O(rows x destinations).
    const int64_t rows = pr0.M / row_len;
    // Whole-lattice pre-test per destination: every vertex lies in the hull of the lattice's 8 extreme corners (first
    // and last vertex give the per-axis extremes); if all 8 are beyond the same face of the destination bound (same
    // margin as below) the pair's count is exactly 0 and the destination drops out of the loops.  Most pairs of a
    // 16-submap atlas do not overlap at all.
    __shared__ int s_act[kMaxGroup];
    __shared__ int s_nact;
    if (threadIdx.x < 32) {
      const int j = threadIdx.x;
      bool keep = false;
      if (j < np) {
        const float* pl = pr0.p + 3 * (pr0.M - 1);
        unsigned common = 0x3fu;
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) {
          const float cp[3] = {(cidx & 1) ? pl[0] : pr0.p[0], (cidx & 2) ? pl[1] : pr0.p[1], (cidx & 4) ? pl[2] : pr0.p[2]};
          float cu[3], cq[3], A2[9], b2[3];
          xform(A1, b1, cp, cu);
#pragma unroll
          for (int i = 0; i < 9; ++i) A2[i] = s_pose[j][12 + i];
#pragma unroll
          for (int i = 0; i < 3; ++i) b2[i] = s_pose[j][21 + i];
          xform(A2, b2, cu, cq);
          unsigned faces = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float lo = s_bound[j][2 * k], hi = s_bound[j][2 * k + 1], eps = s_eps[j][k];
            faces |= (cq[k] < lo - eps ? 1u : 0u) << (2 * k);
            faces |= (cq[k] > hi + eps ? 1u : 0u) << (2 * k + 1);
          }
          common &= faces;
        }
        keep = common == 0u;
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) s_act[__popc(km & ((1u << j) - 1u))] = j;
      if (j == 0) s_nact = __popc(km);
    }
    __syncthreads();
    const int nact = s_nact;
    if (nact == 0) return;
    const float last = (float)(row_len - 1);
    const float inv_last = row_len > 1 ? 1.0f / last : 0.f;
    // whole warps stay in the loop together (the reduction below is warp-wide): iterate on the warp's first row
    for (int64_t r0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); r0 < rows; r0 += stride) {
      const int64_t row = r0 + lane;
      const bool live = row < rows;
      // coordinates from the three per-axis tables inside p (a few KB, cache-resident): the kernel never streams the
      // 12 B/vertex list
      const int64_t iz = live ? row / rows_per_plane : 0, iy = live ? row - iz * rows_per_plane : 0;
      const float py = __ldg(pr0.p + 3 * (iy * row_len) + 1);
      const float pz = __ldg(pr0.p + 3 * (iz * row_len * rows_per_plane) + 2);
      float u0[3], u1[3];
      {
        const float pa[3] = {__ldg(pr0.p), py, pz}, pb[3] = {__ldg(pr0.p + 3 * (row_len - 1)), py, pz};
        xform(A1, b1, pa, u0);
        xform(A1, b1, pb, u1);
      }
      for (int ja = 0; ja < nact; ++ja) {
        const int j = s_act[ja];
        float A2[9], b2[3], bd[6];
#pragma unroll
        for (int i = 0; i < 9; ++i) A2[i] = s_pose[j][12 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) b2[i] = s_pose[j][21 + i];
#pragma unroll
        for (int i = 0; i < 6; ++i) bd[i] = s_bound[j][i];
        unsigned hits = 0;
        if (live) {
          float q0[3], q1[3];
          xform(A2, b2, u0, q0);
          xform(A2, b2, u1, q1);
          // index intervals: [slo, shi] surely inside (shrunk bound), [plo, phi] possibly inside (grown bound)
          float slo = 0.f, shi = last, plo = 0.f, phi = last;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float lo = bd[2 * k], hi = bd[2 * k + 1], eps = s_eps[j][k];
            const float d = (q1[k] - q0[k]) * inv_last;   // change of q_k per index step
            if (fabsf(d) < 1e-12f) {
              if (!(q0[k] >= lo + eps && q0[k] <= hi - eps)) slo = 1.f, shi = 0.f;   // not surely inside anywhere
              if (!(q0[k] >= lo - eps && q0[k] <= hi + eps)) plo = 1.f, phi = 0.f;   // surely outside everywhere
            } else {
              const float inv = 1.0f / d;
              float a = (lo + eps - q0[k]) * inv, b = (hi - eps - q0[k]) * inv;
              slo = fmaxf(slo, fminf(a, b)), shi = fminf(shi, fmaxf(a, b));
              a = (lo - eps - q0[k]) * inv, b = (hi + eps - q0[k]) * inv;
              plo = fmaxf(plo, fminf(a, b)), phi = fminf(phi, fmaxf(a, b));
            }
          }
          // one more index of slack on each side for the rounding of the interval arithmetic itself
          int i_slo = (int)ceilf(slo) + 1, i_shi = (int)floorf(shi) - 1;
          int i_plo = max(0, (int)floorf(plo) - 1), i_phi = min(row_len - 1, (int)ceilf(phi) + 1);
          if (!(phi >= plo)) i_plo = 1, i_phi = 0;          // empty (also catches NaN)
          if (!(shi >= slo) || i_shi < i_slo) i_slo = i_phi + 1, i_shi = i_phi;   // no sure part: test all of [plo, phi]
          i_slo = max(i_slo, i_plo), i_shi = min(i_shi, i_phi);
          if (i_shi >= i_slo) hits = (unsigned)(i_shi - i_slo + 1);
          for (int i = i_plo; i <= i_phi; ++i) {
            if (i == i_slo && i_shi >= i_slo) {   // skip the sure part
              i = i_shi;
              continue;
            }
            const float p[3] = {__ldg(pr0.p + 3 * i), py, pz};
            float uu[3], q[3];
            xform(A1, b1, p, uu);
            xform(A2, b2, uu, q);
            hits += in_bound(q, bd) ? 1u : 0u;
          }
        }
        const unsigned tot = __reduce_add_sync(0xffffffffu, hits);
        if (lane == j) my_cnt += tot;
      }
    }
    if (lane < np && my_cnt) atomicAdd(counts + s_slot[lane], (unsigned long long)my_cnt);
    return;
  }
  // Plain path (arbitrary point lists).  Each thread carries

  // kV vertices through the destination loop so one shared-memory read of a destination's (A2, b2, bound) serves
  // kV transforms (the loop used to be bound by those reads: 18 floats per vertex and destination), and the kV hit
  // bits are summed over the warp with ONE redux per destination.
  constexpr int kV = 4;
  const int64_t n_iter = (pr0.M + stride * kV - 1) / (stride * kV);
  for (int64_t it = 0; it < n_iter; ++it) {
    float u[kV][3];
    unsigned live = 0;
#pragma unroll
    for (int v = 0; v < kV; ++v) {
      const int64_t n = (it * kV + v) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      float p[3] = {0.f, 0.f, 0.f};
      if (n < pr0.M) {
        live |= 1u << v;
        p[0] = pr0.p[3 * n], p[1] = pr0.p[3 * n + 1], p[2] = pr0.p[3 * n + 2];
      }
      xform(A1, b1, p, u[v]);
    }
    for (int j = 0; j < np; ++j) {
      float A2[9], b2[3], bd[6];
#pragma unroll
      for (int i = 0; i < 9; ++i) A2[i] = s_pose[j][12 + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) b2[i] = s_pose[j][21 + i];
#pragma unroll
      for (int i = 0; i < 6; ++i) bd[i] = s_bound[j][i];
      unsigned hits = 0;
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        float q[3];
        xform(A2, b2, u[v], q);
        hits += (((live >> v) & 1u) && in_bound(q, bd)) ? 1u : 0u;
      }
      const unsigned tot = __reduce_add_sync(0xffffffffu, hits);
      if (lane == j) my_cnt += tot;
    }
  }
  if (lane < np && my_cnt) atomicAdd(counts + s_slot[lane], (unsigned long long)my_cnt);
}

__global__ void align_intersection_finalize(const miso_align_pair_t* __restrict__ pairs, int num_pairs,
                                            const unsigned long long* __restrict__ counts, float thresh,
                                            int32_t* __restrict__ enabled) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_pairs) return;
  // overlap_percentage = num_valid / num_all (float32 division of int64 tensors in torch) > thresh
  const int64_t M = pairs[i].M;
  const int slot = pairs[i].reserved;
  float frac = M > 0 ? (float)counts[slot] / (float)M : 0.f;
  enabled[slot] = frac > thresh ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------
// GridAtlas.query_feature (grid_opt/models/grid_atlas.py:374-391): in-bound-masked MEAN over the active submaps
// of each submap's multi-level feature at a world point, in ONE launch over all submaps (the reference loops over
// submaps with a transform, a mask, L grid_sample launches and two accumulations each).  poses (S,12) = per submap
// (A = R^T row-major, b = -R^T t), composed by the caller exactly as transfrom_points_from does.  Submaps are visited
// in index order so the fp32 sums match the reference's accumulation order.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads)
    atlas_features_kernel(const miso_field_t* __restrict__ fields, const int32_t* __restrict__ active, int num_active,
                          const float* __restrict__ poses, const float* __restrict__ x, int64_t N, int levels,
                          float* __restrict__ out) {
  const int F = levels * C;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const float p[3] = {x[3 * n], x[3 * n + 1], x[3 * n + 2]};
    float sum[MISO_MAX_LEVELS * C];
#pragma unroll
    for (int i = 0; i < MISO_MAX_LEVELS * C; ++i) sum[i] = 0.f;
    float wsum = 0.f;
    for (int a = 0; a < num_active; ++a) {
      const int sidx = active[a];
      const miso_field_t& fld = fields[sidx];
      float A[9], b[3], q[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) A[i] = poses[sidx * 12 + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) b[i] = poses[sidx * 12 + 9 + i];
      xform(A, b, p, q);
      if (!in_bound(q, fld.bound)) continue;     // mask_bnd * feats: out-of-bound submaps add exactly zero
      wsum += 1.f;
      float qn[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) qn[d] = normalize_coord(q[d], fld.bound[2 * d], fld.bound[2 * d + 1]);
#pragma unroll
      for (int l = 0; l < MISO_MAX_LEVELS; ++l) {
        if (l >= levels || ((fld.ignore_mask >> l) & 1u)) continue;
        const miso_level_t& lv = fld.level[l];
        Cell c = make_cell(unnormalize_nc(qn[0], lv.X), unnormalize_nc(qn[1], lv.Y), unnormalize_nc(qn[2], lv.Z), lv);
        float f[C];
        gather4<C>(lv, c, f, nullptr, nullptr, nullptr, false);
#pragma unroll
        for (int i = 0; i < C; ++i) sum[l * C + i] += f[i];
      }
    }
    const float den = wsum == 0.f ? 1.0f : wsum;   // torch.where(sum_weights == 0, 1, sum_weights)
    for (int i = 0; i < F; ++i) out[n * F + i] = sum[i] / den;
  }
}

}  // namespace miso

using namespace miso;

// Blocks of disabled (non-intersecting) pairs exit at once and the host cannot know how many pairs are enabled (the
// flags live on the device), so the grid is over-provisioned: ~64 blocks per SM spread over all pairs keeps the ~1/3
// of pairs that do overlap in several waves of small blocks instead of less than one wave of big ones (B200, 16 submaps,
// level 1: 4.99 ms at 8 per SM, 3.51 at 32, 3.14 at 64, 3.15 at 128).
#ifndef MISO_ALIGN_BLOCKS_PER_SM
#define MISO_ALIGN_BLOCKS_PER_SM 64
#endif
extern "C" int miso_align_batch(const miso_field_t* fields, int32_t num_fields, const miso_align_pair_t* pairs,
                                int32_t num_pairs, int64_t max_M, const float* poses, double* out, int32_t flags,
                                miso_stream_t stream) {
  const int want_gn = flags & 1, loss_kind = (flags >> 4) & 3;
  MISO_REQUIRE(loss_kind <= 2, "align_batch: loss kind must be 0 (L2), 1 (L1) or 2 (cos)");
  MISO_REQUIRE(!(want_gn && loss_kind != 0), "align_batch: Gauss-Newton accumulators exist for the L2 loss only");
  MISO_REQUIRE(fields && pairs && poses && out, "align_batch: null argument");
  MISO_REQUIRE(num_fields > 0 && num_pairs >= 0, "align_batch: bad counts");
  if (num_pairs == 0) return MISO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double) * MISO_ALIGN_OUT * num_pairs, s);
  if (max_M <= 0) return check_launch("align_batch(memset)");
  MISO_REQUIRE(num_pairs <= 65535, "align_batch: too many pairs (%d)", num_pairs);
  // channel count is read on the host from nothing (fields live on the device): the ABI fixes C=4
  // per level for alignment (fdim=4, miso.py:122); other widths go through the generic path.
  const int budget = sm_count() * MISO_ALIGN_BLOCKS_PER_SM;
  // at least four voxels per thread: small (coarse-level) lattices pay for every extra block's reduction epilogue
  int bx = grid_for(max_M, kThreads * 4, std::max(1, budget / std::max(1, std::min(num_pairs, budget))));
  dim3 grid(bx, num_pairs);
  if (want_gn)
    align_batch_kernel<4, true, 0><<<grid, kThreads, 0, s>>>(fields, pairs, poses, out);
  else if (loss_kind == 1)
    align_batch_kernel<4, false, 1><<<grid, kThreads, 0, s>>>(fields, pairs, poses, out);
  else if (loss_kind == 2)
    align_batch_kernel<4, false, 2><<<grid, kThreads, 0, s>>>(fields, pairs, poses, out);
  else
    align_batch_kernel<4, false, 0><<<grid, kThreads, 0, s>>>(fields, pairs, poses, out);
  return check_launch("align_batch");
}

extern "C" int miso_align_intersections(const miso_field_t* fields, int32_t num_fields,
                                        const miso_align_pair_t* pairs, int32_t num_pairs, const int32_t* groups,
                                        int32_t num_groups, int64_t max_M, const float* poses, float overlap_thresh,
                                        int32_t* enabled_out, unsigned long long* counts_out, miso_stream_t stream) {
  MISO_REQUIRE(fields && pairs && poses && enabled_out && counts_out && groups, "align_intersections: null argument");
  MISO_REQUIRE(num_fields > 0 && num_pairs >= 0 && num_groups >= 0 && num_groups <= 65535,
               "align_intersections: bad counts");
  if (num_pairs == 0) return MISO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(counts_out, 0, sizeof(unsigned long long) * num_pairs, s);
  if (max_M > 0 && num_groups > 0) {
    int bx = grid_for(max_M, kThreads, std::max(1, (sm_count() * 8) / std::max(1, std::min(num_groups, sm_count() * 8))));
    dim3 grid(bx, num_groups);
    align_intersection_kernel<<<grid, kThreads, 0, s>>>(fields, pairs, reinterpret_cast<const int2*>(groups), poses,
                                                        counts_out);
  }
  align_intersection_finalize<<<(num_pairs + 127) / 128, 128, 0, s>>>(pairs, num_pairs, counts_out, overlap_thresh,
                                                                       enabled_out);
  return check_launch("align_intersections");
}

extern "C" int miso_atlas_features(const miso_field_t* fields, int32_t num_fields, const int32_t* active,
                                   int32_t num_active, const float* poses, const float* x, int64_t N, int32_t levels,
                                   float* feats, miso_stream_t stream) {
  MISO_REQUIRE(fields && active && poses && num_fields > 0 && num_active >= 0, "atlas_features: null argument");
  MISO_REQUIRE(levels >= 1 && levels <= MISO_MAX_LEVELS, "atlas_features: levels %d not in [1,%d]", levels, MISO_MAX_LEVELS);
  MISO_REQUIRE(N >= 0 && (N == 0 || (x && feats)), "atlas_features: null x/feats");
  if (N == 0) return MISO_OK;
  // fields live on the device: the ABI fixes C = 4 per level here, as for alignment (fdim=4, miso.py:122)
  atlas_features_kernel<4><<<grid_for(N, kThreads, sm_count() * 8), kThreads, 0, (cudaStream_t)stream>>>(
      fields, active, num_active, poses, x, N, levels, feats);
  return check_launch("atlas_features");
}
