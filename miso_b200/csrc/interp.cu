// Generic trilinear grid_sample forward / backward / double-backward (the reference's plugin point).
//
// Replaces, for the 3D case MISO uses:
//   fwd      ATen grid_sampler_3d  via F.grid_sample             (grid_opt/models/grid_modules.py:89-94)
//   bwd      aten::grid_sampler_3d_backward                      (third_party/cuda_gridsample_grad2/cuda_gridsample.py:102-107)
//   bwd_bwd  grid_sampler_3d_grad2_kernel                        (third_party/cuda_gridsample_grad2/gridsample_cuda.cu:212-533)
//
// B200 design: one thread per sample point; with a channels-last grid (stride_C == 1, C % 4 == 0)
// every corner is fetched with 128-bit loads (x-adjacent corners share a 32-byte sector) and
// gradients are scattered with 128-bit `red.global.add.v4.f32`, i.e. 8 vector atomics per point
// and channel-quad instead of the reference's 32 scalar ones.  NCDHW (planar) grids and fp64 go
// through the scalar instantiation so the entry points are a complete drop-in.
#include "common.cuh"

namespace miso {

template <typename T>
struct GridArgs {
  const T* input;
  int64_t B, C, D, H, W;
  int64_t sN, sC, sD, sH, sW;
  const T* grid;  // (B,P,3)
  int64_t P;
  int pad, align;
};

// strides of a gradient buffer shaped like `input` (may differ from input's own strides)
struct GStrides {
  int64_t sN, sC, sD, sH, sW;
  __device__ __forceinline__ int64_t koff(int k) const {
    return ((k & 1) ? sW : 0) + ((k & 2) ? sH : 0) + ((k & 4) ? sD : 0);
  }
};

template <typename T>
__device__ __forceinline__ T src_index(T c, int64_t size, int pad, int align, T* mult) {
  // grid_sampler_compute_source_index_set_grad (ATen GridSampler.cuh; used at gridsample_cuda.cu:297-299)
  T r;
  if (align) {
    *mult = (T)(size - 1) / 2;
    r = ((c + 1) / 2) * (T)(size - 1);
  } else {
    *mult = (T)size / 2;
    r = ((c + 1) * (T)size - 1) / 2;
  }
  if (pad == MISO_PAD_BORDER) {
    if (r <= (T)0) {
      r = 0;
      *mult = 0;
    } else if (r >= (T)(size - 1)) {
      r = (T)(size - 1);
      *mult = 0;
    }
  }
  return r;
}
// fp32, align_corners=False: identical rounding sequence to the CPU oracle (no FMA contraction)
template <>
__device__ __forceinline__ float src_index<float>(float c, int64_t size, int pad, int align, float* mult) {
  float r;
  if (align) {
    *mult = (float)(size - 1) / 2;
    r = __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(size - 1));
  } else {
    *mult = (float)size / 2;
    r = unnormalize_nc(c, (int)size);
  }
  if (pad == MISO_PAD_BORDER) {
    if (r <= 0.f) {
      r = 0.f;
      *mult = 0.f;
    } else if (r >= (float)(size - 1)) {
      r = (float)(size - 1);
      *mult = 0.f;
    }
  }
  return r;
}

template <typename T>
struct PCell {
  int64_t x0, y0, z0;
  T fx, fy, fz;
  T mx, my, mz;  // d(index)/d(normalised coord)
  unsigned valid;
  int64_t base;
};

template <typename T>
__device__ __forceinline__ PCell<T> point_cell(const GridArgs<T>& a, int64_t b, int64_t p) {
  const T* g = a.grid + (b * a.P + p) * 3;
  PCell<T> c;
  T ix = src_index<T>(g[0], a.W, a.pad, a.align, &c.mx);
  T iy = src_index<T>(g[1], a.H, a.pad, a.align, &c.my);
  T iz = src_index<T>(g[2], a.D, a.pad, a.align, &c.mz);
  T flx = floor(ix), fly = floor(iy), flz = floor(iz);
  T lo = (T)-2;
  c.x0 = (int64_t)fmin(fmax(flx, lo), (T)(a.W + 1));
  c.y0 = (int64_t)fmin(fmax(fly, lo), (T)(a.H + 1));
  c.z0 = (int64_t)fmin(fmax(flz, lo), (T)(a.D + 1));
  c.fx = ix - flx;
  c.fy = iy - fly;
  c.fz = iz - flz;
  unsigned v = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int64_t x = c.x0 + (k & 1), y = c.y0 + ((k >> 1) & 1), z = c.z0 + (k >> 2);
    bool ok = x >= 0 && x < a.W && y >= 0 && y < a.H && z >= 0 && z < a.D;
    v |= (ok ? 1u : 0u) << k;
  }
  c.valid = v;
  c.base = b * a.sN + c.z0 * a.sD + c.y0 * a.sH + c.x0 * a.sW;
  return c;
}

template <typename T>
__device__ __forceinline__ int64_t koff(const GridArgs<T>& a, int k) {
  return ((k & 1) ? a.sW : 0) + ((k & 2) ? a.sH : 0) + ((k & 4) ? a.sD : 0);
}

template <typename T>
__device__ __forceinline__ void kweights(const PCell<T>& c, int k, T& wx, T& wy, T& wz, T& sx, T& sy, T& sz) {
  wx = (k & 1) ? c.fx : (T)1 - c.fx;
  wy = (k & 2) ? c.fy : (T)1 - c.fy;
  wz = (k & 4) ? c.fz : (T)1 - c.fz;
  sx = (k & 1) ? (T)1 : (T)-1;
  sy = (k & 2) ? (T)1 : (T)-1;
  sz = (k & 4) ? (T)1 : (T)-1;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) grid_sample_fwd_kernel(GridArgs<T> a, T* __restrict__ out, int64_t oB,
                                                                   int64_t oC, int64_t oP) {
  const int64_t total = a.B * a.P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / a.P, p = i - b * a.P;
    PCell<T> c = point_cell(a, b, p);
    T w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      T wx, wy, wz, sx, sy, sz;
      kweights(c, k, wx, wy, wz, sx, sy, sz);
      w[k] = (c.valid >> k) & 1u ? (wx * wy) * wz : (T)0;
    }
    T* o = out + b * oB + p * oP;
    if constexpr (VEC == 4) {
      for (int64_t ch = 0; ch < a.C; ch += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if ((c.valid >> k) & 1u) {
            float4 v = ldg_f4(a.input + c.base + koff(a, k) + ch);
            acc.x = fmaf(v.x, w[k], acc.x);
            acc.y = fmaf(v.y, w[k], acc.y);
            acc.z = fmaf(v.z, w[k], acc.z);
            acc.w = fmaf(v.w, w[k], acc.w);
          }
        }
        if (oC == 1) {
          *reinterpret_cast<float4*>(o + ch) = acc;
        } else {
          o[(ch + 0) * oC] = acc.x;
          o[(ch + 1) * oC] = acc.y;
          o[(ch + 2) * oC] = acc.z;
          o[(ch + 3) * oC] = acc.w;
        }
      }
    } else {
      for (int64_t ch = 0; ch < a.C; ++ch) {
        T acc = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if ((c.valid >> k) & 1u) acc += a.input[c.base + koff(a, k) + ch * a.sC] * w[k];
        o[ch * oC] = acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: grad_input (scatter, accumulated) and grad_grid
// ---------------------------------------------------------------------------------------------
#ifndef MISO_BWD_MIN_BLOCKS
#define MISO_BWD_MIN_BLOCKS 2
#endif
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? MISO_BWD_MIN_BLOCKS : 1)
    grid_sample_bwd_kernel(GridArgs<T> a, const T* __restrict__ go, int64_t gB, int64_t gC, int64_t gP,
                           T* __restrict__ grad_input, GStrides gs, T* __restrict__ grad_grid) {
  const int64_t total = a.B * a.P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / a.P, p = i - b * a.P;
    PCell<T> c = point_cell(a, b, p);
    const T* g = go + b * gB + p * gP;
    const int64_t gbase = b * gs.sN + c.z0 * gs.sD + c.y0 * gs.sH + c.x0 * gs.sW;
    T gix = 0, giy = 0, giz = 0;
    if constexpr (VEC == 4) {
      for (int64_t ch = 0; ch < a.C; ch += 4) {
        float4 gv;
        if (gC == 1) {
          gv = *reinterpret_cast<const float4*>(g + ch);
        } else {
          gv = make_float4(g[ch * gC], g[(ch + 1) * gC], g[(ch + 2) * gC], g[(ch + 3) * gC]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!((c.valid >> k) & 1u)) continue;
          float wx, wy, wz, sx, sy, sz;
          kweights(c, k, wx, wy, wz, sx, sy, sz);
          const int64_t off = c.base + koff(a, k) + ch;
          if (grad_input) {
            float w = (wx * wy) * wz;
            red_add_f4(grad_input + gbase + gs.koff(k) + ch, w * gv.x, w * gv.y, w * gv.z, w * gv.w);
          }
          if (grad_grid) {
            float4 v = ldg_f4(a.input + off);
            float dot = v.x * gv.x + v.y * gv.y + v.z * gv.z + v.w * gv.w;
            gix += dot * (sx * wy * wz);
            giy += dot * (wx * sy * wz);
            giz += dot * (wx * wy * sz);
          }
        }
      }
    } else {
      for (int64_t ch = 0; ch < a.C; ++ch) {
        const T gv = g[ch * gC];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!((c.valid >> k) & 1u)) continue;
          T wx, wy, wz, sx, sy, sz;
          kweights(c, k, wx, wy, wz, sx, sy, sz);
          const int64_t off = c.base + koff(a, k) + ch * a.sC;
          if (grad_input) red_add(grad_input + gbase + gs.koff(k) + ch * gs.sC, (wx * wy) * wz * gv);
          if (grad_grid) {
            T dot = a.input[off] * gv;
            gix += dot * (sx * wy * wz);
            giy += dot * (wx * sy * wz);
            giz += dot * (wx * wy * sz);
          }
        }
      }
    }
    if (grad_grid) {
      T* gg = grad_grid + i * 3;
      gg[0] = gix * c.mx;
      gg[1] = giy * c.my;
      gg[2] = giz * c.mz;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// double backward (gridsample_cuda.cu:212-533): given cotangents gg_input (grid-shaped) and
// gg_grid (per point) of the first backward's outputs, produce gg_output, g_input, g_grid.
// ---------------------------------------------------------------------------------------------
#ifndef MISO_BWD2_MIN_BLOCKS
#define MISO_BWD2_MIN_BLOCKS 2   // 128 registers: 16 warps/SM instead of 8 at 189 registers (1.5-1.7x, profiles/r02_interp_sweep.csv)
#endif
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? MISO_BWD2_MIN_BLOCKS : 1)
    grid_sample_bwd_bwd_kernel(GridArgs<T> a, const T* __restrict__ ggi, int64_t iN, int64_t iC, int64_t iD,
                               int64_t iH, int64_t iW, const T* __restrict__ ggg, const T* __restrict__ go,
                               int64_t gB, int64_t gC, int64_t gP, T* __restrict__ ggo, int64_t oB, int64_t oC,
                               int64_t oP, T* __restrict__ g_input, GStrides gs, T* __restrict__ g_grid) {
  const int64_t total = a.B * a.P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / a.P, p = i - b * a.P;
    PCell<T> c = point_cell(a, b, p);
    T dx = 0, dy = 0, dz = 0;
    if (ggg) {
      dx = ggg[i * 3 + 0] * c.mx;
      dy = ggg[i * 3 + 1] * c.my;
      dz = ggg[i * 3 + 2] * c.mz;
    }
    const int64_t ibase = b * iN + c.z0 * iD + c.y0 * iH + c.x0 * iW;
    const int64_t gbase = b * gs.sN + c.z0 * gs.sD + c.y0 * gs.sH + c.x0 * gs.sW;
    const T* g = go + b * gB + p * gP;
    T gix = 0, giy = 0, giz = 0;
    constexpr int CH = (VEC == 4) ? 4 : 1;
    for (int64_t ch = 0; ch < a.C; ch += CH) {
      T gv[CH], oacc[CH], dxy[CH], dxz[CH], dyz[CH], sgx[CH], sgy[CH], sgz[CH];
#pragma unroll
      for (int e = 0; e < CH; ++e) {
        gv[e] = g[(ch + e) * gC];
        oacc[e] = dxy[e] = dxz[e] = dyz[e] = sgx[e] = sgy[e] = sgz[e] = 0;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!((c.valid >> k) & 1u)) continue;
        T wx, wy, wz, sx, sy, sz;
        kweights(c, k, wx, wy, wz, sx, sy, sz);
        const T w = (wx * wy) * wz;
        const T wdx = sx * wy * wz, wdy = wx * sy * wz, wdz = wx * wy * sz;
        const T tmp = dx * wdx + dy * wdy + dz * wdz;
        T val[CH], g2[CH];
        if constexpr (VEC == 4) {
          const int64_t off = c.base + koff(a, k) + ch;
          float4 v = ldg_f4(a.input + off);
          val[0] = v.x, val[1] = v.y, val[2] = v.z, val[3] = v.w;
          if (ggi) {
            const int64_t ioff = ibase + ((k & 1) ? iW : 0) + ((k & 2) ? iH : 0) + ((k & 4) ? iD : 0);
#pragma unroll
            for (int e = 0; e < 4; ++e) g2[e] = ggi[ioff + (ch + e) * iC];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) g2[e] = 0;
          }
          if (g_input) red_add_f4(g_input + gbase + gs.koff(k) + ch, tmp * gv[0], tmp * gv[1], tmp * gv[2], tmp * gv[3]);
        } else {
          const int64_t off = c.base + koff(a, k) + ch * a.sC;
          val[0] = a.input[off];
          g2[0] = ggi ? ggi[ibase + ((k & 1) ? iW : 0) + ((k & 2) ? iH : 0) + ((k & 4) ? iD : 0) + ch * iC] : (T)0;
          if (g_input) red_add(g_input + gbase + gs.koff(k) + ch * gs.sC, tmp * gv[0]);
        }
#pragma unroll
        for (int e = 0; e < CH; ++e) {
          oacc[e] += g2[e] * w + val[e] * tmp;
          dxy[e] += val[e] * (sx * sy * wz);
          dxz[e] += val[e] * (sx * wy * sz);
          dyz[e] += val[e] * (wx * sy * sz);
          sgx[e] += g2[e] * wdx;
          sgy[e] += g2[e] * wdy;
          sgz[e] += g2[e] * wdz;
        }
      }
#pragma unroll
      for (int e = 0; e < CH; ++e) {
        if (ggo) ggo[b * oB + p * oP + (ch + e) * oC] = oacc[e];
        gix += gv[e] * (sgx[e] + dz * dxz[e] + dy * dxy[e]);
        giy += gv[e] * (sgy[e] + dx * dxy[e] + dz * dyz[e]);
        giz += gv[e] * (sgz[e] + dx * dxz[e] + dy * dyz[e]);
      }
    }
    if (g_grid) {
      g_grid[i * 3 + 0] = gix * c.mx;
      g_grid[i * 3 + 1] = giy * c.my;
      g_grid[i * 3 + 2] = giz * c.mz;
    }
  }
}

template <typename T>
static GridArgs<T> make_args(const void* input, const int64_t* sz, const int64_t* st, const void* grid, int64_t P,
                             int pad, int align) {
  GridArgs<T> a;
  a.input = (const T*)input;
  a.B = sz[0], a.C = sz[1], a.D = sz[2], a.H = sz[3], a.W = sz[4];
  a.sN = st[0], a.sC = st[1], a.sD = st[2], a.sH = st[3], a.sW = st[4];
  a.grid = (const T*)grid;
  a.P = P;
  a.pad = pad;
  a.align = align;
  return a;
}

static GStrides make_gstrides(const int64_t* st) {
  GStrides g;
  g.sN = st[0], g.sC = st[1], g.sD = st[2], g.sH = st[3], g.sW = st[4];
  return g;
}

static bool vec4_ok(int dtype, const void* base, const int64_t* sz, const int64_t* st) {
  if (dtype != MISO_F32) return false;
  if (st[1] != 1 || sz[1] % 4 != 0) return false;
  if (((uintptr_t)base) % 16 != 0) return false;
  for (int d : {0, 2, 3, 4})
    if (st[d] % 4 != 0) return false;
  return true;
}

static int validate(int dtype, const void* input, const int64_t* sz, const int64_t* st, const void* grid, int64_t P,
                    int pad) {
  MISO_REQUIRE(dtype == MISO_F32 || dtype == MISO_F64, "grid_sample3d: dtype must be f32 or f64");
  MISO_REQUIRE(input && sz && st && (grid || P == 0), "grid_sample3d: null input/grid");
  MISO_REQUIRE(pad == MISO_PAD_ZEROS || pad == MISO_PAD_BORDER, "grid_sample3d: padding_mode must be zeros|border");
  for (int d = 0; d < 5; ++d) MISO_REQUIRE(sz[d] > 0, "grid_sample3d: empty input dimension %d", d);
  MISO_REQUIRE(P >= 0, "grid_sample3d: negative point count");
  return MISO_OK;
}

}  // namespace miso

using namespace miso;

extern "C" int miso_grid_sample3d_fwd(int dtype, const void* input, const int64_t in_sizes[5],
                                      const int64_t in_strides[5], const void* grid, int64_t P, void* output,
                                      const int64_t out_strides[3], int padding_mode, int align_corners,
                                      miso_stream_t stream) {
  if (int e = validate(dtype, input, in_sizes, in_strides, grid, P, padding_mode)) return e;
  MISO_REQUIRE(output && out_strides, "grid_sample3d_fwd: null output");
  const int64_t total = in_sizes[0] * P;
  if (total == 0) return MISO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = grid_for(total, kThreads, sm_count() * 16);
  if (dtype == MISO_F32) {
    auto a = make_args<float>(input, in_sizes, in_strides, grid, P, padding_mode, align_corners);
    bool ovec = out_strides[1] != 1 || (out_strides[0] % 4 == 0 && out_strides[2] % 4 == 0 && ((uintptr_t)output) % 16 == 0);
    if (vec4_ok(dtype, input, in_sizes, in_strides) && ovec)
      grid_sample_fwd_kernel<float, 4><<<blocks, kThreads, 0, s>>>(a, (float*)output, out_strides[0], out_strides[1], out_strides[2]);
    else
      grid_sample_fwd_kernel<float, 1><<<blocks, kThreads, 0, s>>>(a, (float*)output, out_strides[0], out_strides[1], out_strides[2]);
  } else {
    auto a = make_args<double>(input, in_sizes, in_strides, grid, P, padding_mode, align_corners);
    grid_sample_fwd_kernel<double, 1><<<blocks, kThreads, 0, s>>>(a, (double*)output, out_strides[0], out_strides[1], out_strides[2]);
  }
  return check_launch("grid_sample3d_fwd");
}

extern "C" int miso_grid_sample3d_bwd(int dtype, const void* grad_output, const int64_t go_strides[3],
                                      const void* input, const int64_t in_sizes[5], const int64_t in_strides[5],
                                      const void* grid, int64_t P, void* grad_input, const int64_t gi_strides[5],
                                      void* grad_grid, int padding_mode, int align_corners, miso_stream_t stream) {
  if (int e = validate(dtype, input, in_sizes, in_strides, grid, P, padding_mode)) return e;
  MISO_REQUIRE(grad_output && go_strides, "grid_sample3d_bwd: null grad_output");
  const GStrides gs = make_gstrides(gi_strides ? gi_strides : in_strides);
  const int64_t total = in_sizes[0] * P;
  if (total == 0 || (!grad_input && !grad_grid)) return MISO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = grid_for(total, kThreads, sm_count() * 16);
  if (dtype == MISO_F32) {
    auto a = make_args<float>(input, in_sizes, in_strides, grid, P, padding_mode, align_corners);
    bool gvec = go_strides[1] != 1 || (go_strides[0] % 4 == 0 && go_strides[2] % 4 == 0 && ((uintptr_t)grad_output) % 16 == 0);
    bool ivec = !grad_input || vec4_ok(dtype, grad_input, in_sizes, gi_strides ? gi_strides : in_strides);
    if (vec4_ok(dtype, input, in_sizes, in_strides) && gvec && ivec)
      grid_sample_bwd_kernel<float, 4><<<blocks, kThreads, 0, s>>>(a, (const float*)grad_output, go_strides[0], go_strides[1], go_strides[2], (float*)grad_input, gs, (float*)grad_grid);
    else
      grid_sample_bwd_kernel<float, 1><<<blocks, kThreads, 0, s>>>(a, (const float*)grad_output, go_strides[0], go_strides[1], go_strides[2], (float*)grad_input, gs, (float*)grad_grid);
  } else {
    auto a = make_args<double>(input, in_sizes, in_strides, grid, P, padding_mode, align_corners);
    grid_sample_bwd_kernel<double, 1><<<blocks, kThreads, 0, s>>>(a, (const double*)grad_output, go_strides[0], go_strides[1], go_strides[2], (double*)grad_input, gs, (double*)grad_grid);
  }
  return check_launch("grid_sample3d_bwd");
}

extern "C" int miso_grid_sample3d_bwd_bwd(int dtype, const void* gg_input, const int64_t ggi_strides[5],
                                          const void* gg_grid, const void* grad_output, const int64_t go_strides[3],
                                          const void* input, const int64_t in_sizes[5], const int64_t in_strides[5],
                                          const void* grid, int64_t P, void* gg_output, const int64_t ggo_strides[3],
                                          void* g_input, const int64_t gi_strides[5], void* g_grid,
                                          int padding_mode, int align_corners, miso_stream_t stream) {
  if (int e = validate(dtype, input, in_sizes, in_strides, grid, P, padding_mode)) return e;
  MISO_REQUIRE(grad_output && go_strides, "grid_sample3d_bwd_bwd: null grad_output");
  MISO_REQUIRE(!gg_input || ggi_strides, "grid_sample3d_bwd_bwd: gg_input given without strides");
  MISO_REQUIRE(!gg_output || ggo_strides, "grid_sample3d_bwd_bwd: gg_output given without strides");
  const int64_t total = in_sizes[0] * P;
  if (total == 0 || (!gg_output && !g_input && !g_grid)) return MISO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = grid_for(total, kThreads, sm_count() * 16);
  static const int64_t zero5[5] = {0, 0, 0, 0, 0};
  static const int64_t zero3[3] = {0, 0, 0};
  const int64_t* is = gg_input ? ggi_strides : zero5;
  const int64_t* os = gg_output ? ggo_strides : zero3;
  const GStrides gs = make_gstrides(gi_strides ? gi_strides : in_strides);
  if (dtype == MISO_F32) {
    auto a = make_args<float>(input, in_sizes, in_strides, grid, P, padding_mode, align_corners);
    bool ivec = !g_input || vec4_ok(dtype, g_input, in_sizes, gi_strides ? gi_strides : in_strides);
    if (vec4_ok(dtype, input, in_sizes, in_strides) && ivec)
      grid_sample_bwd_bwd_kernel<float, 4><<<blocks, kThreads, 0, s>>>(a, (const float*)gg_input, is[0], is[1], is[2], is[3], is[4], (const float*)gg_grid, (const float*)grad_output, go_strides[0], go_strides[1], go_strides[2], (float*)gg_output, os[0], os[1], os[2], (float*)g_input, gs, (float*)g_grid);
    else
      grid_sample_bwd_bwd_kernel<float, 1><<<blocks, kThreads, 0, s>>>(a, (const float*)gg_input, is[0], is[1], is[2], is[3], is[4], (const float*)gg_grid, (const float*)grad_output, go_strides[0], go_strides[1], go_strides[2], (float*)gg_output, os[0], os[1], os[2], (float*)g_input, gs, (float*)g_grid);
  } else {
    auto a = make_args<double>(input, in_sizes, in_strides, grid, P, padding_mode, align_corners);
    grid_sample_bwd_bwd_kernel<double, 1><<<blocks, kThreads, 0, s>>>(a, (const double*)gg_input, is[0], is[1], is[2], is[3], is[4], (const double*)gg_grid, (const double*)grad_output, go_strides[0], go_strides[1], go_strides[2], (double*)gg_output, os[0], os[1], os[2], (double*)g_input, gs, (double*)g_grid);
  }
  return check_launch("grid_sample3d_bwd_bwd");
}
