// Decoder-parameter gradients of the mapping step (included by fused.cu, inside namespace miso).
//
// The reference trains the decoder together with the grids when `decoder.fix: False` (grid_opt/models/grid_net.py:110,
// 126, 346-348): autograd then differentiates sdf = W3 relu(W2 relu(W1 f + b1) + b2) + b3 and -- through
// create_graph=True (diff.py:27-33) -- the eikonal term's grad_x sdf = S G^T W1^T D1 W2^T D2 W3^T w.r.t. the weights.
// With D1, D2 the ReLU patterns, u2 = D2 W3^T, u1 = D1 W2^T u2, J = W1^T u1 (= d sdf / d f), a = d total / d sdf and
// gbar = d total / d J (the eikonal cotangent pulled back through the interpolation Jacobian G), one sample adds
//
//     dW1 += u1 (a f + gbar)^T        db1 += a u1                 t1 = D1 W1 gbar
//     dW2 += u2 (a h1 + t1)^T         db2 += a u2
//     dW3 += a h2 + D2 W2 t1          db3 += a
//
// i.e. two rank-1 updates per sample: contractions over the SAMPLES, done here as small shared-memory GEMMs.  One CTA
// walks 64-sample tiles; every [sample][hidden] operand lives in shared memory with a 68-float pitch (LDS.128 conflict
// free), every thread owns a 4x4 register block of each product, and the three weight-gradient blocks stay in
// registers across all tiles of the CTA.  Per-CTA results go to a workspace row; a second kernel sums the rows into the
// caller's gradient tensors (deterministic, no atomics).  FP32 SIMT throughout: the contraction dimension is the batch,
// so the tensor-core route would need transposed operand staging per tile; the decoder is frozen in every shipped
// config, this path exists so that `fix: False` does not fall off the fused step.
namespace wgrad {

constexpr int kT = 64;      // samples per tile
constexpr int kP = H + 4;   // row pitch of the [sample][hidden] tiles (68 = 4 mod 32)

template <int F>
struct Smem {
  float W2[H * kP];   // [j][k]
  float A[kT * kP];   // h1, later p = a h1 + t1
  float B[kT * kP];   // h2 (its sign pattern is D2)
  float U[kT * kP];   // u1
  float T[kT * kP];   // t1
  float W1[H * (F + 4)];   // [k][i], pitch F + 4 (rows k, k+1, ... land in different banks for LDS.128)
  float b1[H], b2[H], W3[H];
  float f[kT * F], dx[kT * F], dy[kT * F], dz[kT * F];   // features and their world-space derivatives
  float q[kT * F];      // J, later q = a f + gbar
  float gbar[kT * F];
  float sdf[kT], a[kT];
  float red[64];
  float b3;
};

template <int F>
__host__ __device__ constexpr int param_count() { return H * F + H + H * H + H + H + 1; }
template <int F>
__host__ __device__ constexpr int row_stride() { return (param_count<F>() + 63) / 64 * 64; }

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float c) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, c))));
}
__device__ __forceinline__ const float4& f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

}  // namespace wgrad

template <int L, int C>
__global__ void __launch_bounds__(kThreads, 2)
    mapping_wgrad_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                         const __grid_constant__ miso_frames_t fr, const __grid_constant__ MapArgs m,
                         float* __restrict__ rows) {
  using namespace wgrad;
  constexpr int F = L * C, FI = F / 4, kW = F + 4;
  static_assert(L <= 4 && kT * L <= kThreads && F % 4 == 0, "tile gather uses one thread per (sample, level)");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<F>& s = *reinterpret_cast<Smem<F>*>(smem_raw);
  const int t = threadIdx.x, tc = t & 15, tn = t >> 4;

  for (int i = t; i < H * H; i += kThreads) s.W2[(i >> 6) * kP + (i & 63)] = dec.W2[i];
  for (int i = t; i < H * F; i += kThreads) s.W1[(i / F) * kW + i % F] = dec.W1[i];
  if (t < H) s.b1[t] = dec.b1[t], s.b2[t] = dec.b2[t], s.W3[t] = dec.W3[t];
  if (t == 0) s.b3 = dec.b3[0];

  const FieldGeom g = field_geom(fl);
  const bool eik_on = m.cfg.eik_mode != 0 && m.cfg.weight_eik != 0.f;
  const bool eik_filter = m.cfg.eik_trunc_dist >= 0.f;
  const float n_den = (float)(m.cfg.n_total > 0 ? m.cfg.n_total : m.N);
  const float invN = 1.0f / n_den;
  float n_eik = n_den;
  if (eik_on && eik_filter) n_eik = (float)(*m.eik_count);
  const float inv_neik = 1.0f / n_eik;

  float acc2[4][4], acc1[FI], acc3[4], accb2[4], accb1 = 0.f, accb3 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc3[i] = accb2[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < FI; ++i) acc1[i] = 0.f;

  const int64_t ntiles = (m.N + kT - 1) / kT;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n0 = tile * kT;
    __syncthreads();   // the previous tile's readers are done; first trip: the decoder is staged

    // ---- 1. gather: one thread per (sample, level) ---------------------------------------------------------------
    if (t < kT * L) {
      const int pn = t / L, l = t % L;
      const int64_t n = n0 + pn;
      float f[C], dfx[C], dfy[C], dfz[C];
#pragma unroll
      for (int i = 0; i < C; ++i) f[i] = dfx[i] = dfy[i] = dfz[i] = 0.f;
      if (n < m.N && !((fl.ignore_mask >> l) & 1u)) {
        float p[3], xn[3];
        load_point(m.x, fr, n, p);
#pragma unroll
        for (int d = 0; d < 3; ++d) xn[d] = normalize_coord(p[d], g.bmin[d], g.bmax[d]);
        const Cell c = level_cell(fl.level[l], xn);
        gather_level<C, true>(fl.level[l], c, f, dfx, dfy, dfz);
      }
      const float kx = (float)fl.level[l].X * g.inv_len[0], ky = (float)fl.level[l].Y * g.inv_len[1],
                  kz = (float)fl.level[l].Z * g.inv_len[2];
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const int o = pn * F + l * C + i;
        s.f[o] = f[i], s.dx[o] = dfx[i] * kx, s.dy[o] = dfy[i] * ky, s.dz[o] = dfz[i] * kz;
      }
    }
    __syncthreads();

    // ---- 2. layer 1: h1[n][k], n = 4 tn + r, k = tc + 16 q -------------------------------------------------------
    {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = s.b1[tc + 16 * q];
#pragma unroll
      for (int i4 = 0; i4 < FI; ++i4) {
        float4 xa[4], wb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xa[r] = f4(s.f + (4 * tn + r) * F + 4 * i4);
#pragma unroll
        for (int q = 0; q < 4; ++q) wb[q] = f4(s.W1 + (tc + 16 * q) * kW + 4 * i4);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = dot4(xa[r], wb[q], acc[r][q]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) s.A[(4 * tn + r) * kP + tc + 16 * q] = fmaxf(acc[r][q], 0.f);
    }
    __syncthreads();

    // ---- 3. layer 2: h2[n][c], n = 4 tn + r, c = tc + 16 q -------------------------------------------------------
    {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = s.b2[tc + 16 * q];
#pragma unroll 4
      for (int k4 = 0; k4 < H / 4; ++k4) {
        float4 xa[4], wb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xa[r] = f4(s.A + (4 * tn + r) * kP + 4 * k4);
#pragma unroll
        for (int q = 0; q < 4; ++q) wb[q] = f4(s.W2 + (tc + 16 * q) * kP + 4 * k4);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = dot4(xa[r], wb[q], acc[r][q]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) s.B[(4 * tn + r) * kP + tc + 16 * q] = fmaxf(acc[r][q], 0.f);
    }
    __syncthreads();

    // ---- 3b. sdf[n] = W3 h2 + b3: four threads per sample ---------------------------------------------------------
    {
      const int pn = t >> 2, qd = t & 3;
      float sum = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) sum = dot4(f4(s.B + pn * kP + 16 * qd + 4 * j4), f4(s.W3 + 16 * qd + 4 * j4), sum);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (qd == 0) s.sdf[pn] = sum + s.b3;
    }
    // ---- 4. u1[n][k] = D1 sum_j u2[n][j] W2[j][k], n = 4 tn + r, k = 4 tc + kk -----------------------------------
    {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) acc[r][kk] = 0.f;
#pragma unroll 2
      for (int j4 = 0; j4 < H / 4; ++j4) {
        const float4 w3 = f4(s.W3 + 4 * j4);
        float u2[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 xb = f4(s.B + (4 * tn + r) * kP + 4 * j4);
          u2[r][0] = xb.x > 0.f ? w3.x : 0.f, u2[r][1] = xb.y > 0.f ? w3.y : 0.f;
          u2[r][2] = xb.z > 0.f ? w3.z : 0.f, u2[r][3] = xb.w > 0.f ? w3.w : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 w = f4(s.W2 + (4 * j4 + jj) * kP + 4 * tc);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[r][0] = fmaf(u2[r][jj], w.x, acc[r][0]), acc[r][1] = fmaf(u2[r][jj], w.y, acc[r][1]);
            acc[r][2] = fmaf(u2[r][jj], w.z, acc[r][2]), acc[r][3] = fmaf(u2[r][jj], w.w, acc[r][3]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 h = f4(s.A + (4 * tn + r) * kP + 4 * tc);
        *reinterpret_cast<float4*>(s.U + (4 * tn + r) * kP + 4 * tc) =
            make_float4(h.x > 0.f ? acc[r][0] : 0.f, h.y > 0.f ? acc[r][1] : 0.f, h.z > 0.f ? acc[r][2] : 0.f,
                        h.w > 0.f ? acc[r][3] : 0.f);
      }
    }
    __syncthreads();

    // ---- 5. J[n][i] = sum_k W1[k][i] u1[n][k]: four threads per sample, F/4 inputs each ---------------------------
    {
      const int pn = t >> 2, i0 = (t & 3) * FI;
      float j[FI];
#pragma unroll
      for (int ii = 0; ii < FI; ++ii) j[ii] = 0.f;
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        const float u = s.U[pn * kP + k];
#pragma unroll
        for (int ii = 0; ii < FI; ++ii) j[ii] = fmaf(s.W1[k * kW + i0 + ii], u, j[ii]);
      }
#pragma unroll
      for (int ii = 0; ii < FI; ++ii) s.q[pn * F + i0 + ii] = j[ii];
    }
    __syncthreads();

    // ---- 6. loss cotangents of one sample (same terms as mapping_step_kernel above) -------------------------------
    if (t < kT) {
      const int64_t n = n0 + t;
      float a = 0.f, v[3] = {0.f, 0.f, 0.f};
      if (n < m.N) {
        const float pred = s.sdf[t], gt = m.gt_sdf[n];
        if (m.gt_valid[n]) {
          const float w = m.weights ? m.weights[n] : 1.f;
          const float e = pred - gt;
          if (m.cfg.loss_type == 0) a += m.cfg.weight_sdf * w * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
          else a += m.cfg.weight_sdf * w * 2.f * e;
        }
        if (m.cfg.weight_fs != 0.f && m.gt_sign[n] == 1.f) {
          const float up = fmaxf(pred - gt, 0.f), lo = fmaxf(m.cfg.trunc_dist - pred, 0.f);
          if (up > lo) a += m.cfg.weight_fs;
          else if (lo > up) a -= m.cfg.weight_fs;
        }
        a *= invN * m.cfg.grad_scale;
        if (eik_on && (!eik_filter || fabsf(gt) < m.cfg.eik_trunc_dist)) {
          float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
          for (int i = 0; i < F; ++i) {
            const float ji = s.q[t * F + i];
            gx = fmaf(ji, s.dx[t * F + i], gx), gy = fmaf(ji, s.dy[t * F + i], gy), gz = fmaf(ji, s.dz[t * F + i], gz);
          }
          const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
          if (nrm > 0.f) {
            const float k = m.cfg.weight_eik * m.cfg.grad_scale * 2.f * (nrm - 1.f) * inv_neik / nrm;
            v[0] = k * gx, v[1] = k * gy, v[2] = k * gz;
          }
        }
      }
      s.a[t] = a;
      accb3 += a;
#pragma unroll
      for (int i = 0; i < F; ++i) {
        const float gb = fmaf(v[2], s.dz[t * F + i], fmaf(v[1], s.dy[t * F + i], v[0] * s.dx[t * F + i]));
        s.gbar[t * F + i] = gb;
        s.q[t * F + i] = fmaf(a, s.f[t * F + i], gb);
      }
    }
    __syncthreads();

    // ---- 7. t1[n][k] = D1 sum_i W1[k][i] gbar[n][i], k = tc + 16 q;  p = a h1 + t1 (over h1) --------------------
    {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < FI; ++i4) {
        float4 xa[4], wb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xa[r] = f4(s.gbar + (4 * tn + r) * F + 4 * i4);
#pragma unroll
        for (int q = 0; q < 4; ++q) wb[q] = f4(s.W1 + (tc + 16 * q) * kW + 4 * i4);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = dot4(xa[r], wb[q], acc[r][q]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float an = s.a[4 * tn + r];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int o = (4 * tn + r) * kP + tc + 16 * q;
          const float h = s.A[o];
          const float t1 = h > 0.f ? acc[r][q] : 0.f;
          s.T[o] = t1;
          s.A[o] = fmaf(an, h, t1);
        }
      }
    }
    __syncthreads();

    // ---- 8. dW3[c] += sum_n a h2 + D2 (W2 t1), n = 4 tn + r, c = tc + 16 q ----------------------------------------
    {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
#pragma unroll 4
      for (int k4 = 0; k4 < H / 4; ++k4) {
        float4 xa[4], wb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xa[r] = f4(s.T + (4 * tn + r) * kP + 4 * k4);
#pragma unroll
        for (int q = 0; q < 4; ++q) wb[q] = f4(s.W2 + (tc + 16 * q) * kP + 4 * k4);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = dot4(xa[r], wb[q], acc[r][q]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float an = s.a[4 * tn + r];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float h2 = s.B[(4 * tn + r) * kP + tc + 16 * q];
          acc3[q] += fmaf(an, h2, h2 > 0.f ? acc[r][q] : 0.f);
        }
      }
    }
    // ---- 9. dW2[j][k] += sum_n u2[n][j] p[n][k], j = 4 tn + jj, k = 4 tc + kk;  db2[j] += sum_n a u2 --------------
    {
      const float4 w3 = f4(s.W3 + 4 * tn);
#pragma unroll 4
      for (int n = 0; n < kT; ++n) {
        const float4 xb = f4(s.B + n * kP + 4 * tn), pa = f4(s.A + n * kP + 4 * tc);
        const float an = s.a[n];
        const float u2[4] = {xb.x > 0.f ? w3.x : 0.f, xb.y > 0.f ? w3.y : 0.f, xb.z > 0.f ? w3.z : 0.f,
                             xb.w > 0.f ? w3.w : 0.f};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          acc2[jj][0] = fmaf(u2[jj], pa.x, acc2[jj][0]), acc2[jj][1] = fmaf(u2[jj], pa.y, acc2[jj][1]);
          acc2[jj][2] = fmaf(u2[jj], pa.z, acc2[jj][2]), acc2[jj][3] = fmaf(u2[jj], pa.w, acc2[jj][3]);
          accb2[jj] = fmaf(an, u2[jj], accb2[jj]);
        }
      }
    }
    // ---- 10. dW1[k][i] += sum_n u1[n][k] q[n][i], k = t / 4, i = (t % 4) F/4 ...;  db1[k] += sum_n a u1 -----------
    {
      const int k = t >> 2, i0 = (t & 3) * FI;
#pragma unroll 4
      for (int n = 0; n < kT; ++n) {
        const float u = s.U[n * kP + k];
#pragma unroll
        for (int ii = 0; ii < FI; ++ii) acc1[ii] = fmaf(u, s.q[n * F + i0 + ii], acc1[ii]);
        accb1 = fmaf(s.a[n], u, accb1);
      }
    }
  }

  // ---- this CTA's row: dW1 [H F] | db1 [H] | dW2 [H H] | db2 [H] | dW3 [H] | db3 ----------------------------------
  float* o = rows + (size_t)blockIdx.x * row_stride<F>();
  {
    const int k = t >> 2, i0 = (t & 3) * FI;
#pragma unroll
    for (int ii = 0; ii < FI; ++ii) o[k * F + i0 + ii] = acc1[ii];
    if ((t & 3) == 0) o[H * F + k] = accb1;
  }
  float* o2 = o + H * F + H;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    *reinterpret_cast<float4*>(o2 + (4 * tn + jj) * H + 4 * tc) =
        make_float4(acc2[jj][0], acc2[jj][1], acc2[jj][2], acc2[jj][3]);
    if (tc == 0) o2[H * H + 4 * tn + jj] = accb2[jj];
  }
  __syncthreads();
  if (t < 64) s.red[t] = 0.f;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) atomicAdd(&s.red[tc + 16 * q], acc3[q]);
  __syncthreads();
  if (t < 64) o2[H * H + H + t] = s.red[t];
  const float sb3 = block_sum(accb3, s.red);
  if (t == 0) o2[H * H + H + H] = sb3;
}

// Sums the per-CTA rows and ACCUMULATES them into the caller's gradient tensors.
template <int F>
__global__ void __launch_bounds__(kThreads)
    mapping_wgrad_finalize_kernel(const float* __restrict__ rows, int nrows, miso_decoder_grad_t g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wgrad::param_count<F>()) return;
  float sum = 0.f;
  for (int r = 0; r < nrows; ++r) sum += rows[(size_t)r * wgrad::row_stride<F>() + i];
  int o = i;
  float* dst;
  if (o < H * F) dst = g.W1;
  else if ((o -= H * F) < H) dst = g.b1;
  else if ((o -= H) < H * H) dst = g.W2;
  else if ((o -= H * H) < H) dst = g.b2;
  else if ((o -= H) < H) dst = g.W3;
  else { o -= H; dst = g.b3; }
  if (dst) dst[o] += sum;
}
