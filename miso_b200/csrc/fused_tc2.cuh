// Mapping step, tensor-core decoder, TWO threads per point ("tc2").  Included by fused.cu.
//
// Why: the one-thread-per-point kernel (mapping_step_tc_kernel) is latency-bound -- TMEM (512 columns =
// four 128-point tiles of A_hi 64 + D 64) caps the points in flight per SM at 512, and at one thread per
// point that is only 16 warps (4 per scheduler) with ~128 live registers each; ncu shows 31 % issue-slot
// utilisation with the stalls spread over the gather, three MMA round trips and the scatter.  Here a tile
// still holds 128 points (one TMEM lane per point) but is worked on by 256 threads: thread (point, half)
// with half in {0,1} living in warps w and w+4 of the group (same TMEM lane quarter), so the SAME four tiles in
// flight give 32 warps per SM (8 per scheduler) at <= 64 registers:
//   * gather / scatter / eikonal partial: half h owns feature columns [h*F/2, (h+1)*F/2) (for the 2-level
//     C=4 grids: one level each) -- half the corner fetches, half the reductions per thread;
//   * decoder epilogues: half h owns hidden units [32h, 32h+32) of its point's row;
//   * Jacobian J = (D1 g1) W1 is a FOURTH tensor-core product (128 x 16 x 64, 3xTF32): half h masks hidden units
//     [32h, 32h+32) of g1 with its own ReLU-1 sign word and writes them back as the A operand (hi -> TMEM over
//     the dead ReLU-2 mask, lo -> shared memory), then reads back the F/2 Jacobian entries it scatters with.
//     This removes the 64 broadcast LDS.128 + 128 FFMA2 + 64 mask selects per thread of the SIMT product (the
//     shared-memory wavefronts were half of the kernel's L1TEX data-pipe load).
// Partners exchange two small items through shared memory: the partial sdf (dedicated 4 B per thread) and the
// partial grad_x sdf (16 B per thread, parked in the A_lo operand while it is idle).
// Per-tile synchronisation: four 256-thread named barriers (one before each MMA batch) + one for the
// eikonal exchange; MMA completion through one mbarrier per group, as in the one-thread kernel.
#pragma once

#ifndef MISO_PAIR_SCATTER
#define MISO_PAIR_SCATTER 1   // 0: paired gather but one-corner-per-lane reductions (measured 2.5 % slower)
#endif

namespace miso {

constexpr int kTc2MaxSmemPoses = 128;

template <int F, int G>
struct Tc2Smem {
  static constexpr int KP = ((F + 1 + 7) / 8) * 8;
  alignas(128) unsigned char w2_hi[tc::kWeightBytes];
  alignas(128) unsigned char w2_lo[tc::kWeightBytes];
  alignas(128) unsigned char w2t_hi[tc::kWeightBytes];
  alignas(128) unsigned char w2t_lo[tc::kWeightBytes];
  alignas(128) unsigned char a_lo[G][kTcABytes];
  alignas(128) unsigned char w1e_hi[H * KP * 4];   // layer 1: canonical [64][KP], row = {W1[n][0..F), b1[n], 0..}
  alignas(128) unsigned char w1e_lo[H * KP * 4];
  alignas(128) unsigned char w1j_hi[16 * tc::kK * 4];   // Jacobian product: canonical [16][64], row i = W1[:, i] (rows >= F zero)
  alignas(128) unsigned char w1j_lo[16 * tc::kK * 4];
  alignas(16) float2 ep[H];       // {b2, W3}
  alignas(16) float b3[4];        // {b3, a scale, eikonal scale, mode 4: sample count as int bits}
  uint64_t bar[G];
  uint32_t tmem_base;
  float ppx[G * 256];             // partial sdf of each thread, read by its partner
  alignas(16) float poses[kTc2MaxSmemPoses * 12];
};

__device__ __forceinline__ void group_barrier(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(256) : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- MMA issue with the descriptor words held as 32-bit halves: the K step only bumps the low word ----
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) {
  return ((smem_addr >> 4) & 0x3fffu) | (((tc::kLBO >> 4) & 0x3fffu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14); }

template <bool kAcc>
__device__ __forceinline__ void mma_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(blo), "r"(bhi), "r"(idesc), "r"(kAcc ? 1u : 0u)
      : "memory");
}
template <bool kAcc>
__device__ __forceinline__ void mma_ss_w(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                         uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 ad, {%1, %2};\n\t"
      "mov.b64 bd, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(kAcc ? 1u : 0u)
      : "memory");
}

// Gather one 4-channel group of one level (channels [ch, ch+4)) and lerp it (value + index-space derivatives).
__device__ __forceinline__ void gather_group4(const miso_level_t& lv, const CellLite& c, int ch, float* __restrict__ f,
                                              float* __restrict__ dfx, float* __restrict__ dfy,
                                              float* __restrict__ dfz) {
  float4 v[8];
  const float* base = lv.feat + ch;
  if (c.valid == 0xffu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ldg_f4(base + (c.base + corner_delta(lv, k)));
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((c.valid >> k) & 1u) v[k] = ldg_f4(base + (c.base + corner_delta(lv, k)));
    }
  }
  lerp_corners4<true>(v, c, f, dfx, dfy, dfz);
}

__device__ __forceinline__ void scatter_group4(const miso_level_t& lv, const CellLite& c, int ch, unsigned on, float a,
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
  float* gbase = lv.grad + ch;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dz = k >> 2;
    const float sz = dz ? viz : -viz;
    const float coef = fmaf(wz[dz], r[k & 3], sz * q[k & 3]);
    const unsigned ok = on & (c.valid >> k) & 1u;
    float* dst = gbase + (ok ? c.base + corner_delta(lv, k) : 0);
    red_add_f4_if(ok, dst, coef * J[0], coef * J[1], coef * J[2], coef * J[3]);
  }
}

// ---- lane-paired gather / scatter --------------------------------------------------------------------
// A 16-byte corner fetch (or reduction) per lane makes the L1TEX tag stage serve one 128-byte line per lane
// (~2 cycles per line, the kernel's busiest unit).  The two x-neighbours of a corner pair are contiguous in the
// channels-last grid (32 bytes, same line 7 times out of 8), so lanes l and l^1 cooperate: in one instruction
// lane l touches corner (dx=0) and lane l^1 corner (dx=1) of the SAME point; a second instruction does the
// other point.  Each lane therefore always works on dx = hsel (its parity) for its own point and its
// partner's, and one shuffle per word brings home the neighbour it did not fetch.  Lines per instruction halve.
// partner lane = lane ^ kPairXor.  Adjacent lanes hold consecutive samples of a batch (in the reference's RGB-D / LiDAR
// batches: consecutive samples along one ray), which very often sit in the same voxel cell -- the merged
// reductions below depend on that.
constexpr int kPairXor = 1;
__device__ __forceinline__ float4 shfl16_f4(float4 v) {
  float4 r;
  r.x = __shfl_xor_sync(0xffffffffu, v.x, kPairXor);
  r.y = __shfl_xor_sync(0xffffffffu, v.y, kPairXor);
  r.z = __shfl_xor_sync(0xffffffffu, v.z, kPairXor);
  r.w = __shfl_xor_sync(0xffffffffu, v.w, kPairXor);
  return r;
}
__device__ __forceinline__ float4 sel_f4(bool p, float4 a, float4 b) {
  return make_float4(p ? a.x : b.x, p ? a.y : b.y, p ? a.z : b.z, p ? a.w : b.w);
}

__device__ __forceinline__ void gather_group4_paired(const miso_level_t& lv, const CellLite& c, int ch, unsigned hsel,
                                                     float* __restrict__ f, float* __restrict__ dfx,
                                                     float* __restrict__ dfy, float* __restrict__ dfz) {
  const int pbase = __shfl_xor_sync(0xffffffffu, c.base, kPairXor);
  const unsigned pvalid = __shfl_xor_sync(0xffffffffu, c.valid, kPairXor);
  const int base1 = hsel ? pbase : c.base, base2 = hsel ? c.base : pbase;          // point of the LOW lane first
  const unsigned valid1 = hsel ? pvalid : c.valid, valid2 = hsel ? c.valid : pvalid;
  const float* src = lv.feat + ch + (hsel ? (int)lv.sX : 0);
  float4 l1[4], l2[4];
  const bool same_cell = base1 == base2 && valid1 == valid2;   // the pair shares all 8 corners: fetch them once
  if (__all_sync(0xffffffffu, c.valid == 0xffu)) {
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
      const int dlt = ((yz & 1) ? (int)lv.sY : 0) + ((yz & 2) ? (int)lv.sZ : 0);
      l1[yz] = ldg_f4(src + (base1 + dlt));
    }
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
      const int dlt = ((yz & 1) ? (int)lv.sY : 0) + ((yz & 2) ? (int)lv.sZ : 0);
      l2[yz] = l1[yz];
      if (!same_cell) l2[yz] = ldg_f4(src + (base2 + dlt));
    }
  } else {
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
      const int dlt = ((yz & 1) ? (int)lv.sY : 0) + ((yz & 2) ? (int)lv.sZ : 0);
      const unsigned k = 2u * yz + hsel;
      l1[yz] = make_float4(0.f, 0.f, 0.f, 0.f);
      l2[yz] = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((valid1 >> k) & 1u) l1[yz] = ldg_f4(src + (base1 + dlt));
      if ((valid2 >> k) & 1u) l2[yz] = ldg_f4(src + (base2 + dlt));
    }
  }
  // own corner (dx = hsel) stays, the partner's is sent; what arrives is the own point's corner with dx = 1 - hsel
  float4 P[4], Q[4];
#pragma unroll
  for (int yz = 0; yz < 4; ++yz) {
    P[yz] = sel_f4(hsel != 0, l2[yz], l1[yz]);
    Q[yz] = shfl16_f4(sel_f4(hsel != 0, l1[yz], l2[yz]));
  }
  // separable lerp written around the corner this lane holds: v(dx=hsel) = P, v(other) = Q
  const float fxp = hsel ? 1.0f - c.fx : c.fx;
  const float sgn = hsel ? -1.0f : 1.0f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float a[4], d[4];
#pragma unroll
    for (int yz = 0; yz < 4; ++yz) {
      const float p = reinterpret_cast<const float*>(&P[yz])[e], q = reinterpret_cast<const float*>(&Q[yz])[e];
      d[yz] = q - p;
      a[yz] = fmaf(fxp, d[yz], p);
    }
    float ay[2], ey[2], dxy[2];
#pragma unroll
    for (int z = 0; z < 2; ++z) {
      ey[z] = a[2 * z + 1] - a[2 * z];
      ay[z] = fmaf(c.fy, ey[z], a[2 * z]);
      dxy[z] = fmaf(c.fy, d[2 * z + 1] - d[2 * z], d[2 * z]);
    }
    const float ez = ay[1] - ay[0];
    f[e] = fmaf(c.fz, ez, ay[0]);
    dfz[e] = ez;
    dfy[e] = fmaf(c.fz, ey[1] - ey[0], ey[0]);
    dfx[e] = sgn * fmaf(c.fz, dxy[1] - dxy[0], dxy[0]);
  }
}

// coefficients (a*w_c + v . dw_c/di) of one point's four (dy,dz) corners with dx = hsel
__device__ __forceinline__ void corner_coefs4(unsigned hsel, float a, float vix, float viy, float viz, float fx, float fy,
                                              float fz, float (&coef)[4]) {
  const float wx = hsel ? fx : 1.0f - fx;
  const float px = fmaf(a, wx, hsel ? vix : -vix);
  float r[2], q[2];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const float wy = dy ? fy : 1.0f - fy;
    r[dy] = fmaf(wy, px, (dy ? viy : -viy) * wx);
    q[dy] = wx * wy;
  }
#pragma unroll
  for (int yz = 0; yz < 4; ++yz) {
    const int dy = yz & 1, dz = yz >> 1;
    const float wz = dz ? fz : 1.0f - fz;
    coef[yz] = fmaf(wz, r[dy], (dz ? viz : -viz) * q[dy]);
  }
}

// Lanes l and l^1 hold two consecutive samples A (even lane) and B (odd lane); each lane issues the reductions of the
// corners with dx = hsel of BOTH points.  When A and B sit in the same cell (same base, same validity -- consecutive
// ray samples usually do on the coarse level and near surfaces on the fine one) their contributions to a corner are
// summed in registers and ONE reduction is issued: the SM retires reductions at ~2 cycles per lane, so every merged
// pair saves real time, and the L2 sees fewer same-address updates.
__device__ __forceinline__ void scatter_group4_paired(const miso_level_t& lv, const CellLite& c, int ch, unsigned hsel,
                                                      unsigned on, float a, float vix, float viy, float viz,
                                                      const float* __restrict__ J) {
  const unsigned valid = on ? c.valid : 0u;
#define MISO_X16(v) __shfl_xor_sync(0xffffffffu, (v), kPairXor)
  const int pbase = MISO_X16(c.base);
  const unsigned pvalid = MISO_X16(valid);
  const float pa = MISO_X16(a), pvx = MISO_X16(vix), pvy = MISO_X16(viy), pvz = MISO_X16(viz);
  const float pfx = MISO_X16(c.fx), pfy = MISO_X16(c.fy), pfz = MISO_X16(c.fz);
  const float pJ0 = MISO_X16(J[0]), pJ1 = MISO_X16(J[1]), pJ2 = MISO_X16(J[2]), pJ3 = MISO_X16(J[3]);
#undef MISO_X16
  float* gbase = lv.grad + ch;
  const bool h = hsel != 0;
  // point 1 = the EVEN lane's (own for hsel = 0, the partner's for hsel = 1), point 2 = the odd lane's
  const int base1 = h ? pbase : c.base, base2 = h ? c.base : pbase;
  const unsigned valid1 = h ? pvalid : valid, valid2 = h ? valid : pvalid;
  float c1[4], c2[4];
  corner_coefs4(hsel, h ? pa : a, h ? pvx : vix, h ? pvy : viy, h ? pvz : viz, h ? pfx : c.fx, h ? pfy : c.fy,
                h ? pfz : c.fz, c1);
  corner_coefs4(hsel, h ? a : pa, h ? vix : pvx, h ? viy : pvy, h ? viz : pvz, h ? c.fx : pfx, h ? c.fy : pfy,
                h ? c.fz : pfz, c2);
  const float J1a[4] = {h ? pJ0 : J[0], h ? pJ1 : J[1], h ? pJ2 : J[2], h ? pJ3 : J[3]};
  const float J2a[4] = {h ? J[0] : pJ0, h ? J[1] : pJ1, h ? J[2] : pJ2, h ? J[3] : pJ3};
  const bool same = base1 == base2 && valid1 == valid2;
#pragma unroll
  for (int yz = 0; yz < 4; ++yz) {
    const int dy = yz & 1, dz = yz >> 1;
    const int dlt = (dy ? (int)lv.sY : 0) + (dz ? (int)lv.sZ : 0) + (hsel ? (int)lv.sX : 0);
    const unsigned k = 2u * yz + hsel;
    const unsigned ok1 = (valid1 >> k) & 1u;
    const unsigned ok2 = same ? 0u : (valid2 >> k) & 1u;
    const float m2 = same ? c2[yz] : 0.f;   // merged: point 2 rides on point 1's reduction
    red_add_f4_if(ok1, gbase + (ok1 ? base1 + dlt : 0), fmaf(m2, J2a[0], c1[yz] * J1a[0]), fmaf(m2, J2a[1], c1[yz] * J1a[1]),
                  fmaf(m2, J2a[2], c1[yz] * J1a[2]), fmaf(m2, J2a[3], c1[yz] * J1a[3]));
    red_add_f4_if(ok2, gbase + (ok2 ? base2 + dlt : 0), c2[yz] * J2a[0], c2[yz] * J2a[1], c2[yz] * J2a[2], c2[yz] * J2a[3]);
  }
}

// kMode: 4 = mode 0 with the number of samples read from device memory (m.cfg.n_device: batch compacted on the device);
// kMode: 3 = backward pass with a per-point cotangent given by the caller (m.a_ext) over displaced "virtual" points
// (finite-difference eikonal, miso_mapping_step_fd); mode 2 also understands virtual points.
// kMode: 0 = whole mapping step (losses + scatter), 1 = forward with Jacobian / grad_x outputs (miso_sdf_forward with
// jac / gradx), 2 = forward only (dense queries: no derivative gather, no backward products)
template <int L, int C, int G, bool kPaired, int kMode>
__global__ void __launch_bounds__(G * 256, 1)
    mapping_step_tc2_kernel(const __grid_constant__ miso_field_t fl, const __grid_constant__ miso_decoder_t dec,
                            const __grid_constant__ miso_frames_t fr, const __grid_constant__ MapArgs m) {
  constexpr int F = L * C;
  constexpr int FH = F / 2;          // feature columns per half
  constexpr int GH = FH / 4;         // 4-channel groups per half
  constexpr int CG = C / 4;          // 4-channel groups per level
  constexpr int NLH = GH >= CG ? GH / CG : 1;   // distinct levels a half touches
  constexpr int KP = Tc2Smem<F, G>::KP;
  constexpr int kThreadsCta = G * 256;
  static_assert(F % 8 == 0, "tc2 needs an even number of 4-channel groups");
  static_assert(GH % CG == 0 || CG % GH == 0, "a half must cover whole levels or a whole fraction of one level");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float red[32];
  using Smem = Tc2Smem<F, G>;
  Smem* s = reinterpret_cast<Smem*>(smem_raw);

  const int tid = threadIdx.x;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform copies (uniform datapath)
  const int grp = warp_u >> 3;
  const int half = (warp_u >> 2) & 1;
  const int gtid = tid & 255;
  const int pt = gtid & 127;
  const unsigned hsel = (unsigned)tid & (unsigned)kPairXor ? 1u : 0u;   // which x-neighbour this lane fetches for its pair

  // ---- one-time CTA setup --------------------------------------------------------------------------
  if (warp_u == 0) tc::tmem_alloc(&s->tmem_base, 512);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < G; ++i) tc::mbar_init(&s->bar[i], 1);
    tc::fence_mbar_init();
  }
  tc::stage_weights(dec.W2, false, s->w2_hi, s->w2_lo, tid, kThreadsCta);
  for (int i = tid; i < H * H; i += kThreadsCta) {
    const int k = i / H, j = i % H;
    float hi, lo;
    tc::tf32_split(dec.W3[j] * dec.W2[j * H + k], hi, lo);
    const uint32_t off = tc::b_offset(k, j);
    *reinterpret_cast<float*>(s->w2t_hi + off) = hi;
    *reinterpret_cast<float*>(s->w2t_lo + off) = lo;
  }
  for (int i = tid; i < H * KP; i += kThreadsCta) {
    const int n = i / KP, k = i % KP;
    const float w = k < F ? dec.W1[n * F + k] : (k == F ? dec.b1[n] : 0.f);
    float hi, lo;
    tc::tf32_split(w, hi, lo);
    const uint32_t off = tc::b_offset_k(n, k, KP);
    *reinterpret_cast<float*>(s->w1e_hi + off) = hi;
    *reinterpret_cast<float*>(s->w1e_lo + off) = lo;
  }
  for (int i = tid; i < 16 * H; i += kThreadsCta) {
    const int n = i / H, k = i % H;   // B[n = input i][k = hidden] = W1[k][i]
    float hi, lo;
    tc::tf32_split(n < F ? dec.W1[k * F + n] : 0.f, hi, lo);
    const uint32_t off = tc::b_offset(n, k);
    *reinterpret_cast<float*>(s->w1j_hi + off) = hi;
    *reinterpret_cast<float*>(s->w1j_lo + off) = lo;
  }
  for (int i = tid; i < H; i += kThreadsCta) s->ep[i] = make_float2(dec.b2[i], dec.W3[i]);
  if (tid == 0) {
    // loop-invariant scalars live next to b3 (one LDS.128 per tile) instead of in registers
    const bool eik_on0 = m.cfg.eik_mode != 0 && m.cfg.weight_eik != 0.f;
    const float n_den = (float)(m.cfg.n_total > 0 ? m.cfg.n_total : m.N);
    float n_eik = n_den;
    if (eik_on0 && m.cfg.eik_trunc_dist >= 0.f) n_eik = (float)(*m.eik_count);
    s->b3[0] = dec.b3[0];
    s->b3[1] = (1.0f / n_den) * m.cfg.grad_scale;                                       // a scale
    s->b3[2] = m.cfg.weight_eik * m.cfg.grad_scale * 2.f * (1.0f / n_eik);              // eikonal scale
    s->b3[3] = 0.f;
    if constexpr (kMode == 4) s->b3[3] = __int_as_float(min((int)m.N, *m.cfg.n_device));   // device-side sample count
  }
  const bool poses_in_smem = fr.ids != nullptr && fr.num_frames <= kTc2MaxSmemPoses;
  if (poses_in_smem) {
    for (int i = tid; i < fr.num_frames * 9; i += kThreadsCta) s->poses[(i / 9) * 12 + i % 9] = fr.R[i];
    for (int i = tid; i < fr.num_frames * 3; i += kThreadsCta) s->poses[(i / 3) * 12 + 9 + i % 3] = fr.t[i];
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const uint32_t tbase = s->tmem_base + (uint32_t)(grp * 128);
  const uint32_t lane_bits = (uint32_t)((warp_u & 3) * 32) << 16;
  const uint32_t a_lane = tbase + lane_bits, d_lane = tbase + 64 + lane_bits;
  unsigned char* const a_lo_row = s->a_lo[grp] + tc::a_row_offset(pt);
  float4* const xch = reinterpret_cast<float4*>(s->a_lo[grp]);   // partial grad_x sdf exchange slots (A_lo idle)
  float4* const xnbuf = reinterpret_cast<float4*>(s->a_lo[grp] + 4096);   // next tile's normalised coordinates
  uint64_t* const bar = &s->bar[grp];
  uint32_t parity = 0;

  constexpr bool kStep = kMode == 0 || kMode == 3 || kMode == 4;   // scatters gradients
  constexpr bool kExt = kMode == 3;                  // cotangent given by the caller, no loss terms
  constexpr bool kVirt = kMode == 2 || kMode == 3;   // may run over displaced virtual points
  const bool eik_on = kMode == 1 ? true : (kExt ? false : (m.cfg.eik_mode != 0 && m.cfg.weight_eik != 0.f));
  const bool eik_filter = m.cfg.eik_trunc_dist >= 0.f;

  // this half's slice of the feature vector
  const int sg0 = half * GH;            // first 4-channel group

  float acc_sdf = 0.f, acc_fs = 0.f, acc_eik = 0.f;
  const int tile_stride = (int)gridDim.x * G;   // 32-bit point indices: the host routes N >= 2^31 - 2^24 elsewhere
  // device-side sample count (batch compacted by miso_slab_select): the launch is sized for m.N, tiles past it exit.
  // Mode 0 reads the count from the constant bank; mode 4 re-reads it from shared memory (the spare word of b3, staged
  // with the decoder) at each of its three uses per tile instead of keeping it live in this 64-register kernel.
  auto n_limit = [&]() -> int {
    if constexpr (kMode == 4) return __float_as_int(*reinterpret_cast<volatile float*>(&s->b3[3]));
    else return (int)m.N;
  };
  // Point work (load, frame->world, normalise) is done ONCE per point, by half 1, one tile ahead: the result is
  // parked in the idle A_lo operand and picked up by both halves after the next group barrier, so the dependent
  // id -> pose -> transform chain and its global-load latency are off the tile's critical path.
  auto stage_point = [&](int t2) {
    const int n2 = t2 * 128 + pt;
    float p[3] = {0.f, 0.f, 0.f};
    if (n2 < n_limit()) {
      // finite-difference passes (miso_mapping_step_fd) run over 6 N "virtual" points: virtual index k N + i is
      // sample i displaced by +eps (k even) / -eps (k odd) along axis k / 2, in world coordinates (diff.py:18-26)
      int nb = n2, k = -1;
      if constexpr (kVirt) {
        if (m.fd_n > 0) {
          k = n2 / m.fd_n;
          nb = n2 - k * m.fd_n;
        }
      }
      load_point_smem(m.x, fr, nb, poses_in_smem ? s->poses : nullptr, p, (kMode == 0 || kMode == 4) ? m.poison : nullptr);
      if constexpr (kVirt) {
        if (k >= 0) {
          const float d = (k & 1) ? -m.fd_eps : m.fd_eps;
          p[0] += (k >> 1) == 0 ? d : 0.f, p[1] += (k >> 1) == 1 ? d : 0.f, p[2] += (k >> 1) == 2 ? d : 0.f;
        }
      }
      if constexpr (kMode == 1 || kMode == 2) {
        if (m.xw) m.xw[3 * (int64_t)n2] = p[0], m.xw[3 * (int64_t)n2 + 1] = p[1], m.xw[3 * (int64_t)n2 + 2] = p[2];
      }
    }
    float4 q;
    q.x = normalize_coord(p[0], fl.bound[0], fl.bound[1]);
    q.y = normalize_coord(p[1], fl.bound[2], fl.bound[3]);
    q.z = normalize_coord(p[2], fl.bound[4], fl.bound[5]);
    q.w = 0.f;
    xnbuf[pt] = q;
  };
  if (half == 1) stage_point((int)blockIdx.x * G + grp);
  group_barrier(grp);
  for (int tile = (int)blockIdx.x * G + grp; tile * 128 < n_limit(); tile += tile_stride) {
    const int n = tile * 128 + pt;
    const bool active = n < n_limit();
    float xn[3];
    {
      const float4 q = xnbuf[pt];   // normalised coordinates, produced one tile ahead by half 1 (see below)
      xn[0] = q.x, xn[1] = q.y, xn[2] = q.z;
    }
    // ---- gather this half's feature groups ---------------------------------------------------------
    float f[FH], dfx[FH], dfy[FH], dfz[FH];
    CellLite cells[NLH];
#pragma unroll
    for (int j = 0; j < GH; ++j) {
      const int sg = sg0 + j, l = sg / CG, ch = (sg % CG) * 4;
      constexpr bool kWholeLevels = GH >= CG;
      const int ci = kWholeLevels ? j / CG : 0;
      const miso_level_t& lv = fl.level[l];
      if (j == 0 || (kWholeLevels && j % CG == 0)) {
        cells[ci] = make_cell_lite(lv, xn);
        if (!active || ((fl.ignore_mask >> l) & 1u)) cells[ci].valid = 0u;   // contributes zeros, scatters nothing
      }
      CellLite cg = cells[ci];
      if (m.dbg & 2) cg.valid = 0u;
      if constexpr (kPaired) gather_group4_paired(lv, cg, ch, hsel, f + 4 * j, dfx + 4 * j, dfy + 4 * j, dfz + 4 * j);
      else gather_group4(lv, cg, ch, f + 4 * j, dfx + 4 * j, dfy + 4 * j, dfz + 4 * j);
    }
    // ---- layer 1 operand: [f, 1, 0..] split hi | lo, K = KP -----------------------------------------
#pragma unroll
    for (int j = 0; j < GH; ++j) {
      float h[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tc::tf32_split_fast(f[4 * j + i], h[i], lo[i]);
      const uint32_t col = (uint32_t)(half * FH + 4 * j);
      tmem_st4(a_lane + col, __float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]), __float_as_uint(h[3]));
      tmem_st4(a_lane + KP + col, __float_as_uint(lo[0]), __float_as_uint(lo[1]), __float_as_uint(lo[2]),
               __float_as_uint(lo[3]));
    }
    if (half == 1) {   // warp-uniform: bias column (1.0) + zero padding up to KP
#pragma unroll
      for (int c4 = 0; c4 < (KP - F) / 4; ++c4) {
        tmem_st4(a_lane + F + 4 * c4, c4 == 0 ? 0x3f800000u : 0u, 0u, 0u, 0u);
        tmem_st4(a_lane + KP + F + 4 * c4, 0u, 0u, 0u, 0u);
      }
    }
    tc::wait_st();
    tc::fence_before_sync();
    group_barrier(grp);
    if ((warp_u & 7) == 0) {
      if (elect_one()) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc();
        const uint32_t bhi = desc_hi((uint32_t)(KP / 4) * tc::kLBO);
        const uint32_t b_h = desc_lo(tc::smem_u32(s->w1e_hi)), b_l = desc_lo(tc::smem_u32(s->w1e_lo));
#pragma unroll
        for (int ks = 0; ks < KP / 8; ++ks) {
          if (ks == 0) mma_ts_w<false>(tbase + 64, tbase + ks * 8, b_h + ks * 16, bhi, idesc);
          else mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_h + ks * 16, bhi, idesc);
        }
#pragma unroll
        for (int ks = 0; ks < KP / 8; ++ks) mma_ts_w<true>(tbase + 64, tbase + KP + ks * 8, b_h + ks * 16, bhi, idesc);
#pragma unroll
        for (int ks = 0; ks < KP / 8; ++ks) mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_l + ks * 16, bhi, idesc);
        tc::mma_commit(bar);
      }
      __syncwarp();
    }
    if ((warp_u & 7) == 0) tc::mbar_wait(bar, parity);   // one warp polls the mbarrier ...
    group_barrier(grp);                                   // ... the other seven sleep here instead of spinning
    parity ^= 1;
    tc::fence_after_sync();

    // ---- epilogue 1: hidden units [32 half, 32 half + 32): relu, sign word, split -> TMEM (hi) / smem (lo) ----
    unsigned m1 = 0u;   // bit (31 - j) = sign of pre-activation 32*half + j  (set = ReLU off)
    {
      const uint32_t c0 = (uint32_t)(32 * half);
      uint32_t d[8];
      tmem_ld8(d_lane + c0, d);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::wait_ld();
        uint32_t hi[8];
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          m1 = __funnelshift_l(d[i], m1, 1);
          float h;
          tc::tf32_split_fast(fmaxf(__uint_as_float(d[i]), 0.f), h, lo[i]);
          hi[i] = __float_as_uint(h);
        }
        if (c < 3) tmem_ld8(d_lane + c0 + 8 * (c + 1), d);   // next chunk in flight while this one is stored
        tc::tmem_st8(a_lane + c0 + 8 * c, hi);
        unsigned char* row = a_lo_row + (8 * half + 2 * c) * tc::kLBO;
        *reinterpret_cast<float4*>(row) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(row + tc::kLBO) = make_float4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    tc::wait_st();
    tc::fence_proxy_async();
    tc::fence_before_sync();
    group_barrier(grp);
    if ((warp_u & 7) == 0) {
      if (elect_one()) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc();
        const uint32_t dh = desc_hi(tc::kSBO);
        const uint32_t b_h = desc_lo(tc::smem_u32(s->w2_hi)), b_l = desc_lo(tc::smem_u32(s->w2_lo));
        const uint32_t a_l = desc_lo(tc::smem_u32(s->a_lo[grp]));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks == 0) mma_ts_w<false>(tbase + 64, tbase + ks * 8, b_h + ks * 16, dh, idesc);
          else mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_h + ks * 16, dh, idesc);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ss_w<true>(tbase + 64, a_l + ks * 16, dh, b_h + ks * 16, dh, idesc);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_l + ks * 16, dh, idesc);
        tc::mma_commit(bar);
      }
      __syncwarp();
    }
    if ((warp_u & 7) == 0) tc::mbar_wait(bar, parity);   // one warp polls the mbarrier ...
    group_barrier(grp);                                   // ... the other seven sleep here instead of spinning
    parity ^= 1;
    tc::fence_after_sync();

    // ---- epilogue 2 + layer 3 (partial): bias, ReLU, W3 dot over this half's hidden units; A <- 0/1 mask ----
    float pp = 0.f;
    {
      const uint32_t c0 = (uint32_t)(32 * half);
      uint32_t d[8];
      tmem_ld8(d_lane + c0, d);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::wait_ld();
        uint32_t mk[8];
        const float4* epv = reinterpret_cast<const float4*>(s->ep + c0 + 8 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 e = epv[i];   // {b2, W3} of two consecutive hidden units
          const float h2a = __uint_as_float(d[2 * i]) + e.x, h2b = __uint_as_float(d[2 * i + 1]) + e.z;
          pp = fmaf(e.y, fmaxf(h2a, 0.f), pp);
          pp = fmaf(e.w, fmaxf(h2b, 0.f), pp);
          mk[2 * i] = h2a > 0.f ? 0x3f800000u : 0u;
          mk[2 * i + 1] = h2b > 0.f ? 0x3f800000u : 0u;
        }
        if (c < 3) tmem_ld8(d_lane + c0 + 8 * (c + 1), d);
        tc::tmem_st8(a_lane + c0 + 8 * c, mk);
      }
      s->ppx[grp * 256 + gtid] = pp;
    }
    if constexpr (kMode == 2) {
      // forward only: next tile's points, partial-sdf exchange, store
      if (half == 1) stage_point(tile + tile_stride);
      tc::wait_st();
      tc::fence_before_sync();   // the next tile's MMA overwrites D only after every thread has read h2
      group_barrier(grp);
      if (half == 0 && active) m.sdf_out[n] = (pp + s->ppx[grp * 256 + (gtid ^ 128)]) + s->b3[0];
      continue;
    }
    tc::wait_st();
    tc::fence_before_sync();
    group_barrier(grp);
    if ((warp_u & 7) == 0) {
      if (elect_one()) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc();
        const uint32_t dh = desc_hi(tc::kSBO);
        const uint32_t b_h = desc_lo(tc::smem_u32(s->w2t_hi)), b_l = desc_lo(tc::smem_u32(s->w2t_lo));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks == 0) mma_ts_w<false>(tbase + 64, tbase + ks * 8, b_h + ks * 16, dh, idesc);
          else mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_h + ks * 16, dh, idesc);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_l + ks * 16, dh, idesc);
        tc::mma_commit(bar);
      }
      __syncwarp();
    }
    // loss inputs: fetched here so their latency hides behind the MMA round trip
    float gt = 0.f, wgt = 1.f, sgn = 0.f;
    unsigned vld = 0;
    if (kStep && active) {
      if constexpr (kExt) {
        gt = ldg_early_f32(m.a_ext + n);   // backward with a given cotangent: d total / d sdf(n) arrives here
      } else {
        gt = ldg_early_f32(m.gt_sdf + n);
        vld = ldg_early_u8(m.gt_valid + n);
        sgn = ldg_early_f32(m.gt_sign + n);
        if (m.weights) wgt = ldg_early_f32(m.weights + n);
      }
    }
    if ((warp_u & 7) == 0) tc::mbar_wait(bar, parity);   // one warp polls the mbarrier ...
    group_barrier(grp);                                   // ... the other seven sleep here instead of spinning
    parity ^= 1;
    tc::fence_after_sync();

    // ---- Jacobian on the tensor core: J = (relu1' * g1) W1.  This half masks hidden units [32 half, +32) of g1
    // with its own sign word and writes them back as the A operand (hi -> TMEM, lo -> shared memory) ----------
    {
      const uint32_t c0 = (uint32_t)(32 * half);
      uint32_t d[8];
      tmem_ld8(d_lane + c0, d);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::wait_ld();
        uint32_t hi[8];
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k0 = 8 * c + i;   // bit 31-k0 of m1: set = ReLU off
          const float e = ((m1 >> (31 - k0)) & 1u) ? 0.f : __uint_as_float(d[i]);
          float h;
          tc::tf32_split_fast(e, h, lo[i]);
          hi[i] = __float_as_uint(h);
        }
        if (c < 3) tmem_ld8(d_lane + c0 + 8 * (c + 1), d);
        tc::tmem_st8(a_lane + c0 + 8 * c, hi);
        unsigned char* row = a_lo_row + (8 * half + 2 * c) * tc::kLBO;
        *reinterpret_cast<float4*>(row) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(row + tc::kLBO) = make_float4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    tc::wait_st();
    tc::fence_proxy_async();
    tc::fence_before_sync();
    group_barrier(grp);   // also: every thread has finished reading g1, D may be overwritten
    if ((warp_u & 7) == 0) {
      if (elect_one()) {
        tc::fence_after_sync();
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(tc::kTileM >> 4) << 24);
        const uint32_t dh = desc_hi(tc::kSBO);
        const uint32_t b_h = desc_lo(tc::smem_u32(s->w1j_hi)), b_l = desc_lo(tc::smem_u32(s->w1j_lo));
        const uint32_t a_l = desc_lo(tc::smem_u32(s->a_lo[grp]));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks == 0) mma_ts_w<false>(tbase + 64, tbase + ks * 8, b_h + ks * 16, dh, idesc);
          else mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_h + ks * 16, dh, idesc);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ss_w<true>(tbase + 64, a_l + ks * 16, dh, b_h + ks * 16, dh, idesc);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ts_w<true>(tbase + 64, tbase + ks * 8, b_l + ks * 16, dh, idesc);
        tc::mma_commit(bar);
      }
      __syncwarp();
    }
    if ((warp_u & 7) == 0) tc::mbar_wait(bar, parity);   // one warp polls the mbarrier ...
    group_barrier(grp);                                   // ... the other seven sleep here instead of spinning
    parity ^= 1;
    tc::fence_after_sync();
    float J[FH];
#pragma unroll
    for (int j = 0; j < GH; ++j) {
      uint32_t r0, r1, r2, r3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                   : "r"(d_lane + (uint32_t)(half * FH + 4 * j))
                   : "memory");
      tc::wait_ld();
      J[4 * j] = __uint_as_float(r0), J[4 * j + 1] = __uint_as_float(r1);
      J[4 * j + 2] = __uint_as_float(r2), J[4 * j + 3] = __uint_as_float(r3);
    }
    if (half == 1) {
      stage_point(tile + tile_stride);
    } else {
      const int n3 = n + 2 * tile_stride * 128;   // two tiles ahead: pull the per-point inputs towards this SM
      if (n3 < n_limit() && !kVirt) {
        prefetch_l1(m.x + 3 * (int64_t)n3);
        prefetch_l1(m.x + 3 * (int64_t)n3 + 2);
        if (fr.ids) prefetch_l1(fr.ids + n3);
        if constexpr (kMode == 0 || kMode == 4) {
          prefetch_l1(m.gt_sdf + n3);
          prefetch_l1(m.gt_sign + n3);
          prefetch_l1(m.gt_valid + n3);
          if (m.weights) prefetch_l1(m.weights + n3);
        }
      }
    }
    // ---- this half's share of grad_x sdf; exchange with the partner ---------------------------------
    float gxh = 0.f, gyh = 0.f, gzh = 0.f;
    if (eik_on) {
#pragma unroll
      for (int j = 0; j < GH; ++j) {
        const int lj = (sg0 + j) / CG;
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          sx = fmaf(J[4 * j + i], dfx[4 * j + i], sx);
          sy = fmaf(J[4 * j + i], dfy[4 * j + i], sy);
          sz = fmaf(J[4 * j + i], dfz[4 * j + i], sz);
        }
        gxh = fmaf(sx, m.lvl_scale[lj][0], gxh);
        gyh = fmaf(sy, m.lvl_scale[lj][1], gyh);
        gzh = fmaf(sz, m.lvl_scale[lj][2], gzh);
      }
    }
    xch[gtid] = make_float4(0.f, gxh, gyh, gzh);   // one 128-bit store (three 32-bit ones cost 3x the wavefronts)
    tc::fence_before_sync();   // the next tile's MMA overwrites D only after every thread has read g1
    group_barrier(grp);
    const float4 other = xch[gtid ^ 128];
    const float pp_other = s->ppx[grp * 256 + (gtid ^ 128)];
    // identical summation order on both partners: (half 0) + (half 1)
    const float4 sc = *reinterpret_cast<const float4*>(s->b3);   // {b3, a scale, eikonal scale, -}
    const float pred = ((half == 0 ? pp : pp_other) + (half == 0 ? pp_other : pp)) + sc.x;
    const float gx = (half == 0 ? gxh : other.y) + (half == 0 ? other.y : gxh);
    const float gy = (half == 0 ? gyh : other.z) + (half == 0 ? other.z : gyh);
    const float gz = (half == 0 ? gzh : other.w) + (half == 0 ? other.w : gzh);

    if constexpr (kMode == 1) {
      if (active) {
        if (half == 0) {
          m.sdf_out[n] = pred;
          if (m.gradx) m.gradx[3 * (int64_t)n] = gx, m.gradx[3 * (int64_t)n + 1] = gy, m.gradx[3 * (int64_t)n + 2] = gz;
        }
        if (m.jac) {
#pragma unroll
          for (int j = 0; j < GH; ++j)
            *reinterpret_cast<float4*>(m.jac + (int64_t)n * F + half * FH + 4 * j) =
                make_float4(J[4 * j], J[4 * j + 1], J[4 * j + 2], J[4 * j + 3]);
        }
      }
      continue;
    }
    // ---- loss terms (both partners derive a and v; only half 0 accumulates the sums) -------------------
    const float lw = half == 0 ? 1.f : 0.f;
    if (active && half == 0 && m.sdf_out) m.sdf_out[n] = pred;
    float a = 0.f;
    if (vld) {
      const float e = pred - gt;
      if (m.cfg.loss_type == 0) {
        acc_sdf += lw * wgt * fabsf(e);
        a = m.cfg.weight_sdf * wgt * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
      } else {
        acc_sdf += lw * wgt * e * e;
        a = m.cfg.weight_sdf * wgt * 2.f * e;
      }
    }
    if (m.cfg.weight_fs != 0.f && sgn == 1.f) {
      const float up = fmaxf(pred - gt, 0.f), lo = fmaxf(m.cfg.trunc_dist - pred, 0.f);
      acc_fs += lw * fmaxf(up, lo);
      a += up > lo ? m.cfg.weight_fs : (lo > up ? -m.cfg.weight_fs : 0.f);
    }
    a *= sc.y;
    if constexpr (kExt) a = gt;   // backward with a given cotangent (finite-difference passes): no loss terms of its own
    float v[3] = {0.f, 0.f, 0.f};
    if (eik_on) {
      const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
      const float e = nrm - 1.f;
      const bool use = active && (!eik_filter || fabsf(gt) < m.cfg.eik_trunc_dist);
      acc_eik += use ? lw * e * e : 0.f;
      const float k = (use && nrm > 0.f) ? sc.z * e / nrm : 0.f;
      v[0] = k * gx, v[1] = k * gy, v[2] = k * gz;
    }
    const unsigned nz = ((a != 0.f || v[0] != 0.f || v[1] != 0.f || v[2] != 0.f) && !(m.dbg & 1)) ? 1u : 0u;
    // ---- scatter this half's groups ------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < GH; ++j) {
      const int sg = sg0 + j, l = sg / CG, ch = (sg % CG) * 4;
      const int ci = GH >= CG ? j / CG : 0;
      const miso_level_t& lv = fl.level[l];
      const float kx = m.lvl_scale[l][0], ky = m.lvl_scale[l][1], kz = m.lvl_scale[l][2];
      if constexpr (kPaired && MISO_PAIR_SCATTER)
        scatter_group4_paired(lv, cells[ci], ch, hsel, (nz && lv.grad) ? 1u : 0u, a, v[0] * kx, v[1] * ky, v[2] * kz,
                              J + 4 * j);
      else
        scatter_group4(lv, cells[ci], ch, (nz && lv.grad) ? 1u : 0u, a, v[0] * kx, v[1] * ky, v[2] * kz, J + 4 * j);
    }
  }
  if constexpr (kMode == 0 || kMode == 4) {
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
  tc::fence_before_sync();
  __syncthreads();
  if (warp_u == 0) tc::tmem_dealloc(s->tmem_base, 512);
}

}  // namespace miso
