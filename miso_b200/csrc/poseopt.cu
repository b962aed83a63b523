// Pose glue of the batched alignment iteration as three tiny kernels instead of ~100 torch launches.
//
// One iteration of generic_align_multiple_submaps (grid_opt/align/base.py:127-159) around the alignment kernel is:
//   updated_submap_pose for every submap   R = R0 Exp(w), t = t0 + tau     (grid_atlas.py:250-268, utils_geometry.py:78-99)
//   per pair A1 = R_s, b1 = t_s, A2 = R_d^T, b2 = -R_d^T t_d                (utils_geometry.py:214-240)
//   [intersection test + alignment kernel: align.cu]
//   loss_i = mean(r^2) * weight, nan_to_num, total = sum_i loss_i           (miso.py:200-201, base.py:139-146)
//   backward to (w, tau) of every submap through the transforms and so3_exp_map
//   torch.optim.Adam step on (w_i, tau_i), i >= 1                           (base.py:104-111, 147-149)
// With level-0 alignment (32 k samples per pair) the alignment kernel itself takes ~0.1 ms while those torch ops and
// their autograd cost ~1 ms per iteration even inside a CUDA graph.  Here: compose (1 block), pose gradients from the
// kernel's float64 reductions (1 block), Adam (1 block); a multi-GPU run puts one all_reduce of the (S,6) gradient
// buffer between the last two.  so3_exp_map follows pytorch3d's formula (theta = sqrt(clamp(|w|^2, 1e-4))) and its
// derivative is the closed form of that expression (zero d theta / d w inside the clamp, as autograd gives).
#include "common.cuh"

namespace miso {

constexpr int kMaxSubmaps = 64;

struct PoseTables {
  const float* R0;        // (S,9) initial rotations
  const float* t0;        // (S,3)
  float* const* w;        // S pointers to the (1,3) rotation corrections (torch Parameters, updated in place)
  float* const* tau;      // S pointers to the (3,1) translation corrections
  int S;
};

__device__ __forceinline__ void so3_exp_f32(const float w[3], float E[9]) {
  // same operation order as geometry.so3_exp_map / pytorch3d in fp32
  const float n = __fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2]));
  const float th = sqrtf(fmaxf(n, 1e-4f));
  const float inv = 1.0f / th;
  const float f1 = inv * sinf(th);
  const float f2 = inv * inv * (1.0f - cosf(th));
  const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
  float K2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) E[i] = (f1 * K[i] + f2 * K2[i]) + ((i % 4 == 0) ? 1.0f : 0.0f);
}

__global__ void __launch_bounds__(256) compose_poses_kernel(PoseTables T, const int32_t* __restrict__ src,
                                                            const int32_t* __restrict__ dst, int P,
                                                            float* __restrict__ poses24, float* __restrict__ Rt) {
  __shared__ float sR[kMaxSubmaps][9];
  __shared__ float st[kMaxSubmaps][3];
  for (int s = threadIdx.x; s < T.S; s += blockDim.x) {
    const float w[3] = {T.w[s][0], T.w[s][1], T.w[s][2]};
    float E[9];
    so3_exp_f32(w, E);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* r = T.R0 + s * 9 + 3 * i;
        sR[s][3 * i + j] = r[0] * E[j] + r[1] * E[3 + j] + r[2] * E[6 + j];
      }
#pragma unroll
    for (int i = 0; i < 3; ++i) st[s][i] = T.t0[s * 3 + i] + T.tau[s][i];
    if (Rt) {
#pragma unroll
      for (int i = 0; i < 9; ++i) Rt[s * 12 + i] = sR[s][i];
#pragma unroll
      for (int i = 0; i < 3; ++i) Rt[s * 12 + 9 + i] = st[s][i];
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int s = src[p], d = dst[p];
    float* o = poses24 + (int64_t)p * 24;
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = sR[s][i];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[9 + i] = st[s][i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) o[12 + 3 * i + j] = sR[d][3 * j + i];   // A2 = R_d^T
#pragma unroll
    for (int i = 0; i < 3; ++i)   // b2 = -(A2 t_d)
      o[21 + i] = -(sR[d][i] * st[d][0] + sR[d][3 + i] * st[d][1] + sR[d][6 + i] * st[d][2]);
  }
}

// From the alignment kernel's reductions (P, MISO_ALIGN_OUT) to d total / d (w_s, tau_s) of every submap.
__global__ void __launch_bounds__(256)
    pose_grads_kernel(PoseTables T, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int P,
                      const double* __restrict__ out, const float* __restrict__ poses24, const float* __restrict__ Rt, int K,
                      float align_weight,
                      float* __restrict__ grads /* (S,6): dw, dtau */, float* __restrict__ loss_hist,
                      int32_t* __restrict__ iter_counter, float* __restrict__ pair_loss /* optional (P) */,
                      float* __restrict__ contrib /* optional (S): pairs that gave submap s a gradient */) {
  __shared__ double dR[kMaxSubmaps][9];
  __shared__ double dt[kMaxSubmaps][3];
  __shared__ int nc[kMaxSubmaps];
  __shared__ double total;
  for (int i = threadIdx.x; i < T.S; i += blockDim.x) nc[i] = 0;
  for (int i = threadIdx.x; i < T.S * 9; i += blockDim.x) dR[i / 9][i % 9] = 0.0;
  for (int i = threadIdx.x; i < T.S * 3; i += blockDim.x) dt[i / 3][i % 3] = 0.0;
  if (threadIdx.x == 0) total = 0.0;
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const double* o = out + (int64_t)p * MISO_ALIGN_OUT;
    const double S = o[0], cnt = o[1];
    // mean((f_s - f_d)^2) * weight over M_valid x K elements; 0 when nothing is valid (miso.py:180-182, 200-201)
    const double scale = cnt > 0.0 ? (double)align_weight / fmax(cnt * (double)K, 1.0) : 0.0;
    float loss = (float)(S * scale);
    if (!isfinite(loss)) loss = isnan(loss) ? 0.f : (loss > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f);  // nan_to_num
    if (pair_loss) pair_loss[p] = loss;
    atomicAdd(&total, (double)loss);
    if (scale == 0.0 || !isfinite((float)(S * scale))) continue;
    const double* G0 = o + 2;
    const double* G1 = o + 5;
    const double* G2 = o + 14;
    const float* A2f = poses24 + (int64_t)p * 24 + 12;
    double A2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) A2[i] = (double)A2f[i];
    const int s = src[p], d = dst[p];
    atomicAdd(&nc[s], 1);
    atomicAdd(&nc[d], 1);
    // dA1 = A2^T G2 s ; db1 = A2^T G0 s ; dA2 = G1 s ; db2 = G0 s
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double dA1 = (A2[i] * G2[j] + A2[3 + i] * G2[3 + j] + A2[6 + i] * G2[6 + j]) * scale;
        atomicAdd(&dR[s][3 * i + j], dA1);                       // A1 = R_s
      }
      const double db1 = (A2[i] * G0[0] + A2[3 + i] * G0[1] + A2[6 + i] * G0[2]) * scale;
      atomicAdd(&dt[s][i], db1);                                 // b1 = t_s
    }
    // A2 = R_d^T and b2 = -A2 t_d:  dA2_total = G1 s - db2 t_d^T ; dR_d = dA2_total^T ; dt_d = -A2^T db2 = -R_d db2
    const double td[3] = {(double)Rt[d * 12 + 9], (double)Rt[d * 12 + 10], (double)Rt[d * 12 + 11]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double db2 = G0[i] * scale;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double dA2 = G1[3 * i + j] * scale - db2 * td[j];
        atomicAdd(&dR[d][3 * j + i], dA2);                       // transpose
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double v = -(A2[i] * G0[0] + A2[3 + i] * G0[1] + A2[6 + i] * G0[2]) * scale;   // -(A2^T db2)_i
      atomicAdd(&dt[d][i], v);
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < T.S; s += blockDim.x) {
    // R = R0 E  =>  dE = R0^T dR
    double dE[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        dE[3 * i + j] = (double)T.R0[s * 9 + i] * dR[s][j] + (double)T.R0[s * 9 + 3 + i] * dR[s][3 + j] +
                        (double)T.R0[s * 9 + 6 + i] * dR[s][6 + j];
    const double w[3] = {(double)T.w[s][0], (double)T.w[s][1], (double)T.w[s][2]};
    const double n = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const bool clamped = n < 1e-4;            // torch.clamp passes the gradient where n >= min
    const double th = sqrt(clamped ? 1e-4 : n);
    const double sn = sin(th), cs = cos(th);
    const double f1 = sn / th, f2 = (1.0 - cs) / (th * th);
    const double df1 = (th * cs - sn) / (th * th);                         // d f1 / d theta
    const double df2 = (th * sn - 2.0 * (1.0 - cs)) / (th * th * th);      // d f2 / d theta
    const double Km[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
    double K2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) K2[3 * i + j] = Km[3 * i] * Km[j] + Km[3 * i + 1] * Km[3 + j] + Km[3 * i + 2] * Km[6 + j];
    double dEK = 0.0, dEK2 = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) dEK += dE[i] * Km[i], dEK2 += dE[i] * K2[i];
    for (int a = 0; a < 3; ++a) {
      // H_a = hat(e_a)
      double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (a == 0) H[5] = -1.0, H[7] = 1.0;
      if (a == 1) H[2] = 1.0, H[6] = -1.0;
      if (a == 2) H[1] = -1.0, H[3] = 1.0;
      double HK[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          HK[3 * i + j] = (H[3 * i] * Km[j] + H[3 * i + 1] * Km[3 + j] + H[3 * i + 2] * Km[6 + j]) +
                          (Km[3 * i] * H[j] + Km[3 * i + 1] * H[3 + j] + Km[3 * i + 2] * H[6 + j]);
      double g = 0.0;
      for (int i = 0; i < 9; ++i) g += dE[i] * (f1 * H[i] + f2 * HK[i]);
      const double dth = clamped ? 0.0 : w[a] / th;
      g += dth * (df1 * dEK + df2 * dEK2);
      grads[s * 6 + a] = (float)g;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) grads[s * 6 + 3 + i] = (float)dt[s][i];     // t = t0 + tau
    if (contrib) contrib[s] = (float)nc[s];
  }
  if (threadIdx.x == 0 && loss_hist && iter_counter) loss_hist[*iter_counter] = (float)total;
}

// torch.optim.Adam (single-tensor semantics, no amsgrad / weight decay) on (w_s, tau_s) of submaps 1..S-1.
// With `contrib` / `submap_steps`: a submap no pair gave a gradient to in this iteration (contrib[s] == 0) is skipped --
// no moment decay, no step increment, no momentum drift -- exactly like a parameter whose .grad is None in
// torch.optim.Adam (the reference's loss simply does not depend on it then), and bias correction uses the submap's own
// step count.
__global__ void __launch_bounds__(256)
    pose_adam_kernel(PoseTables T, const float* __restrict__ grads, float* __restrict__ m, float* __restrict__ v,
                     int32_t* __restrict__ iter_counter, float lr, float b1, float b2, float eps,
                     const float* __restrict__ contrib, int32_t* __restrict__ submap_steps) {
  const int global_step = *iter_counter + 1;
  for (int i = threadIdx.x; i < T.S * 6; i += blockDim.x) {
    const int s = i / 6, c = i % 6;
    if (s == 0) continue;                       // submap 0 stays fixed (base.py:104-108)
    if (contrib && contrib[s] == 0.f) continue;
    const int step = submap_steps ? submap_steps[s] + 1 : global_step;
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    float* p = c < 3 ? &T.w[s][c] : &T.tau[s][c - 3];
    const float g = grads[i];
    const float mi = m[i] + (g - m[i]) * (1.f - b1);
    const float vi = v[i] * b2 + (1.f - b2) * g * g;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    *p = *p - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
  __syncthreads();
  if (submap_steps)
    for (int s = 1 + threadIdx.x; s < T.S; s += blockDim.x)
      if (!contrib || contrib[s] != 0.f) submap_steps[s] += 1;
  if (threadIdx.x == 0) *iter_counter = global_step;
}

}  // namespace miso

using namespace miso;

static int check_tables(int S, const void* R0, const void* t0, const void* w, const void* tau) {
  MISO_REQUIRE(S >= 1 && S <= kMaxSubmaps, "align pose glue: num_submaps %d not in [1,%d]", S, kMaxSubmaps);
  MISO_REQUIRE(R0 && t0 && w && tau, "align pose glue: null pose table");
  return MISO_OK;
}

extern "C" int miso_align_compose_poses(const float* R0, const float* t0, float* const* w_ptrs, float* const* tau_ptrs,
                                        int32_t num_submaps, const int32_t* src, const int32_t* dst,
                                        int32_t num_pairs, float* poses24, float* Rt_out, miso_stream_t stream) {
  if (int e = check_tables(num_submaps, R0, t0, w_ptrs, tau_ptrs)) return e;
  MISO_REQUIRE(num_pairs >= 0 && (num_pairs == 0 || (src && dst && poses24)), "align_compose_poses: null pair arrays");
  PoseTables T{R0, t0, w_ptrs, tau_ptrs, num_submaps};
  compose_poses_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(T, src, dst, num_pairs, poses24, Rt_out);
  return check_launch("align_compose_poses");
}

extern "C" int miso_align_pose_grads(const float* R0, const float* t0, float* const* w_ptrs, float* const* tau_ptrs,
                                     int32_t num_submaps, const int32_t* src, const int32_t* dst, int32_t num_pairs,
                                     const double* align_out, const float* poses24, const float* Rt, int32_t channels_used,
                                     float align_weight, float* grads, float* loss_hist, int32_t* iter_counter,
                                     float* pair_loss, float* contrib, miso_stream_t stream) {
  if (int e = check_tables(num_submaps, R0, t0, w_ptrs, tau_ptrs)) return e;
  MISO_REQUIRE(grads && Rt && (num_pairs == 0 || (src && dst && align_out && poses24)), "align_pose_grads: null argument");
  MISO_REQUIRE(channels_used > 0, "align_pose_grads: channels_used must be positive");
  PoseTables T{R0, t0, w_ptrs, tau_ptrs, num_submaps};
  pose_grads_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(T, src, dst, num_pairs, align_out, poses24, Rt, channels_used,
                                                       align_weight, grads, loss_hist, iter_counter, pair_loss, contrib);
  return check_launch("align_pose_grads");
}

extern "C" int miso_align_pose_adam(float* const* w_ptrs, float* const* tau_ptrs, int32_t num_submaps,
                                    const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* iter_counter,
                                    float lr, float beta1, float beta2, float eps, const float* contrib,
                                    int32_t* submap_steps, miso_stream_t stream) {
  MISO_REQUIRE(num_submaps >= 1 && num_submaps <= kMaxSubmaps, "align_pose_adam: num_submaps %d not in [1,%d]",
               num_submaps, kMaxSubmaps);
  MISO_REQUIRE(w_ptrs && tau_ptrs && grads && exp_avg && exp_avg_sq && iter_counter, "align_pose_adam: null argument");
  PoseTables T{nullptr, nullptr, w_ptrs, tau_ptrs, num_submaps};
  pose_adam_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(T, grads, exp_avg, exp_avg_sq, iter_counter, lr, beta1, beta2, eps,
                                                        contrib, submap_steps);
  return check_launch("align_pose_adam");
}
