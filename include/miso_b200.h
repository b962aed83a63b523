/*
 * miso_b200.h -- C-ABI of the B200-native (sm_100a) MISO hot path.
 *
 * Plain C: raw device pointers, sizes, element strides, a cudaStream_t (passed as void*), int
 * return code (0 = ok, <0 = error; text via miso_last_error_string()).  No torch types, no
 * exceptions, no allocation: the caller owns every buffer.  Scatter targets ("grad" buffers)
 * are ACCUMULATED into (the caller pre-zeroes them or lets gradients accumulate), which removes
 * the grid-sized zeros_like the reference performs on every call
 * (third_party/cuda_gridsample_grad2/gridsample_cuda.cu:620-622).
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference
 * repository root).  INTEGRATION.md shows the ctypes binding a maintainer of the reference adds.
 */
#ifndef MISO_B200_H_
#define MISO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MISO_ABI_VERSION 1
#define MISO_MAX_LEVELS 4          /* grid.n_levels; shipped configs use 2 (configs/rgbd/scannet.yaml:24) */
#define MISO_MAX_POSES 1024

/* error codes */
#define MISO_OK 0
#define MISO_ERR_INVALID_ARG (-1)
#define MISO_ERR_UNSUPPORTED (-2)
#define MISO_ERR_CUDA (-3)

/* dtype codes for the generic grid_sample entry points */
#define MISO_F32 0
#define MISO_F64 1

/* padding_mode codes -- same indices as cuda_gridsample.py:87 (['zeros','border'].index) */
#define MISO_PAD_ZEROS 0
#define MISO_PAD_BORDER 1

typedef void* miso_stream_t; /* cudaStream_t */

/* Return text of the last error raised on the calling thread ("" if none). */
const char* miso_last_error_string(void);
int miso_abi_version(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int miso_device_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * 1. Generic trilinear grid_sample (the reference's plugin point).
 *
 * Replaces cu.grid_sample_3d / _GridSample3dForward / _GridSample3dBackward
 * (third_party/cuda_gridsample_grad2/cuda_gridsample.py:17-19,76-126) selected by
 * FeatureGrid.grid_sample_func (grid_opt/models/grid_modules.py:63-69), i.e.
 *   fwd      = ATen grid_sampler_3d                (mode bilinear)
 *   bwd      = aten::grid_sampler_3d_backward      (cuda_gridsample.py:102-107)
 *   bwd_bwd  = grid_sampler_3d_grad2_kernel        (gridsample_cuda.cu:212-533)
 *
 * input  : (B,C,D,H,W) with arbitrary element strides in_strides[5] (NCDHW or channels_last_3d;
 *          the 128-bit vector path is taken when stride_C==1, C%4==0 and the base is 16B aligned)
 * grid   : (B,P,3) contiguous, normalised coords in [-1,1], last dim (x,y,z)->(W,H,D)
 * output : (B,C,P) addressed with out_strides[3] = (sB,sC,sP)
 * ------------------------------------------------------------------------------------------ */
int miso_grid_sample3d_fwd(int dtype, const void* input, const int64_t in_sizes[5], const int64_t in_strides[5],
                           const void* grid, int64_t P, void* output, const int64_t out_strides[3],
                           int padding_mode, int align_corners, miso_stream_t stream);

/* grad_input (same sizes as input, element strides gi_strides or, when NULL, input's; accumulated)
 * and grad_grid ((B,P,3) contiguous, overwritten) are each optional (NULL) -- mirrors output_mask of
 * aten::grid_sampler_3d_backward. */
int miso_grid_sample3d_bwd(int dtype, const void* grad_output, const int64_t go_strides[3], const void* input,
                           const int64_t in_sizes[5], const int64_t in_strides[5], const void* grid, int64_t P,
                           void* grad_input, const int64_t gi_strides[5], void* grad_grid, int padding_mode,
                           int align_corners, miso_stream_t stream);

/* Double backward.  Inputs gg_input (grid-shaped, strides ggi_strides) and gg_grid ((B,P,3)) are
 * optional (NULL == zeros).  Outputs gg_output ((B,C,P), ggo_strides, overwritten), g_input
 * (input-shaped, gi_strides or input's when NULL, accumulated), g_grid ((B,P,3), overwritten) are optional. */
int miso_grid_sample3d_bwd_bwd(int dtype, const void* gg_input, const int64_t ggi_strides[5], const void* gg_grid,
                               const void* grad_output, const int64_t go_strides[3], const void* input,
                               const int64_t in_sizes[5], const int64_t in_strides[5], const void* grid, int64_t P,
                               void* gg_output, const int64_t ggo_strides[3], void* g_input,
                               const int64_t gi_strides[5], void* g_grid, int padding_mode, int align_corners,
                               miso_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 2. Fused multiresolution field: L dense levels over one bound + ReLU MLP decoder.
 * ------------------------------------------------------------------------------------------ */
typedef struct miso_level {
  const float* feat; /* level tensor base, logical (1,C,Z,Y,X) (grid_modules.py:55-57) */
  float* grad;       /* same layout, accumulated into; may be NULL when no grid gradient is wanted */
  int32_t X, Y, Z, C;
  int64_t sC, sZ, sY, sX; /* element strides; fused kernels require sC==1 (channels_last_3d) */
} miso_level_t;

typedef struct miso_field {
  int32_t num_levels;
  uint32_t ignore_mask; /* bit l set: level l contributes zeros (GridNet.ignore_level_, utils.py:159-163) */
  float bound[6];       /* xmin,xmax,ymin,ymax,zmin,zmax  (BaseNet.bound, base_net.py:31-36) */
  miso_level_t level[MISO_MAX_LEVELS];
} miso_field_t;

/* MLPNet(input_dim=F, 1, hidden_dim=H, hidden_layers=1, bias=True) (grid_opt/models/modules.py:11-32);
 * nn.Linear layout: W row-major [out][in]. */
typedef struct miso_decoder {
  const float *W1, *b1, *W2, *b2, *W3, *b3;
  int32_t in_dim, hidden_dim;
} miso_decoder_t;

/* Optional frame->world transform fused in front of the field (loss.py:764-774):
 * x_world = R[id] x + t[id].  ids int64 (N), R (K,3,3) row-major, t (K,3).  All NULL = identity. */
typedef struct miso_frames {
  const int64_t* ids;
  const float* R;
  const float* t;
  int32_t num_frames;
} miso_frames_t;

/* grid_interp_regular (grid_opt/utils/utils.py:143-164): feats (N, sum_l C_l) row-major. */
int miso_field_features(const miso_field_t* field, const float* x, int64_t N, float* feats, miso_stream_t stream);

/* GridNet.forward (grid_net.py:306-325) + analytic first-order quantities, one kernel:
 *   sdf (N)      = MLP(concat_l interp_l(x))
 *   jac (N,F)    = d sdf / d feat                          (optional, NULL to skip)
 *   gradx (N,3)  = d sdf / d x  == gradient3d(...,'autograd')   (diff.py:27-33; optional)
 *   xw (N,3)     = transformed coordinates when frames given (optional) */
int miso_sdf_forward(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                     const float* x, int64_t N, float* sdf, float* jac, float* gradx, float* xw,
                     miso_stream_t stream);

/* Backward of the above w.r.t. the grids, a single scatter:
 *   grad_l[corner] += (a*w_c + v . dw_c/dx) * jac_l      a = dL/dsdf (N, optional), v = dL/dgradx (N,3, optional)
 * The v-term is the fused double-backward of the eikonal loss (replaces gridsample_cuda.cu:450-481).
 * hv (N,3) optional output: d/dx of (v . gradx) with jac held fixed (mixed second derivatives,
 * gridsample_cuda.cu:484-531) -- the coordinate cotangent of the eikonal double-backward. */
int miso_sdf_backward(const miso_field_t* field, const float* xw, int64_t N, const float* jac, const float* a,
                      const float* v, float* hv, miso_stream_t stream);

/* Whole mapping step (MisoLossMappingBase.compute loss.py:754-813 + backward), one kernel:
 *   sdf term  (miso_loss_regression loss.py:594-635), L1 or L2, valid mask, weights, mean over N
 *   fs  term  (miso_loss_free_space loss.py:668-700)
 *   eik term  (miso_loss_eikonal loss.py:638-665), analytic gradient, optional |gt|<eik_trunc filter
 * and scatters d(total)/d(grid) into level[].grad.  total = w_sdf*sdf + w_fs*fs + w_eik*eik. */
typedef struct miso_mapping_cfg {
  int32_t loss_type;      /* 0 = L1, 1 = L2 */
  float weight_sdf, weight_fs, weight_eik;
  float trunc_dist;       /* free-space lower bound */
  float eik_trunc_dist;   /* <0: no filter (None) */
  int32_t eik_mode;       /* 0 = off, 1 = analytic (autograd second-order equivalent) */
  float grad_scale;       /* upstream d(total) (normally 1) */
  int64_t n_total;        /* denominator of the means; 0 = N.  Point-sharded multi-GPU fits pass the global
                             batch size so that per-rank gradients / loss terms simply sum (all_reduce). */
  const int32_t* n_device; /* optional DEVICE int32: only the first *n_device (<= N) samples are processed -- the
                             batch was compacted on the device (miso_slab_select) and the host never learns its
                             size; requires n_total > 0.  NULL = all N. */
} miso_mapping_cfg_t;

/* gt arrays are (N) float; valid is uint8/bool (N).  eik_count: device int32 counter holding the
 * number of eikonal samples (written by miso_mapping_count; read by the step when the filter is on).
 * partials: device float[(grid_blocks)*4] workspace, loss_out: device float[4] =
 * {sdf_term, fs_term, eik_term, total} (unweighted terms, weighted total).
 * sdf_out (N) optional.  Returns the number of blocks needed for partials via miso_mapping_workspace. */
int64_t miso_mapping_workspace_floats(void);
int miso_mapping_count(const float* gt_sdf, int64_t N, float eik_trunc_dist, int32_t* eik_count, miso_stream_t stream);
int miso_mapping_step(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                      const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                      const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                      const int32_t* eik_count, float* partials, float* loss_out, float* sdf_out,
                      miso_stream_t stream);

/* Decoder-parameter gradients of miso_mapping_step's total (same arguments, same cfg; analytic eikonal term only):
 * the `decoder.fix: False` case (grid_opt/models/grid_net.py:110,126,346-348), where autograd differentiates the MLP
 * (modules.py:11-40) and, through create_graph=True (diff.py:27-33), the eikonal term's grad_x sdf w.r.t. the weights.
 * d total / d {W1,b1,W2,b2,W3,b3} is ACCUMULATED into `grad` (NULL members are skipped; shapes as miso_decoder_t).
 * A second pass over the batch, independent of miso_mapping_step (which yields the loss terms and the grid gradients);
 * eik_count must hold the value miso_mapping_count wrote for this batch.  workspace: device
 * float[miso_mapping_wgrad_workspace_floats()].  Two launches, deterministic (no atomics on the results). */
typedef struct miso_decoder_grad {
  float *W1, *b1, *W2, *b2, *W3, *b3;
} miso_decoder_grad_t;
int64_t miso_mapping_wgrad_workspace_floats(void);
int miso_mapping_step_wgrad(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                            const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                            const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                            const int32_t* eik_count, const miso_decoder_grad_t* grad, float* workspace,
                            miso_stream_t stream);

/* The same step with the FINITE-DIFFERENCE eikonal term the shipped configs select (grad_method: finitediff,
 * configs/rgbd/scannet.yaml:48-49; grid_opt/diff.py:18-26 + loss.py:638-665):
 *   g_d = (f(x + eps e_d) - f(x - eps e_d)) / (2 eps), eik = mean (|g| - 1)^2, all six evaluations differentiated.
 * Four launches instead of the reference's 14 interpolation + 14 decoder passes and their autograd glue: the step on
 * the N samples (sdf + free-space terms), ONE forward launch over the 6 N displaced points, the eikonal epilogue
 * (term + the six cotangents), ONE backward launch that scatters them.  cfg->eik_mode is ignored (the term is on
 * when weight_eik != 0).  fd_workspace: device float[12 * N].  Returns MISO_ERR_UNSUPPORTED when the
 * two-threads-per-point kernel does not cover the field (levels*channels % 8 != 0, >= 2^31-element grids). */
int miso_mapping_step_fd(const miso_field_t* field, const miso_decoder_t* dec, const miso_frames_t* frames,
                         const float* x, int64_t N, const float* gt_sdf, const uint8_t* gt_valid,
                         const float* gt_sign, const float* weights, const miso_mapping_cfg_t* cfg,
                         const int32_t* eik_count, float* partials, float* loss_out, float* sdf_out,
                         float finite_diff_eps, float* fd_workspace, miso_stream_t stream);

/* Gauss-Newton / LM normal equations of one keyframe in ONE launch (Tracker.lm_step, grid_opt/slam/tracker.py:148-212):
 * x_w = R x + t, r = sdf(x_w) - gt, g = grad_x sdf(x_w), J = [((R x) x g)^T R, g^T], w = 1 (loss_type 0, L2) or
 * gm_scale/(gm_scale + r^2)^2 (loss_type 1, Geman-McClure :139-146); samples with |gt| >= trunc_dist are skipped when
 * trunc_dist >= 0 (:158-164).  Rt = device pointer to 12 floats (R row-major, t).  out (45 doubles, overwritten):
 * H = J^T W J row-major [0,36), b = J^T W r [36,42), in-bound count [42] (fov_overlap numerator :176), used samples [43],
 * sum w r^2 [44]. */
int miso_track_normal_equations(const miso_field_t* field, const miso_decoder_t* dec, const float* x_frame,
                                const float* gt_sdf, int64_t N, const float* Rt, int32_t loss_type, float gm_scale,
                                float trunc_dist, double* out, miso_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3. Latent-space submap alignment (pairwise_loss_latent, grid_opt/align/miso.py:116-211),
 *    batched over all submap pairs of one iteration of generic_align_multiple_submaps
 *    (grid_opt/align/base.py:127-159) in ONE launch.
 *
 * For pair i and each src sample p (M,3):  u = A1 p + b1 (src->world), q = A2 u + b2 (world->dst)
 * (transform_points_to / transfrom_points_from, utils_geometry.py:214-240; A1=R_s, b1=t_s,
 * A2=R_d^T, b2=-R_d^T t_d are composed by the caller exactly as the reference does, so autograd
 * can finish through so3_exp_map), inclusive in-bound test against the dst bound
 * (coords_in_bound :11-27), r = f_src(p)[0:K] - f_dst(q)[0:K], K = C*(levels_used).
 * Reductions (float64) into out[i*MISO_ALIGN_OUT + ...]:
 *   [0] S = sum r^2   [1] count of valid points   [2..4] G0 = sum gamma   [5..13] G1 = sum gamma u^T
 *   [14..22] G2 = sum gamma p^T (row-major 3x3),  gamma = dS/dq           [23] sum_i |r_i|_2
 *   [24..29] Jtr, [30..65] JtJ: Gauss-Newton normal equations of r wrt a left-multiplied dst-frame
 *   twist (tracker.py:179-197 conventions), accumulated only when want_gn != 0.
 * ------------------------------------------------------------------------------------------ */
#define MISO_ALIGN_OUT 72
typedef struct miso_align_pair {
  int32_t src, dst;        /* indices into fields[] */
  int32_t levels_used;     /* level+1: channels [0, C*levels_used) enter the residual (miso.py:133-134) */
  int32_t reserved;        /* miso_align_intersections: output slot / pose row of this pair */
  const float* p;          /* (M,3) src-frame samples: GridAtlas.coordinates_for_alignment (grid_atlas.py:581-587) */
  int64_t M;
  const float* fsrc;       /* optional (M,K) cached f_src(p): constant across iterations */
  uint8_t* mask_out;       /* optional (M): bit-exact in-bound mask (1 = valid) */
  const int32_t* enabled;  /* optional device flag (e.g. intersection test result); NULL = enabled */
  float src_grad_scale;    /* when fields[src].level[].grad != NULL: grad += src_grad_scale * dS/dfeat */
  float dst_grad_scale;
} miso_align_pair_t;

/* fields, pairs, poses ((num_pairs,24) floats: A1 row-major, b1, A2 row-major, b2) and out are
 * DEVICE pointers; out is overwritten.  flags: bit 0 = accumulate the Gauss-Newton block; bits 4-5 = align_loss
 * (miso.py:200-205): 0 'L2' = sum r^2, 1 'L1' = sum_i |r_i|_2, 2 'cos' = sum_i (1 - cosine_similarity(f_s, f_d)).
 * out[0] holds the sum of the per-point loss values and gamma its derivative w.r.t. q, so the caller normalises
 * with weight/(count*K) for L2 and weight/count for L1 / cos.  Feature-grid scatter and the Gauss-Newton block
 * exist for L2 only. */
#define MISO_ALIGN_WANT_GN 1
#define MISO_ALIGN_LOSS_L2 (0 << 4)
#define MISO_ALIGN_LOSS_L1 (1 << 4)
#define MISO_ALIGN_LOSS_COS (2 << 4)
int miso_align_batch(const miso_field_t* fields, int32_t num_fields, const miso_align_pair_t* pairs,
                     int32_t num_pairs, int64_t max_M, const float* poses, double* out, int32_t flags,
                     miso_stream_t stream);

/* check_submap_intersection (grid_atlas.py:405-420) for all pairs in one launch.  Here pairs[i].p / M are
 * the SOURCE submap's finest-level vertex positions, pairs[i].levels_used describes the lattice: X | (Y << 16) when p is
 * FeatureGrid.vertex_positions() of an (Z,Y,X) grid (x fastest; a per-axis outer product), 0 = unstructured point list, pairs[i].reserved is the pair's output slot (also its
 * row in `poses`), and the array is sorted so pairs that share a source are contiguous; groups (num_groups x 2
 * int32: first pair, count <= 32) lets a block read each vertex once and test it against every destination
 * paired with that source.  enabled_out[slot] = (count / M > overlap_thresh); counts_out (num_pairs) uint64. */
int miso_align_intersections(const miso_field_t* fields, int32_t num_fields, const miso_align_pair_t* pairs,
                             int32_t num_pairs, const int32_t* groups, int32_t num_groups, int64_t max_M,
                             const float* poses, float overlap_thresh, int32_t* enabled_out,
                             unsigned long long* counts_out, miso_stream_t stream);

/* GridAtlas.query_feature (grid_opt/models/grid_atlas.py:374-391) for all submaps in one launch: feats (N, levels*4) =
 * in-bound-masked mean over the `active` submaps (device int32 indices into fields[], visited in order) of each
 * submap's multi-level feature at the world points x.  poses (num_fields,12) = per submap (R^T row-major, -R^T t). */
int miso_atlas_features(const miso_field_t* fields, int32_t num_fields, const int32_t* active, int32_t num_active,
                        const float* poses, const float* x, int64_t N, int32_t levels, float* feats,
                        miso_stream_t stream);

/* Pose glue of one alignment iteration (grid_opt/align/base.py:127-159 around the loss) as three single-block
 * kernels: (1) R = R0 Exp(w), t = t0 + tau for every submap (grid_atlas.py:250-268; Exp = pytorch3d so3_exp_map
 * with its 1e-4 clamp) and the per-pair (A1,b1,A2,b2) rows; (2) from miso_align_batch's reductions to
 * loss_i = mean(r^2)*weight (nan_to_num), total -> loss_hist[*iter_counter], and d total/d(w_s, tau_s) -> grads
 * (S,6) through the transforms and the closed-form derivative of Exp; (3) torch.optim.Adam on (w_s, tau_s),
 * s >= 1 (submap 0 fixed, base.py:104-108), state in exp_avg / exp_avg_sq (S,6), step = ++*iter_counter.
 * w_ptrs / tau_ptrs are DEVICE arrays of S device pointers to the (1,3) / (3,1) correction tensors, which are
 * updated in place.  A multi-GPU run all-reduces `grads` (and `contrib`) between (2) and (3).  Rt (S,12) = [R row-major, t].
 * contrib (S floats, optional): number of pairs that gave submap s a gradient this iteration; with it and
 * submap_steps (S int32, zero-initialised) the Adam kernel skips submaps with contrib == 0 -- no moment decay, no step
 * increment -- as torch.optim.Adam skips a parameter whose .grad is None (a submap none of whose pairs intersects any
 * more stops moving, base.py:112-146), and bias-corrects with the submap's own step count. */
int miso_align_compose_poses(const float* R0, const float* t0, float* const* w_ptrs, float* const* tau_ptrs,
                             int32_t num_submaps, const int32_t* src, const int32_t* dst, int32_t num_pairs,
                             float* poses24, float* Rt_out, miso_stream_t stream);
int miso_align_pose_grads(const float* R0, const float* t0, float* const* w_ptrs, float* const* tau_ptrs,
                          int32_t num_submaps, const int32_t* src, const int32_t* dst, int32_t num_pairs,
                          const double* align_out, const float* poses24, const float* Rt, int32_t channels_used,
                          float align_weight, float* grads, float* loss_hist, int32_t* iter_counter,
                          float* pair_loss, float* contrib, miso_stream_t stream);
int miso_align_pose_adam(float* const* w_ptrs, float* const* tau_ptrs, int32_t num_submaps, const float* grads,
                         float* exp_avg, float* exp_avg_sq, int32_t* iter_counter, float lr, float beta1,
                         float beta2, float eps, const float* contrib, int32_t* submap_steps, miso_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 4. Helpers around the path.
 * ------------------------------------------------------------------------------------------ */
/* Domain-decomposed multi-GPU fit: keep the samples of a (replicated) batch whose trilinear cell of one level starts in
 * planes [z_begin, z_end) of that level along `axis` (0 = x, 1 = y, 2 = z; Z planes over [zmin, zmax] on that axis --
 * the caller stores the level with that axis slowest, so a range of planes is one contiguous piece of memory),
 * compacted to the front of the *_out arrays in batch order within 1024-sample chunks; *count (device int32) receives
 * their number.  Same index arithmetic as the fused kernels, so each sample is owned by exactly one rank and touches
 * only planes [z_begin, z_end] of that level.  frames / weights / ids_out / weights_out may be NULL. */
int miso_slab_select(const miso_frames_t* frames, const float* x, int64_t N, float zmin, float zmax, int32_t Z,
                     int32_t axis, int32_t z_begin, int32_t z_end, const float* gt_sdf, const uint8_t* gt_valid, const float* gt_sign,
                     const float* weights, float* x_out, int64_t* ids_out, float* sdf_out, uint8_t* valid_out,
                     float* sign_out, float* weights_out, int32_t* count, miso_stream_t stream);

/* 30-bit (10 bits/axis) Morton key of each point inside bound, for L2-local batch ordering. */
int miso_morton_keys(const float* x, int64_t N, const float bound[6], uint32_t* keys, miso_stream_t stream);

/* x_world = R[id] x + t[id]  (loss.py:764-774 without the per-keyframe host loop). */
int miso_transform_points(const float* x, const int64_t* ids, const float* R, const float* t, int32_t num_frames,
                          int64_t N, float* y, miso_stream_t stream);

/* Compact host batches: the reference's dataset emits sample_frame_ids as int64 and the two masks
 * sdf_valid = |sdf| < trunc, sdf_signs = +-1 beyond +-trunc as separate tensors (sdf_rgbd.py:452-455,
 * submap_dataset.py:57-76): 33 B/point over PCIe.  A dataset that ships int16 ids and only the sdf moves
 * 18 B/point; this entry rebuilds the int64 ids and both masks on the device (bit-identical to the
 * reference's host-side expressions). */
int miso_expand_batch(const int16_t* ids16, const float* sdf, float trunc_dist, int64_t N, int64_t* ids64,
                      uint8_t* valid, float* sign, miso_stream_t stream);

/* torch.optim.Adam (no amsgrad, no weight decay) single-tensor step; optionally zeroes g in the
 * same pass so the next scatter starts from a clean buffer (trainer.py:216-217 + zero_grad). */
int miso_adam_step(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int32_t step, int32_t zero_grad, miso_stream_t stream);

/* Same step with an "ever touched" bitmap (ceil(n/4/32) uint32 words, zero-initialised by the caller, one bit per
 * 4-float voxel; n % 4 == 0, 16-byte aligned tensors): voxels that no sample has ever touched are skipped after
 * reading only their gradient.  Bit-identical results to miso_adam_step. */
int miso_adam_step_tracked(float* p, float* g, float* m, float* v, uint32_t* touched, int64_t n, float lr, float beta1,
                           float beta2, float eps, int32_t step, int32_t zero_grad, miso_stream_t stream);

/* Kernel-variant switches of the fused step (tests / profiling; defaults come from the MISO_* environment
 * variables): keys "mlp_tc" (1 tcgen05 decoder | 0 SIMT), "tc2_groups" (4 | 3 tiles in flight of the
 * two-threads-per-point kernel, 0 = one-thread kernel), "pair" (lane-paired gather/scatter), "fwd_tc2", "dbg"
 * (ablation bits), "force_int64" (take the 64-bit-offset route the reference selects at
 * third_party/cuda_gridsample_grad2/gridsample_cuda.cu:628-660 for any grid).  Process-global, not thread-safe
 * against concurrent launches.  miso_get_tuning returns the value or -1 for an unknown key. */
int miso_set_tuning(const char* key, int32_t value);
int miso_get_tuning(const char* key);

/* The same step (tracked when `touched` != NULL) with the step count kept on the DEVICE: every block derives the
 * bias-correction scalars of step *step_counter + 1 itself and the last block to finish publishes the new count, so
 * the optimizer step is ONE launch whose arguments never change and a whole training step can be captured in a CUDA
 * graph.  scalars: 3 device floats of scratch, zero-initialised once by the caller (scalars[0] is the block ticket).
 * gate (optional device float, the step's total loss): when it is not finite the update is skipped as the reference's
 * trainer does (grid_opt/trainer.py:214-217) -- counter, p, m, v untouched, the gradient cleared when zero_grad. */
int miso_adam_step_dev(float* p, float* g, float* m, float* v, uint32_t* touched, int64_t n, float lr, float beta1,
                       float beta2, float eps, int32_t* step_counter, float* scalars, const float* gate,
                       int32_t zero_grad, miso_stream_t stream);

/* Adam on one boundary plane of a slab-sharded level, fused with both halo exchanges over NVLink peer memory
 * (miso_b200/sharded_fit.py; the multi-GPU form of trainer.py:209-217 + 422-437 for one large grid): gradient =
 * g + g_peer (the same plane of the lower neighbour's gradient buffer, read with peer loads and cleared in place), the
 * new parameters are written to p and to p_peer (the neighbour's copy of the plane, read by its next step).  g_peer /
 * p_peer: pointers obtained from miso_ipc_import (or NULL at the ends of the chain).  Device-side step counter as in
 * miso_adam_step_dev.  The caller orders the ranks with the collectives around the call. */
int miso_adam_step_halo(float* p, float* g, float* m, float* v, int64_t n, float* g_peer, float* p_peer, float lr,
                        float beta1, float beta2, float eps, int32_t* step_counter, float* scalars, miso_stream_t stream);

/* Neighbour-to-neighbour ordering over peer memory, no collective: miso_peer_signal adds 1 to a counter that lives in
 * the neighbour's memory (pointer from miso_ipc_import) once everything enqueued before it on `stream` is complete;
 * miso_peer_wait(sync) holds `stream` until sync[0] (the counter the peer bumps) has reached sync[1], then increments
 * sync[1] (the number of waits so far).  sync: 4 device uint32 {counter, waits, error, -}, zero-initialised; the wait
 * gives up after ~30 s and sets sync[2] instead of hanging the device. */
int miso_peer_signal(uint32_t* flag_peer, miso_stream_t stream);
int miso_peer_wait(uint32_t* sync, miso_stream_t stream);

/* Same-node peer mapping of a device allocation (CUDA IPC).  export: handle (64 bytes) of the allocation `ptr` lives
 * in + ptr's offset from its base, to be sent to the neighbour's process; import: maps it for the CURRENT device
 * (peer access enabled lazily) and returns the pointer; mappings are cached per (device, handle). */
int miso_ipc_export(const void* ptr, unsigned char handle[64], int64_t* offset);
int miso_ipc_import(const unsigned char handle[64], int64_t offset, void** ptr);

/* ------------------------------------------------------------------------------------------
 * 5. Self-test of the tensor-core building block of the fused decoder (tcgen05.mma kind::tf32 with the
 *    3xTF32 split, activations in TMEM, weights in shared memory): D (M,64) = A (M,64) * W^T
 *    (transpose=0, the forward layer h W2^T as nn.Linear computes it) or A * W (transpose=1, the backward
 *    product).  W is a row-major 64x64 fp32 matrix (nn.Linear.weight layout, modules.py:18).
 * ------------------------------------------------------------------------------------------ */
int miso_tc_selftest(const float* A, const float* W, int32_t transpose, float* D, int64_t M, miso_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MISO_B200_H_ */
