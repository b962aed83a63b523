"""Latent-space submap alignment on the B200 path.  Same names / arguments as the reference:

    pairwise_loss_sdf                     grid_opt/align/miso.py:14-113
    pairwise_loss_latent                  grid_opt/align/miso.py:116-211
    generic_align_multiple_submaps        grid_opt/align/base.py:89-163
    align_multiple_submaps_hierarchical   grid_opt/align/miso.py:217-322

`pairwise_loss_latent` keeps its per-pair signature (it is a batch of one).  The multi-pair loop
evaluates EVERY pair of an iteration in one launch (miso_align_batch): transform, inclusive
in-bound mask, both interpolations, residual and the reductions that autograd needs
(sum r^2, count, sum gamma, sum gamma u^T, sum gamma p^T), then finishes through the same tiny
torch ops as the reference (R0 @ so3_exp_map(w), R^T, -R^T t), so `dL/dw`, `dL/dtau` and the Adam
update follow the reference's path.  The per-pair intersection test (grid_atlas.py:405-420) is one
more launch and stays on the device (the reference syncs the host once per pair per iteration).

Deliberate deviation: the reference also back-propagates into both submaps' dense feature grids
(they are left requires_grad=True by Mapper.mapping, mapper.py:72) although no optimizer reads
them; here that scatter is opt-in (`feature_grads=True`).
"""
import ctypes as C
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.optim as optim

from . import _lib
from . import field as _field
from . import geometry as utils_geometry
from .models import GridAtlas


_structs_to_device = _field.structs_to_device


class AlignBatch:
    """Device-side description of one alignment problem: all submaps' fields + a list of
    (src, dst) pairs at one level.  Built once per level; reused by every iteration."""

    LOSS_KINDS = {"L2": 0, "L1": 1, "cos": 2}

    def __init__(self, grid_atlas: GridAtlas, pairs: Sequence[Tuple[int, int]], level: int, fdim: int = 4,
                 subsample_points: Optional[int] = None, cache_src_features: bool = True, want_masks: bool = False,
                 check_intersection: bool = True, lattice_rows: bool = True, align_loss: str = "L2",
                 trunc_factor: Optional[float] = None):
        if align_loss not in self.LOSS_KINDS:
            if align_loss == "InfoNCE":
                raise NotImplementedError("align_loss='InfoNCE' (miso.py:206-208) couples every pair of samples of a "
                                          "batch; the fused per-sample kernel implements L2, L1 and cos")
            raise ValueError(f"Invalid align loss: {align_loss}!")
        self.align_loss = align_loss
        self.loss_flags = self.LOSS_KINDS[align_loss] << 4
        self.atlas = grid_atlas
        self.pairs = list(pairs)
        self.level = level
        self.device = torch.device(grid_atlas.device)
        self.levels_used = level + 1
        S = grid_atlas.num_submaps
        for s in range(S):
            sm = grid_atlas.get_submap(s)
            if sm.fdim != 4 or fdim != 4:
                raise NotImplementedError("fused alignment kernel is built for fdim=4 (miso.py:122 default)")
        self._keep = []
        fields = []
        for s in range(S):
            sm = grid_atlas.get_submap(s)
            mask = sum((1 << l) for l in range(sm.num_levels) if sm.ignore_level_[l])
            fields.append(_field.make_field(sm.level_tensors(), sm._bound_host, None, mask))
        self.fields_dev = _structs_to_device(fields, self.device)
        self.num_fields = S
        P = len(self.pairs)
        self.enabled = torch.ones(max(P, 1), dtype=torch.int32, device=self.device)
        self.counts = torch.zeros(max(P, 1), dtype=torch.int64, device=self.device)
        self.masks: List[Optional[torch.Tensor]] = [None] * P
        K = 4 * self.levels_used
        # per-src caches
        self._coords, self._fsrc, self._verts = {}, {}, {}
        pair_structs, isect_structs = [], []
        self.max_M, self.max_V = 0, 0
        for i, (src, dst) in enumerate(self.pairs):
            assert src < S and dst < S
            if src not in self._coords:
                p = grid_atlas.coordinates_for_alignment(src, level)
                if subsample_points is not None:
                    n = min(subsample_points, p.shape[0])
                    idx = np.random.choice(p.shape[0], n, replace=False)  # miso.py:146-149
                    p = p[torch.from_numpy(idx).to(p.device), :]
                if trunc_factor is not None:
                    # truncation pruning (miso.py:176-183): keep the samples whose SOURCE sdf is within trunc_factor
                    # cells of the surface -- a function of the source submap alone, so it is applied once here
                    sm = grid_atlas.get_submap(src)
                    with torch.no_grad():
                        near = torch.abs(sm(p)) < trunc_factor * sm.cell_sizes[level]
                    p = p[near[:, 0]]
                p = p.detach().contiguous().float()
                self._coords[src] = p
                if cache_src_features and p.shape[0] > 0:
                    sm = grid_atlas.get_submap(src)
                    with torch.no_grad():
                        f = sm.query_feature(p)[:, :K].contiguous()
                    self._fsrc[src] = f
            p = self._coords[src]
            M = p.shape[0]
            self.max_M = max(self.max_M, M)
            ap = _lib.AlignPair()
            ap.src, ap.dst, ap.levels_used = src, dst, self.levels_used
            ap.p, ap.M = p.data_ptr() if M > 0 else None, M
            ap.fsrc = self._fsrc[src].data_ptr() if src in self._fsrc else None
            if want_masks:
                self.masks[i] = torch.zeros(M, dtype=torch.uint8, device=self.device)
                ap.mask_out = self.masks[i].data_ptr() if M > 0 else None
            ap.enabled = self.enabled.data_ptr() + 4 * i if check_intersection else None
            ap.src_grad_scale = 0.0
            ap.dst_grad_scale = 0.0
            pair_structs.append(ap)
            if check_intersection:
                if src not in self._verts:
                    sm = grid_atlas.get_submap(src)
                    self._verts[src] = sm.features[-1].vertex_positions().to(self.device).contiguous().float()
                v = self._verts[src]
                self.max_V = max(self.max_V, v.shape[0])
                ip = _lib.AlignPair()
                # vertex_positions() lists the finest lattice row by row (x fastest, grid_modules.py:111-123): tell the
                # kernel the row length so it can decide whole row segments from their end points (exact count)
                shp = grid_atlas.get_submap(src).features[-1].feature.shape          # (1,C,Z,Y,X)
                X_, Y_, Z_ = int(shp[-1]), int(shp[-2]), int(shp[-3])
                lattice = X_ | (Y_ << 16)
                if v.shape[0] != X_ * Y_ * Z_ or X_ >= 65536 or Y_ >= 32768 or not lattice_rows:
                    lattice = 0
                ip.src, ip.dst, ip.levels_used, ip.reserved = src, dst, lattice, i
                ip.p, ip.M = v.data_ptr(), v.shape[0]
                isect_structs.append(ip)
        self.check_intersection = check_intersection
        self.pairs_dev = _structs_to_device(pair_structs, self.device) if P else None
        self.isect_dev, self.groups_dev, self.num_groups = None, None, 0
        if P and check_intersection:
            # group by source submap (<= 32 pairs per group): one vertex read serves every pair of the group
            isect_structs.sort(key=lambda q: q.src)
            groups, start = [], 0
            for j in range(1, len(isect_structs) + 1):
                if j == len(isect_structs) or isect_structs[j].src != isect_structs[start].src or j - start == 32:
                    groups.append((start, j - start))
                    start = j
            self.isect_dev = _structs_to_device(isect_structs, self.device)
            self.groups_dev = torch.tensor(groups, dtype=torch.int32, device=self.device).contiguous()
            self.num_groups = len(groups)
        self.src_idx = torch.tensor([s for s, _ in self.pairs], dtype=torch.long, device=self.device)
        self.dst_idx = torch.tensor([d for _, d in self.pairs], dtype=torch.long, device=self.device)
        self.K = K
        self.norm_channels = K if align_loss == "L2" else 1

    # ---- poses ------------------------------------------------------------------------------------
    def submap_poses(self):
        """Batched GridAtlas.updated_submap_pose (grid_atlas.py:250-268) for all submaps."""
        a = self.atlas
        R0 = torch.stack(list(a.R_world_submap_list), 0)
        t0 = torch.stack(list(a.t_world_submap_list), 0)
        w = torch.cat(list(a.rotation_corrections), 0)
        tau = torch.stack(list(a.translation_corrections), 0)
        R = torch.matmul(R0, utils_geometry.so3_exp_map(w))
        t = t0 + tau
        return R, t

    def pair_poses(self):
        """(P,24): A1=R_s, b1=t_s, A2=R_d^T, b2=-R_d^T t_d (utils_geometry.py:214-240)."""
        R, t = self.submap_poses()
        A1 = R[self.src_idx]
        b1 = t[self.src_idx]
        A2 = R[self.dst_idx].transpose(1, 2)
        b2 = -torch.matmul(A2, t[self.dst_idx])
        return torch.cat([A1.reshape(-1, 9), b1.reshape(-1, 3), A2.reshape(-1, 9), b2.reshape(-1, 3)], 1)

    # ---- launches -----------------------------------------------------------------------------------
    def update_intersections(self, poses24: torch.Tensor, overlap_thresh: float = 1e-2):
        if not self.check_intersection or not self.pairs:
            return
        lib = _lib.load()
        p = poses24.detach().contiguous().float()
        with torch.cuda.device(self.device):
            _lib.check(lib.miso_align_intersections(
                self.fields_dev.data_ptr(), self.num_fields, self.isect_dev.data_ptr(), len(self.pairs),
                self.groups_dev.data_ptr(), self.num_groups, self.max_V, p.data_ptr(), float(overlap_thresh),
                self.enabled.data_ptr(), self.counts.data_ptr(), _lib.stream_ptr(self.device)), "align_intersections")

    def launch(self, poses24: torch.Tensor, want_gn: bool = False) -> torch.Tensor:
        lib = _lib.load()
        P = len(self.pairs)
        out = torch.empty((P, _lib.MISO_ALIGN_OUT), dtype=torch.float64, device=self.device)
        if P == 0:
            return out
        p = poses24.detach().contiguous().float()
        with torch.cuda.device(self.device):
            _lib.check(lib.miso_align_batch(
                self.fields_dev.data_ptr(), self.num_fields, self.pairs_dev.data_ptr(), P, self.max_M, p.data_ptr(),
                out.data_ptr(), int(want_gn) | self.loss_flags, _lib.stream_ptr(self.device)), "align_batch")
        return out

    def losses(self, align_weight: float = 3000.0, poses24: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(P,) alignment losses (self.align_loss), differentiable w.r.t. the submap pose corrections."""
        if poses24 is None:
            poses24 = self.pair_poses()
        return _AlignBatchFn.apply(poses24, self, float(align_weight))


class _AlignBatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, poses24, batch: AlignBatch, align_weight: float):
        out = batch.launch(poses24)
        S, cnt = out[:, 0], out[:, 1]
        # L2: mean over M_valid x K elements of r^2; L1 / cos: mean over the M_valid samples of |r|_2 / (1 - cos)
        denom = cnt * batch.norm_channels
        # times weight; 0 when nothing is valid (miso.py:180-182,200-205)
        scale = torch.where(cnt > 0, align_weight / denom.clamp(min=1.0), torch.zeros_like(denom))
        loss = (S * scale).to(torch.float32)
        ctx.save_for_backward(poses24, out, scale)
        return loss

    @staticmethod
    def backward(ctx, g):
        poses24, out, scale = ctx.saved_tensors
        P = out.shape[0]
        s = (scale * g.to(torch.float64)).view(P, 1, 1)
        G0 = out[:, 2:5].reshape(P, 3, 1)
        G1 = out[:, 5:14].reshape(P, 3, 3)
        G2 = out[:, 14:23].reshape(P, 3, 3)
        A2 = poses24[:, 12:21].reshape(P, 3, 3).to(torch.float64)
        A2t = A2.transpose(1, 2)
        dA1 = torch.matmul(A2t, G2) * s
        db1 = torch.matmul(A2t, G0) * s
        dA2 = G1 * s
        db2 = G0 * s
        grad = torch.cat([dA1.reshape(P, 9), db1.reshape(P, 3), dA2.reshape(P, 9), db2.reshape(P, 3)], 1)
        return grad.to(poses24.dtype), None, None


class FusedPoseAligner:
    """One alignment iteration (base.py:127-159, latent L2 loss) as FIVE launches: compose poses, intersection
    test, alignment kernel, pose gradients, Adam -- no torch ops, no autograd graph (csrc/poseopt.cu).  The torch
    path (`AlignBatch.losses` + autograd + torch.optim.Adam) needs ~100 small launches for the same iteration,
    which bounds level-0 alignment (32 k samples per pair) at ~1 ms per iteration even inside a CUDA graph.

    The submap pose corrections (GridAtlas.rotation_corrections / translation_corrections) are updated in place
    through a device table of their addresses; submap 0 stays fixed; Adam state lives here.  `allreduce` (multi-GPU,
    pair-sharded) sums the (S,6) gradient buffer over ranks between the gradient and the Adam kernel."""

    def __init__(self, batch: AlignBatch, lr: float = 1e-2, align_weight: float = 3000.0, betas=(0.9, 0.999),
                 eps: float = 1e-8, max_iters: int = 4096, allreduce=None):
        a = batch.atlas
        dev = batch.device
        self.batch, self.lr, self.align_weight, self.betas, self.eps = batch, float(lr), float(align_weight), betas, float(eps)
        self.allreduce = allreduce
        S = a.num_submaps
        self.S, self.P = S, len(batch.pairs)
        self.R0 = torch.stack([r.detach().float() for r in a.R_world_submap_list], 0).reshape(S, 9).contiguous().to(dev)
        self.t0 = torch.stack([t.detach().float() for t in a.t_world_submap_list], 0).reshape(S, 3).contiguous().to(dev)
        for q in list(a.rotation_corrections) + list(a.translation_corrections):
            if not (q.is_cuda and q.dtype == torch.float32 and q.is_contiguous()):
                raise RuntimeError("pose corrections must be contiguous float32 CUDA tensors")
        self.w_ptrs = torch.tensor([q.data_ptr() for q in a.rotation_corrections], dtype=torch.int64, device=dev)
        self.tau_ptrs = torch.tensor([q.data_ptr() for q in a.translation_corrections], dtype=torch.int64, device=dev)
        self.src = batch.src_idx.to(torch.int32).contiguous()
        self.dst = batch.dst_idx.to(torch.int32).contiguous()
        self.poses24 = torch.zeros((max(self.P, 1), 24), dtype=torch.float32, device=dev)
        self.Rt = torch.zeros((S, 12), dtype=torch.float32, device=dev)
        self.out = torch.zeros((max(self.P, 1), _lib.MISO_ALIGN_OUT), dtype=torch.float64, device=dev)
        self.grads = torch.zeros((S, 6), dtype=torch.float32, device=dev)
        # pairs that gave each submap a gradient this iteration; a submap with none is skipped by Adam (grad None in torch)
        self.contrib = torch.zeros(S, dtype=torch.float32, device=dev)
        self.submap_steps = torch.zeros(S, dtype=torch.int32, device=dev)
        self.exp_avg = torch.zeros((S, 6), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros((S, 6), dtype=torch.float32, device=dev)
        self.iter_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_hist = torch.zeros(max_iters, dtype=torch.float32, device=dev)
        self.pair_loss = torch.zeros(max(self.P, 1), dtype=torch.float32, device=dev)
        self.max_iters = max_iters

    def compose(self):
        lib, b = _lib.load(), self.batch
        with torch.cuda.device(b.device):
            _lib.check(lib.miso_align_compose_poses(
                self.R0.data_ptr(), self.t0.data_ptr(), self.w_ptrs.data_ptr(), self.tau_ptrs.data_ptr(), self.S,
                self.src.data_ptr(), self.dst.data_ptr(), self.P, self.poses24.data_ptr(), self.Rt.data_ptr(),
                _lib.stream_ptr(b.device)), "align_compose_poses")
        return self.poses24

    def iteration(self):
        """Everything is enqueued on the current stream; nothing is synchronised (CUDA-graph capturable)."""
        lib, b = _lib.load(), self.batch
        stream = _lib.stream_ptr(b.device)
        self.compose()
        if b.check_intersection:
            b.update_intersections(self.poses24)
        with torch.cuda.device(b.device):
            if self.P:
                _lib.check(lib.miso_align_batch(b.fields_dev.data_ptr(), b.num_fields, b.pairs_dev.data_ptr(), self.P,
                                                b.max_M, self.poses24.data_ptr(), self.out.data_ptr(), b.loss_flags,
                                                stream), "align_batch")
            _lib.check(lib.miso_align_pose_grads(
                self.R0.data_ptr(), self.t0.data_ptr(), self.w_ptrs.data_ptr(), self.tau_ptrs.data_ptr(), self.S,
                self.src.data_ptr(), self.dst.data_ptr(), self.P, self.out.data_ptr(), self.poses24.data_ptr(),
                self.Rt.data_ptr(), b.norm_channels, self.align_weight, self.grads.data_ptr(), self.loss_hist.data_ptr(),
                self.iter_counter.data_ptr(), self.pair_loss.data_ptr(), self.contrib.data_ptr(), stream),
                "align_pose_grads")
            if self.allreduce is not None:
                self.allreduce([self.grads, self.contrib])
            _lib.check(lib.miso_align_pose_adam(
                self.w_ptrs.data_ptr(), self.tau_ptrs.data_ptr(), self.S, self.grads.data_ptr(),
                self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.iter_counter.data_ptr(), self.lr,
                float(self.betas[0]), float(self.betas[1]), self.eps, self.contrib.data_ptr(),
                self.submap_steps.data_ptr(), stream), "align_pose_adam")

    def run(self, num_iters: int, use_cuda_graph: bool = True):
        """`num_iters` iterations; returns the per-iteration total losses (device tensor).  With a CUDA graph the
        first iteration runs eagerly (module loading, NCCL warm-up), the rest are replays of one captured iteration."""
        if int(self.iter_counter.item()) + num_iters > self.max_iters:
            raise RuntimeError("FusedPoseAligner: loss history too short for this many iterations")
        start = int(self.iter_counter.item())
        done = 0
        if use_cuda_graph and num_iters > 2:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.iteration()
            torch.cuda.current_stream().wait_stream(side)
            done = 1
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self.iteration()
            while done < num_iters:
                graph.replay()
                done += 1
        while done < num_iters:
            self.iteration()
            done += 1
        with torch.no_grad():   # the kernels wrote through raw pointers: bump the tensors' version counters
            for q in list(self.batch.atlas.rotation_corrections) + list(self.batch.atlas.translation_corrections):
                q.add_(0.0)
        if self.allreduce is not None:
            total = self.loss_hist[start:start + num_iters].clone()
            self.allreduce([total])
            return total
        return self.loss_hist[start:start + num_iters]


# ------------------------------------------------------------------------------------------------
# reference-named entry points
# ------------------------------------------------------------------------------------------------
def pairwise_loss_latent(grid_atlas: GridAtlas, data_loader, src_id: int, dst_id: int, level: int, fdim=4,
                         align_weight=3000, align_loss="L2", use_bound=True, stability_thresh=0,
                         covariance_thresh=None, subsample_points=None, trunc_factor=None, device="cuda:0"):
    """miso.py:116-211 for one pair on the fused kernel: align_loss 'L2' (the shipped configuration,
    configs/rgbd/scannet.yaml:56-63), 'L1' or 'cos'; truncation pruning (`trunc_factor`) is applied to the source
    samples.  Variants that need data outside the path are rejected loudly: use_bound=False, stability pruning
    (stability grids), covariance pruning (NotImplementedError in the reference too), InfoNCE."""
    loss_key = f"align_latent_level{level}_{src_id}_{dst_id}"
    assert src_id < grid_atlas.num_submaps
    assert dst_id < grid_atlas.num_submaps
    if covariance_thresh is not None:
        raise NotImplementedError
    if not use_bound or stability_thresh > 0:
        raise NotImplementedError("miso_b200.pairwise_loss_latent needs use_bound=True and stability_thresh=0 (the "
                                  "stability grids are outside the hot path, SURVEY.md section 8)")
    batch = AlignBatch(grid_atlas, [(src_id, dst_id)], level, fdim=fdim, subsample_points=subsample_points,
                       cache_src_features=False, check_intersection=False, align_loss=align_loss,
                       trunc_factor=trunc_factor)
    return {loss_key: batch.losses(align_weight)[0]}


def get_batch(data_loader, device="cuda:0"):
    """utils.py:495-498: first batch of the loader, moved to the device."""
    for model_input, gt in data_loader:
        model_input = {k: v.to(device) for k, v in model_input.items()}
        gt = {k: v.to(device) for k, v in gt.items()}
        return model_input, gt
    raise RuntimeError("empty data loader")


def pairwise_loss_sdf(grid_atlas: GridAtlas, data_loader, src_id: int, dst_id: int, align_weight=3000,
                      align_loss="L2", use_bound=True, stability_thresh=0, covariance_thresh=None,
                      subsample_points=None, gm_scale_sdf=0.1, device="cuda:0"):
    """miso.py:14-113: SDF-space alignment residual of one pair on the SOURCE submap's dataset samples.

    Samples of the batch whose keyframe belongs to `src_id` are taken to the submap frame (one batched gather of
    the keyframe poses instead of the per-keyframe loop at :44-53), to the world and into `dst_id`; both submaps are
    evaluated by the fused grid+decoder kernel (GridNet.forward, differentiable w.r.t. the coordinates), so the
    pose gradients flow through d sdf / d x exactly as in the reference.  L2 / L1 / GM as in :100-110."""
    assert src_id < grid_atlas.num_submaps
    assert dst_id < grid_atlas.num_submaps
    if covariance_thresh is not None:
        raise NotImplementedError
    if stability_thresh > 0:
        raise NotImplementedError("stability pruning needs the stability grids, which are outside the hot path "
                                  "(SURVEY.md section 8: unused unless the stability loss is on)")
    model_input, gt = get_batch(data_loader, device)
    submap_from = grid_atlas.get_submap(src_id)
    submap_to = grid_atlas.get_submap(dst_id)
    loss_dict = {}
    kf_all = model_input["sample_frame_ids"][0, :, 0]
    submap_idxs = grid_atlas.submap_id_for_kf_batch(kf_ids=kf_all)
    sample_indices = torch.nonzero(submap_idxs == src_id, as_tuple=False).squeeze(1)
    if sample_indices.numel() == 0:
        return {}
    coords_kf = model_input["coords_frame"][0, sample_indices, :]
    kf_local = kf_all[sample_indices] - grid_atlas.anchor_kf_for_submap(src_id)
    R_kf, t_kf = submap_from.all_kf_poses()                       # (K,3,3), (K,3,1): updated_kf_pose of every keyframe
    coords_from = torch.einsum("nij,nj->ni", R_kf[kf_local], coords_kf) + t_kf[kf_local].squeeze(-1)
    mask_valid = gt["sdf_valid"][0][sample_indices, :]
    R_world_from, t_world_from = grid_atlas.updated_submap_pose(src_id, device)
    R_world_to, t_world_to = grid_atlas.updated_submap_pose(dst_id, device)
    coords_world = utils_geometry.transform_points_to(coords_from, R_world_from, t_world_from)
    coords_to = utils_geometry.transfrom_points_from(coords_world, R_world_to, t_world_to)
    if subsample_points is not None:
        down_points = min(subsample_points, coords_from.shape[0])
        down_indices = torch.from_numpy(np.random.choice(coords_from.shape[0], down_points, replace=False)).to(
            coords_from.device)
        coords_from = coords_from[down_indices, :]
        coords_to = coords_to[down_indices, :]
        mask_valid = mask_valid[down_indices, :]
    if use_bound:
        mask_bnd = utils_geometry.coords_in_bound(coords_to, submap_to.bound.to(coords_to.device))
        assert mask_bnd.shape == mask_valid.shape
        mask_valid = torch.logical_and(mask_bnd, mask_valid)
    valid_indices = torch.nonzero(mask_valid, as_tuple=False)[:, 0]
    p_from = coords_from[valid_indices, :]
    p_to = coords_to[valid_indices, :]
    align_constraint = submap_from(p_from) - submap_to(p_to)
    loss_key = f"align_sdf_{src_id}_{dst_id}"
    if align_loss == "L2":
        loss_dict[loss_key] = torch.mean(align_constraint ** 2) * align_weight
    elif align_loss == "L1":
        loss_dict[loss_key] = torch.mean(torch.linalg.vector_norm(align_constraint, dim=1)) * align_weight
    elif align_loss == "GM":
        e = align_constraint.clone().detach()
        w = gm_scale_sdf / (gm_scale_sdf + e ** 2) ** 2
        loss_dict[loss_key] = torch.mean(w * align_constraint ** 2) * align_weight
    else:
        raise ValueError(f"Invalid align loss: {align_loss}!")
    return loss_dict


def relative_param_change(params_curr, params_prev=None):
    """utils.py:507-516 without the per-iteration .item() host sync: returns a 0-dim tensor (inf first)."""
    if params_prev is None:
        return None
    num_sq = 0
    den_sq = 0
    for c, p in zip(params_curr, params_prev):
        num_sq = num_sq + torch.sum((c - p) ** 2)
        den_sq = den_sq + torch.sum(p ** 2)
    return torch.sqrt(num_sq / den_sq)


def grid_atlas_pose_trust_region_loss(model: GridAtlas, thresh_rad, thresh_m, weight=1e3):
    """base.py:20-27."""
    loss_dict = {}
    for submap_id in range(model.num_submaps):
        rot_norm = torch.linalg.norm(model.rotation_corrections[submap_id])
        loss_dict[f"submap{submap_id}_trust_region_R"] = weight * torch.nn.functional.relu(rot_norm - thresh_rad)
        tran_norm = torch.linalg.norm(model.translation_corrections[submap_id])
        loss_dict[f"submap{submap_id}_trust_region_t"] = weight * torch.nn.functional.relu(tran_norm - thresh_m)
    return loss_dict


def generic_align_multiple_submaps(grid_atlas: GridAtlas, dataset=None, pairwise_loss_tuple=None, num_iters=10,
                                   lr=1e-2, rel_change_thresh=0, submap_pairs=None, check_intersection=True,
                                   pose_reg_weight=0, pose_thresh_rad=1.0, pose_thresh_m=1.0, verbose=True,
                                   save_iterations=False, *, level: int = 0, align_weight=3000.0,
                                   subsample_points=None, pair_filter=None, allreduce=None, use_cuda_graph=False,
                                   fused_pose_glue=True, align_loss="L2", trunc_factor=None):
    """base.py:89-163 with the pair loop replaced by one batched launch per iteration.

    `pairwise_loss_tuple` is accepted for signature compatibility; the loss is the latent `align_loss` ('L2' | 'L1' |
    'cos', miso.py:200-205) at `level`, optionally with truncation pruning of the source samples.  `pair_filter` / `allreduce` are the multi-GPU hooks (miso_b200.dist): a rank evaluates
    only its share of the pairs and the per-submap pose gradients are summed across ranks before Adam.
    Runs `num_iters + 1` iterations like the reference (`while iter <= num_iters`, base.py:127).
    `use_cuda_graph=True` captures one whole iteration (pose composition, intersection test, alignment
    kernel, backward, Adam) into a CUDA graph after 3 eager warm-up iterations and replays it: level 0 has
    only ~32 k samples per pair, so the iteration is launch-bound without it.
    `fused_pose_glue=True` (default; needs pose_reg_weight == 0, no relative-change stop, no saved iterations) runs the
    pose composition, its backward and the Adam step as three single-block kernels (FusedPoseAligner) instead of
    torch ops; `False` keeps the reference's own torch glue (so3_exp_map autograd + torch.optim.Adam)."""
    def pose_params():
        params = []
        for submap_id in range(1, grid_atlas.num_submaps):  # submap 0 stays fixed (base.py:104-108)
            params += list(grid_atlas.params_for_submap_pose(submap_id))
        return params

    if pairwise_loss_tuple is not None and callable(pairwise_loss_tuple[1]):
        return _generic_align_with_loss_func(grid_atlas, dataset, pairwise_loss_tuple, pose_params, num_iters, lr,
                                             rel_change_thresh, submap_pairs, check_intersection, pose_reg_weight,
                                             pose_thresh_rad, pose_thresh_m, save_iterations, pair_filter, allreduce)

    if submap_pairs is None:
        submap_pairs = [(s, d) for s in range(grid_atlas.num_submaps) for d in range(s + 1, grid_atlas.num_submaps)]
    my_pairs = list(submap_pairs) if pair_filter is None else [p for i, p in enumerate(submap_pairs) if pair_filter(i, p)]
    batch = AlignBatch(grid_atlas, my_pairs, level, subsample_points=subsample_points,
                       check_intersection=check_intersection, align_loss=align_loss, trunc_factor=trunc_factor)
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if fused_pose_glue and pose_reg_weight == 0 and rel_change_thresh <= 0 and not save_iterations:
        # whole iteration as five launches (csrc/poseopt.cu): no autograd graph, Adam state held by the aligner
        aligner = FusedPoseAligner(batch, lr=lr, align_weight=align_weight, max_iters=num_iters + 1, allreduce=allreduce)
        losses = aligner.run(num_iters + 1, use_cuda_graph=use_cuda_graph)
        ev1.record()
        torch.cuda.synchronize()
        return {"cpu_time_sec": time.perf_counter() - t0, "gpu_time_sec": ev0.elapsed_time(ev1) / 1e3,
                "iteration_results": dict(), "losses": losses.cpu(), "iterations": num_iters + 1}
    optimizer = optim.Adam([{"params": pose_params(), "lr": lr}], lr=lr, capturable=bool(use_cuda_graph))
    iteration_results = dict()
    params_prev = None
    losses_hist = []
    it = 0
    graph, static_loss = None, None
    # an NCCL all_reduce is capturable: with `allreduce` given the collective becomes a node of the graph
    can_graph = use_cuda_graph and not save_iterations and rel_change_thresh <= 0

    def one_iteration():
        optimizer.zero_grad(set_to_none=False) if graph_params_ready[0] else optimizer.zero_grad()
        poses24 = batch.pair_poses()
        if check_intersection:
            batch.update_intersections(poses24)
        total = torch.nan_to_num(batch.losses(align_weight, poses24)).sum()
        if pose_reg_weight > 0:
            reg = grid_atlas_pose_trust_region_loss(grid_atlas, thresh_rad=pose_thresh_rad, thresh_m=pose_thresh_m,
                                                    weight=pose_reg_weight)
            total = total + sum(reg.values())
        total.backward()
        if allreduce is not None:
            allreduce([p.grad for p in pose_params() if p.grad is not None])
        optimizer.step()
        return total.detach()

    graph_params_ready = [False]
    if can_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            while it <= min(2, num_iters):          # eager warm-up iterations (also materialise .grad)
                losses_hist.append(one_iteration())
                it += 1
        torch.cuda.current_stream().wait_stream(side)
        graph_params_ready[0] = True
        if it <= num_iters:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                static_loss = one_iteration()
            # the capture itself does not execute: every replay is one iteration
            while it <= num_iters:
                graph.replay()
                losses_hist.append(static_loss.clone())
                it += 1
    while graph is None and it <= num_iters:
        if save_iterations:
            R, t = batch.submap_poses()
            T = torch.eye(4, device=R.device).repeat(R.shape[0], 1, 1)
            T[:, :3, :3] = R.detach()
            T[:, :3, 3:] = t.detach()
            iteration_results[it] = T
        optimizer.zero_grad()
        poses24 = batch.pair_poses()
        if check_intersection:
            batch.update_intersections(poses24)
        pair_losses = torch.nan_to_num(batch.losses(align_weight, poses24))  # base.py:139-141
        total_loss = pair_losses.sum()
        if pose_reg_weight > 0:
            reg = grid_atlas_pose_trust_region_loss(grid_atlas, thresh_rad=pose_thresh_rad, thresh_m=pose_thresh_m,
                                                    weight=pose_reg_weight)
            total_loss = total_loss + sum(reg.values())
        total_loss.backward()
        if allreduce is not None:
            allreduce([p.grad for p in pose_params() if p.grad is not None])
        optimizer.step()
        losses_hist.append(total_loss.detach())
        params_curr = [p.clone().detach() for p in pose_params()]
        relchange = relative_param_change(params_curr, params_prev)
        params_prev = params_curr
        if rel_change_thresh > 0 and relchange is not None and float(relchange) < rel_change_thresh:
            break
        it += 1
    ev1.record()
    torch.cuda.synchronize()
    info = {"cpu_time_sec": time.perf_counter() - t0, "gpu_time_sec": ev0.elapsed_time(ev1) / 1e3,
            "iteration_results": iteration_results, "losses": torch.stack(losses_hist).cpu() if losses_hist else None,
            "iterations": it if it <= num_iters else num_iters + 1}
    return info


def _make_loader(dataset):
    """base.py:112: DataLoader(dataset, shuffle=True, batch_size=1); plain iterables of (model_input, gt) pass."""
    if dataset is None:
        return None
    if hasattr(dataset, "__getitem__") and hasattr(dataset, "__len__") and not isinstance(dataset, (list, tuple)):
        return torch.utils.data.DataLoader(dataset, shuffle=True, batch_size=1, num_workers=0)
    return dataset


def _generic_align_with_loss_func(grid_atlas, dataset, pairwise_loss_tuple, pose_params, num_iters, lr,
                                  rel_change_thresh, submap_pairs, check_intersection, pose_reg_weight,
                                  pose_thresh_rad, pose_thresh_m, save_iterations, pair_filter, allreduce):
    """base.py:89-163 verbatim in structure: a caller-supplied `loss_func(grid_atlas, loader, src, dst) -> dict`
    per pair (used by the SDF-space fine-tune, which is off in the shipped configs)."""
    loss_name, loss_func = pairwise_loss_tuple
    optimizer = optim.Adam([{"params": pose_params(), "lr": lr}], lr=lr)
    loader = _make_loader(dataset)
    if submap_pairs is None:
        submap_pairs = [(s, d) for s in range(grid_atlas.num_submaps) for d in range(s + 1, grid_atlas.num_submaps)]
    my_pairs = list(submap_pairs) if pair_filter is None else [p for i, p in enumerate(submap_pairs) if pair_filter(i, p)]
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    iteration_results, losses_hist, params_prev, it = dict(), [], None, 0
    while it <= num_iters:
        if save_iterations:
            with torch.no_grad():
                iteration_results[it] = torch.stack([utils_geometry.pose_matrix(*grid_atlas.updated_submap_pose(i))
                                                     for i in range(grid_atlas.num_submaps)], 0)
        optimizer.zero_grad()
        loss_dict = {}
        for src_id, dst_id in my_pairs:
            if check_intersection:
                with torch.no_grad():
                    if not bool(grid_atlas.check_submap_intersection(src_id, dst_id)):   # base.py:133-135
                        continue
            pair_loss_dict = loss_func(grid_atlas, loader, src_id, dst_id)
            for key, val in pair_loss_dict.items():
                pair_loss_dict[key] = torch.nan_to_num(val)
            loss_dict.update(pair_loss_dict)
        if pose_reg_weight > 0:
            loss_dict.update(grid_atlas_pose_trust_region_loss(grid_atlas, thresh_rad=pose_thresh_rad,
                                                               thresh_m=pose_thresh_m, weight=pose_reg_weight))
        total_loss = sum(loss_dict.values()) if loss_dict else None
        if isinstance(total_loss, torch.Tensor) and total_loss.requires_grad and not torch.isnan(total_loss):
            total_loss.backward()
            if allreduce is not None:
                allreduce([p.grad for p in pose_params() if p.grad is not None])
            optimizer.step()
        losses_hist.append(total_loss.detach() if isinstance(total_loss, torch.Tensor) else torch.zeros((), device=grid_atlas.device))
        params_curr = [p.clone().detach() for p in pose_params()]
        relchange = relative_param_change(params_curr, params_prev)
        params_prev = params_curr
        if rel_change_thresh > 0 and relchange is not None and float(relchange) < rel_change_thresh:
            break
        it += 1
    ev1.record()
    torch.cuda.synchronize()
    return {"cpu_time_sec": time.perf_counter() - t0, "gpu_time_sec": ev0.elapsed_time(ev1) / 1e3,
            "iteration_results": iteration_results, "losses": torch.stack(losses_hist).cpu() if losses_hist else None,
            "iterations": it if it <= num_iters else num_iters + 1}


def align_multiple_submaps_hierarchical(grid_atlas: GridAtlas, dataset=None, level_iters=10, finetune_iters=10,
                                        level_thresh=0.0, lr=1e-2, align_weight=3000, align_loss="L2",
                                        use_bound=True, stability_thresh=0, subsample_points=None,
                                        latent_levels=None, skip_finetune=False, submap_pairs=None,
                                        pose_reg_weight=0, pose_thresh_m=1.0, pose_thresh_rad=1.0, gm_scale_sdf=0.1,
                                        device="cuda:0", verbose=True, save_iterations=False, pair_filter=None,
                                        allreduce=None):
    """miso.py:217-322: latent-space levels (one batched launch per iteration), then -- unless `skip_finetune`
    (True in the shipped configs, scannet.yaml:63) -- the SDF-space fine-tune with pairwise_loss_sdf on `dataset`."""
    if not use_bound or stability_thresh > 0:
        raise NotImplementedError("fused alignment needs use_bound=True and stability_thresh=0")
    if not skip_finetune and dataset is None:
        raise ValueError("the SDF-space fine-tune needs a dataset of (model_input, gt) batches")
    grid_atlas.precompute_coordinates_for_alignment()
    info = dict()
    cpu_total, gpu_total = 0.0, 0.0
    if latent_levels is None:
        latent_levels = range(grid_atlas.num_levels)
    for curr_level in latent_levels:
        loss_name = f"hier_latent_level{curr_level}_{align_loss}"
        level_dict = generic_align_multiple_submaps(
            grid_atlas, dataset, (loss_name, None), num_iters=level_iters, rel_change_thresh=level_thresh, lr=lr,
            submap_pairs=submap_pairs, pose_reg_weight=pose_reg_weight, pose_thresh_m=pose_thresh_m,
            pose_thresh_rad=pose_thresh_rad, verbose=verbose, save_iterations=save_iterations, level=curr_level,
            align_weight=align_weight, subsample_points=subsample_points, pair_filter=pair_filter,
            allreduce=allreduce, align_loss=align_loss)
        cpu_total += level_dict["cpu_time_sec"]
        gpu_total += level_dict["gpu_time_sec"]
        info[loss_name] = level_dict
    if not skip_finetune:
        def loss_func(atlas, loader, src_id, dst_id):
            return pairwise_loss_sdf(atlas, loader, src_id, dst_id, align_weight=align_weight, align_loss=align_loss,
                                     use_bound=use_bound, stability_thresh=stability_thresh,
                                     subsample_points=subsample_points, gm_scale_sdf=gm_scale_sdf, device=device)
        loss_name = f"hier_sdf_{align_loss}"
        final_dict = generic_align_multiple_submaps(
            grid_atlas, dataset, (loss_name, loss_func), lr=lr, num_iters=finetune_iters, submap_pairs=submap_pairs,
            pose_reg_weight=pose_reg_weight, pose_thresh_m=pose_thresh_m, pose_thresh_rad=pose_thresh_rad,
            verbose=verbose, save_iterations=save_iterations, pair_filter=pair_filter, allreduce=allreduce)
        cpu_total += final_dict["cpu_time_sec"]
        gpu_total += final_dict["gpu_time_sec"]
        info[loss_name] = final_dict
    info["cpu_time_sec"] = cpu_total
    info["gpu_time_sec"] = gpu_total
    return info


def gauss_newton_dst_step(batch: AlignBatch, lm_lambda: float = 1e-4):
    """Explicit 6-DoF normal equations of the latent residual w.r.t. a right-multiplied twist of each
    pair's DST pose (conventions of Tracker.lm_step, grid_opt/slam/tracker.py:179-197):
    H = J^T J + lambda I, g = J^T r, delta = solve(H, -g).  Returns (delta (P,6), H (P,6,6), g (P,6))."""
    out = batch.launch(batch.pair_poses(), want_gn=True)
    P = out.shape[0]
    g = out[:, 24:30]
    Hm = out[:, 30:66].reshape(P, 6, 6) + lm_lambda * torch.eye(6, dtype=out.dtype, device=out.device)
    delta = torch.linalg.solve(Hm, -g.unsqueeze(-1)).squeeze(-1)
    return delta, Hm, g
