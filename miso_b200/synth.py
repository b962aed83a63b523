"""Seeded synthetic inputs shaped like the reference's data -- there is no
network for ScanNet / Newer College, so the benchmarks and the parity tests draw from these.

  * rgbd_batch   -- "RGB-D-sampled" points following grid_opt/datasets/sdf_rgbd.py:381-483 and
                    utils_sample.py:195-302: per keyframe `n_rays` pixels of a 480x640 pinhole camera,
                    per ray 1 surface sample + 7 N(depth,0.1) samples + 19 stratified samples in
                    [0.07, depth+0.1]; sdf = |dir| * (depth - z) (bounds_ray, sdf_rgbd.py:525-534);
                    valid = |sdf| < trunc, sign = +-1 beyond the truncation band.  The scene is an
                    axis-aligned room so depth is closed-form.
  * lidar_batch  -- "LiDAR-sampled" points following grid_opt/datasets/sdf_3d_lidar.py:214-347.
  * model_cfg    -- cfg['model'] dict shaped like configs/rgbd/scannet.yaml:7-29.
  * latent_atlas -- N overlapping submaps whose grids sample one smooth global latent field, for the
                    alignment workload (demo/align_submaps.py:240-314).
Everything runs on the device it is asked for (CPU for the oracle, CUDA for the product).
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

SCANNET_SUBMAP_BOUND = [[-10.0, 10.0], [-5.0, 5.0], [-10.0, 10.0]]   # system.submap_local_bound, scannet.yaml:70
NCD_QUAD_BOUND = [[-45.0, 45.0], [-45.0, 45.0], [-5.0, 15.0]]        # ncd_quad.yaml:68


def model_cfg(bound, n_levels=2, feature_dim=4, base_cell_size=0.5, per_level_scale=5, init_stddev=0.0,
              hidden_dim=64, hidden_layers=1, num_poses=1, fix=True, second_order=True) -> dict:
    return {
        "name": "grid_net", "spatial_dim": 3,
        "decoder": {"type": "mlp", "hidden_dim": hidden_dim, "hidden_layers": hidden_layers, "out_dim": 1,
                    "pos_invariant": True, "fix": fix, "pretrained_model": None},
        "grid": {"type": "regular", "feature_dim": feature_dim, "init_stddev": init_stddev, "bound": bound,
                 "base_cell_size": base_cell_size, "per_level_scale": per_level_scale, "n_levels": n_levels,
                 "second_order_grid_sample": second_order},
        "pose": {"optimize": False, "num_poses": num_poses},
    }


def _rot_yaw_pitch(yaw, pitch):
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    return (Ry @ Rx).astype(np.float32)


def keyframe_poses(num_kf: int, bound=SCANNET_SUBMAP_BOUND, seed=55, margin=2.0):
    """Random keyframe poses (R (K,3,3), t (K,3,1)) inside the room, camera z looking roughly horizontal."""
    rng = np.random.RandomState(seed)
    b = np.asarray(bound, dtype=np.float32)
    lo, hi = b[:, 0] + margin, b[:, 1] - margin
    R = np.stack([_rot_yaw_pitch(rng.uniform(-math.pi, math.pi), rng.uniform(-0.3, 0.3)) for _ in range(num_kf)])
    t = rng.uniform(lo, hi, size=(num_kf, 3)).astype(np.float32)[..., None]
    return torch.from_numpy(R), torch.from_numpy(t)


def rgbd_batch(n_points: int, num_kf: int = 49, bound=SCANNET_SUBMAP_BOUND, trunc_dist=0.15, seed=55,
               wall_margin=1.0, n_surf=8, n_strat=19, dist_behind=0.1, min_depth=0.07,
               poses: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Returns (model_input, gt, (R,t)) as CPU tensors with the reference's batch layout
    (leading batch dim of 1 as produced by DataLoader(batch_size=1), submap_dataset.py:57-76)."""
    rng = np.random.RandomState(seed)
    R, t = poses if poses is not None else keyframe_poses(num_kf, bound, seed)
    num_kf = R.shape[0]
    per_ray = n_surf + n_strat
    n_rays_total = (n_points + per_ray - 1) // per_ray
    kf = rng.randint(0, num_kf, size=n_rays_total)
    u = rng.uniform(0, 640, size=n_rays_total)
    v = rng.uniform(0, 480, size=n_rays_total)
    dirs_c = np.stack([(u - 320.0) / 577.0, (v - 240.0) / 577.0, np.ones_like(u)], 1).astype(np.float32)
    Rn, tn = R.numpy(), t.numpy()[..., 0]
    d_w = np.einsum("nij,nj->ni", Rn[kf], dirs_c)
    o_w = tn[kf]
    b = np.asarray(bound, dtype=np.float32)
    wall_lo, wall_hi = b[:, 0] + wall_margin, b[:, 1] - wall_margin
    with np.errstate(divide="ignore", invalid="ignore"):
        t_hi = np.where(d_w > 0, (wall_hi - o_w) / d_w, np.inf)
        t_lo = np.where(d_w < 0, (wall_lo - o_w) / d_w, np.inf)
    depth = np.minimum(t_hi, t_lo).min(axis=1).astype(np.float32)   # z-depth along the optical axis
    depth = np.clip(depth, 0.2, 12.0)
    z = np.empty((n_rays_total, per_ray), dtype=np.float32)
    z[:, 0] = depth
    z[:, 1:n_surf] = depth[:, None] + rng.normal(0, 0.1, size=(n_rays_total, n_surf - 1))
    z[:, 1:n_surf] = np.clip(z[:, 1:n_surf], min_depth, depth[:, None] + dist_behind)
    edges = np.linspace(0, 1, n_strat + 1)[None, :-1] + rng.uniform(0, 1.0 / n_strat, size=(n_rays_total, n_strat))
    z[:, n_surf:] = min_depth + edges * (depth[:, None] + dist_behind - min_depth)
    pts_c = dirs_c[:, None, :] * z[..., None]                        # camera/keyframe frame
    sdf = np.linalg.norm(dirs_c, axis=1, keepdims=True) * (depth[:, None] - z)
    coords = pts_c.reshape(-1, 3)[:n_points]
    sdf = sdf.reshape(-1, 1)[:n_points].astype(np.float32)
    ids = np.repeat(kf, per_ray)[:n_points].astype(np.int64)[:, None]
    gt_sdf = torch.from_numpy(sdf)
    valid = torch.abs(gt_sdf) < trunc_dist
    signs = torch.zeros_like(gt_sdf)
    signs[gt_sdf < -trunc_dist] = -1
    signs[gt_sdf > trunc_dist] = 1
    model_input = {"coords_frame": torch.from_numpy(np.ascontiguousarray(coords))[None],
                   "sample_frame_ids": torch.from_numpy(ids)[None],
                   "weights": torch.ones_like(gt_sdf)[None]}
    gt = {"sdf": gt_sdf[None], "sdf_valid": valid[None], "sdf_signs": signs[None]}
    return model_input, gt, (R, t)


def lidar_batch(n_points: int, num_kf: int = 8, bound=NCD_QUAD_BOUND, trunc_dist=0.5, seed=55, max_range=60.0):
    """LiDAR-style samples (sdf_3d_lidar.py:214-347): per surface point 1 surface + 4 near (N(0,0.25) along
    the ray) + 2 free-space (ratio U[0.5, 1-0.5/d]) + 1 behind (0.25 + U*0.5); weights 1 + 0.4 - 0.8 d/60
    for surface/near samples.  Scene: ground plane z = -1 and vertical walls at +-40 m."""
    rng = np.random.RandomState(seed)
    per = 8
    n_surf = (n_points + per - 1) // per
    b = np.asarray(bound, dtype=np.float32)
    kf = rng.randint(0, num_kf, size=n_surf)
    pos = np.stack([rng.uniform(-20, 20, num_kf), rng.uniform(-20, 20, num_kf), rng.uniform(0.5, 2.0, num_kf)], 1)
    yaw = rng.uniform(-math.pi, math.pi, num_kf)
    R = np.stack([np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]]) for a in yaw])
    az = rng.uniform(-math.pi, math.pi, n_surf)
    el = rng.uniform(-0.4, 0.25, n_surf)
    d_l = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], 1)
    d_w = np.einsum("nij,nj->ni", R[kf], d_l)
    o_w = pos[kf]
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(d_w[:, 2] < 0, (-1.0 - o_w[:, 2]) / d_w[:, 2], np.inf)
        t_wx = np.where(d_w[:, 0] > 0, (40 - o_w[:, 0]) / d_w[:, 0], (-40 - o_w[:, 0]) / d_w[:, 0])
        t_wy = np.where(d_w[:, 1] > 0, (40 - o_w[:, 1]) / d_w[:, 1], (-40 - o_w[:, 1]) / d_w[:, 1])
    dist = np.minimum(np.minimum(t_ground, t_wx), t_wy)
    dist = np.clip(np.nan_to_num(dist, nan=max_range, posinf=max_range), 1.0, max_range)
    offs = np.empty((n_surf, per))
    offs[:, 0] = 0.0
    offs[:, 1:5] = rng.normal(0, 0.25, size=(n_surf, 4))
    ratio = rng.uniform(0.5, np.maximum(0.5, 1 - 0.5 / dist)[:, None], size=(n_surf, 2))
    offs[:, 5:7] = -(1 - ratio) * dist[:, None]
    offs[:, 7] = 0.25 + rng.uniform(0, 1, n_surf) * 0.5
    r = dist[:, None] + offs
    pts_l = d_l[:, None, :] * r[..., None]
    sdf = (-offs).astype(np.float32)
    w = np.ones((n_surf, per), dtype=np.float32)
    w[:, :5] = (1 + 0.4 - 0.8 * dist / max_range)[:, None]
    coords = pts_l.reshape(-1, 3)[:n_points].astype(np.float32)
    gt_sdf = torch.from_numpy(sdf.reshape(-1, 1)[:n_points])
    weights = torch.from_numpy(w.reshape(-1, 1)[:n_points])
    ids = torch.from_numpy(np.repeat(kf, per)[:n_points].astype(np.int64)[:, None])
    valid = torch.abs(gt_sdf) < trunc_dist
    signs = torch.zeros_like(gt_sdf)
    signs[gt_sdf < -trunc_dist] = -1
    signs[gt_sdf > trunc_dist] = 1
    model_input = {"coords_frame": torch.from_numpy(coords)[None], "sample_frame_ids": ids[None],
                   "weights": weights[None]}
    gt = {"sdf": gt_sdf[None], "sdf_valid": valid[None], "sdf_signs": signs[None]}
    return model_input, gt, (torch.from_numpy(R.astype(np.float32)), torch.from_numpy(pos.astype(np.float32))[..., None])


def decoder_weights(in_dim=8, hidden=64, seed=0):
    """MLPNet(in_dim,1,64,1,bias=True) default nn.Linear init under torch.manual_seed(seed): the shipped
    decoder_indoor.pt / decoder_quad.pt are external downloads (README.md:55), so weights are synthetic."""
    g = torch.Generator().manual_seed(seed)
    def lin(o, i):
        k = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * k, (torch.rand(o, generator=g) * 2 - 1) * k
    W1, b1 = lin(hidden, in_dim)
    W2, b2 = lin(hidden, hidden)
    W3, b3 = lin(1, hidden)
    return {"network.0.weight": W1, "network.0.bias": b1, "network.2.weight": W2, "network.2.bias": b2,
            "network.4.weight": W3, "network.4.bias": b3}


# ------------------------------------------------------------------------------------------------
# alignment workload
# ------------------------------------------------------------------------------------------------
def latent_field(xw: torch.Tensor, channels: int = 8, seed: int = 7) -> torch.Tensor:
    """Smooth global latent field R^3 -> R^channels (sum of a few low-frequency sinusoids)."""
    g = torch.Generator().manual_seed(seed)
    K = 3
    freq = (torch.rand(channels, K, 3, generator=g) * 2 - 1) * 0.35
    phase = torch.rand(channels, K, generator=g) * 2 * math.pi
    amp = torch.rand(channels, K, generator=g) * 0.5 + 0.25
    freq, phase, amp = freq.to(xw), phase.to(xw), amp.to(xw)
    arg = torch.einsum("nd,ckd->nck", xw, freq) + phase[None]
    return (amp[None] * torch.sin(arg)).sum(-1)


def submap_layout(num_submaps: int, spacing=(12.0, 6.0), seed=3):
    """True world poses of submaps tiled on a floor plan with >= 30 % overlap of their local bounds
    (bound x-extent 20 m, z-extent 20 m): a ceil(sqrt(n)) grid with small random yaw."""
    rng = np.random.RandomState(seed)
    cols = int(math.ceil(math.sqrt(num_submaps)))
    Rs, ts = [], []
    for i in range(num_submaps):
        r, c = divmod(i, cols)
        yaw = rng.uniform(-0.15, 0.15) if i > 0 else 0.0
        Rs.append(torch.from_numpy(_rot_yaw_pitch(yaw, 0.0)))
        ts.append(torch.tensor([[c * spacing[0]], [0.0], [r * spacing[1]]], dtype=torch.float32))
    return Rs, ts


def perturb_poses(Rs, ts, rot_deg=10.0, trans_m=0.5, seed=55):
    """demo/align_submaps.py:267-273: perturb submaps i >= 1 with Gaussian rotation / translation noise."""
    rng = np.random.RandomState(seed)
    outR, outt = [Rs[0].clone()], [ts[0].clone()]
    for i in range(1, len(Rs)):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = math.radians(rot_deg) * rng.normal() * 0.5
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        dR = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
        outR.append((Rs[i].double() @ torch.from_numpy(dR)).float())
        outt.append(ts[i] + torch.from_numpy(rng.normal(0, trans_m, size=(3, 1)).astype(np.float32)) * 0.5)
    return outR, outt


def fill_submap_from_field(level_shapes: Sequence[Tuple[int, ...]], bound, R_true, t_true, fdim=4, seed=7,
                           device="cpu") -> List[torch.Tensor]:
    """Level tensors (1,C,Z,Y,X) whose voxel (z,y,x) holds latent_field(world position of the voxel centre)
    channels [l*fdim, (l+1)*fdim)."""
    b = torch.tensor(bound, dtype=torch.float32, device=device)
    out = []
    for l, shp in enumerate(level_shapes):
        _, Cc, Z, Y, X = shp
        zs = (torch.arange(Z, device=device) + 0.5) / Z
        ys = (torch.arange(Y, device=device) + 0.5) / Y
        xs = (torch.arange(X, device=device) + 0.5) / X
        zz, yy, xx = torch.meshgrid(zs, ys, xs, indexing="ij")
        p = torch.stack([xx, yy, zz], -1).reshape(-1, 3) * (b[:, 1] - b[:, 0]) + b[:, 0]
        pw = p @ R_true.to(device).T + t_true.to(device).T
        f = latent_field(pw, channels=fdim * len(level_shapes), seed=seed)[:, l * fdim:(l + 1) * fdim]
        out.append(f.reshape(Z, Y, X, Cc).permute(3, 0, 1, 2).unsqueeze(0).contiguous())
    return out
