"""Host-side mirror of the reference's model interface for the hot path: same class names,
constructor arguments, attribute names and state-dict keys as

    grid_opt/models/grid_modules.py  (FeatureGrid            :41-123)
    grid_opt/models/modules.py       (MLPNet                 :11-40)
    grid_opt/models/grid_net.py      (GridNet                :17-352)
    grid_opt/models/grid_atlas.py    (GridAtlas              :18-587, the parts alignment uses)
    grid_opt/utils/utils.py          (normalize_coordinates :22-51, grid_interp_regular :143-164,
                                      grid_decode :194-208, all_grid_positions :294-307)

so a user of the reference finds the same objects; the compute underneath is the CUDA path.
What changes on purpose: grid parameters are stored channels_last_3d (logical shape `(1,C,Z,Y,X)`
unchanged), `grid_sample_func` is always `miso_b200.cuda_gridsample.grid_sample_3d`, and
`GridNet.forward` takes the fused kernel when the decoder is fixed (cfg decoder.fix: True).
"""
import math
import os
from copy import deepcopy
from typing import Tuple

import numpy as np
import torch
from torch import nn

from . import cuda_gridsample as cu
from . import field as _field
from . import geometry as utils_geometry


# ------------------------------------------------------------------------------------------------
# utils.py mirrors
# ------------------------------------------------------------------------------------------------
def normalize_coordinates(queries: torch.Tensor, bounds: torch.Tensor):
    """utils.py:22-51 (same operation order: 2*(x-bmin)/(bmax-bmin) - 1)."""
    d = bounds.shape[0]
    assert queries.shape[-1] == d
    if queries.dim() == 2:
        bounds_min = bounds[:, 0].view(1, -1)
        bounds_max = bounds[:, 1].view(1, -1)
    elif queries.dim() == 3:
        bounds_min = bounds[:, 0].view(1, 1, -1)
        bounds_max = bounds[:, 1].view(1, 1, -1)
    else:
        raise ValueError("queries tensor must be either 2D or 3D")
    return 2 * (queries - bounds_min) / (bounds_max - bounds_min) - 1


def denormalize_coordinates(normalized_queries: torch.Tensor, bounds: torch.Tensor):
    """utils.py:53-80."""
    if normalized_queries.dim() == 2:
        bounds_min = bounds[:, 0].view(1, -1)
        bounds_max = bounds[:, 1].view(1, -1)
    elif normalized_queries.dim() == 3:
        bounds_min = bounds[:, 0].view(1, 1, -1)
        bounds_max = bounds[:, 1].view(1, 1, -1)
    else:
        raise ValueError("normalized_queries tensor must be either 2D or 3D")
    return (normalized_queries + 1) / 2 * (bounds_max - bounds_min) + bounds_min


def all_grid_positions(features):
    """utils.py:294-307: voxel-centre coordinates, shape (1,Z,Y,X,3), last dim (x,y,z)."""
    B, Cc, D, H, W = features.shape
    half_dx = 0.5 * 1 / D
    half_dy = 0.5 * 1 / H
    half_dz = 0.5 * 1 / W
    xs = 2 * torch.linspace(half_dx, 1 - half_dx, D) - 1.
    ys = 2 * torch.linspace(half_dy, 1 - half_dy, H) - 1.
    zs = 2 * torch.linspace(half_dz, 1 - half_dz, W) - 1.
    xv, yv, zv = torch.meshgrid([xs, ys, zs], indexing="ij")
    grid = torch.stack((zv, yv, xv), axis=-1)
    return grid.unsqueeze(0)


def grid_interp_regular(reg_grids, x, ignore_level=None):
    """utils.py:143-164."""
    num_levels = len(reg_grids)
    if ignore_level is None:
        ignore_level = np.zeros(num_levels).astype(bool)
    level_feats = []
    for level in range(num_levels):
        feats = reg_grids[level].interpolate(x)
        if not ignore_level[level]:
            level_feats.append(feats)
        else:
            level_feats.append(torch.zeros_like(feats))
    return torch.cat(level_feats, dim=1)


def grid_decode(feats, x, decoder=None, pos_invariant=True):
    """utils.py:194-208."""
    assert feats.ndim == 2
    if decoder is not None:
        inputs = feats if pos_invariant else torch.cat((feats, x), dim=1)
        preds = decoder(inputs)
    else:
        preds = feats
    return preds


# ------------------------------------------------------------------------------------------------
# grid_modules.py / modules.py mirrors
# ------------------------------------------------------------------------------------------------
class FeatureGrid(nn.Module):
    """Dense 3D feature grid `(1,C,Z,Y,X)` (grid_modules.py:41-123)."""

    def __init__(self, d, fdim, bound, cell_size, name="grid", dtype=torch.float32, initial_feature=None,
                 init_stddev=0.0, second_order_grid_sample=False):
        super().__init__()
        if d != 3:
            raise NotImplementedError("miso_b200 implements spatial_dim 3 only (shipped MISO configs)")
        self.d = d
        self.fdim = fdim
        self.bound = bound
        self.cell_size = cell_size
        self.dtype = dtype
        self.name = name
        assert self.bound.shape == (d, 2)
        grid_len = (self.bound[:, 1] - self.bound[:, 0]).cpu().numpy()
        grid_size = np.ceil(grid_len / cell_size).astype(int)
        feature_shape = (1, self.fdim, int(grid_size[2]), int(grid_size[1]), int(grid_size[0]))
        if initial_feature is None:
            initial_feature = torch.randn(feature_shape, dtype=self.dtype) * init_stddev
        assert initial_feature.shape == feature_shape
        self.feature = torch.nn.Parameter(initial_feature.contiguous(memory_format=torch.channels_last_3d))
        # the reference picks F.grid_sample or the CUDA double-backward plugin here (:63-69);
        # the B200 op covers both, so the flag only records the request
        self.second_order_grid_sample = second_order_grid_sample
        self.grid_sample_func = cu.grid_sample_3d
        self._bound_host = _field.bound_to_list(bound)

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        # .to(device) / .cuda() re-allocate: keep the channels-last layout and follow with `bound`
        _field.to_channels_last_3d_(self.feature)
        if isinstance(self.bound, torch.Tensor):
            self.bound = fn(self.bound)
        return self

    def interpolate(self, x):
        """grid_modules.py:72-95."""
        x = normalize_coordinates(x, self.bound)
        N = x.shape[0]
        sample_coords = x.reshape(1, N, 1, 1, 3)
        feats = self.grid_sample_func(self.feature, sample_coords, align_corners=False,
                                      padding_mode="zeros")[0, :, :, 0, 0].transpose(0, 1)
        return feats

    def norm(self):
        return self.feature.norm()

    def num_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def lock(self):
        for param in self.parameters():
            param.requires_grad = False

    def unlock(self):
        for param in self.parameters():
            param.requires_grad = True

    def zero_features(self):
        with torch.no_grad():
            self.feature.zero_()

    def randn_features(self, std):
        with torch.no_grad():
            new_feat = torch.randn(self.feature.shape, dtype=self.dtype) * std
            self.feature.copy_(new_feat.to(self.feature))

    def vertex_positions(self, denormalize=True) -> torch.Tensor:
        """grid_modules.py:111-123 (computed on the host exactly like the reference, then moved)."""
        pos_nrm = all_grid_positions(self.feature)
        pos_nrm = torch.flatten(pos_nrm.squeeze(0), start_dim=0, end_dim=-2)
        if denormalize:
            return denormalize_coordinates(pos_nrm, self.bound.to(pos_nrm))
        return pos_nrm


class MLPNet(nn.Module):
    """modules.py:11-40."""

    def __init__(self, input_dim, output_dim, hidden_dim=64, hidden_layers=1, bias=False, acti_func=nn.ReLU,
                 pretrained_path=None, no_optimize=False):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.layers = [nn.Linear(input_dim, hidden_dim, bias=bias), acti_func()]
        for _ in range(hidden_layers):
            self.layers.append(nn.Linear(hidden_dim, hidden_dim, bias=bias))
            self.layers.append(acti_func())
        self.layers.append(nn.Linear(hidden_dim, output_dim, bias=bias))
        self.network = nn.Sequential(*self.layers)
        if pretrained_path is not None:
            self.load(pretrained_path)
        if no_optimize:
            for param in self.parameters():
                param.requires_grad = False

    def forward(self, x):
        return self.network(x)

    def save(self, filepath):
        torch.save(self.state_dict(), filepath)

    def load(self, filepath):
        self.load_state_dict(torch.load(filepath))


class BaseNet(nn.Module):
    """models/base_net.py:11-40."""

    def __init__(self, cfg: dict, device="cpu", dtype=torch.float32):
        super().__init__()
        self.cfg = cfg
        self.d = self.cfg["spatial_dim"]
        self.device = device
        self.dtype = dtype
        assert self.d == 2 or self.d == 3
        self.bound = torch.tensor(np.asarray(cfg["grid"]["bound"]), device=device, dtype=dtype)
        assert self.bound.shape == (self.d, 2)


class GridNet(BaseNet):
    """One submap: L FeatureGrid levels + MLP decoder + per-keyframe pose corrections
    (grid_net.py:17-352).  `forward` / `query_feature` / `params_at_level` keep their signatures."""

    def __init__(self, cfg: dict, device="cuda:0", dtype=torch.float32, initial_features=dict()):
        super().__init__(cfg, device, dtype)
        self.initial_features = initial_features
        self.init_grid(cfg)
        self.init_decoder(cfg)
        self.init_poses(cfg)
        self.to(device)
        self._bound_host = _field.bound_to_list(cfg["grid"]["bound"])
        self._spec_cache = None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self.bound = fn(self.bound)
        self._spec_cache = None
        return self

    def save(self, ckpt_dir, ckpt_prefix):
        self.decoder.save(os.path.join(ckpt_dir, f"{ckpt_prefix}_decoder.pt"))

    def init_grid(self, cfg):
        self.num_levels = cfg["grid"]["n_levels"]
        self.second_order_grid_sample = bool(cfg["grid"].get("second_order_grid_sample", False))
        base_cell_size = cfg["grid"]["base_cell_size"]
        scale_factor = cfg["grid"]["per_level_scale"]
        self.fdim = cfg["grid"]["feature_dim"]
        self.features = nn.ModuleList()
        self.feature_stability = nn.ModuleList()
        self.grid_type = cfg["grid"]["type"]
        if self.grid_type != "regular":
            raise NotImplementedError("miso_b200 implements grid.type 'regular' (the shipped configs); "
                                      "the VM variants are out of scope (SURVEY.md section 2, row 1)")
        self.cell_sizes = []
        bound_cpu = self.bound.detach().cpu()
        for level in range(self.num_levels):
            cell_size = base_cell_size / (scale_factor ** level)
            self.cell_sizes.append(cell_size)
            init_feature = self.initial_features.get(level, None)
            self.features.append(FeatureGrid(d=self.d, fdim=self.fdim, bound=bound_cpu, cell_size=cell_size,
                                             name=f"feat-{level}", dtype=self.dtype, initial_feature=init_feature,
                                             init_stddev=cfg["grid"]["init_stddev"],
                                             second_order_grid_sample=self.second_order_grid_sample))
            self.feature_stability.append(FeatureGrid(d=self.d, fdim=1, bound=bound_cpu, cell_size=cell_size,
                                                      name=f"stab-{level}", dtype=self.dtype, initial_feature=None,
                                                      init_stddev=0.0,
                                                      second_order_grid_sample=self.second_order_grid_sample))
        self.ignore_level_ = np.zeros(self.num_levels).astype(bool)

    def init_decoder(self, cfg):
        self.decoder_hidden_dim = cfg["decoder"]["hidden_dim"]
        self.decoder_hidden_layers = cfg["decoder"]["hidden_layers"]
        self.decoder_out_dim = cfg["decoder"]["out_dim"]
        self.pos_invariant = cfg["decoder"]["pos_invariant"]
        self.decoder_fixed = cfg["decoder"]["fix"]
        self.decoder_type = cfg["decoder"]["type"]
        input_dim = self.num_levels * self.fdim
        if not self.pos_invariant:
            input_dim += self.d
        if self.decoder_type == "mlp":
            self.decoder = MLPNet(input_dim=input_dim, output_dim=self.decoder_out_dim,
                                  hidden_dim=self.decoder_hidden_dim, hidden_layers=self.decoder_hidden_layers,
                                  bias=True, pretrained_path=cfg["decoder"]["pretrained_model"],
                                  no_optimize=self.decoder_fixed)
        elif self.decoder_type == "none":
            self.decoder = None
        else:
            raise ValueError(f"Unknown decoder type: {self.decoder_type}")

    def init_poses(self, cfg):
        self.num_poses = cfg["pose"]["num_poses"]
        self.optimize_pose = cfg["pose"]["optimize"]
        self.rotation_corrections = torch.nn.Parameter(torch.zeros(self.num_poses, 3).float(),
                                                       requires_grad=self.optimize_pose)
        self.translation_corrections = torch.nn.Parameter(torch.zeros(self.num_poses, 3, 1).float(),
                                                          requires_grad=self.optimize_pose)
        self.pose_estimates_known = [False] * self.num_poses
        self.register_buffer("Rwk", utils_geometry.identity_rotations(self.num_poses))
        self.register_buffer("twk", torch.zeros(size=(self.num_poses, 3, 1)))
        self.locked_pose_indices = set()
        self._pose_key_to_id = dict()

    # ---- level / pose bookkeeping (grid_net.py:159-262) ------------------------------------------
    def ignore_level(self, l):
        self.ignore_level_[l] = True
        self._spec_cache = None

    def include_level(self, l):
        self.ignore_level_[l] = False
        self._spec_cache = None

    def lock_level(self, l):
        self.features[l].lock()
        self.feature_stability[l].lock()

    def unlock_level(self, l):
        self.features[l].unlock()
        self.feature_stability[l].unlock()

    def lock_feature(self):
        for level in range(self.num_levels):
            self.lock_level(level)

    def unlock_feature(self):
        for level in range(self.num_levels):
            self.unlock_level(level)

    def lock_pose(self):
        self.rotation_corrections.requires_grad_(False)
        self.translation_corrections.requires_grad_(False)
        self.lock_all_pose_indices()

    def unlock_pose(self):
        self.rotation_corrections.requires_grad_(True)
        self.translation_corrections.requires_grad_(True)
        self.unlock_all_pose_indices()

    def lock_pose_index(self, pose_index: int):
        self.locked_pose_indices.add(pose_index)

    def lock_all_pose_indices(self):
        self.locked_pose_indices = set(range(self.num_poses))

    def unlock_pose_index(self, pose_index: int):
        self.locked_pose_indices.remove(pose_index)

    def unlock_all_pose_indices(self):
        self.locked_pose_indices.clear()

    def pose_correction(self, kf_id: int):
        r = self.rotation_corrections[[kf_id], :]
        t = self.translation_corrections[kf_id, :, :]
        if kf_id in self.locked_pose_indices:
            r = r.clone().detach()
            t = t.clone().detach()
        return r, t

    def set_initial_kf_pose(self, kf_id: int, Rwk: torch.Tensor, twk: torch.Tensor, kf_key=None):
        assert Rwk.shape == (3, 3)
        assert twk.shape == (3, 1)
        assert kf_id < self.num_poses, f"KF ID {kf_id} exceeds the number of poses {self.num_poses}!"
        self.pose_estimates_known[kf_id] = True
        self.Rwk[kf_id, :, :] = Rwk.to(self.Rwk)
        self.twk[kf_id, :, :] = twk.to(self.twk)
        with torch.no_grad():
            self.rotation_corrections[kf_id, :].zero_()
            self.translation_corrections[kf_id, :, :].zero_()
        if kf_key is not None:
            self._pose_key_to_id[kf_key] = kf_id

    def pose_key_to_id(self, kf_key):
        assert kf_key in self._pose_key_to_id, f"Key {kf_key} not found in pose key to ID mapping!"
        return self._pose_key_to_id[kf_key]

    def initial_kf_pose(self, kf_id: int):
        assert self.pose_estimates_known[kf_id], f"Initial pose estimate for KF {kf_id} is not available!"
        return self.Rwk[kf_id, :, :], self.twk[kf_id, :, :]

    def initial_kf_pose_in_world(self, kf_id: int):
        return self.initial_kf_pose(kf_id)

    def updated_kf_pose(self, kf_id: int):
        Rwk, twk = self.initial_kf_pose_in_world(kf_id)
        Dr, Dt = self.pose_correction(kf_id)
        return utils_geometry.apply_pose_correction(Rwk, twk, Dr, Dt)

    def updated_kf_pose_in_world(self, kf_id: int):
        return self.updated_kf_pose(kf_id)

    def updated_kf_pose_from_key(self, kf_key):
        return self.updated_kf_pose(self.pose_key_to_id(kf_key))

    def all_kf_poses(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Batched `updated_kf_pose` for every keyframe: (K,3,3), (K,3,1) -- one so3_exp_map instead of a
        per-keyframe Python loop (loss.py:764-774)."""
        R = torch.matmul(self.Rwk, utils_geometry.so3_exp_map(self.rotation_corrections))
        t = self.twk + self.translation_corrections
        return R, t

    def zero_features(self):
        for grid in self.features:
            grid.zero_features()

    def randn_features(self, std):
        for grid in self.features:
            grid.randn_features(std)

    # ---- the hot path -----------------------------------------------------------------------------
    def fused_spec(self):
        """FieldSpec when the fused kernels apply (decoder fixed, 64-wide single hidden layer,
        uniform channel count in {4,8,16}); None -> generic path."""
        if self._spec_cache is not None:
            return self._spec_cache if self._spec_cache != "no" else None
        ok = (self.decoder is not None and self.decoder_type == "mlp" and self.pos_invariant
              and self.decoder_hidden_dim == 64 and self.decoder_hidden_layers == 1 and self.decoder_out_dim == 1
              and not any(p.requires_grad for p in self.decoder.parameters())
              and (self.num_levels, self.fdim) in {(1, 4), (2, 4), (3, 4), (4, 4), (1, 8), (2, 8), (1, 16)}
              and self.features[0].feature.is_cuda)
        if not ok:
            self._spec_cache = "no"
            return None
        mask = 0
        for l in range(self.num_levels):
            if self.ignore_level_[l]:
                mask |= 1 << l
        self._spec_cache = _field.FieldSpec(self._bound_host, _field.DecoderSpec.from_mlp(self.decoder), mask)
        return self._spec_cache

    def level_tensors(self):
        return [g.feature for g in self.features]

    def query_feature(self, x: torch.Tensor):
        """grid_net.py:288-297."""
        assert x.ndim == 2, f"Invalid input coords shape {x.shape}!"
        assert x.shape[-1] == self.d
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(f.requires_grad for f in self.level_tensors()))
        if (not needs_graph and self.fdim % 4 == 0 and self.fdim <= 16 and x.is_cuda):
            mask = sum((1 << l) for l in range(self.num_levels) if self.ignore_level_[l])
            return _field.field_features_raw(self.level_tensors(), self._bound_host, x, mask)
        return grid_interp_regular(self.features, x, self.ignore_level_)

    def query_stability(self, x: torch.Tensor):
        assert x.ndim == 2, f"Invalid input coords shape {x.shape}!"
        assert x.shape[-1] == self.d
        return grid_interp_regular(self.feature_stability, x, None)

    def forward(self, x: torch.Tensor, noise_std=0):
        """grid_net.py:306-325."""
        spec = self.fused_spec()
        if spec is not None:
            needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(f.requires_grad for f in self.level_tensors()))
            if needs_graph:
                pred, _ = _field.fused_sdf(x, self.level_tensors(), spec)
            else:
                # inference / dense queries (utils_sdf.extract_fields): forward only, no Jacobian pass
                sdf, _, _, _ = _field.sdf_forward_raw(self.level_tensors(), spec, x, want_jac=False, want_gradx=False)
                pred = sdf.unsqueeze(1)
        else:
            feats = grid_interp_regular(self.features, x, self.ignore_level_)
            pred = grid_decode(feats, x, self.decoder, self.pos_invariant)
        if noise_std > 0:
            pred = pred + torch.randn(pred.shape, device=x.device) * noise_std
        return pred

    def forward_with_gradient(self, x: torch.Tensor):
        """(sdf (N,1), grad_x sdf (N,3)) from ONE fused launch; both outputs are differentiable wrt the
        grids (the gradient output carries the eikonal double-backward)."""
        spec = self.fused_spec()
        if spec is None:
            x = x if x.requires_grad else x.clone().requires_grad_(True)
            y = self.forward(x)
            g = torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True)[0]
            return y, g
        return _field.fused_sdf(x, self.level_tensors(), spec)

    def params_for_poses(self):
        return [self.rotation_corrections, self.translation_corrections]

    def params_for_features(self, stop_level=None):
        if stop_level is None:
            stop_level = self.num_levels
        assert stop_level <= self.num_levels
        params = []
        for level in range(stop_level):
            params += list(self.features[level].parameters())
        return params

    def params_at_level(self, level):
        """grid_net.py:339-351."""
        params = []
        target_levels = [level] if level < self.num_levels else range(self.num_levels)
        for l in target_levels:
            params += list(self.features[l].parameters())
            params += list(self.feature_stability[l].parameters())
        if not self.decoder_fixed:
            params += list(self.decoder.parameters())
        if self.optimize_pose:
            params += self.params_for_poses()
        return params


class GridAtlas(BaseNet):
    """N submaps + per-submap world pose (grid_atlas.py:18-587): the parts the alignment path uses.
    Keyframe <-> submap maps are plain integer tables (bit-exact by construction)."""

    def __init__(self, cfg: dict, device="cuda:0", dtype=torch.float32):
        super().__init__(cfg, device, dtype)
        self.cfg = cfg
        self.submaps = torch.nn.ModuleList()
        self.rotation_corrections = torch.nn.ParameterList()
        self.translation_corrections = torch.nn.ParameterList()
        self.R_world_submap_list = []
        self.t_world_submap_list = []
        self._submap_anchor_kf = []
        self._kf_id_to_submap_id = []
        self._submap_id_to_kf_ids = dict()
        self.curr_submap_id = -1
        self.curr_kf_id = -1
        self.num_levels = cfg["grid"]["n_levels"]
        self._coords_for_alignment = dict()

    # ---- construction (grid_atlas.py:96-169) ------------------------------------------------------
    def anchor_kf_for_submap(self, submap_id: int):
        return self._submap_anchor_kf[submap_id]

    def add_kf(self, Rsk: torch.Tensor, tsk: torch.Tensor):
        assert Rsk.shape == (3, 3)
        assert tsk.shape == (3, 1)
        assert self.curr_submap_id >= 0, "No submap is created yet. Create a submap first."
        submap_id = self.curr_submap_id
        kf_id_global = self.curr_kf_id + 1
        kf_id_submap = kf_id_global - self.anchor_kf_for_submap(self.curr_submap_id)
        self._kf_id_to_submap_id.append(submap_id)
        self.get_submap(submap_id).set_initial_kf_pose(kf_id_submap, Rsk, tsk, kf_key=f"KF{kf_id_global}")
        self._submap_id_to_kf_ids[submap_id].add(kf_id_global)
        self.curr_kf_id = kf_id_global
        return kf_id_global

    def add_submap(self, local_bound: torch.Tensor, Rws: torch.Tensor, tws: torch.Tensor, num_poses=1,
                   optimize_poses=True):
        assert Rws.shape == (3, 3)
        assert tws.shape == (3, 1)
        submap_id = len(self.submaps)
        cfg_model = deepcopy(self.cfg)
        cfg_model["grid"]["bound"] = local_bound.numpy()
        cfg_model["pose"]["num_poses"] = num_poses
        cfg_model["pose"]["optimize"] = optimize_poses
        self.submaps.append(GridNet(cfg=cfg_model, device=self.device, dtype=self.dtype))
        self.R_world_submap_list.append(Rws.to(self.device))
        self.t_world_submap_list.append(tws.to(self.device))
        anchor_kf = self.curr_kf_id + 1
        self._submap_anchor_kf.append(anchor_kf)
        self.rotation_corrections.append(torch.nn.Parameter(torch.zeros(1, 3).float().to(self.device),
                                                            requires_grad=True))
        self.translation_corrections.append(torch.nn.Parameter(torch.zeros(3, 1).float().to(self.device),
                                                               requires_grad=True))
        self.active_submaps = range(self.num_submaps)
        self.curr_submap_id = submap_id
        self._submap_id_to_kf_ids[submap_id] = set()
        self._submap_id_to_kf_ids[submap_id].add(anchor_kf)

    def add_existing_submap(self, submap: GridNet, Rws: torch.Tensor, tws: torch.Tensor):
        """Attach an already-built GridNet (multi-GPU build_submaps gathers submaps to rank 0 in id order,
        grid_atlas.py:145-150)."""
        self.submaps.append(submap)
        self.R_world_submap_list.append(Rws.to(self.device))
        self.t_world_submap_list.append(tws.to(self.device))
        self._submap_anchor_kf.append(self.curr_kf_id + 1)
        self.rotation_corrections.append(torch.nn.Parameter(torch.zeros(1, 3).float().to(self.device)))
        self.translation_corrections.append(torch.nn.Parameter(torch.zeros(3, 1).float().to(self.device)))
        self.active_submaps = range(self.num_submaps)
        self.curr_submap_id = len(self.submaps) - 1
        self._submap_id_to_kf_ids[self.curr_submap_id] = set()

    def set_submap_pose(self, submap_id: int, Rws: torch.Tensor, tws: torch.Tensor):
        """grid_atlas.py:171-187 (also resets the corrections to zero)."""
        assert Rws.shape == (3, 3)
        assert tws.shape == (3, 1)
        with torch.no_grad():
            self.R_world_submap_list[submap_id].copy_(Rws.to(self.device))
            self.t_world_submap_list[submap_id].copy_(tws.to(self.device))
            self.rotation_corrections[submap_id].zero_()
            self.translation_corrections[submap_id].zero_()

    def set_submap_pose_correction(self, submap_id: int, R_delta: torch.Tensor, t_delta: torch.Tensor):
        assert R_delta.shape == (1, 3)
        assert t_delta.shape == (3, 1)
        with torch.no_grad():
            self.rotation_corrections[submap_id].copy_(R_delta)
            self.translation_corrections[submap_id].copy_(t_delta)

    @property
    def num_submaps(self):
        return len(self.submaps)

    @property
    def num_keyframes(self):
        return self.curr_kf_id + 1

    def submap_id_for_kf(self, kf_id: int):
        return self._kf_id_to_submap_id[kf_id]

    def submap_id_for_kf_batch(self, kf_ids: torch.Tensor) -> torch.Tensor:
        """grid_atlas.py:226-236: integer gather."""
        table = torch.tensor(self._kf_id_to_submap_id, device=kf_ids.device)
        return table[kf_ids]

    def updated_kf_pose_in_submap(self, kf_id: int, submap_id: int):
        """grid_atlas.py:286-300."""
        expect_submap_id = self.submap_id_for_kf(kf_id)
        assert expect_submap_id == submap_id, f"Wrong submap for KF {kf_id}! Expect {expect_submap_id}, got {submap_id}."
        kf_id_submap = kf_id - self.anchor_kf_for_submap(submap_id)
        return self.get_submap(submap_id).updated_kf_pose(kf_id_submap)

    def initial_submap_pose(self, submap_id: int):
        return self.R_world_submap_list[submap_id], self.t_world_submap_list[submap_id]

    def updated_submap_pose(self, submap_id: int, device=None):
        """grid_atlas.py:250-268."""
        R, t = self.initial_submap_pose(submap_id)
        R, t = utils_geometry.apply_pose_correction(R=R, t=t, R_delta=self.rotation_corrections[submap_id],
                                                    t_delta=self.translation_corrections[submap_id])
        if device is not None:
            R, t = R.to(device), t.to(device)
        return R, t

    def params_for_submap_pose(self, submap_id: int):
        return [self.rotation_corrections[submap_id], self.translation_corrections[submap_id]]

    def get_submap(self, submap_id: int) -> GridNet:
        assert submap_id >= 0 and submap_id < self.num_submaps
        return self.submaps[submap_id]

    def _fused_query_ok(self, x_world):
        if not x_world.is_cuda or len(self.active_submaps) == 0:
            return False
        if torch.is_grad_enabled() and (x_world.requires_grad or any(
                f.requires_grad for i in self.active_submaps for f in self.get_submap(i).level_tensors())):
            return False
        subs = [self.get_submap(i) for i in range(self.num_submaps)]
        return all(sm.fdim == 4 and sm.num_levels == subs[0].num_levels and sm.num_levels <= 4 and
                   sm.features[0].feature.is_cuda for sm in subs)

    def query_feature(self, x_world: torch.Tensor):
        """grid_atlas.py:374-391: in-bound-masked mean of the submaps' features at world points.  Without autograd
        (fusion / meshing queries) every submap is visited inside ONE launch (miso_atlas_features)."""
        if self._fused_query_ok(x_world):
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            dev = x_world.device
            key = tuple((f.data_ptr(), tuple(f.stride())) for i in range(self.num_submaps)
                        for f in self.get_submap(i).level_tensors()) + tuple(
                tuple(self.get_submap(i).ignore_level_) for i in range(self.num_submaps))
            cache = getattr(self, "_atlas_fields_cache", None)
            if cache is None or cache[0] != key:
                fields = []
                for i in range(self.num_submaps):
                    sm = self.get_submap(i)
                    mask = sum((1 << l) for l in range(sm.num_levels) if sm.ignore_level_[l])
                    fields.append(_field.make_field(sm.level_tensors(), sm._bound_host, None, mask))
                cache = (key, _field.structs_to_device(fields, dev))
                self._atlas_fields_cache = cache
            with torch.no_grad():
                poses = []
                for i in range(self.num_submaps):
                    R, t = self.updated_submap_pose(i)
                    Rt_ = R.T                                        # transfrom_points_from (utils_geometry.py:227-240)
                    poses.append(torch.cat([Rt_.reshape(9), (-Rt_ @ t).reshape(3)]))
                poses = torch.stack(poses, 0).contiguous().float()
            active = torch.tensor(list(self.active_submaps), dtype=torch.int32, device=dev)
            x = _field._prep_x(x_world)
            L = self.get_submap(0).num_levels
            out = torch.empty((x.shape[0], 4 * L), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.miso_atlas_features(cache[1].data_ptr(), self.num_submaps, active.data_ptr(),
                                                   int(active.numel()), poses.data_ptr(), x.data_ptr(), x.shape[0], L,
                                                   out.data_ptr(), _lib.stream_ptr(dev)), "atlas_features")
            return out
        sum_feats = 0
        sum_weights = 0
        for submap_id in self.active_submaps:
            submap = self.get_submap(submap_id)
            R_world_submap, t_world_submap = self.updated_submap_pose(submap_id)
            x_submap = utils_geometry.transfrom_points_from(x_world, R_world_submap, t_world_submap)
            mask_bnd = utils_geometry.coords_in_bound(x_submap, submap.bound)
            submap_feats = submap.query_feature(x_submap)
            sum_feats = sum_feats + mask_bnd * submap_feats
            sum_weights = sum_weights + mask_bnd
        sum_weights = torch.where(sum_weights == 0, torch.ones_like(sum_weights), sum_weights).float()
        return sum_feats / sum_weights

    def forward(self, x_world: torch.Tensor, noise_std=0):
        """grid_atlas.py:393-399: decode the mean feature with submap 0's decoder."""
        mean_feats = self.query_feature(x_world)
        pred = grid_decode(mean_feats, None, self.submaps[0].decoder, True)
        if noise_std > 0:
            pred = pred + torch.randn(pred.shape, device=x_world.device) * noise_std
        return pred

    def check_submap_intersection(self, src_id: int, dst_id: int, overlap_thresh=1e-2):
        """grid_atlas.py:405-420 (torch ops; the batched kernel version lives in miso_b200.align)."""
        submap_src = self.get_submap(src_id)
        submap_dst = self.get_submap(dst_id)
        corners_src = submap_src.features[-1].vertex_positions().to(self.device)
        R_world_src, t_world_src = self.updated_submap_pose(src_id)
        R_world_dst, t_world_dst = self.updated_submap_pose(dst_id)
        corners_world = utils_geometry.transform_points_to(corners_src, R_world_src, t_world_src)
        corners_dst = utils_geometry.transfrom_points_from(corners_world, R_world_dst, t_world_dst)
        mask_bnd = utils_geometry.coords_in_bound(corners_dst, submap_dst.bound)
        num_valid = torch.count_nonzero(mask_bnd)
        return num_valid / corners_src.shape[0] > overlap_thresh

    def precompute_coordinates_for_alignment(self, norm_thresh=1e-5):
        """grid_atlas.py:565-579: voxel centres of every level whose feature norm exceeds the threshold,
        in the reference's flatten order (Z,Y,X with X fastest)."""
        self._coords_for_alignment = dict()
        for level in range(self.num_levels):
            for submap_id in range(self.num_submaps):
                submap = self.get_submap(submap_id)
                coords = submap.features[level].vertex_positions().to(self.device)
                with torch.no_grad():
                    feature = submap.query_feature(coords)
                featnorm_from = torch.linalg.norm(feature, dim=1, keepdim=True).detach()
                mask_feat = featnorm_from > norm_thresh
                valid_indices = torch.nonzero(mask_feat, as_tuple=False)[:, 0]
                self._coords_for_alignment[f"submap{submap_id}_level{level}"] = coords[valid_indices, :].detach()

    def coordinates_for_alignment(self, submap_id: int, level: int):
        assert submap_id >= 0 and submap_id < self.num_submaps
        assert level >= 0 and level < self.num_levels
        key = f"submap{submap_id}_level{level}"
        if key not in self._coords_for_alignment:
            raise ValueError(f"Coordinates for alignment not found for submap {submap_id} and level {level}. "
                             "Did you call precompute_coordinates_for_alignment()?")
        return self._coords_for_alignment[key]
