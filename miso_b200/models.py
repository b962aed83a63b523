"""Parameter holders of the hot path: dense feature levels, the ReLU decoder, keyframe / submap pose tables.

The classes answer to the reference's names and keep its constructor arguments, method names and state-dict keys
(`features.<l>.feature`, `feature_stability.<l>.feature`, `decoder.network.<i>.{weight,bias}`,
`rotation_corrections`, `translation_corrections`, `Rwk`, `twk`) so that checkpoints and call sites of

    grid_opt/models/grid_modules.py  FeatureGrid   :41-123
    grid_opt/models/modules.py       MLPNet        :11-40
    grid_opt/models/grid_net.py      GridNet       :17-352
    grid_opt/models/grid_atlas.py    GridAtlas     :18-587 (construction, poses, alignment samples, queries)

carry over; the bodies are written for this package: levels live channels_last_3d, every query goes to the CUDA
library (fused kernels when the decoder is frozen, the per-level plugin otherwise), pose tables are evaluated in
batch with a per-keyframe trainable mask instead of a Python loop, and the keyframe <-> submap maps are integer
tables.  To use the kernels from the reference's own classes instead, see `miso_b200.adapter` / INTEGRATION.md.
"""
import copy
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import cuda_gridsample as cu
from . import field as _field
from . import geometry as geo

FUSED_LEVEL_CHANNEL_SHAPES = {(1, 4), (2, 4), (3, 4), (4, 4), (1, 8), (2, 8), (1, 16)}


# ------------------------------------------------------------------------------------------------
# coordinate helpers (grid_opt/utils/utils.py:22-80, :294-307)
# ------------------------------------------------------------------------------------------------
def _lo_hi(bounds: torch.Tensor, like: torch.Tensor):
    """bounds (d,2) -> broadcastable (lo, hi) rows for a (..., d) coordinate tensor."""
    if like.dim() not in (2, 3):
        raise ValueError("coordinates must be (N,d) or (B,N,d)")
    shape = (1,) * (like.dim() - 1) + (-1,)
    return bounds[:, 0].reshape(shape), bounds[:, 1].reshape(shape)


def normalize_coordinates(queries: torch.Tensor, bounds: torch.Tensor) -> torch.Tensor:
    """World -> [-1,1]^d.  The operation order `2 (x - lo) / (hi - lo) - 1` is the reference's (utils.py:49) and the
    kernels' (`normalize_coord`, csrc/common.cuh): floor() decisions depend on it."""
    lo, hi = _lo_hi(bounds, queries)
    return 2 * (queries - lo) / (hi - lo) - 1


def denormalize_coordinates(unit: torch.Tensor, bounds: torch.Tensor) -> torch.Tensor:
    """[-1,1]^d -> world, `(u + 1) / 2 (hi - lo) + lo` (utils.py:53-80)."""
    lo, hi = _lo_hi(bounds, unit)
    return (unit + 1) / 2 * (hi - lo) + lo


def _axis_centres(n: int) -> torch.Tensor:
    """Normalised centres of n voxels along one axis, with the rounding sequence of utils.py:294-307
    (linspace over [1/2n, 1 - 1/2n], then 2u - 1) so alignment samples are bit-identical to the reference's."""
    half = 0.5 / n
    return 2 * torch.linspace(half, 1 - half, n) - 1.0


def all_grid_positions(features: torch.Tensor) -> torch.Tensor:
    """(1,Z,Y,X,3) normalised voxel centres of a (1,C,Z,Y,X) level, last dim (x,y,z)."""
    Z, Y, X = features.shape[2:]
    out = torch.empty((1, Z, Y, X, 3))
    out[..., 0] = _axis_centres(X).reshape(1, 1, 1, X)
    out[..., 1] = _axis_centres(Y).reshape(1, 1, Y, 1)
    out[..., 2] = _axis_centres(Z).reshape(1, Z, 1, 1)
    return out


def level_dims(bound, cell_size: float) -> Tuple[int, int, int]:
    """Voxels per axis (X,Y,Z) = ceil(extent / cell) in float32, as the reference sizes a level (grid_modules.py:47-53)."""
    b = np.asarray(bound.detach().cpu() if isinstance(bound, torch.Tensor) else bound, dtype=np.float32)
    n = np.ceil((b[:, 1] - b[:, 0]) / cell_size).astype(int)
    return int(n[0]), int(n[1]), int(n[2])


def grid_interp_regular(levels: Sequence["FeatureGrid"], x: torch.Tensor, ignore_level=None) -> torch.Tensor:
    """Per-level interpolation + concat; an ignored level keeps its slot filled with zeros (utils.py:143-164)."""
    cols = []
    for l, lvl in enumerate(levels):
        f = lvl.interpolate(x)
        cols.append(f * 0 if (ignore_level is not None and ignore_level[l]) else f)
    return torch.cat(cols, dim=1)


def grid_decode(feats: torch.Tensor, x: Optional[torch.Tensor], decoder=None, pos_invariant=True) -> torch.Tensor:
    """utils.py:194-208: features (optionally with the position appended) through the decoder, or unchanged."""
    if feats.ndim != 2:
        raise AssertionError(f"features must be (N,F), got {tuple(feats.shape)}")
    if decoder is None:
        return feats
    return decoder(feats if pos_invariant else torch.cat((feats, x), dim=1))


def _set_trainable(module: nn.Module, flag: bool):
    for p in module.parameters():
        p.requires_grad_(flag)


# ------------------------------------------------------------------------------------------------
# one dense level
# ------------------------------------------------------------------------------------------------
class FeatureGrid(nn.Module):
    """Dense feature level, logical shape (1,C,Z,Y,X), stored channels_last_3d (one voxel = C contiguous floats)."""

    def __init__(self, d, fdim, bound, cell_size, name="grid", dtype=torch.float32, initial_feature=None,
                 init_stddev=0.0, second_order_grid_sample=False):
        super().__init__()
        if d != 3:
            raise NotImplementedError("miso_b200 implements spatial_dim 3 only (every shipped MISO config)")
        self.d, self.fdim, self.cell_size, self.dtype, self.name = d, fdim, cell_size, dtype, name
        self.bound = bound if isinstance(bound, torch.Tensor) else torch.as_tensor(np.asarray(bound), dtype=dtype)
        if tuple(self.bound.shape) != (3, 2):
            raise AssertionError(f"bound must be (3,2), got {tuple(self.bound.shape)}")
        X, Y, Z = level_dims(self.bound, cell_size)
        shape = (1, fdim, Z, Y, X)
        if initial_feature is None:
            initial_feature = torch.randn(shape, dtype=dtype) * init_stddev
        elif tuple(initial_feature.shape) != shape:
            raise AssertionError(f"initial feature {tuple(initial_feature.shape)} does not fit level shape {shape}")
        self.feature = nn.Parameter(initial_feature.contiguous(memory_format=torch.channels_last_3d))
        # the reference switches between F.grid_sample and its double-backward plugin here (grid_modules.py:63-69);
        # the CUDA op below is both, the flag is only remembered
        self.second_order_grid_sample = second_order_grid_sample
        self.grid_sample_func = cu.grid_sample_3d
        self._bound_host = _field.bound_to_list(self.bound)

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        _field.to_channels_last_3d_(self.feature)    # .to()/.cuda() re-allocate: restore the layout
        self.bound = fn(self.bound)
        return self

    def interpolate(self, x: torch.Tensor) -> torch.Tensor:
        """(N,3) world points -> (N,C) trilinear features, zeros outside, align_corners=False (grid_modules.py:72-95)."""
        unit = normalize_coordinates(x, self.bound).reshape(1, -1, 1, 1, 3)
        sampled = self.grid_sample_func(self.feature, unit, padding_mode="zeros", align_corners=False)
        return sampled[0, :, :, 0, 0].t()

    def norm(self):
        return self.feature.norm()

    def num_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def lock(self):
        _set_trainable(self, False)

    def unlock(self):
        _set_trainable(self, True)

    @torch.no_grad()
    def zero_features(self):
        self.feature.zero_()

    @torch.no_grad()
    def randn_features(self, std: float):
        self.feature.copy_((torch.randn(self.feature.shape, dtype=self.dtype) * std).to(self.feature.device))

    def vertex_positions(self, denormalize=True) -> torch.Tensor:
        """(Z*Y*X, 3) voxel centres, x fastest (grid_modules.py:111-123); evaluated on the host like the reference."""
        unit = all_grid_positions(self.feature).reshape(-1, 3)
        return denormalize_coordinates(unit, self.bound.to(unit)) if denormalize else unit


class MLPNet(nn.Module):
    """Linear -> act -> (Linear -> act) x hidden_layers -> Linear, exposed as `.network` (modules.py:11-40)."""

    def __init__(self, input_dim, output_dim, hidden_dim=64, hidden_layers=1, bias=False, acti_func=nn.ReLU,
                 pretrained_path=None, no_optimize=False):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        widths = [input_dim] + [hidden_dim] * (hidden_layers + 1)
        stack: List[nn.Module] = []
        for w_in, w_out in zip(widths[:-1], widths[1:]):
            stack += [nn.Linear(w_in, w_out, bias=bias), acti_func()]
        stack.append(nn.Linear(hidden_dim, output_dim, bias=bias))
        self.layers = stack
        self.network = nn.Sequential(*stack)
        if pretrained_path is not None:
            self.load(pretrained_path)
        if no_optimize:
            _set_trainable(self, False)

    def forward(self, x):
        return self.network(x)

    def save(self, filepath):
        torch.save(self.state_dict(), filepath)

    def load(self, filepath):
        self.load_state_dict(torch.load(filepath))


class BaseNet(nn.Module):
    """cfg + spatial bound shared by GridNet / GridAtlas (models/base_net.py:11-40)."""

    def __init__(self, cfg: dict, device="cpu", dtype=torch.float32):
        super().__init__()
        self.cfg, self.device, self.dtype = cfg, device, dtype
        self.d = cfg["spatial_dim"]
        self.bound = torch.tensor(np.asarray(cfg["grid"]["bound"]), device=device, dtype=dtype)
        if self.d != 3 or tuple(self.bound.shape) != (3, 2):
            raise AssertionError("miso_b200 models are 3-D with a (3,2) bound")


# ------------------------------------------------------------------------------------------------
# one submap
# ------------------------------------------------------------------------------------------------
class GridNet(BaseNet):
    """One submap: L levels (+ their 1-channel stability levels), the decoder, K keyframe poses with corrections."""

    def __init__(self, cfg: dict, device="cuda:0", dtype=torch.float32, initial_features=dict()):
        super().__init__(cfg, device, dtype)
        self.initial_features = initial_features
        self._build_levels(cfg["grid"])
        self._build_decoder(cfg["decoder"])
        self._build_pose_table(cfg["pose"])
        self.to(device)
        self._bound_host = _field.bound_to_list(cfg["grid"]["bound"])
        self._decoder_spec = None

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.bound = fn(self.bound)
        self._decoder_spec = None
        return self

    # ---- construction -----------------------------------------------------------------------------
    def _build_levels(self, g: dict):
        if g["type"] != "regular":
            raise NotImplementedError("grid.type 'regular' only (the shipped configs); VM grids are out of scope")
        self.grid_type = g["type"]
        self.num_levels, self.fdim = g["n_levels"], g["feature_dim"]
        self.second_order_grid_sample = bool(g.get("second_order_grid_sample", False))
        self.cell_sizes = [g["base_cell_size"] / g["per_level_scale"] ** l for l in range(self.num_levels)]
        host_bound = self.bound.detach().cpu()

        def level(l, channels, tag, init, std):
            return FeatureGrid(d=3, fdim=channels, bound=host_bound, cell_size=self.cell_sizes[l], name=f"{tag}-{l}",
                               dtype=self.dtype, initial_feature=init, init_stddev=std,
                               second_order_grid_sample=self.second_order_grid_sample)

        self.features = nn.ModuleList(level(l, self.fdim, "feat", self.initial_features.get(l), g["init_stddev"])
                                      for l in range(self.num_levels))
        self.feature_stability = nn.ModuleList(level(l, 1, "stab", None, 0.0) for l in range(self.num_levels))
        self.ignore_level_ = np.zeros(self.num_levels, dtype=bool)

    def _build_decoder(self, d: dict):
        self.decoder_hidden_dim, self.decoder_hidden_layers = d["hidden_dim"], d["hidden_layers"]
        self.decoder_out_dim, self.pos_invariant = d["out_dim"], d["pos_invariant"]
        self.decoder_fixed, self.decoder_type = d["fix"], d["type"]
        if self.decoder_type == "none":
            self.decoder = None
        elif self.decoder_type == "mlp":
            width_in = self.num_levels * self.fdim + (0 if self.pos_invariant else 3)
            self.decoder = MLPNet(width_in, self.decoder_out_dim, hidden_dim=self.decoder_hidden_dim,
                                  hidden_layers=self.decoder_hidden_layers, bias=True,
                                  pretrained_path=d["pretrained_model"], no_optimize=self.decoder_fixed)
        else:
            raise ValueError(f"Unknown decoder type: {self.decoder_type}")

    def _build_pose_table(self, p: dict):
        K = self.num_poses = p["num_poses"]
        self.optimize_pose = p["optimize"]
        self.rotation_corrections = nn.Parameter(torch.zeros(K, 3), requires_grad=self.optimize_pose)
        self.translation_corrections = nn.Parameter(torch.zeros(K, 3, 1), requires_grad=self.optimize_pose)
        self.register_buffer("Rwk", geo.identity_rotations(K))
        self.register_buffer("twk", torch.zeros(K, 3, 1))
        self.pose_estimates_known = [False] * K
        self.locked_pose_indices = set()
        self._pose_key_to_id: Dict[str, int] = {}

    def save(self, ckpt_dir, ckpt_prefix):
        self.decoder.save(os.path.join(ckpt_dir, f"{ckpt_prefix}_decoder.pt"))

    # ---- levels -------------------------------------------------------------------------------------
    def ignore_level(self, l):
        self.ignore_level_[int(l)] = True     # the level keeps its slot in the feature vector, filled with zeros

    def include_level(self, l):
        self.ignore_level_[int(l)] = False

    def _ignore_mask(self) -> int:
        return sum(1 << l for l in range(self.num_levels) if self.ignore_level_[l])

    def _level_trainable(self, levels, flag: bool):
        for l in levels:
            _set_trainable(self.features[l], flag)
            _set_trainable(self.feature_stability[l], flag)

    def lock_level(self, l):
        self._level_trainable((l,), False)

    def unlock_level(self, l):
        self._level_trainable((l,), True)

    def lock_feature(self):
        self._level_trainable(range(self.num_levels), False)

    def unlock_feature(self):
        self._level_trainable(range(self.num_levels), True)

    def zero_features(self):
        for lvl in self.features:
            lvl.zero_features()

    def randn_features(self, std):
        for lvl in self.features:
            lvl.randn_features(std)

    def level_tensors(self) -> List[torch.Tensor]:
        return [lvl.feature for lvl in self.features]

    # ---- keyframe poses (grid_net.py:159-262) ---------------------------------------------------------
    def _poses_trainable(self, flag: bool):
        for p in self.params_for_poses():
            p.requires_grad_(flag)
        self.locked_pose_indices = set() if flag else set(range(self.num_poses))

    def lock_pose(self):
        self._poses_trainable(False)

    def unlock_pose(self):
        self._poses_trainable(True)

    def lock_pose_index(self, pose_index: int):
        self.locked_pose_indices |= {int(pose_index)}

    def unlock_pose_index(self, pose_index: int):
        self.locked_pose_indices.remove(int(pose_index))    # KeyError when it was not locked, like the reference

    def lock_all_pose_indices(self):
        self.locked_pose_indices = set(range(self.num_poses))

    def unlock_all_pose_indices(self):
        self.locked_pose_indices = set()

    def pose_correction(self, kf_id: int):
        """(1,3) rotation and (3,1) translation correction of one keyframe; a locked keyframe's are detached."""
        dr, dt = self.rotation_corrections[kf_id:kf_id + 1], self.translation_corrections[kf_id]
        return (dr.detach().clone(), dt.detach().clone()) if kf_id in self.locked_pose_indices else (dr, dt)

    def set_initial_kf_pose(self, kf_id: int, Rwk: torch.Tensor, twk: torch.Tensor, kf_key=None):
        if tuple(Rwk.shape) != (3, 3) or tuple(twk.shape) != (3, 1):
            raise AssertionError("keyframe pose must be R (3,3), t (3,1)")
        assert kf_id < self.num_poses, f"KF ID {kf_id} exceeds the number of poses {self.num_poses}!"
        with torch.no_grad():
            self.Rwk[kf_id] = Rwk.to(self.Rwk)
            self.twk[kf_id] = twk.to(self.twk)
            self.rotation_corrections[kf_id].zero_()
            self.translation_corrections[kf_id].zero_()
        self.pose_estimates_known[kf_id] = True
        if kf_key is not None:
            self._pose_key_to_id[str(kf_key)] = int(kf_id)

    def pose_key_to_id(self, kf_key):
        assert kf_key in self._pose_key_to_id, f"Key {kf_key} not found in pose key to ID mapping!"
        return self._pose_key_to_id[kf_key]

    def initial_kf_pose(self, kf_id: int):
        assert self.pose_estimates_known[kf_id], f"Initial pose estimate for KF {kf_id} is not available!"
        return self.Rwk[kf_id], self.twk[kf_id]

    initial_kf_pose_in_world = initial_kf_pose

    def updated_kf_pose(self, kf_id: int):
        R0, t0 = self.initial_kf_pose(kf_id)
        return geo.apply_pose_correction(R0, t0, *self.pose_correction(kf_id))

    updated_kf_pose_in_world = updated_kf_pose

    def updated_kf_pose_from_key(self, kf_key):
        return self.updated_kf_pose(self.pose_key_to_id(kf_key))

    def all_kf_poses(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """`updated_kf_pose` of every keyframe at once: (K,3,3), (K,3,1).  Rows listed in `locked_pose_indices` use
        detached corrections, exactly like `pose_correction` -- so a batch that mixes samples of locked and unlocked
        keyframes (the reference's track_window: unlock_pose, lock_all_pose_indices, unlock_pose_index) sends no
        gradient into the locked ones."""
        dr, dt = self.rotation_corrections, self.translation_corrections
        if self.locked_pose_indices and (dr.requires_grad or dt.requires_grad):
            frozen = torch.zeros(self.num_poses, dtype=torch.bool, device=dr.device)
            frozen[sorted(self.locked_pose_indices)] = True
            dr = torch.where(frozen[:, None], dr.detach(), dr)
            dt = torch.where(frozen[:, None, None], dt.detach(), dt)
        return torch.matmul(self.Rwk, geo.so3_exp_map(dr)), self.twk + dt

    # ---- queries: the hot path ------------------------------------------------------------------------
    def fused_spec(self, trainable_decoder_ok: bool = False) -> Optional[_field.FieldSpec]:
        """Descriptor for the fused kernels, or None when they do not apply (trainable / non-64-wide / positional
        decoder, unsupported (levels, channels), not on CUDA).  The cheap predicate is re-evaluated on every call
        -- un-freezing the decoder or swapping it takes effect immediately -- and only the packed decoder weights
        are cached, keyed on the identity and storage of the weight tensors.  `trainable_decoder_ok`: the mapping
        step also covers a trainable decoder (its parameter gradients come from miso_mapping_step_wgrad); the
        forward-only users keep the autograd route for it."""
        dec = self.decoder
        if (dec is None or self.decoder_type != "mlp" or not self.pos_invariant or self.decoder_hidden_dim != 64
                or self.decoder_hidden_layers != 1 or self.decoder_out_dim != 1
                or (self.num_levels, self.fdim) not in FUSED_LEVEL_CHANNEL_SHAPES
                or not self.features[0].feature.is_cuda
                or (not trainable_decoder_ok and any(p.requires_grad for p in dec.parameters()))):
            return None
        key = tuple((id(p), p.data_ptr(), p._version) for p in dec.parameters())
        if self._decoder_spec is None or self._decoder_spec[0] != key:
            self._decoder_spec = (key, _field.DecoderSpec.from_mlp(dec))
        return _field.FieldSpec(self._bound_host, self._decoder_spec[1], self._ignore_mask())

    def _wants_graph(self, x: torch.Tensor) -> bool:
        return torch.is_grad_enabled() and (x.requires_grad or any(f.requires_grad for f in self.level_tensors()))

    def _check_points(self, x: torch.Tensor):
        assert x.ndim == 2 and x.shape[-1] == 3, f"Invalid input coords shape {tuple(x.shape)}!"

    def query_feature(self, x: torch.Tensor) -> torch.Tensor:
        """(N,3) -> (N, L*C) concatenated level features (grid_net.py:288-297): one fused launch when no autograd
        graph is needed, the twice-differentiable per-level op otherwise."""
        self._check_points(x)
        if x.is_cuda and self.fdim in (4, 8, 12, 16) and not self._wants_graph(x):
            return _field.field_features_raw(self.level_tensors(), self._bound_host, x, self._ignore_mask())
        return grid_interp_regular(self.features, x, self.ignore_level_)

    def query_stability(self, x: torch.Tensor) -> torch.Tensor:
        self._check_points(x)
        return grid_interp_regular(self.feature_stability, x)

    def forward(self, x: torch.Tensor, noise_std=0) -> torch.Tensor:
        """(N,3) -> (N,out) decoded prediction (grid_net.py:306-325)."""
        spec = self.fused_spec()
        if spec is None:
            out = grid_decode(grid_interp_regular(self.features, x, self.ignore_level_), x, self.decoder,
                              self.pos_invariant)
        elif self._wants_graph(x):
            out, _ = _field.fused_sdf(x, self.level_tensors(), spec)
        else:   # inference / dense queries: values only, no Jacobian pass
            out = _field.sdf_forward_raw(self.level_tensors(), spec, x, want_jac=False, want_gradx=False)[0].unsqueeze(1)
        return out + torch.randn_like(out) * noise_std if noise_std > 0 else out

    def forward_with_gradient(self, x: torch.Tensor):
        """(sdf (N,1), grad_x sdf (N,3)) from ONE fused launch; both are differentiable w.r.t. the levels (the
        gradient output carries the eikonal double-backward).  Falls back to autograd.grad over `forward`."""
        spec = self.fused_spec()
        if spec is not None:
            return _field.fused_sdf(x, self.level_tensors(), spec)
        xr = x if x.requires_grad else x.clone().requires_grad_(True)
        y = self.forward(xr)
        return y, torch.autograd.grad(y, xr, grad_outputs=torch.ones_like(y), create_graph=True)[0]

    # ---- parameter groups (grid_net.py:327-351) ---------------------------------------------------------
    def params_for_poses(self):
        return [self.rotation_corrections, self.translation_corrections]

    def params_for_features(self, stop_level=None):
        stop = self.num_levels if stop_level is None else stop_level
        assert stop <= self.num_levels
        return [p for l in range(stop) for p in self.features[l].parameters()]

    def params_at_level(self, level):
        """Level `level`'s feature + stability tensors (every level when level == num_levels: the joint stage), plus
        the decoder when trainable and the pose corrections when optimised."""
        chosen = range(self.num_levels) if level >= self.num_levels else (level,)
        out = [p for l in chosen for grid in (self.features[l], self.feature_stability[l]) for p in grid.parameters()]
        if not self.decoder_fixed:
            out += list(self.decoder.parameters())
        if self.optimize_pose:
            out += self.params_for_poses()
        return out


# ------------------------------------------------------------------------------------------------
# many submaps
# ------------------------------------------------------------------------------------------------
class GridAtlas(BaseNet):
    """Submaps with a world pose each (+ se(3) correction parameters) and the integer keyframe <-> submap tables."""

    def __init__(self, cfg: dict, device="cuda:0", dtype=torch.float32):
        super().__init__(cfg, device, dtype)
        self.submaps = nn.ModuleList()
        self.rotation_corrections = nn.ParameterList()
        self.translation_corrections = nn.ParameterList()
        self.R_world_submap_list: List[torch.Tensor] = []
        self.t_world_submap_list: List[torch.Tensor] = []
        self._submap_anchor_kf: List[int] = []
        self._kf_id_to_submap_id: List[int] = []
        self._submap_id_to_kf_ids: Dict[int, set] = {}
        self.curr_submap_id = self.curr_kf_id = -1
        self.num_levels = cfg["grid"]["n_levels"]
        self.active_submaps = range(0)
        self._coords_for_alignment: Dict[str, torch.Tensor] = {}

    # ---- bookkeeping ------------------------------------------------------------------------------------
    @property
    def num_submaps(self) -> int:
        return len(self.submaps)

    @property
    def num_keyframes(self) -> int:
        return self.curr_kf_id + 1

    def get_submap(self, submap_id: int) -> GridNet:
        assert 0 <= submap_id < self.num_submaps
        return self.submaps[submap_id]

    def anchor_kf_for_submap(self, submap_id: int) -> int:
        return self._submap_anchor_kf[submap_id]

    def submap_id_for_kf(self, kf_id: int) -> int:
        return self._kf_id_to_submap_id[kf_id]

    def submap_id_for_kf_batch(self, kf_ids: torch.Tensor) -> torch.Tensor:
        """Integer gather through the keyframe -> submap table (grid_atlas.py:226-236)."""
        return torch.as_tensor(self._kf_id_to_submap_id, device=kf_ids.device)[kf_ids]

    def _register_submap(self, net: GridNet, Rws, tws):
        if tuple(Rws.shape) != (3, 3) or tuple(tws.shape) != (3, 1):
            raise AssertionError("submap pose must be R (3,3), t (3,1)")
        sid = self.num_submaps
        self.submaps.append(net)
        dev = self.device
        self.R_world_submap_list += [Rws.to(dev)]
        self.t_world_submap_list += [tws.to(dev)]
        self.rotation_corrections.append(nn.Parameter(torch.zeros(1, 3, device=dev)))
        self.translation_corrections.append(nn.Parameter(torch.zeros(3, 1, device=dev)))
        first_kf = self.curr_kf_id + 1
        self._submap_anchor_kf.append(first_kf)
        self._submap_id_to_kf_ids[sid] = set()
        self.active_submaps = range(self.num_submaps)
        self.curr_submap_id = sid
        return sid, first_kf

    def add_submap(self, local_bound: torch.Tensor, Rws: torch.Tensor, tws: torch.Tensor, num_poses=1,
                   optimize_poses=True):
        """New empty submap over `local_bound` posed at (Rws, tws) (grid_atlas.py:120-169)."""
        sub_cfg = copy.deepcopy(self.cfg)
        sub_cfg["grid"]["bound"] = np.asarray(local_bound)
        sub_cfg["pose"].update(num_poses=num_poses, optimize=optimize_poses)
        sid, first_kf = self._register_submap(GridNet(cfg=sub_cfg, device=self.device, dtype=self.dtype), Rws, tws)
        self._submap_id_to_kf_ids[sid].add(first_kf)

    def add_existing_submap(self, submap: GridNet, Rws: torch.Tensor, tws: torch.Tensor):
        """Attach a GridNet built elsewhere (multi-GPU build_submaps gathers them to rank 0 in id order)."""
        self._register_submap(submap, Rws, tws)

    def add_kf(self, Rsk: torch.Tensor, tsk: torch.Tensor) -> int:
        """Next global keyframe, posed (Rsk, tsk) inside the CURRENT submap (grid_atlas.py:96-118)."""
        assert self.curr_submap_id >= 0, "No submap is created yet. Create a submap first."
        sid, kf = self.curr_submap_id, self.curr_kf_id + 1
        self.get_submap(sid).set_initial_kf_pose(kf - self._submap_anchor_kf[sid], Rsk, tsk, kf_key=f"KF{kf}")
        self._kf_id_to_submap_id.append(sid)
        self._submap_id_to_kf_ids[sid].add(kf)
        self.curr_kf_id = kf
        return kf

    # ---- submap poses -----------------------------------------------------------------------------------
    @torch.no_grad()
    def set_submap_pose(self, submap_id: int, Rws: torch.Tensor, tws: torch.Tensor):
        """Overwrite the initial pose and reset the correction to zero (grid_atlas.py:171-187)."""
        if tuple(Rws.shape) != (3, 3) or tuple(tws.shape) != (3, 1):
            raise AssertionError("submap pose must be R (3,3), t (3,1)")
        R0, t0 = self.initial_submap_pose(submap_id)
        R0.copy_(Rws.to(R0))
        t0.copy_(tws.to(t0))
        for corr in self.params_for_submap_pose(submap_id):
            corr.zero_()

    @torch.no_grad()
    def set_submap_pose_correction(self, submap_id: int, R_delta: torch.Tensor, t_delta: torch.Tensor):
        if tuple(R_delta.shape) != (1, 3) or tuple(t_delta.shape) != (3, 1):
            raise AssertionError("correction must be w (1,3), tau (3,1)")
        w, tau = self.params_for_submap_pose(submap_id)
        w.copy_(R_delta)
        tau.copy_(t_delta)

    def initial_submap_pose(self, submap_id: int):
        return self.R_world_submap_list[submap_id], self.t_world_submap_list[submap_id]

    def updated_submap_pose(self, submap_id: int, device=None):
        """(R0 Exp(w), t0 + tau) of one submap (grid_atlas.py:250-268); stays in the autograd graph."""
        R, t = geo.apply_pose_correction(*self.initial_submap_pose(submap_id), self.rotation_corrections[submap_id],
                                         self.translation_corrections[submap_id])
        return (R, t) if device is None else (R.to(device), t.to(device))

    def params_for_submap_pose(self, submap_id: int):
        return [self.rotation_corrections[submap_id], self.translation_corrections[submap_id]]

    def updated_kf_pose_in_submap(self, kf_id: int, submap_id: int):
        owner = self.submap_id_for_kf(kf_id)
        assert owner == submap_id, f"Wrong submap for KF {kf_id}! Expect {owner}, got {submap_id}."
        return self.get_submap(submap_id).updated_kf_pose(kf_id - self._submap_anchor_kf[submap_id])

    # ---- queries ------------------------------------------------------------------------------------------
    def _world_to_submap_rows(self) -> torch.Tensor:
        """(S,12) rows [R^T row-major | -R^T t]: the world -> submap transform of every submap."""
        rows = []
        for sid in range(self.num_submaps):
            R, t = self.updated_submap_pose(sid)
            rows.append(torch.cat([R.T.reshape(9), (-R.T @ t).reshape(3)]))
        return torch.stack(rows).float().contiguous()

    def _one_launch_query_ok(self, x_world: torch.Tensor) -> bool:
        subs = list(self.submaps)
        if not x_world.is_cuda or len(self.active_submaps) == 0:
            return False
        if torch.is_grad_enabled() and (x_world.requires_grad or any(
                f.requires_grad for sid in self.active_submaps for f in subs[sid].level_tensors())):
            return False
        L = subs[0].num_levels
        return L <= 4 and all(s.fdim == 4 and s.num_levels == L and s.features[0].feature.is_cuda for s in subs)

    def query_feature(self, x_world: torch.Tensor) -> torch.Tensor:
        """Mean over the active submaps whose bound contains the point of that submap's features
        (grid_atlas.py:374-391).  Without autograd all submaps are visited inside one launch."""
        if self._one_launch_query_ok(x_world):
            from . import _lib
            lib, dev, subs = _lib.load(), x_world.device, list(self.submaps)
            key = tuple((f.data_ptr(), f.stride()) for s in subs for f in s.level_tensors()) + tuple(
                s._ignore_mask() for s in subs)
            cached = getattr(self, "_atlas_fields_cache", None)
            if cached is None or cached[0] != key:
                descr = [_field.make_field(s.level_tensors(), s._bound_host, None, s._ignore_mask()) for s in subs]
                cached = self._atlas_fields_cache = (key, _field.structs_to_device(descr, dev))
            with torch.no_grad():
                rows = self._world_to_submap_rows()
            active = torch.tensor(list(self.active_submaps), dtype=torch.int32, device=dev)
            x = _field._prep_x(x_world)
            L = subs[0].num_levels
            out = torch.empty((x.shape[0], 4 * L), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.miso_atlas_features(cached[1].data_ptr(), len(subs), active.data_ptr(), active.numel(),
                                                   rows.data_ptr(), x.data_ptr(), x.shape[0], L, out.data_ptr(),
                                                   _lib.stream_ptr(dev)), "atlas_features")
            return out
        total, hits = 0, 0
        for sid in self.active_submaps:
            sub = self.get_submap(sid)
            x_local = geo.transfrom_points_from(x_world, *self.updated_submap_pose(sid))
            inside = geo.coords_in_bound(x_local, sub.bound)
            total = total + inside * sub.query_feature(x_local)
            hits = hits + inside
        return total / torch.where(hits == 0, torch.ones_like(hits), hits).float()

    def forward(self, x_world: torch.Tensor, noise_std=0) -> torch.Tensor:
        """Decode the mean feature with submap 0's decoder (grid_atlas.py:393-399)."""
        out = grid_decode(self.query_feature(x_world), None, self.submaps[0].decoder, True)
        return out + torch.randn_like(out) * noise_std if noise_std > 0 else out

    # ---- alignment support ------------------------------------------------------------------------------------
    def check_submap_intersection(self, src_id: int, dst_id: int, overlap_thresh=1e-2):
        """Share of src's finest-level voxel centres that land inside dst's bound > thresh (grid_atlas.py:405-420).
        Torch ops; `miso_b200.align` runs the batched kernel version inside the alignment loop."""
        centres = self.get_submap(src_id).features[-1].vertex_positions().to(self.device)
        in_world = geo.transform_points_to(centres, *self.updated_submap_pose(src_id))
        in_dst = geo.transfrom_points_from(in_world, *self.updated_submap_pose(dst_id))
        inside = geo.coords_in_bound(in_dst, self.get_submap(dst_id).bound)
        return torch.count_nonzero(inside) / centres.shape[0] > overlap_thresh

    def precompute_coordinates_for_alignment(self, norm_thresh=1e-5):
        """Per (submap, level): the voxel centres of that level whose full feature vector has norm > thresh, in
        (Z,Y,X) order with X fastest (grid_atlas.py:565-579)."""
        self._coords_for_alignment = {}
        for level in range(self.num_levels):
            for sid, sub in enumerate(self.submaps):
                centres = sub.features[level].vertex_positions().to(self.device)
                with torch.no_grad():
                    strength = torch.linalg.norm(sub.query_feature(centres), dim=1)
                keep = torch.nonzero(strength > norm_thresh)[:, 0]
                self._coords_for_alignment[f"submap{sid}_level{level}"] = centres[keep].detach()

    def coordinates_for_alignment(self, submap_id: int, level: int) -> torch.Tensor:
        assert 0 <= submap_id < self.num_submaps and 0 <= level < self.num_levels
        try:
            return self._coords_for_alignment[f"submap{submap_id}_level{level}"]
        except KeyError:
            raise ValueError(f"Coordinates for alignment not found for submap {submap_id} and level {level}. "
                             "Did you call precompute_coordinates_for_alignment()?") from None
