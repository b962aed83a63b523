"""Adapter: run the reference's OWN classes on the CUDA library (INTEGRATION.md section 2 as code).

`patch_reference()` needs `grid_opt` importable (the reference tree on sys.path).  It leaves every reference class in
place and re-routes, for CUDA tensors only, the five calls that make up the hot path:

    FeatureGrid.interpolate        grid_opt/models/grid_modules.py:72-95   -> miso_b200.cuda_gridsample.grid_sample_3d
    GridNet.query_feature          grid_opt/models/grid_net.py:288-297     -> miso_field_features
    GridNet.forward                grid_opt/models/grid_net.py:306-325     -> miso_sdf_forward / _backward (decoder fixed)
    diff.gradient3d ('autograd')   grid_opt/diff.py:27-33                  -> the fused kernel's analytic gradient
    MisoLossMappingBase.compute    grid_opt/loss.py:754-813                -> miso_mapping_step (+ _fd)

Anything the fused path does not cover (CPU tensors, trainable decoder, unlocked poses, VM grids, 2-D) falls through to
the reference's original method, untouched -- that is the reference running its own code, not a fallback of this
package.  `unpatch_reference()` restores the originals.  The grid parameters are re-laid channels_last_3d in place on
first use (same logical shape, values and state-dict keys).
"""
import torch

from . import cuda_gridsample as cu
from . import field as _field
from . import geometry as geo
from . import loss as _loss

_ORIG = {}


def _levels(net):
    feats = [g.feature for g in net.features]
    for f in feats:
        if f.is_cuda and f.stride(1) != 1:
            _field.to_channels_last_3d_(f)
    return feats


def _fused_spec(net, trainable_decoder_ok=False):
    dec = getattr(net, "decoder", None)
    feats = net.features
    ok = (dec is not None and getattr(net, "decoder_type", "mlp") == "mlp" and net.pos_invariant
          and net.decoder_hidden_dim == 64 and net.decoder_hidden_layers == 1 and net.decoder_out_dim == 1
          and (net.num_levels, net.fdim) in {(1, 4), (2, 4), (3, 4), (4, 4), (1, 8), (2, 8), (1, 16)}
          and feats[0].feature.is_cuda
          and (trainable_decoder_ok or not any(p.requires_grad for p in dec.parameters()))
          and getattr(net, "grid_type", "regular") == "regular")
    if not ok:
        return None
    mask = sum(1 << l for l in range(net.num_levels) if net.ignore_level_[l])
    return _field.FieldSpec(_field.bound_to_list(net.bound), _field.DecoderSpec.from_mlp(dec), mask)


def _wants_graph(net, x):
    return torch.is_grad_enabled() and (x.requires_grad or any(g.feature.requires_grad for g in net.features))


def _interpolate(self, x):
    if not (x.is_cuda and self.feature.is_cuda and self.feature.ndim == 5):
        return _ORIG["interpolate"](self, x)
    lo, hi = self.bound[:, 0].view(1, -1).to(x), self.bound[:, 1].view(1, -1).to(x)
    unit = (2 * (x - lo) / (hi - lo) - 1).reshape(1, -1, 1, 1, 3)
    return cu.grid_sample_3d(self.feature, unit, padding_mode="zeros", align_corners=False)[0, :, :, 0, 0].t()


def _query_feature(self, x):
    if not (x.is_cuda and self.d == 3 and self.fdim in (4, 8, 12, 16)) or _wants_graph(self, x):
        return _ORIG["query_feature"](self, x)
    mask = sum(1 << l for l in range(self.num_levels) if self.ignore_level_[l])
    return _field.field_features_raw(_levels(self), _field.bound_to_list(self.bound), x, mask)


def _forward(self, x, noise_std=0):
    spec = _fused_spec(self) if x.is_cuda and self.d == 3 else None
    if spec is None:
        return _ORIG["forward"](self, x, noise_std)
    if _wants_graph(self, x):
        out, _ = _field.fused_sdf(x, _levels(self), spec)
    else:
        out = _field.sdf_forward_raw(_levels(self), spec, x, want_jac=False, want_gradx=False)[0].unsqueeze(1)
    return out + torch.randn_like(out) * noise_std if noise_std > 0 else out


def _forward_with_gradient(self, x):
    return _field.fused_sdf(x, _levels(self), _fused_spec(self))


def _all_kf_poses(self):
    dr, dt = self.rotation_corrections, self.translation_corrections
    locked = getattr(self, "locked_pose_indices", ())
    if locked and (dr.requires_grad or dt.requires_grad):
        frozen = torch.zeros(self.num_poses, dtype=torch.bool, device=dr.device)
        frozen[sorted(locked)] = True
        dr = torch.where(frozen[:, None], dr.detach(), dr)
        dt = torch.where(frozen[:, None, None], dt.detach(), dt)
    return torch.matmul(self.Rwk, geo.so3_exp_map(dr)), self.twk + dt


def _gradient3d(x, f, method="finitediff", finite_diff_eps=1e-2, create_graph=True):
    if method == "autograd" and x.is_cuda and hasattr(f, "features") and _fused_spec(f) is not None:
        grad = _forward_with_gradient(f, x)[1]
        return grad if create_graph else grad.detach()
    return _ORIG["gradient3d"](x, f, method=method, finite_diff_eps=finite_diff_eps, create_graph=create_graph)


class _FusedLossView(_loss.MisoLossMapping):
    """The reference loss object's hyper-parameters viewed through this package's fused `compute`."""

    def __init__(self, ref_loss):
        self.__dict__.update({k: getattr(ref_loss, k) for k in (
            "loss_type", "trunc_dist", "weight_sdf", "weight_eik", "weight_fs", "finite_diff_eps", "grad_method",
            "eik_trunc_dist")})
        self.use_stability = getattr(ref_loss, "use_stability", False)
        self.weight_clip = getattr(ref_loss, "weight_clip", 0)
        self.last_terms, self.check_frame_ids = None, True


def _compute(self, model, model_input, gt):
    x = model_input["coords_frame"]
    view = _FusedLossView(self)
    if (x.is_cuda and not view.use_stability and view.weight_clip == 0 and hasattr(model, "_pose_key_to_id")
            and view._fused_ok(model)):
        return view.compute(model, model_input, gt)
    return _ORIG["compute"](self, model, model_input, gt)


def patch_reference():
    """Install the re-routes on the imported `grid_opt` package.  Idempotent.  Returns the patched names."""
    import grid_opt.diff as gd
    import grid_opt.loss as gl
    import grid_opt.models.grid_modules as gm
    import grid_opt.models.grid_net as gn
    if _ORIG:
        return sorted(_ORIG)
    _ORIG.update(interpolate=gm.FeatureGrid.interpolate, query_feature=gn.GridNet.query_feature,
                 forward=gn.GridNet.forward, gradient3d=gd.gradient3d, compute=gl.MisoLossMappingBase.compute)
    gm.FeatureGrid.interpolate = _interpolate
    gn.GridNet.query_feature = _query_feature
    gn.GridNet.forward = _forward
    gn.GridNet.forward_with_gradient = _forward_with_gradient
    gn.GridNet.fused_spec = _fused_spec
    gn.GridNet.level_tensors = _levels
    gn.GridNet.all_kf_poses = _all_kf_poses
    gd.gradient3d = _gradient3d
    if hasattr(gl, "gradient3d"):
        _ORIG["loss.gradient3d"] = gl.gradient3d
        gl.gradient3d = _gradient3d
    gl.MisoLossMappingBase.compute = _compute
    return sorted(_ORIG)


def unpatch_reference():
    import grid_opt.diff as gd
    import grid_opt.loss as gl
    import grid_opt.models.grid_modules as gm
    import grid_opt.models.grid_net as gn
    if not _ORIG:
        return
    gm.FeatureGrid.interpolate = _ORIG["interpolate"]
    gn.GridNet.query_feature = _ORIG["query_feature"]
    gn.GridNet.forward = _ORIG["forward"]
    gd.gradient3d = _ORIG["gradient3d"]
    if "loss.gradient3d" in _ORIG:
        gl.gradient3d = _ORIG["loss.gradient3d"]
    gl.MisoLossMappingBase.compute = _ORIG["compute"]
    for name in ("forward_with_gradient", "fused_spec", "level_tensors", "all_kf_poses"):
        if hasattr(gn.GridNet, name):
            delattr(gn.GridNet, name)
    _ORIG.clear()
