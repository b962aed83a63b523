"""Mapping losses of the hot path, same names and semantics as grid_opt/loss.py:
`miso_loss_regression` (:594-635), `miso_loss_eikonal` (:638-665), `miso_loss_free_space`
(:668-700), `MisoLossMappingBase.compute` / `MisoLossMapping` (:703-813, :847-853).

Two execution paths, both on the GPU:
  * fused   -- ONE kernel does frame->world transform, interpolation of every level, the MLP, its
               analytic spatial gradient, the three loss terms and the scatter of d(total)/d(grid)
               (miso_mapping_step).  Taken when the model exposes a fused spec (decoder fixed) and
               keyframe poses are locked.  grad_method 'finitediff' (the shipped default) runs as
               four launches (miso_mapping_step_fd): step, six displaced forwards in one launch,
               eikonal epilogue, six backward scatters in one launch.
  * generic -- the reference's op sequence on torch tensors; the per-level interpolation is the
               twice-differentiable miso_b200.cuda_gridsample op, the eikonal gradient comes from
               miso_b200.diff.gradient3d.  Handles trainable decoders, unlocked poses, finite
               differences, Cosine loss.

Reference defect mirrored deliberately: the reference asserts on an
attribute `use_clip` that is never set (loss.py:788), so `weight_eik > 0` raises there; here the
eikonal term simply works.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from . import field as _field
from .diff import gradient3d


def miso_loss_regression(pred, targ, valid_mask=None, sample_weights=None, loss_type="L1"):
    """mean_i( w_i * [valid_i == 1] * rho(pred_i, targ_i) ), rho = squared / absolute error summed over the feature
    dim, or 1 - cosine similarity; the mean runs over ALL samples (loss.py:594-635)."""
    if pred.shape != targ.shape:
        raise AssertionError(f"pred {tuple(pred.shape)} vs targ {tuple(targ.shape)}")
    n = pred.shape[0]
    diff = pred - targ
    if loss_type == "L1":
        per_sample = diff.abs().sum(dim=1, keepdim=True)
    elif loss_type == "L2":
        per_sample = (diff ** 2).sum(dim=1, keepdim=True)
    elif loss_type == "Cosine":
        per_sample = (1.0 - F.cosine_similarity(pred, targ, dim=1, eps=1e-8)).unsqueeze(1)
    else:
        raise ValueError(f"Invalid loss type: {loss_type}")
    if valid_mask is not None:
        assert valid_mask.shape == (n, 1)
        per_sample = torch.where(valid_mask == 1, per_sample, torch.zeros_like(per_sample))
    if sample_weights is not None:
        assert sample_weights.shape == (n, 1)
        per_sample = sample_weights * per_sample
    return per_sample.mean()


def miso_loss_eikonal(model, coords_world, gt_sdf, eik_trunc_dist, grad_method, finite_diff_eps):
    """mean (|grad_x f| - 1)^2 over the samples with |gt| < eik_trunc_dist (all samples when None), loss.py:638-665."""
    pts = coords_world if eik_trunc_dist is None else coords_world[(gt_sdf.abs() < eik_trunc_dist)[:, 0]]
    pts = pts.clone().requires_grad_(True)
    g = gradient3d(pts, model, method=grad_method, finite_diff_eps=finite_diff_eps, create_graph=True)
    return ((g.norm(dim=-1) - 1) ** 2).mean()


def miso_loss_free_space(pred_sdf, gt_sdf, gt_sdf_sign, trunc_dist):
    """On samples in front of the surface (sign == 1): max(relu(pred - gt), relu(trunc - pred)); mean over all
    samples (loss.py:668-700)."""
    assert trunc_dist is not None
    front = gt_sdf_sign == 1
    zero = torch.zeros_like(pred_sdf)
    over = torch.where(front, F.relu(pred_sdf - gt_sdf), zero)
    under = torch.where(front, F.relu(trunc_dist - pred_sdf), zero)
    return torch.maximum(over, under).mean()


# ------------------------------------------------------------------------------------------------
# fused step
# ------------------------------------------------------------------------------------------------
# bench.py hook: when set to a list, (start, end) CUDA events are recorded around every fused step launch
PROFILE_EVENTS = None


class MappingWorkspace:
    """Per-device scratch of the fused step (never allocated inside the kernel)."""
    _cache = {}

    def __init__(self, device):
        lib = _lib.load()
        self.partials = torch.zeros(int(lib.miso_mapping_workspace_floats()), dtype=torch.float32, device=device)
        self.eik_count = torch.zeros(1, dtype=torch.int32, device=device)
        self.fd = None   # (12 N) floats of the finite-difference step, grown on demand
        self.wgrad = None   # per-CTA rows of the decoder-gradient pass, allocated on first use

    @classmethod
    def get(cls, device):
        key = (device.type, device.index)
        if key not in cls._cache:
            cls._cache[key] = cls(device)
        return cls._cache[key]


def mapping_step_raw(feats, grads, spec, frames, x, gt_sdf, gt_valid, gt_sign, weights, *, loss_type, weight_sdf,
                     weight_fs, weight_eik, trunc_dist, eik_trunc_dist, eik_on, grad_scale=1.0, sdf_out=None,
                     n_total=0, count_allreduce=None, fd_eps=None, n_device=None, count_on=None, dec_grads=None,
                     loss_out=None):
    """Launch the fused mapping step.  Returns a (4,) float tensor [sdf, fs, eik, total] (unweighted
    terms, weighted total).  Gradients are ACCUMULATED into `grads` (None entries are skipped).
    `fd_eps` selects the finite-difference eikonal term (miso_mapping_step_fd) instead of the analytic one.
    `n_device` (device int32 tensor) limits the step to the first *n_device samples (batch compacted on the device,
    miso_b200.sharded_fit); `count_on` is then the FULL batch's gt sdf for the |gt| < eik_trunc count.
    `dec_grads` (six float32 tensors shaped like W1, b1, W2, b2, W3, b3, or None entries): the trainable-decoder case --
    d total / d decoder parameters is accumulated into them by a second pass (miso_mapping_step_wgrad)."""
    lib = _lib.load()
    dev = x.device
    N = x.shape[0]
    ws = MappingWorkspace.get(dev)
    cfg = _lib.MappingCfg()
    cfg.loss_type = {"L1": 0, "L2": 1}[loss_type]
    cfg.weight_sdf, cfg.weight_fs, cfg.weight_eik = float(weight_sdf), float(weight_fs), float(weight_eik)
    cfg.trunc_dist = float(trunc_dist if trunc_dist is not None else 0.0)
    cfg.eik_trunc_dist = float(eik_trunc_dist) if eik_trunc_dist is not None else -1.0
    cfg.eik_mode = 1 if (eik_on and weight_eik > 0) else 0
    cfg.grad_scale = float(grad_scale)
    cfg.n_total = int(n_total)
    cfg.n_device = n_device.data_ptr() if n_device is not None else None
    fld = _field.make_field(feats, spec.bound, grads, spec.ignore_mask)
    dec = spec.decoder.struct()
    fr = frames.struct() if frames is not None else None
    if loss_out is None:
        loss_out = torch.empty(4, dtype=torch.float32, device=dev)
    stream = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):
        if cfg.eik_mode == 1 and eik_trunc_dist is not None:
            cnt_src = gt_sdf if count_on is None else count_on
            _lib.check(lib.miso_mapping_count(cnt_src.data_ptr(), cnt_src.numel(), cfg.eik_trunc_dist,
                                              ws.eik_count.data_ptr(), stream), "mapping_count")
            if count_allreduce is not None:
                count_allreduce([ws.eik_count])   # point-sharded fit: the eikonal mean runs over all ranks
        profile = PROFILE_EVENTS is not None and not torch.cuda.is_current_stream_capturing()
        if profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(dev))
        if fd_eps is not None and cfg.eik_mode == 1:
            if ws.fd is None or ws.fd.numel() < 12 * N:
                ws.fd = torch.empty(12 * N, dtype=torch.float32, device=dev)
            _lib.check(lib.miso_mapping_step_fd(
                C.byref(fld), C.byref(dec), C.byref(fr) if fr is not None else None, x.data_ptr(), N,
                gt_sdf.data_ptr(), gt_valid.data_ptr(), gt_sign.data_ptr(), _lib.ptr(weights), C.byref(cfg),
                ws.eik_count.data_ptr(), ws.partials.data_ptr(), loss_out.data_ptr(), _lib.ptr(sdf_out),
                float(fd_eps), ws.fd.data_ptr(), stream), "mapping_step_fd")
        else:
            _lib.check(lib.miso_mapping_step(
                C.byref(fld), C.byref(dec), C.byref(fr) if fr is not None else None, x.data_ptr(), N,
                gt_sdf.data_ptr(), gt_valid.data_ptr(), gt_sign.data_ptr(), _lib.ptr(weights), C.byref(cfg),
                ws.eik_count.data_ptr(), ws.partials.data_ptr(), loss_out.data_ptr(), _lib.ptr(sdf_out), stream),
                "mapping_step")
        if dec_grads is not None:
            if fd_eps is not None and cfg.eik_mode == 1:
                raise RuntimeError("decoder gradients are fused for the analytic eikonal term only")
            if n_device is not None:
                raise RuntimeError("decoder gradients are not available for device-compacted batches")
            if ws.wgrad is None:
                ws.wgrad = torch.empty(int(lib.miso_mapping_wgrad_workspace_floats()), dtype=torch.float32, device=dev)
            dg = _lib.DecoderGrad()
            for name, t, ref in zip(("W1", "b1", "W2", "b2", "W3", "b3"), dec_grads, spec.decoder.tensors):
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != ref.numel()):
                    raise RuntimeError(f"decoder gradient {name}: expected a contiguous float32 tensor of {ref.numel()} elements")
                setattr(dg, name, _lib.ptr(t))
            _lib.check(lib.miso_mapping_step_wgrad(
                C.byref(fld), C.byref(dec), C.byref(fr) if fr is not None else None, x.data_ptr(), N,
                gt_sdf.data_ptr(), gt_valid.data_ptr(), gt_sign.data_ptr(), _lib.ptr(weights), C.byref(cfg),
                ws.eik_count.data_ptr(), C.byref(dg), ws.wgrad.data_ptr(), stream), "mapping_step_wgrad")
        if profile:
            e1.record(torch.cuda.current_stream(dev))
            PROFILE_EVENTS.append((e0, e1))
    return loss_out


def _flat_f32(t):
    t = t.detach().reshape(-1)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _flat_u8(t):
    t = t.detach().reshape(-1)
    if t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8)
    return (t != 0).to(torch.uint8).contiguous()


class _FusedMappingLoss(torch.autograd.Function):
    """Forward runs the whole step (including the gradient scatter); backward hands the stored
    gradients to autograd scaled by the upstream cotangent.  Outputs are the three WEIGHTED terms so
    the caller's `sum(loss_dict.values())` has unit cotangents; unequal cotangents poison the result
    with NaN instead of returning a silently wrong gradient (a single kernel cannot un-mix them)."""

    @staticmethod
    def forward(ctx, x, gt_sdf, gt_valid, gt_sign, weights, spec, frames, cfg, n_levels, *tensors):
        # tensors = the level grids, then (trainable decoder only) W1, b1, W2, b2, W3, b3
        feats, dec = tensors[:n_levels], tensors[n_levels:]
        grads = [torch.zeros_like(f) if f.requires_grad else None for f in feats]
        dgrads = [torch.zeros_like(p) if p.requires_grad else None for p in dec]
        out = mapping_step_raw(feats, grads, spec, frames, x, gt_sdf, gt_valid, gt_sign, weights,
                               dec_grads=dgrads if any(g is not None for g in dgrads) else None, **cfg)
        ctx.grads = grads + dgrads
        ctx.set_materialize_grads(False)  # terms the caller drops arrive as None, not as zeros
        w = out.new_tensor([cfg["weight_sdf"], cfg["weight_fs"], cfg["weight_eik"] if cfg["eik_on"] else 0.0])
        terms = out[:3] * w
        ctx.mark_non_differentiable(out)
        return terms[0], terms[1], terms[2], out

    @staticmethod
    def backward(ctx, g0, g1, g2, _g_out):
        grads = ctx.grads
        ctx.grads = None
        gs = [g for g in (g0, g1, g2) if g is not None]
        if not gs:
            return (None,) * (9 + len(grads))
        s = gs[0]
        for g in gs[1:]:
            s = torch.where(g == s, s, torch.full_like(s, float("nan")))
        outs = []
        for G in grads:
            outs.append(None if G is None else G.mul_(s))
        return (None,) * 9 + tuple(outs)


class MisoLossMappingBase:
    """loss.py:703-813.  Same constructor arguments and `compute(model, model_input, gt) -> dict`."""

    def __init__(self, loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0, trunc_dist=0,
                 finite_diff_eps=1e-2, grad_method="autograd", eik_trunc_dist=0.1, use_stability=False,
                 weight_clip=0):
        self.loss_type = loss_type
        self.trunc_dist = trunc_dist
        self.weight_sdf = weight_sdf
        self.weight_eik = weight_eik
        self.weight_fs = weight_fs
        self.finite_diff_eps = finite_diff_eps
        self.grad_method = grad_method
        self.eik_trunc_dist = eik_trunc_dist
        self.use_stability = use_stability
        self.weight_clip = weight_clip
        if use_stability or weight_clip > 0:
            raise NotImplementedError("stability / CLIP terms are outside the hot path (SURVEY.md section 8)")
        self.last_terms = None
        # `compute` checks (one host sync) that every sample's keyframe has a pose, as the reference's per-keyframe
        # loop does by construction; the trainer's sync-free `step_into_grads` relies on the kernel's NaN poison
        self.check_frame_ids = True

    def validate_frame_ids(self, model, sample_frame_ids):
        pass

    # -- pose table -------------------------------------------------------------------------------
    def frame_table(self, model):
        """(R (K,3,3), t (K,3,1), lut) with lut[global kf id] = row; replaces the per-keyframe
        query_kf_pose loop + np.unique host sync of loss.py:764-774."""
        raise NotImplementedError

    def _fused_ok(self, model):
        if getattr(model, "fused_spec", None) is None or model.fused_spec(trainable_decoder_ok=True) is None:
            return False
        if self.loss_type not in ("L1", "L2"):
            return False
        if self.weight_eik > 0 and self.grad_method != "autograd":
            # finite differences run fused (miso_mapping_step_fd) where the two-threads-per-point kernel applies;
            # the decoder-gradient pass (miso_mapping_step_wgrad) covers the analytic term only
            if (self.grad_method != "finitediff" or not self._fd_kernel_covers(model)
                    or self._trainable_decoder(model)):
                return False
        R, t, _ = self.frame_table(model)
        return not (R.requires_grad or t.requires_grad)

    @staticmethod
    def _trainable_decoder(model):
        dec = getattr(model, "decoder", None)
        return dec is not None and any(p.requires_grad for p in dec.parameters())

    @staticmethod
    def _fd_kernel_covers(model) -> bool:
        shape_ok = (model.num_levels, model.fdim) in {(2, 4), (4, 4), (1, 8), (2, 8), (1, 16)}
        return shape_ok and all(f.numel() + 4 * f.stride(2) < 2 ** 31 - 1 for f in model.level_tensors())

    def compute(self, model, model_input: dict, gt: dict) -> dict:
        coords_frame = model_input["coords_frame"][0]
        sample_frame_ids = model_input["sample_frame_ids"][0, :, 0]
        sample_weights = model_input["weights"][0]
        gt_sdf = gt["sdf"][0]
        gt_sdf_valid = gt["sdf_valid"][0]
        gt_sdf_sign = gt["sdf_signs"][0]
        assert coords_frame.ndim == 2 and gt_sdf.ndim == 2
        assert sample_weights.shape == gt_sdf.shape
        if self.check_frame_ids:
            self.validate_frame_ids(model, sample_frame_ids)
        if self._fused_ok(model):
            return self._compute_fused(model, coords_frame, sample_frame_ids, sample_weights, gt_sdf, gt_sdf_valid,
                                       gt_sdf_sign)
        return self._compute_generic(model, coords_frame, sample_frame_ids, sample_weights, gt_sdf, gt_sdf_valid,
                                     gt_sdf_sign)

    def _step_cfg(self):
        return dict(loss_type=self.loss_type, weight_sdf=self.weight_sdf, weight_fs=self.weight_fs,
                    weight_eik=self.weight_eik, trunc_dist=self.trunc_dist, eik_trunc_dist=self.eik_trunc_dist,
                    eik_on=self.weight_eik > 0,
                    fd_eps=self.finite_diff_eps if (self.weight_eik > 0 and self.grad_method == "finitediff") else None)

    def _frames(self, model, sample_frame_ids):
        R, t, lut = self.frame_table(model)
        ids = lut[sample_frame_ids.reshape(-1)] if lut is not None else sample_frame_ids.reshape(-1)
        return _field.FramesSpec(ids, R, t)

    def _compute_fused(self, model, coords_frame, ids, weights, gt_sdf, gt_valid, gt_sign):
        spec = model.fused_spec(trainable_decoder_ok=True)
        frames = self._frames(model, ids)
        x = _field._prep_x(coords_frame)
        levels = model.level_tensors()
        dec_params = _field.decoder_parameters(model.decoder) if self._trainable_decoder(model) else []
        t_sdf, t_fs, t_eik, raw = _FusedMappingLoss.apply(
            x, _flat_f32(gt_sdf), _flat_u8(gt_valid), _flat_f32(gt_sign), _flat_f32(weights), spec, frames,
            self._step_cfg(), len(levels), *levels, *dec_params)
        self.last_terms = raw
        loss_dict = {f"sdf_{self.loss_type}": t_sdf}
        if self.weight_eik > 0:
            loss_dict["eik"] = t_eik
        if self.weight_fs > 0:
            loss_dict["free_space"] = t_fs
        return loss_dict

    def _compute_generic(self, model, coords_frame, ids, weights, gt_sdf, gt_valid, gt_sign):
        R, t, lut = self.frame_table(model)
        loc = lut[ids] if lut is not None else ids
        coords_world = torch.einsum("nij,nj->ni", R[loc], coords_frame) + t[loc].squeeze(-1)
        pred_sdf = model(coords_world)[:, [0]]
        sdf_loss = miso_loss_regression(pred=pred_sdf, targ=gt_sdf, valid_mask=gt_valid, sample_weights=weights,
                                        loss_type=self.loss_type)
        loss_dict = {f"sdf_{self.loss_type}": self.weight_sdf * sdf_loss}
        if self.weight_eik > 0:
            eik_loss = miso_loss_eikonal(model=model, coords_world=coords_world, gt_sdf=gt_sdf,
                                         eik_trunc_dist=self.eik_trunc_dist, grad_method=self.grad_method,
                                         finite_diff_eps=self.finite_diff_eps)
            loss_dict["eik"] = eik_loss * self.weight_eik
        if self.weight_fs > 0:
            fs_loss = miso_loss_free_space(pred_sdf=pred_sdf, gt_sdf=gt_sdf, gt_sdf_sign=gt_sign,
                                           trunc_dist=self.trunc_dist)
            loss_dict["free_space"] = fs_loss * self.weight_fs
        return loss_dict

    # -- autograd-free variant used by miso_b200.trainer (gradients go straight into param.grad) -----
    def step_into_grads(self, model, model_input: dict, gt: dict, active_levels=None, n_total=0,
                        count_allreduce=None):
        """Runs the fused step accumulating into `feature.grad` of the active levels.  Returns the (4,)
        loss tensor [sdf, fs, eik, total].  `n_total` / `count_allreduce` are the point-sharded multi-GPU
        hooks (miso_b200.dist): means run over the global batch so per-rank results sum."""
        if not self._fused_ok(model):
            raise RuntimeError("step_into_grads needs the fused path (64-wide MLP decoder, locked poses; a trainable "
                               "decoder needs grad_method='autograd' when weight_eik > 0)")
        coords_frame = model_input["coords_frame"][0]
        ids = model_input["sample_frame_ids"][0, :, 0]
        feats = model.level_tensors()
        grads = []
        for l, f in enumerate(feats):
            active = f.requires_grad and (active_levels is None or l in active_levels)
            if active and f.grad is None:
                f.grad = torch.zeros_like(f)
            grads.append(f.grad if active else None)
        dec_grads = None
        if self._trainable_decoder(model):
            dec_grads = []
            for p in _field.decoder_parameters(model.decoder):
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)
                dec_grads.append(p.grad if p.requires_grad else None)
        raw = mapping_step_raw(feats, grads, model.fused_spec(trainable_decoder_ok=True), self._frames(model, ids),
                               _field._prep_x(coords_frame), _flat_f32(gt["sdf"][0]), _flat_u8(gt["sdf_valid"][0]),
                               _flat_f32(gt["sdf_signs"][0]), _flat_f32(model_input["weights"][0]),
                               n_total=n_total, count_allreduce=count_allreduce, dec_grads=dec_grads,
                               **self._step_cfg())
        self.last_terms = raw
        return raw


class MisoLossMapping(MisoLossMappingBase):
    """loss.py:847-853: mapping within a single submap (GridNet); keyframe k is addressed by the key
    f'KF{k}' (grid_net.py:232-235)."""

    def frame_table(self, model):
        """Pose tables indexed by GLOBAL keyframe id (so the kernel consumes `sample_frame_ids` directly).  An id
        with no registered 'KF<id>' key gets a NaN row: the kernels turn a sample that hits one (or an id outside the
        table) into a NaN step loss instead of training it with some other keyframe's pose, and `compute` raises the
        reference's assertion for it (grid_net.py:243).  While poses are locked the table is cached and only
        rebuilt when a pose tensor's version counter moves."""
        keys = model._pose_key_to_id
        params = (model.rotation_corrections, model.translation_corrections, model.Rwk, model.twk)
        locked = not (model.rotation_corrections.requires_grad or model.translation_corrections.requires_grad)
        stamp = (len(keys), tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        cache = getattr(model, "_miso_frame_cache", None)
        if locked and cache is not None and cache[0] == stamp:
            return cache[1], cache[2], None
        R, t = model.all_kf_poses()
        n = max(1 + max([int(k[2:]) for k in keys], default=-1), 1)
        row = torch.full((n,), -1, dtype=torch.int64)
        for k, v in keys.items():
            row[int(k[2:])] = v
        known = (row >= 0).to(R.device)
        row = row.clamp(min=0).to(R.device)
        Rg = R[row]
        tg = torch.where(known[:, None, None], t[row], torch.full_like(t[row], float("nan")))
        model._miso_frame_known = known
        if locked:
            model._miso_frame_cache = (stamp, Rg.detach(), tg.detach())
        return Rg, tg, None

    def validate_frame_ids(self, model, sample_frame_ids):
        """Host-side check (one device sync) that every sample's keyframe id has a registered pose; raises the
        reference's message (grid_net.py:243) for the first one that has not."""
        self.frame_table(model)
        known = model._miso_frame_known
        ids = sample_frame_ids.reshape(-1)
        inside = (ids >= 0) & (ids < known.numel())
        ok = inside & known[ids.clamp(0, known.numel() - 1)]
        if not bool(ok.all()):
            bad = int(ids[~ok][0])
            raise AssertionError(f"Key KF{bad} not found in pose key to ID mapping!")
