"""Gauss-Newton / Levenberg-Marquardt keyframe tracking step, mirror of `Tracker.lm_step`
(grid_opt/slam/tracker.py:148-212): residual r = sdf(R x + t) - gt, Jacobian
J = [ ((R x) x g)^T R , g^T ] with g = grad_x sdf, Geman-McClure weights, H = J^T W J + lambda I,
b = J^T W r, delta = solve(H, -b), w += delta[:3], tau += delta[3:].

On the B200 path the whole right-hand side is ONE launch (`miso_track_normal_equations`): transform,
interpolation, decoder + analytic spatial gradient (no autograd backward pass), the |gt| < trunc sample filter,
Geman-McClure weights, the 6-vector J and the float64 reductions H = J^T W J, b = J^T W r and the in-bound count.
What is left on the torch side is the 6x6 solve and the pose update.  `normal_equations_torch` keeps the previous
formulation (fused sdf+gradient launch, then torch ops on (N,6) tensors) as the in-repo cross-check."""
import ctypes as C
import math

import torch

from . import _lib
from . import field as _field
from . import geometry as utils_geometry
from .models import GridNet


class Tracker:
    def __init__(self, grid: GridNet, loss_type="GM", gm_scale_sdf=0.1, lm_lambda=1e-4, trunc_dist=None,
                 lm_max_iter=30, lm_tol_deg=0.01, lm_tol_m=0.001):
        self.grid = grid
        self.loss_type = loss_type
        self.gm_scale_sdf = gm_scale_sdf
        self.lm_lambda = lm_lambda
        self.trunc_dist = trunc_dist
        self.lm_max_iter = lm_max_iter
        self.lm_tol_deg = lm_tol_deg
        self.lm_tol_m = lm_tol_m

    def residual_weights(self, r: torch.Tensor):
        """tracker.py:139-146."""
        if self.loss_type == "L2":
            return torch.ones_like(r)
        if self.loss_type == "GM":
            return self.gm_scale_sdf / (self.gm_scale_sdf + r ** 2) ** 2
        raise ValueError(f"Unknown loss type: {self.loss_type}.")

    def normal_equations(self, coords_frame, gt_sdf, Rwf, twf, trunc_dist=None):
        """(H (6,6), g (6,1), fov_overlap tensor) for one keyframe from ONE kernel launch; coords_frame (N,3),
        gt_sdf (N,1).  `trunc_dist` applies the |gt| < trunc filter of lm_step (tracker.py:158-164) inside the kernel."""
        grid = self.grid
        spec = grid.fused_spec()
        if spec is None:
            raise RuntimeError("Tracker needs the fused field (fixed decoder)")
        lib = _lib.load()
        x = _field._prep_x(coords_frame)
        gt = gt_sdf.detach().reshape(-1).contiguous().float()
        dev = x.device
        Rt = torch.cat([Rwf.detach().reshape(9), twf.detach().reshape(3)]).contiguous().float()
        out = torch.empty(45, dtype=torch.float64, device=dev)
        fld = _field.make_field(grid.level_tensors(), spec.bound, None, spec.ignore_mask)
        dec = spec.decoder.struct()
        with torch.cuda.device(dev):
            _lib.check(lib.miso_track_normal_equations(
                C.byref(fld), C.byref(dec), x.data_ptr(), gt.data_ptr(), x.shape[0], Rt.data_ptr(),
                {"L2": 0, "GM": 1}[self.loss_type], float(self.gm_scale_sdf),
                float(trunc_dist) if trunc_dist is not None else -1.0, out.data_ptr(), _lib.stream_ptr(dev)),
                "track_normal_equations")
        H = out[:36].reshape(6, 6).float() + self.lm_lambda * torch.eye(6, device=dev)
        g = out[36:42].reshape(6, 1).float()
        fov = (out[42] / out[43].clamp(min=1.0)).float()
        return H, g, fov

    def normal_equations_torch(self, coords_frame, gt_sdf, Rwf, twf):
        """Same quantities from the fused sdf+gradient launch and torch ops on (N,6) tensors (cross-check)."""
        grid = self.grid
        spec = grid.fused_spec()
        if spec is None:
            raise RuntimeError("Tracker needs the fused field (fixed decoder)")
        coords_world = utils_geometry.transform_points_to(coords_frame, Rwf, twf)
        mask_bnd = utils_geometry.coords_in_bound(coords_world, grid.bound)
        sdf, _, grad_world, _ = _field.sdf_forward_raw(grid.level_tensors(), spec, coords_world, want_jac=False,
                                                       want_gradx=True)
        Rxi = utils_geometry.transform_points_to(coords_frame, Rwf, torch.zeros_like(twf))
        cT = torch.cross(Rxi, grad_world, dim=1)            # hat(Rx) g  (tracker.py:181-183)
        J = torch.cat((cT @ Rwf, grad_world), dim=1)        # (N,6) = [J_R, J_t]
        r = sdf.unsqueeze(1) - gt_sdf
        w = self.residual_weights(r)
        H = J.T @ (w * J) + self.lm_lambda * torch.eye(6, device=J.device)
        g = J.T @ (w * r)
        return H, g, mask_bnd.float().mean()

    @torch.no_grad()
    def lm_step(self, optimize_kf: int, model_input: dict, gt: dict):
        """One LM step for keyframe `optimize_kf` on a batch holding only that keyframe's samples
        (the reference selects them with dataset.select_keyframes, tracker.py:152-153)."""
        coords_frame = model_input["coords_frame"][0]
        frame_ids = model_input["sample_frame_ids"][0]
        gt_sdf = gt["sdf"][0]
        gt_sdf_valid = gt["sdf_valid"][0]
        grid = self.grid
        Rwf, twf = grid.updated_kf_pose_from_key(f"KF{optimize_kf}")
        Rwf, twf = Rwf.detach(), twf.detach()
        # the |gt| < trunc selection (tracker.py:158-164) happens inside the kernel: no nonzero() / gather passes
        H, g, fov = self.normal_equations(coords_frame, gt_sdf, Rwf, twf, trunc_dist=self.trunc_dist)
        delta = torch.linalg.solve(H, -g)
        delta_R, delta_t = delta[:3], delta[3:]
        kf_id = grid.pose_key_to_id(f"KF{optimize_kf}")
        grid.rotation_corrections[kf_id] += delta_R.squeeze()
        grid.translation_corrections[kf_id] += delta_t
        return {"delta_R_deg": math.degrees(torch.linalg.norm(delta_R).item()),
                "delta_t_norm": torch.linalg.norm(delta_t).item(), "grad_norm": torch.linalg.norm(g).item(),
                "fov_overlap": float(fov)}

    def track(self, optimize_kf: int, model_input: dict, gt: dict):
        """tracker.py:125-137: iterate lm_step until the update is below tolerance."""
        info = None
        for step in range(self.lm_max_iter):
            info = self.lm_step(optimize_kf, model_input, gt)
            if info["delta_R_deg"] < self.lm_tol_deg and info["delta_t_norm"] < self.lm_tol_m:
                break
        return info
