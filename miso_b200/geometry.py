"""SE(3) glue of the hot path.  Function names and argument meaning follow the reference's
grid_opt/utils/utils_geometry.py (:11-27 in-bound test, :61-99 pose helpers, :214-240 point transforms) so call
sites carry over; the bodies are small torch expressions that stay in the autograd graph, which is how pose
gradients finish through the exponential map exactly as in the reference.

`so3_exp_map` / `hat` restate pytorch3d (not vendored and un-pinned by the reference, environment.yaml:114):
Rodrigues' formula with the angle clamped from below, theta = sqrt(max(|w|^2, 1e-4)).
"""
import torch


def _check(name, tensor, shape):
    if tuple(tensor.shape) != shape:
        raise AssertionError(f"{name} must have shape {shape}, got {tuple(tensor.shape)}")


def hat(v: torch.Tensor) -> torch.Tensor:
    """(N,3) -> (N,3,3) cross-product matrices [v]_x."""
    K = v.new_zeros(v.shape[0], 3, 3)
    K[:, 2, 1], K[:, 0, 2], K[:, 1, 0] = v[:, 0], v[:, 1], v[:, 2]
    return K - K.transpose(1, 2)


def so3_exp_map(log_rot: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """(N,3) axis-angle -> (N,3,3): I + sin(t)/t K + (1-cos t)/t^2 K^2 with t = sqrt(clamp(|w|^2, eps))."""
    theta = (log_rot * log_rot).sum(1).clamp(min=eps).sqrt()
    inv = 1.0 / theta
    a = (inv * theta.sin()).reshape(-1, 1, 1)
    b = (inv * inv * (1.0 - theta.cos())).reshape(-1, 1, 1)
    K = hat(log_rot)
    return a * K + b * torch.bmm(K, K) + torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)


def identity_rotations(n: int) -> torch.Tensor:
    return torch.eye(3).expand(n, 3, 3).clone()


def coords_in_bound(coords: torch.Tensor, bound: torch.Tensor) -> torch.Tensor:
    """(N,d) points vs (d,2) bound -> (N,1) bool, both faces inclusive."""
    lo, hi = bound[:, 0], bound[:, 1]
    return torch.logical_and(coords >= lo, coords <= hi).all(dim=1, keepdim=True)


def apply_pose_correction(R, t, R_delta, t_delta):
    """(R Exp(R_delta), t + t_delta) for R (3,3), t (3,1), R_delta (1,3), t_delta (3,1)."""
    for name, ten, shp in (("R", R, (3, 3)), ("t", t, (3, 1)), ("R_delta", R_delta, (1, 3)), ("t_delta", t_delta, (3, 1))):
        _check(name, ten, shp)
    return R @ so3_exp_map(R_delta)[0], t + t_delta


def transform_points_to(points_src, R_dst_src, t_dst_src):
    """Rows of points_src (N,3) mapped src -> dst: p R^T + t^T."""
    _check("R_dst_src", R_dst_src, (3, 3))
    _check("t_dst_src", t_dst_src, (3, 1))
    return points_src @ R_dst_src.T + t_dst_src.T


def transfrom_points_from(points_dst, R_dst_src, t_dst_src):
    """Inverse direction (the reference's spelling is part of its interface): dst -> src with the inverted pose
    (R^T, -R^T t), composed in this order so rounding matches the reference."""
    _check("R_dst_src", R_dst_src, (3, 3))
    _check("t_dst_src", t_dst_src, (3, 1))
    Rt = R_dst_src.T
    return transform_points_to(points_dst, Rt, -Rt @ t_dst_src)


def pose_matrix(R, t):
    """4x4 homogeneous matrix [[R, t], [0, 1]]."""
    _check("R", R, (3, 3))
    _check("t", t, (3, 1))
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).to(R)
    return torch.cat([torch.cat([R, t], dim=1), bottom], dim=0)
