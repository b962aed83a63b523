"""SE(3) glue of the hot path, same names / argument meaning as the reference's
grid_opt/utils/utils_geometry.py (:11-27, :78-99, :214-240).  These are tiny torch ops that stay in
the autograd graph so pose gradients finish through `so3_exp_map` exactly as in the reference.

`so3_exp_map` / `hat` restate pytorch3d (absent and un-pinned in the reference,
environment.yaml:114): Rodrigues with `theta = sqrt(clamp(|w|^2, 1e-4))`.
"""
import torch


def hat(v: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.so3.hat: (N,3) -> (N,3,3) skew-symmetric matrices."""
    x, y, z = v.unbind(1)
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], 1), torch.stack([z, o, -x], 1), torch.stack([-y, x, o], 1)], 1)


def so3_exp_map(log_rot: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    nrms = (log_rot * log_rot).sum(1)
    rot_angles = torch.clamp(nrms, eps).sqrt()
    rot_angles_inv = 1.0 / rot_angles
    fac1 = rot_angles_inv * rot_angles.sin()
    fac2 = rot_angles_inv * rot_angles_inv * (1.0 - rot_angles.cos())
    skews = hat(log_rot)
    skews_square = torch.bmm(skews, skews)
    eye = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]
    return fac1[:, None, None] * skews + fac2[:, None, None] * skews_square + eye


def coords_in_bound(coords: torch.Tensor, bound: torch.Tensor):
    """utils_geometry.py:11-27 -- inclusive test, returns (N,1) bool."""
    inside_min = coords >= bound[:, 0]
    inside_max = coords <= bound[:, 1]
    return (inside_min & inside_max).all(dim=1).unsqueeze(1)


def apply_pose_correction(R, t, R_delta, t_delta):
    """utils_geometry.py:78-99: (R Exp(R_delta), t + t_delta)."""
    assert R.shape == (3, 3)
    assert t.shape == (3, 1)
    assert R_delta.shape == (1, 3)
    assert t_delta.shape == (3, 1)
    return torch.matmul(R, so3_exp_map(R_delta)[0]), t + t_delta


def transform_points_to(points_src, R_dst_src, t_dst_src):
    """utils_geometry.py:214-225."""
    assert R_dst_src.shape == (3, 3)
    assert t_dst_src.shape == (3, 1)
    return points_src @ (R_dst_src.T) + t_dst_src.T


def transfrom_points_from(points_dst, R_dst_src, t_dst_src):
    """utils_geometry.py:227-240 (the reference's spelling is kept)."""
    assert R_dst_src.shape == (3, 3)
    assert t_dst_src.shape == (3, 1)
    R_src_dst = R_dst_src.T
    t_src_dst = -R_dst_src.T @ t_dst_src
    return transform_points_to(points_dst, R_src_dst, t_src_dst)


def pose_matrix(R, t):
    """utils_geometry.py:61-76."""
    assert R.shape == (3, 3)
    assert t.shape == (3, 1)
    pose = torch.eye(4).to(R)
    pose[:3, :3] = R
    pose[:3, [3]] = t
    return pose


def identity_rotations(n: int) -> torch.Tensor:
    return torch.eye(3).unsqueeze(0).repeat(n, 1, 1)
