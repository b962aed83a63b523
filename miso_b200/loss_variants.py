"""Eikonal / SDF loss variants the reference's other trainers put on top of the same hot path
(SURVEY.md section 8, row a9/a10).  Same names and argument meaning as the reference:

    TsdfLoss3D                     grid_opt/loss.py:71-144  (sdf + sign terms; eikonal on N uniform random points in the
                                                              model bound drawn with np.random.uniform, weight 50)
    full_sdf_loss / sdf_loss / tot_loss   grid_opt/loss_isdf.py:280-365 (iSDF: free-space + truncation terms, per-sample
                                                              abs(|grad|-1) zeroed where bounds < eik_apply_dist)
    isdf_loss                      grid_opt/loss_isdf.py:93-152 compute_default without the surface-normal term

Every SDF value and spatial gradient comes from the fused grid+decoder kernel (GridNet.forward /
forward_with_gradient: value and analytic gradient in ONE launch, the gradient output carrying the eikonal
double-backward); what is left here are elementwise epilogues on device tensors."""
import numpy as np
import torch

from .diff import gradient3d


class TsdfLoss3D:
    """loss.py:71-144."""

    def __init__(self, sdf_weight=3e3, sign_weight=1e2, eik_weight=5e1, trunc_dist=0.15, grad_method="autograd",
                 finite_diff_eps=1e-2):
        self.sdf_weight = sdf_weight
        self.sign_weight = sign_weight
        self.eik_weight = eik_weight
        self.trunc_dist = trunc_dist
        self.grad_method = grad_method
        self.finite_diff_eps = finite_diff_eps

    @staticmethod
    def _hinge_where(mask, value):
        """mean(max(0, value where mask else 0)) -- the free-space hinge of :105-127."""
        z = torch.zeros_like(value)
        return torch.mean(torch.maximum(z, torch.where(mask, value, z)))

    def _uniform_points_in_bound(self, model, n, like):
        # same draw order as the reference (:128-133: all x, then all y, then all z from numpy's global RNG) so a seeded
        # RNG reproduces its points
        b = model.bound.detach().cpu().numpy()
        cols = [np.random.uniform(b[d, 0], b[d, 1], n).reshape(n, 1) for d in range(3)]
        return torch.from_numpy(np.concatenate(cols, axis=1)).to(like)

    def compute(self, model, model_input, gt):
        coords, target = model_input["coords"][0], gt["sdf"][0]
        valid, sign = gt["sdf_valid"][0], gt["sdf_sign"][0]
        assert coords.ndim == 2 and target.ndim == 2
        pred = model(coords)                                             # fused grid + decoder launch
        err = torch.where(valid == 1, pred - target, torch.zeros_like(pred))
        terms = {"sdf": self.sdf_weight * torch.mean(err ** 2)}
        if self.sign_weight > 0:
            assert self.trunc_dist is not None
            terms["pos_space"] = self.sign_weight * self._hinge_where(sign == 1, self.trunc_dist - pred)
            terms["neg_space"] = self.sign_weight * self._hinge_where(sign == -1, pred + self.trunc_dist)
        if self.eik_weight > 0:
            x = self._uniform_points_in_bound(model, target.shape[0], target).requires_grad_(True)
            g = gradient3d(x, model, method=self.grad_method, finite_diff_eps=self.finite_diff_eps, create_graph=True)
            terms["eik"] = self.eik_weight * torch.mean((g.norm(dim=-1) - 1) ** 2)
        return terms


def full_sdf_loss(sdf, target_sdf, free_space_factor=5.0):
    """loss_isdf.py:280-296."""
    free_space_loss_mat = torch.max(torch.nn.functional.relu(sdf - target_sdf),
                                    torch.exp(-free_space_factor * sdf) - 1.0)
    trunc_loss_mat = sdf - target_sdf
    return free_space_loss_mat, trunc_loss_mat


def sdf_loss(sdf, bounds, t, loss_type="L1", p75=0.05):
    """loss_isdf.py:299-333 (the deprecated GM branch raises there too)."""
    free_space_loss_mat, trunc_loss_mat = full_sdf_loss(sdf, bounds)
    free_space_ixs = bounds > t
    free_space_loss_mat = torch.where(free_space_ixs, free_space_loss_mat, torch.zeros_like(free_space_loss_mat))
    trunc_loss_mat = torch.where(free_space_ixs, torch.zeros_like(trunc_loss_mat), trunc_loss_mat)
    sdf_loss_mat = free_space_loss_mat + trunc_loss_mat
    if loss_type == "L1":
        sdf_loss_mat = torch.abs(sdf_loss_mat)
    elif loss_type == "L2":
        sdf_loss_mat = torch.square(sdf_loss_mat)
    elif loss_type == "GM":
        raise ValueError("GM loss is deprecated.")
    else:
        raise ValueError("Must be L1 or L2")
    return sdf_loss_mat, free_space_ixs


def tot_loss(sdf_loss_mat, grad_loss_mat, eik_loss_mat, free_space_ixs, bounds, eik_apply_dist, trunc_weight,
             grad_weight, eik_weight):
    """loss_isdf.py:335-365.  Returns (tot_loss, tot_loss_mat, losses); the per-term entries of `losses` stay
    device tensors (the reference calls .item() on each: three host syncs per step)."""
    sdf_loss_mat = torch.where(free_space_ixs, sdf_loss_mat, sdf_loss_mat * trunc_weight)
    losses = {"sdf_loss": sdf_loss_mat.mean().detach()}
    tot_loss_mat = sdf_loss_mat
    if grad_loss_mat is not None:
        tot_loss_mat = tot_loss_mat + grad_weight * grad_loss_mat
        losses["grad_loss"] = grad_loss_mat.mean().detach()
    if eik_loss_mat is not None:
        eik_loss_mat = torch.where(bounds.squeeze(-1) < eik_apply_dist, torch.zeros_like(eik_loss_mat), eik_loss_mat)
        eik_loss_mat = eik_loss_mat * eik_weight
        tot_loss_mat = tot_loss_mat + eik_loss_mat
        losses["eikonal_loss"] = eik_loss_mat.mean().detach()
    tot = tot_loss_mat.mean()
    losses["total_loss"] = tot
    return tot, tot_loss_mat, losses


def isdf_loss(model, pc, bounds, trunc_distance, trunc_weight, eik_weight, eik_apply_dist, loss_type="L1",
              grad_method="autograd", finite_diff_eps=1e-2):
    """iSDFLoss.compute_default (loss_isdf.py:93-152) without the surface-normal term: `pc` (N,3) points,
    `bounds` (N,1) upper bounds on |sdf|.  With grad_method='autograd' the value and its gradient come from one
    fused launch.  Returns (total_loss, losses)."""
    assert pc.ndim == 2 and bounds.shape == (pc.shape[0], 1)
    eik_loss_mat = None
    if eik_weight != 0:
        if grad_method == "autograd" and getattr(model, "forward_with_gradient", None) is not None:
            x = pc if pc.requires_grad else pc.clone().requires_grad_(True)
            sdf, g = model.forward_with_gradient(x)
        else:
            x = pc.clone().requires_grad_(True) if grad_method == "autograd" else pc
            sdf = model(x)
            g = gradient3d(x, model, method=grad_method, finite_diff_eps=finite_diff_eps, create_graph=True)
        eik_loss_mat = torch.abs(g.norm(2, dim=-1) - 1).unsqueeze(-1)
    else:
        sdf = model(pc)
    sdf_loss_mat, free_space_ixs = sdf_loss(sdf, bounds, trunc_distance, loss_type=loss_type)
    if eik_loss_mat is not None:
        eik_loss_mat = eik_loss_mat.squeeze(-1)
        # the reference's tensors are (1,N,1)/(1,N): per-sample addition; here (N,1)/(N,)
        total, _, losses = tot_loss(sdf_loss_mat.squeeze(-1), None, eik_loss_mat, free_space_ixs.squeeze(-1), bounds,
                                    eik_apply_dist, trunc_weight, 0.0, eik_weight)
    else:
        total, _, losses = tot_loss(sdf_loss_mat, None, None, free_space_ixs, bounds, eik_apply_dist, trunc_weight,
                                    0.0, eik_weight)
    return total, losses
