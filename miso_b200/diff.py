"""Mirror of grid_opt/diff.py:14-38 -- `gradient3d(x, f, method, finite_diff_eps, create_graph)`.

'autograd' on a model that exposes `forward_with_gradient` (miso_b200.models.GridNet with a fixed
decoder) returns the analytic gradient the fused kernel produced in the same launch as the SDF;
it is differentiable w.r.t. the grids (this is the eikonal double-backward).  Anything else goes
through torch.autograd exactly like the reference.
"""
import torch


def gradient3d(x, f, method="finitediff", finite_diff_eps=1e-2, create_graph=True):
    assert x.ndim == 2
    assert x.shape[-1] == 3
    if method == "finitediff":
        eps_x = torch.tensor([finite_diff_eps, 0.0, 0.0], device=x.device, dtype=x.dtype)
        eps_y = torch.tensor([0.0, finite_diff_eps, 0.0], device=x.device, dtype=x.dtype)
        eps_z = torch.tensor([0.0, 0.0, finite_diff_eps], device=x.device, dtype=x.dtype)
        grad = torch.cat([f(x + eps_x) - f(x - eps_x),
                          f(x + eps_y) - f(x - eps_y),
                          f(x + eps_z) - f(x - eps_z)], dim=-1)
        grad = grad / (finite_diff_eps * 2.0)
    elif method == "autograd":
        assert x.requires_grad, "requires_grad need to be true for autograd!"
        fused = getattr(f, "forward_with_gradient", None)
        if fused is not None and getattr(f, "fused_spec", lambda: None)() is not None:
            _, grad = fused(x)
            if not create_graph:
                grad = grad.detach()
        else:
            y = f(x)
            grad = torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=create_graph)[0]
    else:
        raise ValueError("Unknown method: {}".format(method))
    return grad
