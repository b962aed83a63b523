"""`gradient3d(x, f, method, finite_diff_eps, create_graph)` with the reference's signature (grid_opt/diff.py:14-38).

'autograd' on a model that exposes `forward_with_gradient` and a fused spec (miso_b200.models.GridNet with a frozen
decoder) returns the analytic gradient the fused kernel produced in the same launch as the SDF; that output is
differentiable w.r.t. the levels -- it IS the eikonal double-backward.  Any other callable goes through
torch.autograd.  'finitediff' is the central difference over the three axes; all six evaluations stay in the
autograd graph (first-order backward passes), as in the reference.
"""
import torch


def _central_differences(x, f, eps):
    """(f(x + eps e_d) - f(x - eps e_d)) / (2 eps) for d = x, y, z -> (N,3)."""
    steps = torch.eye(3, device=x.device, dtype=x.dtype) * eps
    cols = [f(x + steps[d]) - f(x - steps[d]) for d in range(3)]
    return torch.cat(cols, dim=-1) / (eps * 2.0)


def gradient3d(x, f, method="finitediff", finite_diff_eps=1e-2, create_graph=True):
    if x.ndim != 2 or x.shape[-1] != 3:
        raise AssertionError(f"gradient3d expects (N,3) points, got {tuple(x.shape)}")
    if method == "finitediff":
        return _central_differences(x, f, finite_diff_eps)
    if method != "autograd":
        raise ValueError("Unknown method: {}".format(method))
    assert x.requires_grad, "requires_grad need to be true for autograd!"
    if getattr(f, "forward_with_gradient", None) is not None and getattr(f, "fused_spec", lambda: None)() is not None:
        grad = f.forward_with_gradient(x)[1]
        return grad if create_graph else grad.detach()
    y = f(x)
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=create_graph)[0]
