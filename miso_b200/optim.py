"""Adam over dense grid tensors as one fused pass per tensor (miso_adam_step): reads p, g, m, v and
writes p, m, v -- and zeroes g in the same pass, so the next scatter starts from a clean buffer
without a separate grid-sized memset.  Same update rule and defaults as torch.optim.Adam
(no amsgrad / weight decay), which the reference uses for the grids (grid_opt/trainer.py:422-437,
lr 1e-3 configs/rgbd/scannet.yaml:42).
"""
from typing import Iterable

import torch

from . import _lib


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 zero_grad_in_step=True, track_touched=True, device_step=False):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.zero_grad_in_step = zero_grad_in_step
        self.track_touched = track_touched
        # device_step: the per-tensor step counter lives on the GPU (miso_adam_step_dev), so `step()` enqueues the same
        # launches every time and can be captured in a CUDA graph; state["step"] then mirrors the number of calls
        self.device_step = device_step
        self.state = {}
        self.param_groups = [{"params": self.params, "lr": self.lr}]

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            st = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            if self.track_touched and p.numel() % 4 == 0 and p.data_ptr() % 16 == 0:
                # one bit per 4-float voxel: never-touched voxels are skipped after reading only their gradient
                st["touched"] = torch.zeros((p.numel() // 4 + 31) // 32, dtype=torch.int32, device=p.device)
            if self.device_step:
                st["step_dev"] = torch.zeros(1, dtype=torch.int32, device=p.device)
                st["scalars"] = torch.zeros(3, dtype=torch.float32, device=p.device)
            self.state[p] = st
        return st

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, gate=None):
        """`gate` (device_step only): a device scalar, normally the step's total loss -- a non-finite value skips the
        update on the device (no host sync), as the reference's trainer does on a NaN total (trainer.py:214-217)."""
        lib = _lib.load()
        for p in self.params:
            g = p.grad
            if g is None or not p.requires_grad:
                continue  # torch.optim.Adam skips parameters without a gradient
            _lib.require_cuda(p, g)
            if p.dtype != torch.float32 or g.dtype != torch.float32:
                raise RuntimeError("FusedAdam is float32 only")
            if g.stride() != p.stride():
                g = p.grad = _restride_like(g, p)
            st = self._state(p)
            st["step"] += 1
            m, v = st["exp_avg"], st["exp_avg_sq"]
            if m.stride() != p.stride() or v.stride() != p.stride():
                raise RuntimeError("FusedAdam: optimizer state layout diverged from the parameter layout")
            if self.device_step:
                tracked = "touched" in st and g.data_ptr() % 16 == 0
                with torch.cuda.device(p.device):
                    _lib.check(lib.miso_adam_step_dev(
                        p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                        st["touched"].data_ptr() if tracked else None, p.numel(), self.lr, self.betas[0], self.betas[1],
                        self.eps, st["step_dev"].data_ptr(), st["scalars"].data_ptr(),
                        gate.data_ptr() if gate is not None else None, int(self.zero_grad_in_step),
                        _lib.stream_ptr(p.device)), "adam_step")
                continue
            if "touched" in st and g.data_ptr() % 16 == 0:
                with torch.cuda.device(p.device):
                    _lib.check(lib.miso_adam_step_tracked(
                        p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), st["touched"].data_ptr(), p.numel(),
                        self.lr, self.betas[0], self.betas[1], self.eps, st["step"], int(self.zero_grad_in_step),
                        _lib.stream_ptr(p.device)), "adam_step")
                continue
            with torch.cuda.device(p.device):
                _lib.check(lib.miso_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                              self.lr, self.betas[0], self.betas[1], self.eps, st["step"],
                                              int(self.zero_grad_in_step), _lib.stream_ptr(p.device)), "adam_step")


def _restride_like(g: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(p)
    out.copy_(g)
    return out
