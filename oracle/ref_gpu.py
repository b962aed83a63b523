"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's GPU path for the interpolation, restated around the
reference's OWN compiled double-backward kernel (oracle/_ref/gridsample_grad2.so, built unmodified from
third_party/cuda_gridsample_grad2/gridsample_cuda.{cpp,cu} by oracle/build_ref.py).

`grid_sample_3d` follows the plugin FeatureGrid.grid_sample_func selects when `second_order_grid_sample` is on
(grid_opt/models/grid_modules.py:63-69 -> third_party/cuda_gridsample_grad2/cuda_gridsample.py:17-19,76-126):
forward = F.grid_sample (ATen grid_sampler_3d), backward = aten::grid_sampler_3d_backward with output_mask,
double backward = the extension's grad2_3d.  It is the kernel-for-kernel bar the CUDA path of this repo is timed
and value-checked against on the B200 box; nothing under miso_b200/ imports it.
"""
import torch
import torch.nn.functional as F

from . import build_ref

_ext = None


def available() -> bool:
    global _ext
    if _ext is None:
        _ext = build_ref.load_module() or False
    return bool(_ext)


def grid_sample_3d(input, grid, padding_mode="zeros", align_corners=True):
    assert padding_mode in ("zeros", "border")
    return _Fwd.apply(input, grid, padding_mode, align_corners)


class _Fwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid, padding_mode, align_corners):
        out = F.grid_sample(input, grid, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)
        ctx.save_for_backward(input, grid)
        ctx.pad, ctx.ac = ("zeros", "border").index(padding_mode), align_corners
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, grid = ctx.saved_tensors
        gi, gg = _Bwd.apply(grad_output, input, grid, ctx.pad, ctx.ac)
        return gi, gg, None, None


class _Bwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, grid, pad, ac):
        mask = (ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        gi, gg = torch.ops.aten.grid_sampler_3d_backward(grad_output, input, grid, 0, pad, ac, mask)
        ctx.save_for_backward(grad_output, input, grid)
        ctx.pad, ctx.ac = pad, ac
        return gi, gg

    @staticmethod
    def backward(ctx, gg_input, gg_grid):
        grad_output, input, grid = ctx.saved_tensors
        assert available(), "oracle/_ref/gridsample_grad2.so missing: run oracle/build_ref.py in the build container"
        if gg_input is None:      # the reference's kernel reads this tensor unconditionally (gridsample_cuda.cu:620-622)
            gg_input = torch.zeros_like(input)
        if gg_grid is None:
            gg_grid = torch.zeros_like(grid)
        out = _ext.grad2_3d(gg_input.contiguous(), gg_grid.contiguous(), grad_output.contiguous(), input, grid,
                            bool(ctx.pad), bool(ctx.ac))
        return out[0], out[1], out[2], None, None


def interp_plugin(feature, xn):
    """FeatureGrid.interpolate body (grid_modules.py:84-95) through the reference's plugin."""
    N = xn.shape[0]
    return grid_sample_3d(feature, xn.reshape(1, N, 1, 1, 3), padding_mode="zeros",
                          align_corners=False)[0, :, :, 0, 0].transpose(0, 1)
