"""Drop-in for the reference's `cuda_gridsample` module
(third_party/cuda_gridsample_grad2/cuda_gridsample.py), the plugin selected by
`FeatureGrid.grid_sample_func` (grid_opt/models/grid_modules.py:63-69).

Same public names and call signature:

    grid_sample_3d(input, grid, padding_mode='zeros', align_corners=True) -> (B,C,Do,Ho,Wo)

twice differentiable w.r.t. `input` and `grid` through the same two-level
`_GridSample3dForward` / `_GridSample3dBackward` autograd.Function structure
(cuda_gridsample.py:76-126), but every level (forward, backward, double backward) runs a
hand-written sm_100a kernel behind the C-ABI in include/miso_b200.h.  The reference asserts
`.is_cuda` only in the double backward (cuda_gridsample.py:118); here all three require CUDA --
there is no CPU path.

Differences that are deliberate:
  * the output's memory layout is point-major ((B,P,C) buffer viewed as (B,C,Do,Ho,Wo)), so
    MISO's `[0,:,:,0,0].transpose(0,1)` (grid_modules.py:94) is a contiguous (N,C) tensor;
  * `grad_input` keeps the memory format of `input` (channels_last_3d stays channels_last_3d).
"""
import torch

from . import _lib

__all__ = ["grid_sample_3d", "grid_sample_2d"]


def grid_sample_2d(input, grid, padding_mode="zeros", align_corners=True):
    # shipped MISO configs are spatial_dim: 3 (configs/rgbd/scannet.yaml:9); SURVEY.md marks 2D out of scope
    raise NotImplementedError("miso_b200 implements the 3D path only (MISO configs use spatial_dim: 3)")


def grid_sample_3d(input, grid, padding_mode="zeros", align_corners=True):
    assert padding_mode in _lib.PAD_MODES  # same assertion as cuda_gridsample.py:18
    return _GridSample3dForward.apply(input, grid, padding_mode, align_corners)


def _dtype_code(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise RuntimeError(f"miso_b200.grid_sample_3d supports float32/float64, got {t.dtype}")


def _check(input, grid):
    assert input.ndim == 5
    assert grid.ndim == 5
    assert input.shape[0] == grid.shape[0]
    assert grid.shape[4] == 3
    _lib.require_cuda(input, grid)
    if input.dtype != grid.dtype:
        raise RuntimeError("grid_sample_3d: input and grid must have the same dtype")
    if input.device != grid.device:
        raise RuntimeError("grid_sample_3d: input and grid must be on the same device")


def _out_strides(t_bpc):
    """(B,P,C) buffer -> strides in the (sB, sC, sP) order the ABI expects."""
    sB, sP, sC = t_bpc.stride()
    return _lib.i64c([sB, sC, sP])


class _GridSample3dForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid, padding_mode=0, align_corners=True):
        _check(input, grid)
        lib = _lib.load()
        B, Cc = input.shape[0], input.shape[1]
        Do, Ho, Wo = grid.shape[1:4]
        P = Do * Ho * Wo
        g = grid.detach().contiguous()
        inp = input.detach()
        out = torch.empty((B, P, Cc), dtype=input.dtype, device=input.device)
        with _lib.on_device(input.device):
            if P > 0:
                _lib.check(lib.miso_grid_sample3d_fwd(
                    _dtype_code(input), inp.data_ptr(), _lib.i64c(inp.shape), _lib.i64c(inp.stride()), g.data_ptr(), P,
                    out.data_ptr(), _out_strides(out), _lib.PAD_MODES.index(padding_mode), int(bool(align_corners)),
                    _lib.stream_ptr(input.device)), "grid_sample3d_fwd")
        ctx.save_for_backward(input, grid)
        ctx.padding_mode = _lib.PAD_MODES.index(padding_mode)
        ctx.align_corners = bool(align_corners)
        return out.permute(0, 2, 1).unflatten(2, (Do, Ho, Wo))

    @staticmethod
    def backward(ctx, grad_output):
        input, grid = ctx.saved_tensors
        grad_input, grad_grid = _GridSample3dBackward.apply(grad_output, input, grid, ctx.padding_mode,
                                                            ctx.align_corners)
        return grad_input, grad_grid, None, None


def _as_bcp(t5):
    """(B,C,Do,Ho,Wo) -> ((B,C,P) view-or-copy, ABI strides (sB,sC,sP)).  `reshape` only copies when
    the three spatial dims cannot be collapsed into one stride."""
    t = t5.reshape(t5.shape[0], t5.shape[1], -1) if t5.numel() > 0 else t5.new_zeros((t5.shape[0], t5.shape[1], 0))
    return t, _lib.i64c(t.stride())


class _GridSample3dBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, grid, padding_mode=0, align_corners=True):
        lib = _lib.load()
        _lib.require_cuda(grad_output, input, grid)
        need_input, need_grid = ctx.needs_input_grad[1], ctx.needs_input_grad[2]  # output_mask, cuda_gridsample.py:104
        Do, Ho, Wo = grid.shape[1:4]
        P = Do * Ho * Wo
        g = grid.detach().contiguous()
        inp = input.detach()
        go, go_st = _as_bcp(grad_output.detach())
        grad_input = torch.zeros_like(inp) if need_input else None  # preserve_format keeps channels_last_3d
        grad_grid = torch.empty_like(g) if need_grid else None
        with _lib.on_device(input.device):
            if P > 0:
              _lib.check(lib.miso_grid_sample3d_bwd(
                _dtype_code(input), go.data_ptr(), go_st, inp.data_ptr(), _lib.i64c(inp.shape), _lib.i64c(inp.stride()),
                g.data_ptr(), P, _lib.ptr(grad_input),
                _lib.i64c(grad_input.stride()) if grad_input is not None else None, _lib.ptr(grad_grid),
                padding_mode, int(align_corners),
                _lib.stream_ptr(input.device)), "grid_sample3d_bwd")
        ctx.save_for_backward(grad_output, input, grid)
        ctx.padding_mode = padding_mode
        ctx.align_corners = align_corners
        ctx.set_materialize_grads(False)  # an undefined grad2_grad_input stays None instead of a grid-sized zeros
        return grad_input, grad_grid

    @staticmethod
    def backward(ctx, grad2_grad_input, grad2_grad_grid):
        grad_output, input, grid = ctx.saved_tensors
        lib = _lib.load()
        _lib.require_cuda(grad_output, input, grid, grad2_grad_input, grad2_grad_grid)
        need_go, need_input, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if grad2_grad_input is None and grad2_grad_grid is None:
            return None, None, None, None, None
        B, Cc = input.shape[0], input.shape[1]
        Do, Ho, Wo = grid.shape[1:4]
        P = Do * Ho * Wo
        g = grid.detach().contiguous()
        inp = input.detach()
        go, go_st = _as_bcp(grad_output.detach())
        ggi = grad2_grad_input.detach() if grad2_grad_input is not None else None
        ggg = grad2_grad_grid.detach().contiguous() if grad2_grad_grid is not None else None
        gg_out = torch.empty((B, P, Cc), dtype=inp.dtype, device=inp.device) if need_go else None
        g_input = torch.zeros_like(inp) if need_input else None
        g_grid = torch.empty_like(g) if need_grid else None
        with _lib.on_device(input.device):
            if P > 0:
              _lib.check(lib.miso_grid_sample3d_bwd_bwd(
                _dtype_code(input), _lib.ptr(ggi), _lib.i64c(ggi.stride()) if ggi is not None else None, _lib.ptr(ggg),
                go.data_ptr(), go_st, inp.data_ptr(), _lib.i64c(inp.shape), _lib.i64c(inp.stride()), g.data_ptr(), P,
                _lib.ptr(gg_out), _out_strides(gg_out) if gg_out is not None else None, _lib.ptr(g_input),
                _lib.i64c(g_input.stride()) if g_input is not None else None, _lib.ptr(g_grid), ctx.padding_mode, int(ctx.align_corners), _lib.stream_ptr(input.device)),
                "grid_sample3d_bwd_bwd")
        if gg_out is not None:
            gg_out = gg_out.permute(0, 2, 1).unflatten(2, (Do, Ho, Wo))
        return gg_out, g_input, g_grid, None, None
