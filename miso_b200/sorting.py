"""Morton (Z-order) ordering of a point batch for L2 / L1 locality of the corner gathers and the
gradient scatter.  Keys come from the library's `miso_morton_keys` kernel on the world-frame coordinates
(`miso_transform_points`); the sort itself is torch.sort (CUB radix sort) -- plumbing, not the hot path.
The permutation is internal: callers that need results in the original order scatter them back."""
import ctypes as C

import torch

from . import _lib


def morton_keys(x_world: torch.Tensor, bound) -> torch.Tensor:
    lib = _lib.load()
    _lib.require_cuda(x_world)
    x = x_world.detach().contiguous().float()
    keys = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    b = (C.c_float * 6)(*[float(v) for v in bound])
    with torch.cuda.device(x.device):
        _lib.check(lib.miso_morton_keys(x.data_ptr(), x.shape[0], b, keys.data_ptr(), _lib.stream_ptr(x.device)),
                   "morton_keys")
    return keys


def transform_points(x: torch.Tensor, ids: torch.Tensor, R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """x_world = R[id] x + t[id] in one launch (loss.py:764-774 without the per-keyframe loop)."""
    lib = _lib.load()
    _lib.require_cuda(x, ids, R, t)
    x = x.detach().contiguous().float()
    ids = ids.detach().reshape(-1).contiguous().long()
    R = R.detach().reshape(-1, 3, 3).contiguous().float()
    t = t.detach().reshape(-1, 3).contiguous().float()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.miso_transform_points(x.data_ptr(), ids.data_ptr(), R.data_ptr(), t.data_ptr(), R.shape[0],
                                             x.shape[0], y.data_ptr(), _lib.stream_ptr(x.device)), "transform_points")
    return y


def morton_order(coords_frame, frame_ids, R, t, bound) -> torch.Tensor:
    """Permutation that sorts the batch by the Morton key of its world-frame position."""
    xw = transform_points(coords_frame, frame_ids, R, t)
    keys = morton_keys(xw, bound).to(torch.int64) & 0xFFFFFFFF
    return torch.argsort(keys)
