"""Host side of the fused multiresolution field (grid levels + ReLU MLP decoder).

Builds the plain-C descriptors of include/miso_b200.h from torch tensors and exposes the fused
kernels as `torch.autograd.Function`s, so the reference's call signatures
(`GridNet.forward`, `GridNet.query_feature`, `gradient3d`; grid_opt/models/grid_net.py:288-325,
grid_opt/diff.py:14-38) keep working with autograd, including `create_graph=True` (the eikonal
double-backward).
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib


def to_channels_last_3d_(param: torch.Tensor) -> torch.Tensor:
    """Re-lay a `(1,C,Z,Y,X)` grid parameter as channels_last_3d IN PLACE (logical shape, state-dict
    key and values unchanged).  This is the layout the fused kernels need for 128-bit corner loads."""
    with torch.no_grad():
        param.data = param.data.contiguous(memory_format=torch.channels_last_3d)
        if param.grad is not None:
            param.grad.data = param.grad.data.contiguous(memory_format=torch.channels_last_3d)
    return param


def _level(feat: torch.Tensor, grad: Optional[torch.Tensor]) -> _lib.Level:
    if feat.ndim != 5 or feat.shape[0] != 1:
        raise RuntimeError(f"grid level must be (1,C,Z,Y,X), got {tuple(feat.shape)}")
    if feat.dtype != torch.float32:
        raise RuntimeError("fused kernels are float32 (configs.py:80)")
    _, Cc, Z, Y, X = feat.shape
    _, sC, sZ, sY, sX = feat.stride()
    if grad is not None and (grad.shape != feat.shape or grad.stride() != feat.stride()):
        raise RuntimeError("grid gradient buffer must have the shape and strides of the grid")
    lv = _lib.Level()
    lv.feat = feat.data_ptr()
    lv.grad = grad.data_ptr() if grad is not None else None
    lv.X, lv.Y, lv.Z, lv.C = X, Y, Z, Cc
    lv.sC, lv.sZ, lv.sY, lv.sX = sC, sZ, sY, sX
    return lv


def make_field(feats: Sequence[torch.Tensor], bound: Sequence[float], grads: Optional[Sequence] = None,
               ignore_mask: int = 0) -> _lib.Field:
    if not 1 <= len(feats) <= _lib.MISO_MAX_LEVELS:
        raise RuntimeError(f"1..{_lib.MISO_MAX_LEVELS} levels supported, got {len(feats)}")
    f = _lib.Field()
    f.num_levels = len(feats)
    f.ignore_mask = int(ignore_mask)
    for i, b in enumerate(bound):
        f.bound[i] = float(b)
    for l, feat in enumerate(feats):
        f.level[l] = _level(feat.detach(), grads[l] if grads is not None else None)
    return f


def structs_to_device(structs, device) -> torch.Tensor:
    """Pack ctypes structs (miso_field_t, miso_align_pair_t ...) into one device byte buffer."""
    raw = b"".join(bytes(s) for s in structs)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def bound_to_list(bound) -> List[float]:
    """(3,2) tensor / array / nested list -> [xmin,xmax,ymin,ymax,zmin,zmax] python floats.
    Done once at module construction: reading a CUDA tensor here would be a host sync."""
    if isinstance(bound, torch.Tensor):
        bound = bound.detach().cpu().tolist()
    out = [float(v) for row in bound for v in row]
    assert len(out) == 6
    return out


class DecoderSpec:
    """Weights of MLPNet(F, 1, hidden_dim=64, hidden_layers=1, bias=True) (modules.py:11-32)."""

    def __init__(self, W1, b1, W2, b2, W3, b3):
        self.tensors = [t.detach() for t in (W1, b1, W2, b2, W3, b3)]
        for t in self.tensors:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("decoder weights must be contiguous float32")
        self.in_dim = W1.shape[1]
        self.hidden = W1.shape[0]
        if W2.shape != (self.hidden, self.hidden) or W3.shape != (1, self.hidden):
            raise RuntimeError("fused decoder must be Linear(F,H)+ReLU+Linear(H,H)+ReLU+Linear(H,1)")

    @staticmethod
    def from_mlp(mlp) -> "DecoderSpec":
        linears = [m for m in mlp.network if isinstance(m, torch.nn.Linear)]
        relus = [m for m in mlp.network if isinstance(m, torch.nn.ReLU)]
        if len(linears) != 3 or len(relus) != 2 or any(l.bias is None for l in linears):
            raise RuntimeError("fused path supports MLPNet(hidden_layers=1, bias=True, ReLU) only")
        return DecoderSpec(linears[0].weight, linears[0].bias, linears[1].weight, linears[1].bias,
                           linears[2].weight, linears[2].bias)

    def struct(self) -> _lib.Decoder:
        d = _lib.Decoder()
        d.W1, d.b1, d.W2, d.b2, d.W3, d.b3 = [t.data_ptr() for t in self.tensors]
        d.in_dim, d.hidden_dim = self.in_dim, self.hidden
        return d


def decoder_parameters(mlp) -> list:
    """[W1, b1, W2, b2, W3, b3] of an MLPNet, the live parameters (order of miso_decoder_t / miso_decoder_grad_t)."""
    linears = [m for m in mlp.network if isinstance(m, torch.nn.Linear)]
    return [t for l in linears for t in (l.weight, l.bias)]


class FramesSpec:
    """Per-sample keyframe ids + per-keyframe poses (loss.py:764-774), all on the device."""

    def __init__(self, ids: torch.Tensor, R: torch.Tensor, t: torch.Tensor):
        self.ids = ids.detach().reshape(-1).contiguous()
        if self.ids.dtype != torch.int64:
            self.ids = self.ids.long()
        self.R = R.detach().reshape(-1, 3, 3).contiguous().float()
        self.t = t.detach().reshape(-1, 3).contiguous().float()
        if self.R.shape[0] != self.t.shape[0]:
            raise RuntimeError("frames: R and t disagree on the number of poses")

    def struct(self) -> _lib.Frames:
        f = _lib.Frames()
        f.ids, f.R, f.t, f.num_frames = self.ids.data_ptr(), self.R.data_ptr(), self.t.data_ptr(), self.R.shape[0]
        return f


class FieldSpec:
    """Everything non-differentiable the fused Functions need besides the level tensors."""

    def __init__(self, bound: Sequence[float], decoder: Optional[DecoderSpec], ignore_mask: int = 0):
        self.bound = list(bound)
        self.decoder = decoder
        self.ignore_mask = int(ignore_mask)


def _prep_x(x: torch.Tensor) -> torch.Tensor:
    if x.ndim != 2 or x.shape[-1] != 3:
        raise AssertionError(f"Invalid input coords shape {tuple(x.shape)}!")  # grid_net.py:289-290
    _lib.require_cuda(x)
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


# ------------------------------------------------------------------------------------------------
# features only: utils.grid_interp_regular
# ------------------------------------------------------------------------------------------------
def field_features_raw(feats, bound, x, ignore_mask=0) -> torch.Tensor:
    lib = _lib.load()
    x = _prep_x(x)
    N = x.shape[0]
    Ftot = sum(f.shape[1] for f in feats)
    out = torch.empty((N, Ftot), dtype=torch.float32, device=x.device)
    fld = make_field(feats, bound, None, ignore_mask)
    with torch.cuda.device(x.device):
        _lib.check(lib.miso_field_features(C.byref(fld), x.data_ptr(), N, out.data_ptr(), _lib.stream_ptr(x.device)),
                   "field_features")
    return out


# ------------------------------------------------------------------------------------------------
# fused sdf (+ analytic spatial gradient) with first- and second-order autograd
# ------------------------------------------------------------------------------------------------
def sdf_forward_raw(feats, spec: FieldSpec, x, frames: Optional[FramesSpec] = None, want_jac=True,
                    want_gradx=True, want_xw=False):
    """One launch: returns (sdf (N,), jac (N,F)|None, gradx (N,3)|None, xw (N,3)|None)."""
    lib = _lib.load()
    x = _prep_x(x)
    N = x.shape[0]
    Ftot = sum(f.shape[1] for f in feats)
    dev = x.device
    sdf = torch.empty((N,), dtype=torch.float32, device=dev)
    jac = torch.empty((N, Ftot), dtype=torch.float32, device=dev) if want_jac else None
    gradx = torch.empty((N, 3), dtype=torch.float32, device=dev) if want_gradx else None
    xw = torch.empty((N, 3), dtype=torch.float32, device=dev) if want_xw else None
    fld = make_field(feats, spec.bound, None, spec.ignore_mask)
    dec = spec.decoder.struct()
    fr = frames.struct() if frames is not None else None
    if frames is not None and frames.ids.shape[0] != N:
        raise RuntimeError("frames.ids must have one entry per sample")
    with torch.cuda.device(dev):
        _lib.check(lib.miso_sdf_forward(C.byref(fld), C.byref(dec), C.byref(fr) if fr is not None else None,
                                        x.data_ptr(), N, sdf.data_ptr(), _lib.ptr(jac), _lib.ptr(gradx), _lib.ptr(xw),
                                        _lib.stream_ptr(dev)), "sdf_forward")
    return sdf, jac, gradx, xw


def sdf_backward_raw(feats, grads, spec: FieldSpec, xw, jac, a, v, want_hv=False):
    """One launch: scatter (a*w + v.dw)*J into `grads` (accumulating); optionally returns hv (N,3)."""
    lib = _lib.load()
    N = xw.shape[0]
    hv = torch.empty((N, 3), dtype=torch.float32, device=xw.device) if want_hv else None
    fld = make_field(feats, spec.bound, grads, spec.ignore_mask)
    a = a.detach().reshape(-1).contiguous().float() if a is not None else None
    v = v.detach().reshape(-1, 3).contiguous().float() if v is not None else None
    with torch.cuda.device(xw.device):
        _lib.check(lib.miso_sdf_backward(C.byref(fld), xw.data_ptr(), N, jac.data_ptr(), _lib.ptr(a), _lib.ptr(v),
                                         _lib.ptr(hv), _lib.stream_ptr(xw.device)), "sdf_backward")
    return hv


class _FusedSDF(torch.autograd.Function):
    """(x, *level tensors) -> (sdf (N,1), grad_x sdf (N,3)).  Decoder fixed (cfg decoder.fix: True)."""

    @staticmethod
    def forward(ctx, x, spec, *feats):
        sdf, jac, gradx, _ = sdf_forward_raw(feats, spec, x)
        ctx.spec = spec
        ctx.save_for_backward(x, jac, gradx, *feats)
        ctx.set_materialize_grads(False)
        return sdf.unsqueeze(1), gradx

    @staticmethod
    def backward(ctx, g_sdf, g_gradx):
        x, jac, gradx, *feats = ctx.saved_tensors
        if g_sdf is None and g_gradx is None:
            return (None, None) + (None,) * len(feats)
        outs = _FusedSDFBackward.apply(g_sdf, g_gradx, x, jac, gradx, ctx.spec, *feats)
        return (outs[0], None) + tuple(outs[1:])


class _FusedSDFBackward(torch.autograd.Function):
    """First backward as a differentiable op: (a, v, x, ...) -> (dL/dx, *dL/dlevel)."""

    @staticmethod
    def forward(ctx, a, v, x, jac, gradx, spec, *feats):
        need_x = ctx.needs_input_grad[2]
        need_feat = [ctx.needs_input_grad[6 + l] for l in range(len(feats))]
        xd = _prep_x(x)
        grads = [torch.zeros_like(f) if nf else None for f, nf in zip(feats, need_feat)]
        hv = None
        if any(need_feat) or (need_x and v is not None):
            hv = sdf_backward_raw(feats, grads, spec, xd, jac, a, v, want_hv=need_x and v is not None)
        gx = None
        if need_x:
            gx = a.reshape(-1, 1) * gradx if a is not None else torch.zeros_like(gradx)
            if hv is not None:
                gx = gx + hv
        ctx.spec = spec
        ctx.has_v = v is not None
        ctx.save_for_backward(a, x, jac, gradx, *feats)
        ctx.set_materialize_grads(False)
        return (gx,) + tuple(grads)

    @staticmethod
    def backward(ctx, ggx, *ggfeats):
        a, x, jac, gradx, *feats = ctx.saved_tensors
        nf = len(feats)
        if any(g is not None for g in ggfeats):
            raise NotImplementedError("miso_b200 fused field: differentiating the grid gradients again is not "
                                      "supported; use miso_b200.cuda_gridsample.grid_sample_3d (generic path)")
        if ggx is None:
            return (None,) * (6 + nf)
        need_a, need_v, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_feat = [ctx.needs_input_grad[6 + l] for l in range(nf)]
        if ctx.has_v and (any(need_feat) or need_x or need_v):
            raise NotImplementedError("miso_b200 fused field: third-order terms (a cotangent on grad_x sdf inside "
                                      "a create_graph backward) are not supported; use the generic path")
        xd = _prep_x(x)
        ga = (ggx * gradx).sum(dim=1, keepdim=True) if need_a else None
        if ga is not None and a is not None:
            ga = ga.reshape(a.shape)
        grads = [torch.zeros_like(f) if n else None for f, n in zip(feats, need_feat)]
        gx = None
        if (any(need_feat) or need_x) and a is not None:
            v2 = a.reshape(-1, 1) * ggx
            gx = sdf_backward_raw(feats, grads, ctx.spec, xd, jac, None, v2, want_hv=need_x)
        return (ga, None, gx, None, None, None) + tuple(grads)


def fused_sdf(x: torch.Tensor, feats: Sequence[torch.Tensor], spec: FieldSpec):
    """Differentiable fused evaluation: returns (sdf (N,1), grad_x sdf (N,3))."""
    if spec.decoder is None:
        raise RuntimeError("fused_sdf needs a decoder")
    return _FusedSDF.apply(x, spec, *feats)


class _FieldFeatures(torch.autograd.Function):
    """utils.grid_interp_regular as one launch; first-order autograd wrt the level tensors."""

    @staticmethod
    def forward(ctx, x, spec, *feats):
        out = field_features_raw(feats, spec.bound, x, spec.ignore_mask)
        ctx.spec = spec
        ctx.save_for_backward(x, *feats)
        return out

    @staticmethod
    def backward(ctx, g):
        # Route through the generic, twice-differentiable per-level op so any order of autograd works.
        raise NotImplementedError("use miso_b200.models.FeatureGrid.interpolate for differentiable features")


def field_features(x, feats, spec: FieldSpec) -> torch.Tensor:
    """Non-differentiable fast path of GridNet.query_feature (used under torch.no_grad)."""
    return field_features_raw(feats, spec.bound, x, spec.ignore_mask)
