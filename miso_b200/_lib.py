"""ctypes binding of libmiso_b200.so (include/miso_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a RuntimeError is
raised.  Nothing here imports `oracle/`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmiso_b200.so")

MISO_MAX_LEVELS = 4
MISO_ALIGN_OUT = 72
F32, F64 = 0, 1
PAD_MODES = ["zeros", "border"]  # same index convention as the reference (cuda_gridsample.py:87)

c_i64p = C.POINTER(C.c_int64)


class Level(C.Structure):
    _fields_ = [("feat", C.c_void_p), ("grad", C.c_void_p),
                ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32), ("C", C.c_int32),
                ("sC", C.c_int64), ("sZ", C.c_int64), ("sY", C.c_int64), ("sX", C.c_int64)]


class Field(C.Structure):
    _fields_ = [("num_levels", C.c_int32), ("ignore_mask", C.c_uint32), ("bound", C.c_float * 6),
                ("level", Level * MISO_MAX_LEVELS)]


class Decoder(C.Structure):
    _fields_ = [("W1", C.c_void_p), ("b1", C.c_void_p), ("W2", C.c_void_p), ("b2", C.c_void_p),
                ("W3", C.c_void_p), ("b3", C.c_void_p), ("in_dim", C.c_int32), ("hidden_dim", C.c_int32)]


class DecoderGrad(C.Structure):
    _fields_ = [("W1", C.c_void_p), ("b1", C.c_void_p), ("W2", C.c_void_p), ("b2", C.c_void_p),
                ("W3", C.c_void_p), ("b3", C.c_void_p)]


class Frames(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("R", C.c_void_p), ("t", C.c_void_p), ("num_frames", C.c_int32)]


class MappingCfg(C.Structure):
    _fields_ = [("loss_type", C.c_int32), ("weight_sdf", C.c_float), ("weight_fs", C.c_float),
                ("weight_eik", C.c_float), ("trunc_dist", C.c_float), ("eik_trunc_dist", C.c_float),
                ("eik_mode", C.c_int32), ("grad_scale", C.c_float), ("n_total", C.c_int64), ("n_device", C.c_void_p)]


class AlignPair(C.Structure):
    _fields_ = [("src", C.c_int32), ("dst", C.c_int32), ("levels_used", C.c_int32), ("reserved", C.c_int32),
                ("p", C.c_void_p), ("M", C.c_int64), ("fsrc", C.c_void_p), ("mask_out", C.c_void_p),
                ("enabled", C.c_void_p), ("src_grad_scale", C.c_float), ("dst_grad_scale", C.c_float)]


_SIGNATURES = {
    "miso_last_error_string": (C.c_char_p, []),
    "miso_abi_version": (C.c_int, []),
    "miso_device_sm_count": (C.c_int, []),
    "miso_grid_sample3d_fwd": (C.c_int, [C.c_int, C.c_void_p, c_i64p, c_i64p, C.c_void_p, C.c_int64, C.c_void_p,
                                         c_i64p, C.c_int, C.c_int, C.c_void_p]),
    "miso_grid_sample3d_bwd": (C.c_int, [C.c_int, C.c_void_p, c_i64p, C.c_void_p, c_i64p, c_i64p, C.c_void_p,
                                         C.c_int64, C.c_void_p, c_i64p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "miso_grid_sample3d_bwd_bwd": (C.c_int, [C.c_int, C.c_void_p, c_i64p, C.c_void_p, C.c_void_p, c_i64p,
                                             C.c_void_p, c_i64p, c_i64p, C.c_void_p, C.c_int64, C.c_void_p, c_i64p,
                                             C.c_void_p, c_i64p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "miso_field_features": (C.c_int, [C.POINTER(Field), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "miso_sdf_forward": (C.c_int, [C.POINTER(Field), C.POINTER(Decoder), C.POINTER(Frames), C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_sdf_backward": (C.c_int, [C.POINTER(Field), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "miso_mapping_workspace_floats": (C.c_int64, []),
    "miso_mapping_count": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "miso_mapping_step": (C.c_int, [C.POINTER(Field), C.POINTER(Decoder), C.POINTER(Frames), C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MappingCfg),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_mapping_step_fd": (C.c_int, [C.POINTER(Field), C.POINTER(Decoder), C.POINTER(Frames), C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MappingCfg),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "miso_mapping_wgrad_workspace_floats": (C.c_int64, []),
    "miso_mapping_step_wgrad": (C.c_int, [C.POINTER(Field), C.POINTER(Decoder), C.POINTER(Frames), C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MappingCfg),
                                          C.c_void_p, C.POINTER(DecoderGrad), C.c_void_p, C.c_void_p]),
    "miso_align_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_void_p]),
    "miso_align_intersections": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_atlas_features": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int32, C.c_void_p, C.c_void_p]),
    "miso_align_compose_poses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_align_pose_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_align_pose_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_slab_select": (C.c_int, [C.POINTER(Frames), C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_morton_keys": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "miso_transform_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                        C.c_void_p, C.c_void_p]),
    "miso_track_normal_equations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                              C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "miso_expand_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "miso_tc_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "miso_adam_step_tracked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                         C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p]),
    "miso_adam_step_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "miso_peer_signal": (C.c_int, [C.c_void_p, C.c_void_p]),
    "miso_peer_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "miso_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "miso_ipc_import": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "miso_adam_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_void_p]),
    "miso_set_tuning": (C.c_int, [C.c_char_p, C.c_int32]),
    "miso_get_tuning": (C.c_int, [C.c_char_p]),
    "miso_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the C-ABI library (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"miso_b200: {LIB_PATH} is missing -- build it with `python -m miso_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.miso_abi_version() != 1:
        raise RuntimeError("miso_b200: ABI version mismatch between _lib.py and libmiso_b200.so")
    _lib = lib
    return lib


# kernels launched by each entry point (for bench.py's `gpu_launches` claim)
KERNELS_PER_CALL = {"grid_sample3d_fwd": 1, "grid_sample3d_bwd": 1, "grid_sample3d_bwd_bwd": 1, "field_features": 1,
                    "sdf_forward": 1, "sdf_backward": 1, "mapping_count": 1, "mapping_step": 2, "mapping_step_fd": 6, "align_batch": 1,
                    "align_intersections": 2, "morton_keys": 1, "transform_points": 1, "adam_step": 1}
LAUNCHES = {"total": 0}


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().miso_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"miso_b200 {what} failed ({rc}): {msg}")
    LAUNCHES["total"] += KERNELS_PER_CALL.get(what, 1)


class tuning:
    """Context manager over miso_set_tuning: `with _lib.tuning(tc2_groups=3, pair=0): ...` (tests / profiling)."""

    def __init__(self, **kv):
        self.kv = kv
        self.old = {}

    def __enter__(self):
        lib = load()
        for k, v in self.kv.items():
            self.old[k] = lib.miso_get_tuning(k.encode())
            rc = lib.miso_set_tuning(k.encode(), int(v))
            if rc != 0:
                raise RuntimeError(lib.miso_last_error_string().decode())
        return self

    def __exit__(self, *exc):
        lib = load()
        for k, v in self.old.items():
            lib.miso_set_tuning(k.encode(), int(v))
        return False


def stream_ptr(device=None) -> int:
    """Raw cudaStream_t of torch's current stream on `device` (the C-level getter: ~10x cheaper than building a
    torch.cuda.Stream object per call, which matters for the per-level plugin at small point counts)."""
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = device.index if isinstance(device, torch.device) else int(device)
        if idx is None:
            idx = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class _NoCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NOCTX = _NoCtx()


def on_device(device):
    """Context that makes `device` current for the launch -- a no-op object when it already is (the common
    one-GPU-per-process case; torch.cuda.device() costs several microseconds per entry)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return _NOCTX if idx == torch.cuda.current_device() else torch.cuda.device(idx)


_I64_CACHE = {}


def i64c(vals):
    """Cached ctypes int64 array for a shape / stride tuple (read-only use)."""
    key = tuple(int(v) for v in vals)
    arr = _I64_CACHE.get(key)
    if arr is None:
        if len(_I64_CACHE) > 4096:
            _I64_CACHE.clear()
        arr = _I64_CACHE[key] = (C.c_int64 * len(key))(*key)
    return arr


def ptr(t):
    return None if t is None else t.data_ptr()


def i64(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("miso_b200: all tensors must live on a CUDA device (no CPU path exists); "
                               f"got a tensor on {t.device}")
