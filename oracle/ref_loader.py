"""TEST INFRASTRUCTURE ONLY -- loads the unmodified reference (`/root/reference`) on CPU.

Only usable in the build container (the GPU box has no `/root/reference`).  Used by
`oracle/gen_golden.py` to produce the committed fixtures under `tests/golden/` and by the
`-m "not gpu"` tests that pin `oracle/oracle.py` against the reference itself.

The reference imports GUI / IO packages that are not installed here (SURVEY.md probe table);
they are stubbed with empty modules.  `pytorch3d` is absent and un-pinned in the reference
(`environment.yaml:114`); the two functions the hot path uses are restated here from the
published algorithm (pytorch3d >= 0.6 `so3_exp_map`: Rodrigues with `clamp(|w|^2, 1e-4)`).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MISO_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = {"matplotlib", "open3d", "trimesh", "mcubes", "pysdf", "evo", "pytorch3d",
               "plotly", "sdf", "cv2_stub_never"}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "grid_opt"))


class _StubModule(types.ModuleType):
    """Module whose every attribute is another stub (callable, subscriptable)."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _StubModule(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _StubModule(self.__name__ + "()")

    def __getitem__(self, k):
        return _StubModule(self.__name__ + "[]")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def hat(v: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.so3.hat: (N,3) -> (N,3,3) skew-symmetric."""
    N = v.shape[0]
    h = torch.zeros((N, 3, 3), dtype=v.dtype, device=v.device)
    x, y, z = v.unbind(1)
    h[:, 0, 1] = -z
    h[:, 0, 2] = y
    h[:, 1, 0] = z
    h[:, 1, 2] = -x
    h[:, 2, 0] = -y
    h[:, 2, 1] = x
    return h


def so3_exp_map(log_rot: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """pytorch3d.transforms.so3_exp_map restated (published algorithm, pytorch3d >= 0.6):
    nrms = sum(w^2); theta = sqrt(clamp(nrms, eps)); R = f1*K + f2*K@K + I,
    f1 = sin(theta)/theta, f2 = (1-cos(theta))/theta^2."""
    nrms = (log_rot * log_rot).sum(1)
    rot_angles = torch.clamp(nrms, eps).sqrt()
    rot_angles_inv = 1.0 / rot_angles
    fac1 = rot_angles_inv * rot_angles.sin()
    fac2 = rot_angles_inv * rot_angles_inv * (1.0 - rot_angles.cos())
    skews = hat(log_rot)
    skews_square = torch.bmm(skews, skews)
    R = (fac1[:, None, None] * skews + fac2[:, None, None] * skews_square
         + torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None])
    return R


_loaded = False


def load_reference():
    """Make `import grid_opt` resolve to the unmodified reference.  Idempotent."""
    global _loaded
    if _loaded:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT} (only present in the build container)")
    sys.meta_path.insert(0, _StubFinder())
    import pytorch3d.transforms as p3t  # stub
    p3t.so3_exp_map = so3_exp_map
    import pytorch3d.transforms.so3 as p3so3
    p3so3.hat = hat
    p3t.so3 = p3so3
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _loaded = True


def reference_model_cfg(bound, n_levels=2, feature_dim=4, base_cell_size=0.5, per_level_scale=5,
                        init_stddev=0.0, hidden_dim=64, hidden_layers=1, num_poses=1, fix=True,
                        second_order=False):
    """cfg['model'] dict shaped like configs/rgbd/scannet.yaml:7-29."""
    return {
        "name": "grid_net", "spatial_dim": 3,
        "decoder": {"type": "mlp", "hidden_dim": hidden_dim, "hidden_layers": hidden_layers, "out_dim": 1,
                    "pos_invariant": True, "fix": fix, "pretrained_model": None},
        "grid": {"type": "regular", "feature_dim": feature_dim, "init_stddev": init_stddev, "bound": bound,
                 "base_cell_size": base_cell_size, "per_level_scale": per_level_scale, "n_levels": n_levels,
                 "second_order_grid_sample": second_order},
        "pose": {"optimize": False, "num_poses": num_poses},
    }
