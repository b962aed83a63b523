"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/miso_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "miso_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(miso_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for required in ["miso_grid_sample3d_fwd", "miso_grid_sample3d_bwd", "miso_grid_sample3d_bwd_bwd",
                     "miso_field_features", "miso_sdf_forward", "miso_sdf_backward", "miso_mapping_step",
                     "miso_align_batch", "miso_align_intersections", "miso_adam_step", "miso_morton_keys",
                     "miso_transform_points", "miso_last_error_string"]:
        assert required in fns


def test_library_exports_every_declared_symbol(lib):
    from miso_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(header_functions())
    assert lib.miso_abi_version() == 1
    assert lib.miso_last_error_string() is not None


def test_struct_layouts_match_header():
    """ctypes mirrors must have the sizes the C compiler gives the header structs."""
    import subprocess, tempfile
    from miso_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "miso_b200.h"
int main(){printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(miso_level_t), sizeof(miso_field_t), sizeof(miso_decoder_t),
 sizeof(miso_frames_t), sizeof(miso_mapping_cfg_t), sizeof(miso_align_pair_t), (size_t)MISO_ALIGN_OUT);return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(s) for s in (_lib.Level, _lib.Field, _lib.Decoder, _lib.Frames, _lib.MappingCfg, _lib.AlignPair)]
    assert got == sizes[:6]
    assert sizes[6] == _lib.MISO_ALIGN_OUT


def test_sass_is_sm100a_with_vector_reds_and_ffma2():
    """Evidence the shipped binary is the hand-written sm_100a path: packed FFMA2 in the MLP and 128-bit
    reductions in the scatter."""
    import shutil, subprocess
    from miso_b200 import _lib, build
    build.build(verbose=False)
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "FFMA2" in sass
    assert "REDG.E.ADD.F32x4" in sass, "no 128-bit vector reduction found"
