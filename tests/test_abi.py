"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/veloxseg_abi.h declares.
No compute calls (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from veloxseg_b200.csrc.build import build
    return build()


def _declared():
    src = open(os.path.join(ROOT, "include", "veloxseg_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib_path):
    from veloxseg_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    c = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(c, n), n
    c.vx_version.restype = ctypes.c_int
    assert c.vx_version() >= 100


def test_bad_descriptor_is_reported_not_crashed(lib_path):
    from veloxseg_b200 import _lib
    lib = _lib.VxLib(lib_path)
    d = _lib.JlcDesc(1, 10, 4, 4, 4, 3, 2, 1e-5, 0.0, 0, 0)     # 10 channels do not split into 3 groups
    assert lib.workspace("jlc", d) == 0
    rc = lib.c.vx_jlc_fwd(ctypes.byref(d), None, None, None, 0, None)
    assert rc < 0 and "jlc" in lib.last_error()


def test_no_cpu_fallback():
    import torch
    from veloxseg_b200 import ops
    x = torch.randn(1, 8, 4, 4, 4)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.gram(x)


def test_sass_is_sm100(lib_path):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_conv_entry_point_reports_unsupported_geometry(lib_path):
    """vx_conv_* has no library fallback: a geometry outside the VeloxSeg layers is refused with a message, before any
    buffer is touched; the workspace query of the tcgen05 convolution is non-zero and of the others zero."""
    from veloxseg_b200 import _lib
    lib = _lib.VxLib(lib_path)
    d = _lib.ConvDesc(1, 16, 64, 4, 4, 4, 2, 2, 0, 1, 4)          # transposed conv with a pixel shuffle: not a VeloxSeg layer
    dummy = (ctypes.c_float * 4)()
    ptrs = (ctypes.c_void_p * 3)(ctypes.addressof(dummy), ctypes.addressof(dummy), None)
    outs = (ctypes.c_void_p * 3)(ctypes.addressof(dummy), ctypes.addressof(dummy), None)
    rc = lib.c.vx_conv_fwd(ctypes.byref(d), ptrs, outs, None, 0, None)
    assert rc == -4 and "unsupported" in lib.last_error()
    assert lib.workspace("conv", _lib.ConvDesc(4, 16, 128, 24, 24, 24, 3, 1, 1, 0, 4)) > 0
    assert lib.workspace("conv", _lib.ConvDesc(4, 2, 16, 96, 96, 96, 7, 4, 3, 0, 0)) == 0


def test_model_path_has_no_library_convolution():
    """nn.py calls no torch / cuDNN convolution (everything goes through ops.conv3d)."""
    src = open(os.path.join(ROOT, "veloxseg_b200", "nn.py")).read()
    for needle in ("F.conv3d", "F.conv_transpose3d", "torch.nn.functional", "self.proj(", "self.down(", "self.up(", "seq(x)"):
        assert needle not in src, needle
