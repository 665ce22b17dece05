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


def test_candidate_kernels_are_off_by_default(lib_path):
    """The tcgen05 candidates written without GPU access must not be reachable unless switched on: the dense-conv entry point
    refuses before it looks at its buffers, and the Python layer routes out_conv1 to the library convolution."""
    import os
    from veloxseg_b200 import _lib, ops
    lib = _lib.VxLib(lib_path)
    d = _lib.DenseConvDesc(1, 16, 64, 4, 4, 4, 0)
    dummy = (ctypes.c_float * 4)()
    ptrs = (ctypes.c_void_p * 3)(ctypes.addressof(dummy), ctypes.addressof(dummy), None)
    outs = (ctypes.c_void_p * 1)(ctypes.addressof(dummy))
    rc = lib.c.vx_dense_conv_fwd(ctypes.byref(d), ptrs, outs, None)
    assert rc < 0 and "off" in lib.last_error()
    if os.environ.get("VX_DENSE_CONV_TC", "0") != "1":
        assert ops.dense_conv_tc_enabled() is False
