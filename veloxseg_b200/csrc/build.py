"""Builds libveloxseg_sm100.so (nvcc, sm_100a) in-tree.  `python -m veloxseg_b200.csrc.build [--emu]`.

--emu builds tools/emu/libveloxseg_emu.so instead: the same kernel sources compiled by g++ against the
test-only CPU shim in tools/emu (developer tooling, never loaded by the package).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["api.cu", "pointwise.cu", "pw_tc.cu", "pw_ffn_tc.cu", "pw_wgrad_tc.cu", "jlc.cu", "conv3_tc.cu", "conv_simt.cu", "ops_misc.cu", "pwa.cu", "resize.cu", "segloss.cu", "stem.cu", "diag.cu"]
LIB = os.path.join(os.path.dirname(HERE), "libveloxseg_sm100.so")
EMU_LIB = os.path.join(ROOT, "tools", "emu", "libveloxseg_emu.so")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "veloxseg_abi.h"))
    return hs


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    hdrs = _headers()
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
             "--use_fast_math" if os.environ.get("VX_FAST_MATH") else "-DVX_PRECISE"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    jobs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            jobs.append([nvcc] + flags + ["-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        for o in outs:
            print(o)
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        _run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB


def build_emu(force=False):
    emu = os.path.join(ROOT, "tools", "emu")
    objdir = os.path.join(emu, "build")
    os.makedirs(objdir, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    hdrs = _headers() + [os.path.join(emu, "cuda_emu.h")]
    flags = ["-O2", "-g", "-std=c++20", "-fPIC", "-DVX_EMU", "-I", emu, "-pthread", "-Wno-unused-value"]
    jobs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            jobs.append(["g++"] + flags + ["-x", "c++", "-c", s, "-o", o])
    o = os.path.join(objdir, "cuda_emu.o")
    if force or _stale(o, [os.path.join(emu, "cuda_emu.cpp")] + hdrs):
        jobs.append(["g++"] + flags + ["-c", os.path.join(emu, "cuda_emu.cpp"), "-o", o])
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, jobs))
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in srcs] + [o]
    if force or jobs or _stale(EMU_LIB, objs):
        _run(["g++", "-shared", "-pthread", "-o", EMU_LIB] + objs)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
