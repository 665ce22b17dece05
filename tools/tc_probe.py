#!/usr/bin/env python
"""Bring-up probe for the tcgen05 pointwise kernel: runs the modal mixer (whose `t` output is the raw W.x + b) at
level-1/2 shapes with the tensor-core kernel in each descriptor variant and with the SIMT kernel, against an fp64
einsum.  GPU only.  `python tools/tc_probe.py [variants...]`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from veloxseg_b200 import _lib, ops  # noqa: E402

lib = _lib.get_lib()
dev = "cuda:0"
variants = [0]
shapes = [((16, 16), 16, (24, 24, 24), 2), ((32, 32), 32, (12, 12, 12), 2), ((16,), 48, (24, 24, 24), 1),
          ((64, 64), 64, (8, 8, 8), 1), ((40,), 24, (16, 16, 12), 3)]


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


for chs, Co, shp, B in shapes:
    torch.manual_seed(0)
    streams = [torch.randn(B, c, *shp, device=dev) for c in chs]
    W = torch.randn(Co, sum(chs), device=dev) * 0.2
    b = torch.randn(Co, device=dev) * 0.1
    ref = torch.einsum("oc,bcs->bos", W.double(), torch.cat(streams, 1).flatten(2).double()) + b.double()[None, :, None]
    st = torch.cuda.current_stream().cuda_stream
    lib.set_option(1, 0)
    y, t, stats = ops.mixer_fwd_raw(lib, st, streams, W, b, None)
    torch.cuda.synchronize()
    line = "K%s N%d S%s B%d  simt %.2e" % (chs, Co, shp, B, rel(t.flatten(2), ref))
    lib.set_option(1, 1)
    for v in variants:
        try:
            y, t, stats = ops.mixer_fwd_raw(lib, st, streams, W, b, None)
            torch.cuda.synchronize()
            line += "  tc[v%d] %.2e" % (v, rel(t.flatten(2), ref))
        except Exception as e:  # noqa: BLE001
            line += "  tc[v%d] ERR %s" % (v, str(e)[:60])
            break
    print(line, flush=True)
