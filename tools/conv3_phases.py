#!/usr/bin/env python
"""Phase timestamps (clock64 of CTA 0) of conv3_fwd_tc_kernel at the out_conv1 shape: where a CTA's time goes.  GPU only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from veloxseg_b200 import _lib, ops  # noqa: E402

lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
x = torch.randn(4, 16, 24, 24, 24, device="cuda")
w = torch.randn(128, 16, 3, 3, 3, device="cuda") * 0.05
for _ in range(3):
    ops.conv_fwd_raw(lib, st, x, w, None, 3, 1, 1, False, 4)
lib.set_option(13, 1)
ops.conv_fwd_raw(lib, st, x, w, None, 3, 1, 1, False, 4)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.c.vx_conv3_trace(buf, 64)
lib.set_option(13, 0)
t0 = buf[0]
names = {0: "start", 1: "barriers+tmem alloc done", 2: "brick staged + sync", 3: "issue thread done", 4: "block 0 complete", 5: "last block complete", 6: "epilogue done"}
for k, n in names.items():
    print("%-28s %8d cycles" % (n, buf[k] - t0))
for g in range(9):
    print("group %d: weights arrived %8d   MMAs issued %8d" % (g, buf[8 + 2 * g] - t0, buf[9 + 2 * g] - t0))
