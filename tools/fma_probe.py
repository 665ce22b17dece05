#!/usr/bin/env python
"""fp32 FMA issue rates of the device: immediate/constant-operand FFMA, 3-register FFMA, packed FFMA2 (fma.rn.f32x2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from veloxseg_b200 import _lib
lib = _lib.get_lib()
scratch = torch.tensor([0.0, 1.0000001, 0.9999999, 1.0000002], device="cuda")
st = torch.cuda.current_stream().cuda_stream
for kind, name, mult in ((0, "FFMA (constant operands)", 1), (2, "FFMA (3 registers)", 1), (3, "FFMA2 fma.rn.f32x2 (3 register pairs)", 2)):
    best = 1e9
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lib.c.vx_microbench(kind, 2000, scratch.data_ptr(), st); b.record(); torch.cuda.synchronize()
        if i: best = min(best, a.elapsed_time(b))
    fl = 2.0 * 8 * 32 * 2000 * 148 * 8 * 256 * mult
    print("%-44s %.3f ms  %.1f TFLOP/s" % (name, best, fl / (best * 1e-3) / 1e12))
