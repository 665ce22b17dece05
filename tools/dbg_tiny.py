import os, sys
sys.path.insert(0, "/root/repo")
import torch
from veloxseg_b200 import _lib, ops
lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
cases = [("up3", 64, 32, 2, 2, 0, True, (2, 2, 2)), ("up2", 32, 16, 2, 2, 0, True, (4, 4, 4)), ("up1", 16, 8, 2, 2, 0, True, (8, 8, 8)),
         ("head2", 16, 2, 1, 1, 0, False, (8, 8, 8)), ("head3", 32, 2, 1, 1, 0, False, (4, 4, 4)), ("head4", 64, 2, 1, 1, 0, False, (2, 2, 2)),
         ("out1", 8, 128, 3, 1, 1, False, (16, 16, 16))]
for B in (1, 2):
    for name, ci, co, k, s, p, tr, ext in cases:
        x = torch.randn(B, ci, *ext, device="cuda")
        w = torch.randn(*((ci, co) if tr else (co, ci)), k, k, k, device="cuda") * 0.05
        try:
            y = ops.conv_fwd_raw(lib, st, x, w, None, k, s, p, tr, 0)
            torch.cuda.synchronize()
            dy = torch.randn_like(y)
            ops.conv_bwd_raw(lib, st, dy, x, w, k, s, p, tr, 0, need_db=not tr)
            torch.cuda.synchronize()
            print(B, name, "ok")
        except Exception as e:
            print(B, name, "FAIL", str(e)[:150])
            torch.cuda.synchronize()
