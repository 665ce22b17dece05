"""FD-vs-analytic gap of the PWA block with dropout at several step sizes / seeds / drop rates (developer probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import veloxseg_oracle as O
from tests._util import pwa_params
from tests.test_gpu_ops import PWA_LEVELS
from veloxseg_b200 import _lib, ops
DEV = "cuda:0"
lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
for lvl in ("autopet_L1", "autopet_L2"):
    size, C, mb, heads, mdh, M, e = PWA_LEVELS[lvl]
    geo = O.pwa_geometry(size, C, mb, [1, 1, 1], 2, heads, mdh)
    for (p_att, p_proj) in ((0.3, 0.2), (0.3, 0.0), (0.0, 0.2), (0.0, 0.0)):
        for seed in (77, 78):
            torch.manual_seed(11)
            xs = [torch.randn(1, C, *size, device=DEV) for _ in range(M)]
            flat, pd, table, index = pwa_params(M, C, geo, e, seed=4)
            flat, table, index = [p.to(DEV) for p in flat], table.to(DEV), index.to(DEV)
            fwd = lambda inp: ops.pwa_block_fwd_raw(lib, st, inp, flat, table, index, geo, e, p_att, p_proj, True, seed)
            zs, saved = fwd(xs)
            dzs = [torch.randn_like(z) for z in zs]
            dxs, dps, dtable = ops.pwa_block_bwd_raw(lib, st, dzs, xs, flat, table, index, saved, geo, e, p_att, p_proj, True, seed)
            d = [torch.randn_like(x) for x in xs]
            an = sum(float((gx.double() * dd.double()).sum()) for gx, dd in zip(dxs, d))
            out = []
            for eps in (3e-3, 1e-3, 3e-4):
                zp, _ = fwd([x + eps * dd for x, dd in zip(xs, d)])
                zm, _ = fwd([x - eps * dd for x, dd in zip(xs, d)])
                fd = sum(float(((a - b).double() / (2 * eps) * g.double()).sum()) for a, b, g in zip(zp, zm, dzs))
                out.append("eps %.0e: fd %.2f gap %.2f%%" % (eps, fd, 100 * abs(fd - an) / max(abs(fd), abs(an), 1.0)))
            print(lvl, "p_att %.1f p_proj %.1f seed %d  an %.2f | " % (p_att, p_proj, seed, an) + " | ".join(out), flush=True)
