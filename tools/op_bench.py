#!/usr/bin/env python
"""Op-level timing at the real per-level shapes (BASELINE.json configs[4] style sweep): each op through the C ABI in a
back-to-back loop between two CUDA events on the launching stream.  `--only jlc` etc. narrows it (used under ncu)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import veloxseg_oracle as O  # noqa: E402  (geometry helper only)
from tests._util import jlc_params, pwa_params  # noqa: E402
from veloxseg_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
LEVELS = [((24, 24, 24), 16, 4, 3), ((12, 12, 12), 32, 4, 3), ((6, 6, 6), 64, 8, 2), ((3, 3, 3), 128, 8, 2)]
PWA = [((24, 24, 24), 16, [3, 3, 3], 1, 4, 3), ((12, 12, 12), 32, [6, 6, 6], 2, 8, 3), ((6, 6, 6), 64, [3, 3, 3], 2, 8, 2),
       ((3, 3, 3), 128, [3, 3, 3], 4, 16, 2)]


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--drop", type=float, default=0.0)
    ap.add_argument("--profile", action="store_true", help="print the library's per-kernel event table per op")
    args = ap.parse_args()
    lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
    lib.set_option(1, int(os.environ.get("VX_PW_TC", "1") == "1"))
    if "VX_PW_SMALL_MAX_S" in os.environ:
        lib.set_option(2, int(os.environ["VX_PW_SMALL_MAX_S"]))
    for opt, env in ((4, "VX_JLC_TILE_FWD"), (5, "VX_JLC_TILE_WGRAD")):
        if env in os.environ:
            tz, ty = [int(v) for v in os.environ[env].split(",")]
            lib.set_option(opt, tz << 8 | ty)
    if "VX_JLC_VX" in os.environ:
        lib.set_option(6, int(os.environ["VX_JLC_VX"]))
    if "VX_JLC_SMALL_MAX_S" in os.environ:
        lib.set_option(8, int(os.environ["VX_JLC_SMALL_MAX_S"]))
    if "VX_JLC_KS" in os.environ:
        lib.set_option(15, int(os.environ["VX_JLC_KS"]))
    if "VX_PW_TC_MIN_S" in os.environ:
        lib.set_option(3, int(os.environ["VX_PW_TC_MIN_S"]))
    B, res = args.B, []
    torch.manual_seed(0)

    def run(name, fwd, bwd=None):
        if args.only and args.only not in name:
            return
        if args.profile:
            lib.profile(True)
            fwd()
            if bwd:
                bwd()
            torch.cuda.synchronize()
            for r in sorted(lib.profile_report(), key=lambda r: -r[3]):
                print("   %-40s %-40s n=%d %.1f us" % (r[0], r[1][:40], r[2], 1e3 * r[3] / r[2]))
            lib.profile(False)
        row = {"op": name, "fwd_us": round(timeit(fwd, args.iters), 1)}
        if bwd:
            row["bwd_us"] = round(timeit(bwd, args.iters), 1)
        res.append(row)
        print(json.dumps(row), flush=True)

    for li, (shape, C, groups, e) in enumerate(LEVELS):
        x = torch.randn(B, C, *shape, device=DEV)
        params = [p.to(DEV) for p in jlc_params(C, groups, e, seed=1)]
        tr = args.drop > 0
        y, z, o, hpre, stats = ops.jlc_fwd_raw(lib, st, x, params, groups, e, args.drop, tr, 7)
        dy = torch.randn_like(y)
        run(f"jlc_L{li + 1}", lambda: ops.jlc_fwd_raw(lib, st, x, params, groups, e, args.drop, tr, 7),
            lambda: ops.jlc_bwd_raw(lib, st, dy, x, z, o, hpre, stats, params, groups, e, args.drop, tr, 7))
        streams = [torch.randn(B, C, *shape, device=DEV) for _ in range(2)]
        W, b_ = torch.randn(C, 2 * C, device=DEV) * 0.1, torch.zeros(C, device=DEV)
        add = torch.randn(B, C, *shape, device=DEV)
        yy, t, sst = ops.mixer_fwd_raw(lib, st, streams, W, b_, add)
        run(f"mixer_L{li + 1}", lambda: ops.mixer_fwd_raw(lib, st, streams, W, b_, add),
            lambda: ops.mixer_bwd_raw(lib, st, dy, streams, W, t, sst))
    for li, (size, C, mb, heads, mdh, e) in enumerate(PWA):
        geo = O.pwa_geometry(size, C, mb, [1, 1, 1], 2, heads, mdh)
        xs = [torch.randn(B, C, *size, device=DEV) for _ in range(2)]
        flat, _, table, index = pwa_params(2, C, geo, e, seed=2)
        flat, table, index = [p.to(DEV) for p in flat], table.to(DEV), index.to(DEV)
        tr = args.drop > 0
        zs, saved = ops.pwa_block_fwd_raw(lib, st, xs, flat, table, index, geo, e, args.drop, args.drop, tr, 5)
        dzs = [torch.randn_like(zz) for zz in zs]
        run(f"pwa_L{li + 1}", lambda: ops.pwa_block_fwd_raw(lib, st, xs, flat, table, index, geo, e, args.drop, args.drop, tr, 5),
            lambda: ops.pwa_block_bwd_raw(lib, st, dzs, xs, flat, table, index, saved, geo, e, args.drop, args.drop, tr, 5))
    # convolutions of the glue layers (AutoPET-II shapes): (name, C_in, C_out, k, s, p, transposed, shuffle, input extent)
    CONVS = [("conv3_out1", 16, 128, 3, 1, 1, False, 4, (24, 24, 24)), ("conv3_rc", 16, 64, 3, 1, 1, False, 4, (24, 24, 24)),
             ("down1", 2, 16, 7, 4, 3, False, 0, (96, 96, 96)), ("down2", 16, 32, 3, 2, 1, False, 0, (24, 24, 24)),
             ("down3", 32, 64, 3, 2, 1, False, 0, (12, 12, 12)), ("down4", 64, 128, 3, 2, 1, False, 0, (6, 6, 6)),
             ("up3", 128, 64, 2, 2, 0, True, 0, (3, 3, 3)), ("up2", 64, 32, 2, 2, 0, True, 0, (6, 6, 6)), ("up1", 32, 16, 2, 2, 0, True, 0, (12, 12, 12)),
             ("head2", 32, 2, 1, 1, 0, False, 0, (12, 12, 12)), ("head3", 64, 2, 1, 1, 0, False, 0, (6, 6, 6)), ("head4", 128, 2, 1, 1, 0, False, 0, (3, 3, 3))]
    for name, ci, co, k, s_, p_, tr, sh, ext in CONVS:
        if args.only and args.only not in "conv_" + name:
            continue
        xc = torch.randn(B, ci, *ext, device=DEV)
        wc = torch.randn(*((ci, co) if tr else (co, ci)), k, k, k, device=DEV) * 0.05
        yc = ops.conv_fwd_raw(lib, st, xc, wc, None, k, s_, p_, tr, sh)
        dyc = torch.randn_like(yc)
        need_dx = name != "down1"
        run("conv_" + name, lambda: ops.conv_fwd_raw(lib, st, xc, wc, None, k, s_, p_, tr, sh),
            lambda: ops.conv_bwd_raw(lib, st, dyc, xc, wc, k, s_, p_, tr, sh, need_dx=need_dx, need_db=False))
    f = torch.randn(B, 16, 24, 24, 24, device=DEV)
    dG = torch.randn(B, 16, 16, device=DEV)
    run("gram", lambda: ops.gram_fwd_raw(lib, st, f), lambda: ops.gram_bwd_raw(lib, st, dG, f))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"op_bench_B{B}{'_' + args.only if args.only else ''}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
