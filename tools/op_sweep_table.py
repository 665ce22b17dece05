#!/usr/bin/env python
"""BASELINE.json configs[4]: op-level sweep table from tools/op_bench.py runs (gpurun_out/op_sweep_B*.jsonl), with the
algorithmic bytes / flops of SURVEY.md section 8d (AutoPET-II level shapes, fp32, M = 2)."""
import glob
import json
import re

HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6469.9
FMA_TF = 74.4          # 148 SMs x 128 lanes x 2 x 1.965 GHz
LV = {1: (13824, 16, 4, 3), 2: (1728, 32, 8, 3), 3: (216, 64, 8, 2), 4: (27, 128, 16, 2)}     # S, C, c_g, e
PWA = {1: (1, 585, 54, 4, 4, 16, 16), 2: (2, 9, 432, 8, 8, 32, 32), 3: (2, 9, 54, 8, 16, 32, 64), 4: (4, 1, 54, 16, 32, 64, 128)}  # h, Ns, L, cq, cv, cqk_tot, cv_tot


def model(op, B):
    kind, lv = op.split("_L")
    lv = int(lv)
    S, C, cg, e = LV[lv]
    if kind == "jlc":
        P = C * cg * 153 + 3 * C + 2 * e * C * C + e * C + C
        return B * 2 * C * S * 4 + 4 * P, B * 2 * S * C * (153 * cg + 2 * e * C)
    if kind == "mixer":
        K = 2 * C
        return B * S * (K + 2 * C) * 4 + 4 * (K * C + C), B * 2 * S * K * C
    if kind == "pwa":
        h, Ns, L, cq, cv, cqk, cvt = PWA[lv]
        M = 2
        P = M * (2 * C + (2 * cqk + cvt) * (C + 1) + C * (cvt + 1) + 2 * C + 2 * e * C * C + e * C + C)
        fl = 2 * (M * S * C * (2 * cqk + cvt) + h * Ns * L * L * (cq + cv) + M * S * cvt * C) + 4 * M * S * e * C * C
        return B * 2 * M * C * S * 4 + 4 * P, B * fl
    return None, None


print("| op | B | fwd us | bwd us | fwd alg GB/s (%% of %.0f) | fwd GFLOP/s (%% of fp32 FMA %.1f T) |" % (HBM, FMA_TF))
print("|---|---|---|---|---|---|")
for f in sorted(glob.glob("gpurun_out/op_sweep_B*.jsonl"), key=lambda p: int(re.search(r"B(\d+)", p).group(1))):
    B = int(re.search(r"B(\d+)", f).group(1))
    for line in open(f):
        r = json.loads(line)
        if "_L" not in r["op"]:
            continue
        by, fl = model(r["op"], B)
        gbs = by / (r["fwd_us"] * 1e-6) / 1e9
        gf = fl / (r["fwd_us"] * 1e-6) / 1e9
        print("| %s | %d | %.1f | %.1f | %.0f (%.1f%%) | %.0f (%.1f%%) |" % (r["op"], B, r["fwd_us"], r["bwd_us"], gbs, 100 * gbs / HBM, gf, 100 * gf / (FMA_TF * 1e3)))
