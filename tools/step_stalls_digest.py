#!/usr/bin/env python
"""Per-kernel digest of tools/gpu_ncu_step_stalls.sh: launches, summed time, executed warp instructions per warp, SASS-fetch
stall (no_instruction) and the other leading stall reasons, aggregated by kernel name over one train step."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
K = {"t": "gpu__time_duration.sum", "inst": "smsp__inst_executed.sum", "noinst": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
     "long": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "short": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active"}


def num(r, k):
    i = ix.get(K[k])
    try:
        return float(r[i].replace(",", "")) if i is not None else 0.0
    except ValueError:
        return 0.0


agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("vx::", "")
    g = [int(x) for x in re.findall(r"\d+", r[ix["Grid Size"]])]
    b = [int(x) for x in re.findall(r"\d+", r[ix["Block Size"]])]
    warps = g[0] * g[1] * g[2] * ((b[0] * b[1] * b[2] + 31) // 32)
    a = agg.setdefault(name, collections.defaultdict(float))
    t = num(r, "t")
    a["n"] += 1; a["t"] += t; a["ipw"] += num(r, "inst") / max(warps, 1)
    for k in ("noinst", "long", "short", "barrier", "wait", "issue", "warps"):
        a[k] += num(r, k) * t            # time-weighted
tot = sum(a["t"] for a in agg.values())
print("%d kernels, %d launches, %.2f ms summed" % (len(agg), sum(a["n"] for a in agg.values()), tot / 1e3))
print("%-44s %4s %8s %6s %7s | stalled warps per issue: %6s %6s %6s %6s %6s | %6s %6s" %
      ("kernel", "n", "us", "avg", "inst/w", "noinst", "long", "short", "barr", "wait", "issue%", "warps%"))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    t = a["t"] or 1.0
    print("%-44s %4d %8.1f %6.1f %7.0f | %30.1f %6.1f %6.1f %6.1f %6.1f | %6.1f %6.1f" %
          (name[:44], a["n"], a["t"], a["t"] / a["n"], a["ipw"] / a["n"], a["noinst"] / t, a["long"] / t, a["short"] / t,
           a["barrier"] / t, a["wait"] / t, a["issue"] / t, a["warps"] / t))
