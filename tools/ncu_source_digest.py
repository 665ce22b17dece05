#!/usr/bin/env python
"""Per-source-line digest of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`: for every CUDA source line the
stall samples and executed instructions of the SASS attributed to it.  usage: ncu_source_digest.py file.csv [top]"""
import collections
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = collections.OrderedDict()
fname, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or fname is None or not r[0].isdigit():
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    def num(name):
        i = ix.get(name)
        try:
            return float(r[i]) if i is not None and r[i] not in ("", "-") else 0.0
        except (ValueError, IndexError):
            return 0.0
    key = (fname, int(r[0]))
    a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0.0, "inst": 0.0})
    a["samples"] += num("# Samples")          # the CUDA-line rows carry the sums of their SASS lines
    a["inst"] += num("Instructions Executed")
tot_s = sum(a["samples"] for a in agg.values()) or 1.0
tot_i = sum(a["inst"] for a in agg.values()) or 1.0
print("total samples %.0f, warp instructions %.0f" % (tot_s, tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * a["samples"] / tot_s, 100 * a["inst"] / tot_i, f, ln, a["src"][:110]))
