#!/usr/bin/env python
"""Device timeline of the library's kernels in the profiled eager steps (gpurun_out/timeline.json, written by bench.py):
how many of our kernels run side by side over time, and which kernels own the stretches where only one (or none) runs --
the serial part of the step that more overlap cannot hide."""
import collections
import json
import sys

tl = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"))
tl = [(s, k.strip("()"), a, d) for s, k, a, d in tl]
# split into the profiled steps: a gap > 20 ms (the spin kernel) starts a new one
tl.sort(key=lambda r: r[2])
steps, cur = [], [tl[0]]
for r in tl[1:]:
    if r[2] - (cur[-1][2] + cur[-1][3]) > 20000:
        steps.append(cur); cur = []
    cur.append(r)
steps.append(cur)
st = steps[-1]
t0 = min(r[2] for r in st); t1 = max(r[2] + r[3] for r in st)
print("steps %d; last step: %d launches, span %.2f ms, summed kernel time %.2f ms" % (len(steps), len(st), (t1 - t0) / 1e3, sum(r[3] for r in st) / 1e3))
ev = []
for i, r in enumerate(st):
    ev.append((r[2], 1, i)); ev.append((r[2] + r[3], -1, i))
ev.sort()
live, last = set(), t0
conc = collections.Counter()
solo = collections.Counter()
idle_after = collections.Counter()
prev_end_name = None
for t, kind, i in ev:
    dt = t - last
    if dt > 0:
        conc[min(len(live), 4)] += dt
        if len(live) == 1:
            j = next(iter(live)); solo[st[j][1] + " @ " + st[j][0]] += dt
        if len(live) == 0 and prev_end_name:
            idle_after[prev_end_name] += dt
    last = t
    if kind == 1:
        live.add(i)
    else:
        live.discard(i); prev_end_name = st[i][1]
tot = sum(conc.values())
print("concurrency (our kernels only): " + ", ".join("%d%s: %.2f ms (%.0f%%)" % (k, "+" if k == 4 else "", v / 1e3, 100 * v / tot) for k, v in sorted(conc.items())))
print("time with exactly one of our kernels running, by kernel @ op (top 30):")
for k, v in solo.most_common(30):
    print("   %8.1f us  %s" % (v, k))
print("idle gaps (none of our kernels running: cuDNN / torch kernels or launch gaps), by the kernel that ended before (top 12):")
for k, v in idle_after.most_common(12):
    print("   %8.1f us  after %s" % (v, k))
