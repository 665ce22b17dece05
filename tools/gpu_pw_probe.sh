#!/bin/bash
# warm per-launch device times of the pointwise kernels inside the mixer / JLC / PWA ops, tensor-core path on and off
mkdir -p gpurun_out
for tc in 1 0 2; do
  if [ $tc = 2 ]; then export VX_PW_SMALL_MAX_S=100000; fi
  VX_PW_TC=$tc timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -k "regex:pw_" -c 4000 \
     --csv --log-file gpurun_out/pw_probe_tc$tc.csv python tools/op_bench.py --iters 2 > gpurun_out/pw_probe_tc$tc.log 2>&1
done
python - <<'PY'
import csv, collections
for tc in (1, 0, 2):
    lines = [l for l in open(f"gpurun_out/pw_probe_tc{tc}.csv") if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = (r["Kernel Name"].split("(")[0].replace("void ", ""), r["Grid Size"])
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
    print("== tensor-core path", tc)
    for k, v in agg.items():
        print("%-28s %-16s n=%3d avg %7.1f us" % (k[0], k[1], v[0], v[1] / v[0] / 1e3))
PY
