#!/bin/bash
# sweep the JLC conv tile (z, y) extents: warm per-launch device times of the conv kernels per level
mkdir -p gpurun_out
export VX_JLC_VX=${VX_JLC_VX:-0}
for t in ${TILES:-"0,0" "8,8" "8,4" "4,4" "6,6" "3,3"}; do
  VX_JLC_TILE_FWD=$t VX_JLC_TILE_WGRAD=$t timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none \
     -k "regex:jlc_conv" -c 400 --csv --log-file gpurun_out/jlc_tile.csv python tools/op_bench.py --iters 1 --only jlc > gpurun_out/jlc_tile.log 2>&1
  python - "$t" <<'PY'
import csv, collections, sys
lines = [l for l in open("gpurun_out/jlc_tile.csv") if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    k = (r["Kernel Name"].split("(")[0].replace("void ", "").replace("vx::jlc_conv_", "").replace("_kernel", ""), r["Grid Size"], r["Block Size"])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
print("tile", sys.argv[1], " | ".join("%s %s%s %.0f" % (k[0], k[1].replace(" ", ""), k[2].replace(" ", ""), v[1] / v[0] / 1e3) for k, v in agg.items()))
PY
done
