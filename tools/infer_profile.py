#!/usr/bin/env python
"""Per-kernel table of one eval forward (Hecktor configuration, batch of sliding windows), event-timed launch by launch.
usage: python tools/infer_profile.py [batch]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from veloxseg_b200 import _lib
from veloxseg_b200.configs import MODEL_CONFIGS
from veloxseg_b200.nn import VeloxSeg

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = MODEL_CONFIGS["hecktor2022"]
torch.manual_seed(12345)
model = VeloxSeg(**cfg).to(dev).eval()
x = torch.randn((B, 2) + tuple(cfg["input_size"]), device=dev)
lib = _lib.get_lib()
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    lib.profile(True)
    model(x)
    torch.cuda.synchronize()
rows = lib.profile_report()
lib.profile(False)
by = collections.OrderedDict()
tot = 0.0
for scope, kern, cnt, ms, nb, nf in rows:
    k = by.setdefault(kern, [0, 0.0]); k[0] += cnt; k[1] += ms; tot += ms
print("eval forward B=%d: %d launches, %.3f ms summed (event-timed, incl. ~5 us per launch of event overhead)" % (B, sum(v[0] for v in by.values()), tot))
for kern, (cnt, ms) in sorted(by.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%8.1f us  n=%3d  avg %6.1f  %s" % (ms * 1e3, cnt, ms * 1e3 / cnt, kern))
print("-- by op scope")
sc = collections.OrderedDict()
for scope, kern, cnt, ms, nb, nf in rows:
    k = sc.setdefault(scope, [0, 0.0]); k[0] += cnt; k[1] += ms
for scope, (cnt, ms) in sorted(sc.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%8.1f us  n=%3d  %s" % (ms * 1e3, cnt, scope))
