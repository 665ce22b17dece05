#!/usr/bin/env python
"""Where the sliding-window time goes on one GPU (Hecktor configuration, 2x320x320x256 volume): graph replays alone, replays
with the window gather and the accumulation, the whole call with staged IO.  usage: python tools/infer_breakdown.py [sw_batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from veloxseg_b200.configs import MODEL_CONFIGS, TRAIN
from veloxseg_b200.inference import GraphedPredictor, sliding_window_labels, window_starts
from veloxseg_b200.nn import VeloxSeg

dev = "cuda:0"
sw = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = MODEL_CONFIGS["hecktor2022"]
torch.manual_seed(12345)
model = VeloxSeg(**cfg).to(dev).eval()
roi = cfg["input_size"]
vol_h = torch.randn((1, 2, 320, 320, 256), generator=torch.Generator().manual_seed(5)).pin_memory()
pred = GraphedPredictor(model, sw, 2, roi, dev)
starts = window_starts(vol_h.shape[2:], roi, TRAIN["sw_overlap"])
nb = -(-len(starts) // sw)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


win = torch.randn((sw, 2) + tuple(roi), device=dev)
print("sw_batch %d: %d windows, %d batches" % (sw, len(starts), nb))
print("graph replays only          %.2f ms" % timed(lambda: [pred.graph.replay() for _ in range(nb)]))
print("predictor(win) (copy+replay) %.2f ms" % timed(lambda: [pred(win) for _ in range(nb)]))
vol = vol_h.to(dev)
acc = torch.zeros((2, 320, 320, 256), device=dev)


def loop():
    for g in range(0, len(starts), sw):
        ids = starts[g:g + sw]
        w = torch.cat([vol[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] for a, b, c in ids])
        y = pred(w)
        y = y[0] if isinstance(y, (list, tuple)) else y
        for k, (a, b, c) in enumerate(ids):
            acc[:, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += y[k]


print("gather + replay + accumulate %.2f ms" % timed(loop))
print("H2D of the whole volume      %.2f ms" % timed(lambda: vol.copy_(vol_h, non_blocking=True)))
out = torch.empty(vol_h.shape[2:], dtype=torch.uint8).pin_memory()
print("sliding_window_labels        %.2f ms" % timed(lambda: sliding_window_labels(vol_h, pred, roi, dev, sw_batch_size=sw, overlap=TRAIN["sw_overlap"], out_host=out)))
