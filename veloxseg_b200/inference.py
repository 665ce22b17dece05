"""Sliding-window inference, single GPU or with windows sharded across ranks (new capability).

Semantics restate monai.inferers.sliding_window_inference (MONAI 1.5.0, un-vendored dependency; the reference's
utils/inference_runtime.py:4-19 is a pass-through to it) for the reference's call: mode="constant" (importance map of
ones), padding_mode="constant" 0, overlap 0.25, roi = training patch size.  The slice table is integer and
deterministic; windows are independent (every norm in VeloxSeg is per-sample), so rank r takes windows r, r+W, ... and
the partial sums `sum_w logits` are combined with one all-reduce; the count map needs no communication.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


def scan_interval(image: Sequence[int], roi: Sequence[int], overlap: float) -> List[int]:
    return [int(r) if r == i else max(int(r * (1 - overlap)), 1) for i, r in zip(image, roi)]


def window_starts(image: Sequence[int], roi: Sequence[int], overlap: float) -> List[Tuple[int, int, int]]:
    """dense_patch_slices: per axis ceil(image/interval) candidates up to the first that reaches the border, the last
    one clamped back to fit; enumeration with the first spatial axis slowest."""
    per_axis = []
    for i, r, s in zip(image, roi, scan_interval(image, roi, overlap)):
        num = int(math.ceil(i / s))
        n = next((d + 1 for d in range(num) if d * s + r >= i), 1)
        per_axis.append([d * s - max(d * s + r - i, 0) for d in range(n)])
    return [(a, b, c) for a in per_axis[0] for b in per_axis[1] for c in per_axis[2]]


def count_map(image: Sequence[int], roi: Sequence[int], starts, device, dtype=torch.float32) -> torch.Tensor:
    cnt = torch.zeros((1, 1) + tuple(image), dtype=dtype, device=device)
    for a, b, c in starts:
        cnt[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += 1
    return cnt


def _logits(y):
    """Net.forward of the reference's inference wrapper takes outputs[0] of a list (utils/inference_petct.py:46-51)."""
    return y[0] if isinstance(y, (list, tuple)) else y


@torch.no_grad()
def sliding_window_predict(inputs: torch.Tensor, predictor: Callable, roi_size: Sequence[int], sw_batch_size: int = 2,
                           overlap: float = 0.25, group=None, shard: bool = True) -> torch.Tensor:
    """(1|B, C, X, Y, Z) -> (B, n_cls, X, Y, Z).  With an initialised process group and shard=True every rank must
    call this with the same `inputs`; all ranks return the full result."""
    B = inputs.shape[0]
    size = list(inputs.shape[2:])
    pads = [max(r - s, 0) for r, s in zip(roi_size, size)]
    lo = [p // 2 for p in pads]
    if any(pads):
        inputs = F.pad(inputs, [lo[2], pads[2] - lo[2], lo[1], pads[1] - lo[1], lo[0], pads[0] - lo[0]])
    image = list(inputs.shape[2:])
    starts = window_starts(image, roi_size, overlap)
    nwin = len(starts)
    world = dist.get_world_size(group) if (shard and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = [i for i in range(B * nwin) if i % world == rank]
    out = None
    for g in range(0, len(mine), sw_batch_size):
        ids = mine[g:g + sw_batch_size]
        win = torch.cat([inputs[i // nwin:i // nwin + 1, :, starts[i % nwin][0]:starts[i % nwin][0] + roi_size[0],
                                starts[i % nwin][1]:starts[i % nwin][1] + roi_size[1],
                                starts[i % nwin][2]:starts[i % nwin][2] + roi_size[2]] for i in ids])
        y = _logits(predictor(win))
        if out is None:
            out = torch.zeros((B, y.shape[1]) + tuple(image), dtype=y.dtype, device=y.device)
        for k, i in enumerate(ids):
            a, b, c = starts[i % nwin]
            out[i // nwin, :, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]] += y[k]
    if world > 1:
        if out is None:     # more ranks than windows
            n_cls = _logits(predictor(inputs[:1, :, :roi_size[0], :roi_size[1], :roi_size[2]])).shape[1]
            out = torch.zeros((B, n_cls) + tuple(image), dtype=inputs.dtype, device=inputs.device)
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    out = out / count_map(image, roi_size, starts, out.device, out.dtype)
    return out[:, :, lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]]


class GraphedPredictor:
    """Eval forward of a fixed window-batch shape captured once into a CUDA graph and replayed per batch of windows
    (an eager eval forward is ~300 launches for ~2 ms of GPU work: host-bound).  Smaller final batches are padded."""

    def __init__(self, model: torch.nn.Module, batch: int, channels: int, roi_size: Sequence[int], device):
        self.model = model.eval()
        self.x = torch.zeros((batch, channels) + tuple(roi_size), device=device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                self.model(self.x)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.y = _logits(self.model(self.x))

    def __call__(self, win: torch.Tensor) -> torch.Tensor:
        k = win.shape[0]
        self.x[:k].copy_(win)
        self.graph.replay()
        return self.y[:k]
