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


def axis_starts(image: Sequence[int], roi: Sequence[int], overlap: float) -> List[List[int]]:
    """dense_patch_slices, one axis at a time: ceil(image/interval) candidates up to the first that reaches the border,
    the last one clamped back to fit."""
    per_axis = []
    for i, r, s in zip(image, roi, scan_interval(image, roi, overlap)):
        num = int(math.ceil(i / s))
        n = next((d + 1 for d in range(num) if d * s + r >= i), 1)
        per_axis.append([d * s - max(d * s + r - i, 0) for d in range(n)])
    return per_axis


def window_starts(image: Sequence[int], roi: Sequence[int], overlap: float) -> List[Tuple[int, int, int]]:
    """The window list is the Cartesian product of the per-axis starts, first spatial axis slowest."""
    per_axis = axis_starts(image, roi, overlap)
    return [(a, b, c) for a in per_axis[0] for b in per_axis[1] for c in per_axis[2]]


def count_map(image: Sequence[int], roi: Sequence[int], starts, device, dtype=torch.float32) -> torch.Tensor:
    cnt = torch.zeros((1, 1) + tuple(image), dtype=dtype, device=device)
    for a, b, c in starts:
        cnt[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += 1
    return cnt


def axis_counts(image: Sequence[int], roi: Sequence[int], per_axis, device, dtype=torch.float32) -> List[torch.Tensor]:
    """count_map factorised: the windows are a Cartesian product, so count(x, y, z) = cx[x] * cy[y] * cz[z] (small
    integers, exact in fp32) -- three 1-D tables instead of one slice-add over the volume per window."""
    out = []
    for n, r, st in zip(image, roi, per_axis):
        c = [0] * n
        for s0 in st:
            for j in range(s0, s0 + r):
                c[j] += 1
        out.append(torch.tensor(c, dtype=dtype, device=device))
    return out


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


def upload_blocks(size: Sequence[int], roi: Sequence[int], overlap: float) -> List[Tuple[int, int, int, int]]:
    """Upload schedule of `sliding_window_labels` on one GPU: blocks (x0, x1, y0, y1) of (plane slab x row slab), in the order the
    windows (first axis outermost, then second) first touch them.  The slab edges are the upper edges of the window rows, so a
    window [a, a + r0) x [b, b + r1) lies inside the blocks whose key (x1, y1) is <= (a + r0, b + r1) in lexicographic order --
    which is also the upload order (`blocks_ready` below)."""
    per = axis_starts(size, roi, overlap)
    xe = sorted({a + roi[0] for a in per[0]})
    ye = sorted({b + roi[1] for b in per[1]})
    out, x0 = [], 0
    for x1 in xe:
        y0 = 0
        for y1 in ye:
            out.append((x0, x1, y0, y1))
            y0 = y1
        x0 = x1
    return out


def blocks_ready(pending: list, need: Tuple[int, int]) -> list:
    """Pops and returns the leading entries of `pending` ([((x1, y1), payload), ...] in upload order) a batch of windows whose
    largest (a + r0, b + r1) is `need` has to wait for."""
    done = []
    while pending and pending[0][0] <= need:
        done.append(pending.pop(0))
    return done


_COPY_STREAMS = {}


def _copy_stream(device):
    """One upload stream per device, created on first use."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


@torch.no_grad()
def sliding_window_labels(volume_host: torch.Tensor, predictor: Callable, roi_size: Sequence[int], device,
                          sw_batch_size: int = 2, overlap: float = 0.25, group=None, dst: int = 0,
                          out_host: torch.Tensor | None = None):
    """Host volume (1, C, X, Y, Z) in -> host uint8 label map (X, Y, Z) = argmax of the blended logits, on rank `dst`
    (None on the other ranks).  Same arithmetic as `sliding_window_predict(...).argmax(1)` (utils/inference_petct.py:214-233
    takes the arg-max of the sliding-window logits); what changes is where the bytes travel when windows are sharded:

    * in:  rank r copies 1/world of the (contiguous) host volume over its own PCIe link and the ranks all-gather the
           pieces over NVLink, instead of every rank pulling the whole volume from host memory at once;
    * out: the partial sums are reduce-scattered in slabs of the first spatial axis, each rank divides by the count map
           and takes the arg-max of its slab only, and the 1-byte labels (not the n_cls fp32 logits) are gathered;
    * the count map is the outer product of three per-axis tables (`axis_counts`).

    Every rank calls this with the same host volume (as the reference's per-rank data loader would supply it).
    """
    assert volume_host.dim() == 5 and volume_host.shape[0] == 1, "one volume per call (the reference infers batch 1)"
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    device = torch.device(device)
    C = volume_host.shape[1]
    size = list(volume_host.shape[2:])
    host_flat = volume_host.contiguous().view(-1)
    n = host_flat.numel()
    chunk = -(-n // world)
    lo_e, hi_e = min(rank * chunk, n), min((rank + 1) * chunk, n)
    if world > 1:
        piece = torch.empty(chunk, dtype=volume_host.dtype, device=device)
        piece[:hi_e - lo_e].copy_(host_flat[lo_e:hi_e], non_blocking=True)
        flat = torch.empty(world * chunk, dtype=volume_host.dtype, device=device)
        dist.all_gather_into_tensor(flat, piece, group=group)
    pads = [max(r - s, 0) for r, s in zip(roi_size, size)]
    lo = [p // 2 for p in pads]
    slab_ready = None           # world == 1: [(first-axis extent covered, event)] of the staged upload
    if world > 1:
        inputs = flat[:n].view(1, C, *size)
    elif device.type == "cuda" and not any(pads) and volume_host.is_pinned() and volume_host.is_contiguous():
        # one GPU: the windows are visited in (first axis, second axis) order, so the volume is uploaded on a copy stream in
        # blocks (slab of planes x slab of rows: what the next row of windows adds) and a batch of windows waits only for the
        # blocks it reads -- the rest of the upload runs under the forward passes of the earlier windows
        from ._lib import get_lib
        lib = get_lib()
        inputs = torch.empty((1, C, *size), dtype=volume_host.dtype, device=device)
        cur = torch.cuda.current_stream(device)
        cs = _copy_stream(device)
        cs.wait_stream(cur)
        esz = volume_host.element_size()
        plane, rowb = size[1] * size[2] * esz, size[2] * esz
        slab_ready = []
        for x0, x1, y0, y1 in upload_blocks(size, roi_size, overlap):
            for c in range(C):
                off = ((c * size[0] + x0) * size[1] + y0) * size[2] * esz
                lib.check(lib.c.vx_copy_block_async(inputs.data_ptr() + off, volume_host.data_ptr() + off, plane, (y1 - y0) * rowb,
                                                    x1 - x0, 0, cs.cuda_stream), "vx_copy_block_async")
            ev = torch.cuda.Event()
            ev.record(cs)
            slab_ready.append(((x1, y1), ev))
    else:
        inputs = host_flat.to(device, non_blocking=True).view(1, C, *size)
    if any(pads):
        inputs = F.pad(inputs, [lo[2], pads[2] - lo[2], lo[1], pads[1] - lo[1], lo[0], pads[0] - lo[0]])
    image = list(inputs.shape[2:])
    per_axis = axis_starts(image, roi_size, overlap)
    starts = [(a, b, c) for a in per_axis[0] for b in per_axis[1] for c in per_axis[2]]
    xs = -(-image[0] // world)                      # slab thickness; the accumulator is padded to world * xs planes
    mine = [i for i in range(len(starts)) if i % world == rank]
    acc = None
    # one GPU with a pinned result buffer: planes no later window touches are finished (count division, arg-max) and sent back
    # while the remaining windows run
    stream_out = slab_ready is not None and out_host is not None and out_host.is_pinned()
    fin, counts, keep = 0, None, []

    def finish_planes(x1):
        nonlocal fin, counts
        if counts is None:
            counts = axis_counts(image, roi_size, per_axis, acc.device, acc.dtype)
        cx, cy, cz = counts
        part = acc[:, fin:x1] / (cx[fin:x1, None, None] * cy[None, :, None] * cz[None, None, :])
        lab = part.argmax(0).to(torch.uint8)
        cur, cs = torch.cuda.current_stream(device), _copy_stream(device)
        ev = torch.cuda.Event()
        ev.record(cur)
        cs.wait_event(ev)
        with torch.cuda.stream(cs):
            out_host[fin:x1].copy_(lab, non_blocking=True)
        keep.append(lab)            # alive until the final synchronisation
        fin = x1

    for g in range(0, len(mine), sw_batch_size):
        ids = mine[g:g + sw_batch_size]
        if slab_ready is not None:
            # blocks are uploaded in (first axis, second axis) order = the order of the windows: everything up to the last
            # block this batch touches
            need = max((starts[i][0] + roi_size[0], starts[i][1] + roi_size[1]) for i in ids)
            for _, ev in blocks_ready(slab_ready, need):
                torch.cuda.current_stream(device).wait_event(ev)
        win = torch.cat([inputs[:, :, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]]
                         for a, b, c in (starts[i] for i in ids)])
        y = _logits(predictor(win))
        if acc is None:
            acc = torch.zeros((y.shape[1], world * xs, image[1], image[2]), dtype=y.dtype, device=y.device)
        for k, i in enumerate(ids):
            a, b, c = starts[i]
            acc[:, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]] += y[k]
        if stream_out:
            rest = mine[g + sw_batch_size:]
            done_to = min(starts[i][0] for i in rest) if rest else image[0]
            if done_to > fin:
                finish_planes(done_to)
    if stream_out and acc is not None:
        _copy_stream(device).synchronize()
        return out_host
    if acc is None:         # more ranks than windows
        n_cls = _logits(predictor(inputs[:, :, :roi_size[0], :roi_size[1], :roi_size[2]])).shape[1]
        acc = torch.zeros((n_cls, world * xs, image[1], image[2]), dtype=inputs.dtype, device=device)
    n_cls = acc.shape[0]
    if world > 1:
        send = acc.view(n_cls, world, xs, image[1], image[2]).permute(1, 0, 2, 3, 4).contiguous()
        slab = torch.empty((n_cls, xs, image[1], image[2]), dtype=acc.dtype, device=acc.device)
        dist.reduce_scatter_tensor(slab.view(-1), send.view(-1), op=dist.ReduceOp.SUM, group=group)
    else:
        slab = acc
    cx, cy, cz = axis_counts(image, roi_size, per_axis, slab.device, slab.dtype)
    cx = torch.cat([cx, cx.new_ones(world * xs - image[0])])[rank * xs:(rank + 1) * xs]     # padding planes: count 1
    slab = slab / (cx[:, None, None] * cy[None, :, None] * cz[None, None, :])
    lab = slab.argmax(0).to(torch.uint8)
    if world > 1:
        full = torch.empty((world * xs, image[1], image[2]), dtype=torch.uint8, device=lab.device)
        dist.all_gather_into_tensor(full.view(-1), lab.contiguous().view(-1), group=group)
        if rank != dst:
            return None
    else:
        full = lab
    full = full[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]]
    if out_host is None:
        return full.cpu()
    out_host.copy_(full, non_blocking=True)
    if full.is_cuda:
        torch.cuda.current_stream(full.device).synchronize()
    return out_host


class GraphedPredictor:
    """Eval forward of a fixed window-batch shape captured once into a CUDA graph and replayed per batch of windows
    (an eager eval forward is ~300 launches for ~2 ms of GPU work: host-bound).  Smaller final batches are padded."""

    def __init__(self, model: torch.nn.Module, batch: int, channels: int, roi_size: Sequence[int], device):
        self.model = model.eval()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):             # the library launches on the current device
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
        with torch.cuda.device(self.device):
            self.x[:k].copy_(win)
            self.graph.replay()
        return self.y[:k]
