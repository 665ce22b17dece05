"""Data-parallel training step (new capability; the reference trains on one device, utils/train_autopet.py:219-268).

Step semantics follow the reference loop: inputs = cat(modalities) (B, sum(in_ch), *patch) -> model(train) ->
Loss(output, labels, sr_labels=inputs) -> backward -> AdamW step (lr 2.5e-4, wd 0.01).  One process per GPU; every
norm in the model is per-sample, so ranks exchange nothing but gradients: `GradBuckets` lays the parameters out in a
few flat fp32 buckets (reverse registration order ~ backward completion order), exposes `.grad` as views into them and
launches one asynchronous all-reduce per bucket from the autograd hooks while the rest of backward is still running
(eager path).  The default path captures the whole step -- forward, backward, the gradient all-reduce (NCCL is captured
into the graph) and AdamW -- as ONE CUDA graph per rank.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .loss import Loss


class GradBuckets:
    """Flat gradient buckets with all-reduce launched as each bucket's last gradient is produced."""

    def __init__(self, params: List[torch.nn.Parameter], bucket_bytes: int = 4 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        order = list(reversed(self.params))                 # decoders finish first in backward
        self.buckets: List[torch.Tensor] = []
        self.bucket_of, self.pending, self.total = {}, [], []
        if self.world == 1:          # nothing to exchange: leave .grad to autograd (no accumulate kernels at all)
            self.works = []
            return
        cur, cur_bytes = [], 0
        groups = []
        for p in order:
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        for bi, g in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in g), dtype=g[0].dtype, device=g[0].device)
            off = 0
            for p in g:
                p.grad = flat[off:off + p.numel()].view_as(p)        # gradients accumulate straight into the bucket
                off += p.numel()
                self.bucket_of[p] = bi
            self.buckets.append(flat)
            self.total.append(len(g))
        self.pending = list(self.total)
        self.works: List[Optional[object]] = [None] * len(self.buckets)
        # The model runs branches on forked streams and autograd replays every AccumulateGrad on its forward stream, so the
        # gradients of one bucket are produced on several streams: each hook records an event on ITS stream and the
        # all-reduce waits for all of the bucket's events (NCCL itself orders only against the launching stream).
        self.events: List[list] = [[] for _ in self.buckets]
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._ready)

    def zero(self):
        if self.world == 1:
            for p in self.params:
                p.grad = None
            return
        for b in self.buckets:
            b.zero_()
        self.pending = list(self.total)
        self.works = [None] * len(self.buckets)
        self.events = [[] for _ in self.buckets]

    def _ready(self, p):
        bi = self.bucket_of[p]
        self.pending[bi] -= 1
        if p.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(p.device))
            self.events[bi].append(ev)
        if self.pending[bi] == 0:
            if p.is_cuda:
                cur = torch.cuda.current_stream(p.device)
                for ev in self.events[bi]:
                    cur.wait_event(ev)
            self.works[bi] = dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for the outstanding all-reduces and turn sums into means (equal B_local on every rank)."""
        if self.world == 1:
            return
        for bi, w in enumerate(self.works):
            if w is None:       # a bucket whose parameters received no gradient this step
                if self.buckets[bi].is_cuda:
                    cur = torch.cuda.current_stream(self.buckets[bi].device)
                    for ev in self.events[bi]:
                        cur.wait_event(ev)
                w = dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            w.wait()
            self.buckets[bi].div_(self.world)


class VxAdamW:
    """torch.optim.AdamW (the reference's optimiser: decoupled weight decay, bias correction, betas (0.9, 0.999), eps
    1e-8) as ONE launch of libveloxseg over every parameter tensor.  State = two flat fp32 moment buffers, a device step
    counter and a device (lr, weight_decay) pair, so a captured CUDA graph replays it unchanged and `set_lr` (the
    reference's schedulers, utils/train_*.py) takes effect inside the replayed graph.

    The (parameter, gradient) pointer table lives on the device and is refreshed from pinned host memory only when it
    changes (autograd hands out new gradient tensors in eager mode).  Uploads rotate through a ring of pinned tables, each
    guarded by an event, so a table is never rewritten while its asynchronous copy may still be pending; a table uploaded
    during stream capture is private to that graph and never rewritten (the graph's copy node re-reads it at every replay).
    `grad_source(flat, params)` points the gradient column into a flat buffer instead (the all-reduced one of TrainStep)."""
    CHUNK = 1024
    RING = 4

    def __init__(self, params, lr: float, weight_decay: float, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        self.lr, self.wd, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        self.grad_scale = float(grad_scale)
        total, offs, chunks = 0, [], []
        for i, p in enumerate(self.params):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("VxAdamW: parameters must be contiguous float32")
            offs.append(total)
            chunks += [(i, s) for s in range(0, p.numel(), self.CHUNK)]
            total += p.numel()
        self.offsets = offs
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.hyper = torch.tensor([self.lr, self.wd], dtype=torch.float32, device=dev)
        self.chunks = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).to(dev)
        self.n_chunks = len(chunks)
        self._ring = [torch.zeros(len(self.params), 4, dtype=torch.int64).pin_memory() for _ in range(self.RING)]
        self._ring_ev = [None] * self.RING
        self._ring_i = 0
        self._graph_tabs = []            # pinned tables owned by captured graphs: kept alive, never rewritten
        self._tab_dev = torch.zeros(len(self.params), 4, dtype=torch.int64, device=dev)
        self._tab_last = None            # host copy of what _tab_dev holds (None: unknown)
        self._grad_src = None

    # ---- hyper-parameters / state (torch.optim.Optimizer-like surface)
    @property
    def param_groups(self):
        return [{"params": self.params, "lr": self.lr, "weight_decay": self.wd, "betas": self.betas, "eps": self.eps}]

    def set_lr(self, lr: float, weight_decay: Optional[float] = None):
        """Takes effect at the next step, also inside an already captured graph (the kernel reads the device pair)."""
        self.lr = float(lr)
        if weight_decay is not None:
            self.wd = float(weight_decay)
        self.hyper.copy_(torch.tensor([self.lr, self.wd], dtype=torch.float32))

    def state_tensors(self):
        return [self.exp_avg, self.exp_avg_sq, self.step_t]

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_t.clone(),
                "lr": self.lr, "weight_decay": self.wd, "betas": tuple(self.betas), "eps": self.eps}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_t.copy_(sd["step"])
        self.betas, self.eps = tuple(sd["betas"]), float(sd["eps"])
        self.set_lr(sd["lr"], sd["weight_decay"])

    def grad_source(self, flat: Optional[torch.Tensor], params=None, append: bool = False):
        """Read the gradients of `params` (in order) from consecutive slices of `flat` instead of `p.grad`; None resets;
        `append=True` adds a second flat buffer (parameters not named by any buffer are skipped by the step)."""
        if flat is None:
            self._grad_src = None
            return
        if flat.dtype != torch.float32 or not flat.is_contiguous():
            raise TypeError("VxAdamW: the flat gradient buffer must be contiguous float32")
        ptrs, off = {}, 0
        for p in params:
            ptrs[id(p)] = flat.data_ptr() + 4 * off
            off += p.numel()
        if off != flat.numel():
            raise ValueError("VxAdamW.grad_source: flat buffer size does not match the parameter list")
        if append and self._grad_src is not None:
            self._grad_src.update(ptrs)
        else:
            self._grad_src = ptrs

    def _table(self):
        rows = []
        for i, p in enumerate(self.params):
            if self._grad_src is not None:
                gp = self._grad_src.get(id(p), 0)
            else:
                g = p.grad
                if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                    raise TypeError("VxAdamW: gradients must be contiguous float32")
                gp = g.data_ptr() if g is not None else 0
            rows.append((p.data_ptr(), gp, self.offsets[i], p.numel()))
        return rows

    def step(self):
        from . import _lib
        from ._lib import AdamwDesc
        dev = self.step_t.device
        rows = self._table()
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            tab = torch.tensor(rows, dtype=torch.int64).pin_memory()       # private to this graph, immutable
            self._graph_tabs.append(tab)
            self._tab_dev.copy_(tab, non_blocking=True)
            self._tab_last = None                                          # replays rewrite _tab_dev behind our back
        elif rows != self._tab_last:
            i = self._ring_i
            self._ring_i = (i + 1) % self.RING
            if self._ring_ev[i] is not None:
                self._ring_ev[i].synchronize()                             # its previous upload has been consumed
            self._ring[i].copy_(torch.tensor(rows, dtype=torch.int64))
            self._tab_dev.copy_(self._ring[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._ring_ev[i] = ev
            self._tab_last = rows if not self._graph_tabs else None        # a graph replay may overwrite the device table
        lib = _lib.get_lib()
        d = AdamwDesc(self.n_chunks, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.grad_scale, 1)
        lib.call("vx_adamw_step", d, [self._tab_dev, self.chunks, self.hyper], [self.exp_avg, self.exp_avg_sq, self.step_t],
                 torch.cuda.current_stream(dev).cuda_stream)


class TrainStep:
    """`step(inputs, labels)` = one optimisation step; accepts host (pinned) or device tensors and returns the loss
    as a Python float only when asked (`sync=True`), mirroring the reference's per-step `loss.item()`.

    On CUDA the step is captured once into CUDA graphs and replayed: ~1 200 kernel launches per step collapse into one
    or two graph launches, which matters because the fused step is short enough for host launch overhead to dominate
    (eager enqueue of one step takes about twice as long as the GPU needs to run it).  Dropout stays random across replays
    because the kernels add a device-resident offset (ops.advance_seed, captured in the graph) to their seeds.

      world == 1   one graph: forward, loss, backward, AdamW.
      world  > 1   one graph: forward, loss, backward, the gradients gathered into one flat fp32 buffer (9 MB), the NCCL
                   all-reduce of that buffer CAPTURED in the graph, AdamW reading the summed gradients straight from the flat
                   buffer with grad_scale = 1 / world (no scatter back, no second graph, no host launch between them).
                   `VX_DP_GRAPH=split` keeps the older form (graph A: forward / backward / gather; eager all-reduce;
                   graph B: AdamW) for A/B measurements and for NCCL builds that cannot be captured.
    `use_graph=False` keeps the eager path, where `GradBuckets` overlaps bucketed all-reduces with backward.
    `set_lr` changes the learning rate of the next step, also of an already captured graph."""

    def __init__(self, model: torch.nn.Module, num_modal: int, device, lr: float = 2.5e-4, weight_decay: float = 0.01,
                 deep_weights=(1, 1, 1, 1), rc_weight: float = 0.5, feature_weight: float = 2.0,
                 bucket_bytes: int = 4 << 20, use_graph: Optional[bool] = None):
        self.device = torch.device(device)
        self.model = model.to(self.device).train()
        self.loss_fn = Loss(num_modal, deep_weights, rc_weight, feature_weight)
        cuda = self.device.type == "cuda"
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.use_graph = cuda if use_graph is None else bool(use_graph)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        # eager data-parallel path only: gradients accumulate straight into flat buckets, all-reduce from grad hooks
        self.buckets = GradBuckets(self.params, bucket_bytes) if not self.use_graph else None
        # libveloxseg's one-launch AdamW on CUDA (VX_TORCH_ADAMW=1 keeps torch's fused multi-tensor optimiser for A/B)
        import os
        self.split_graph = os.environ.get("VX_DP_GRAPH", "one") == "split"
        # VX_DP_OVERLAP=1: two-phase backward with the decoder-side all-reduce under the encoder backward (world > 1, one-graph
        # form).  Measured SLOWER on 2 x B200 (6.93 vs 6.59 ms/step, profiles/r3j_dp_overlap_2gpu.txt): phase 1 has to join every
        # forked decoder stream before the encoder backward may start, which costs more than the ~0.15 ms of exposed all-reduce
        # it hides.  Off by default; kept as an A/B switch (parity: tests/test_gpu_dp.py passes with it on).
        self.overlap = os.environ.get("VX_DP_OVERLAP", "0") == "1" and hasattr(self.model, "encoder")
        if cuda and os.environ.get("VX_TORCH_ADAMW", "0") != "1":
            self.opt = VxAdamW(self.model.parameters(), lr=lr, weight_decay=weight_decay)
        else:
            self.opt = torch.optim.AdamW(self.model.parameters(), lr=lr, weight_decay=weight_decay, fused=cuda,
                                         capturable=cuda and self.use_graph)
        self._graph = None
        self._graph_b = None
        self.graph_launches = 0        # kernels of libveloxseg_sm100 recorded in the captured step

    # ---- pieces shared by the eager and the captured paths
    def _fwd_bwd(self, x, y):
        from . import ops
        ops.advance_seed(self.device)
        out = self.model(x)
        loss = self.loss_fn(out, y, x)
        loss.backward()
        return loss.detach()

    def _gather_grads(self):
        self._gparams = [p for p in self.params if p.grad is not None]
        grads = [p.grad for p in self._gparams]
        return grads, torch.cat([g.reshape(-1) for g in grads])

    def _scatter_mean(self, grads, flat):
        flat.div_(self.world)
        torch._foreach_copy_(grads, [f.view_as(g) for f, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def close(self):
        """Releases the captured graphs.  With world > 1 the step graph holds the captured NCCL all-reduce: release it (or drop
        the TrainStep) BEFORE `dist.destroy_process_group()`, which otherwise waits on the communicator the graph still owns
        (observed as a hang at interpreter exit in tests/dp_worker.py on 2 GPUs)."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        for name in ("_graph", "_graph_b"):
            g = getattr(self, name, None)
            if g is not None:
                g.reset()
            setattr(self, name, None)
        self._staged = None

    def set_lr(self, lr: float):
        if isinstance(self.opt, VxAdamW):
            self.opt.set_lr(lr)
        else:
            for g in self.opt.param_groups:
                if torch.is_tensor(g["lr"]):
                    g["lr"].fill_(lr)
                else:
                    g["lr"] = lr

    def _fwd_bwd_overlapped(self, x, y):
        """world > 1: backward in two phases at the encoder / decoder boundary.  Phase 1 differentiates the loss with respect
        to the decoder-side parameters (segmentation decoder, reconstruction teachers, heads) and the boundary tensors; the
        all-reduce of those gradients (about half of the 9 MB) is then in flight on a communication stream while phase 2
        runs the encoder backward.  Returns (loss, [(flat, params), ...]) with both all-reduces issued."""
        from . import ops
        ops.advance_seed(self.device)
        self.model.split_boundary = True
        try:
            out = self.model(x)
        finally:
            self.model.split_boundary = False
        loss = self.loss_fn(out, y, x)
        enc_ids = {id(p) for p in self.model.encoder.parameters()}
        dec_params = [p for p in self.params if id(p) not in enc_ids]
        enc_params = [p for p in self.params if id(p) in enc_ids]
        boundary, boundary_src = self.model._boundary, self.model._boundary_src
        g = torch.autograd.grad(loss, dec_params + boundary, allow_unused=True)
        gd, gb = g[:len(dec_params)], g[len(dec_params):]
        live = [(p, gi) for p, gi in zip(dec_params, gd) if gi is not None]
        flat_a = torch.cat([gi.reshape(-1) for _, gi in live])
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        comm = self._comm_stream
        comm.wait_stream(cur)
        flat_a.record_stream(comm)
        with torch.cuda.stream(comm):
            dist.all_reduce(flat_a, op=dist.ReduceOp.SUM)
        bt = [(t, gt) for t, gt in zip(boundary_src, gb) if gt is not None and t.requires_grad]
        torch.autograd.backward([t for t, _ in bt], [gt for _, gt in bt])
        enc_live = [p for p in enc_params if p.grad is not None]
        flat_b = torch.cat([p.grad.reshape(-1) for p in enc_live])
        dist.all_reduce(flat_b, op=dist.ReduceOp.SUM)
        cur.wait_stream(comm)
        return loss.detach(), [(flat_a, [p for p, _ in live]), (flat_b, enc_live)]

    def _reduce_and_step_flat(self):
        """world > 1, own optimiser: gather -> all-reduce(sum) -> AdamW reads the flat buffer scaled by 1 / world."""
        self._gparams = [p for p in self.params if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for p in self._gparams])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        self.opt.grad_scale = 1.0 / self.world
        self.opt.grad_source(flat, self._gparams)
        self.opt.step()
        return flat

    def _step_eager(self, x, y):
        """One full step without graphs (also the warm-up before capture and the profiler's per-kernel pass)."""
        if self.buckets is not None:
            self.buckets.zero()
            loss = self._fwd_bwd(x, y)
            self.buckets.finish()
        else:
            for p in self.params:
                p.grad = None
            if self.world > 1 and isinstance(self.opt, VxAdamW) and not self.split_graph and self.overlap:
                loss, parts = self._fwd_bwd_overlapped(x, y)
                self.opt.grad_scale = 1.0 / self.world
                for i, (flat, ps) in enumerate(parts):
                    self.opt.grad_source(flat, ps, append=i > 0)
                self.opt.step()
                self._flat = parts
                return loss
            loss = self._fwd_bwd(x, y)
            if self.world > 1:
                if isinstance(self.opt, VxAdamW) and not self.split_graph:
                    self._flat = self._reduce_and_step_flat()
                    return loss
                grads, flat = self._gather_grads()
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                self._scatter_mean(grads, flat)
        self.opt.step()
        return loss

    def _capture(self, inputs, labels):
        from . import _lib
        self._sx = torch.empty(inputs.shape, dtype=inputs.dtype, device=self.device)
        self._sy = torch.empty(labels.shape, dtype=labels.dtype, device=self.device)
        self._sx.copy_(inputs)
        self._sy.copy_(labels)
        # the warm-up iterations below are real optimiser steps: snapshot and restore so that capture is side-effect free
        snap_p = [p.detach().clone() for p in self.model.parameters()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step_eager(self._sx, self._sy)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        lib = _lib.get_lib()
        n0 = lib.c.vx_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        if self.world == 1 or (isinstance(self.opt, VxAdamW) and not self.split_graph):
            with torch.cuda.graph(self._graph):       # world > 1: the NCCL all-reduce is one of the captured nodes
                self._sloss = self._step_eager(self._sx, self._sy)
        else:
            for p in self.params:
                p.grad = None
            with torch.cuda.graph(self._graph):
                self._sloss = self._fwd_bwd(self._sx, self._sy)
                self._grads, self._flat = self._gather_grads()
            self._graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_b, pool=self._graph.pool()):
                self._scatter_mean(self._grads, self._flat)
                self.opt.step()
        self.graph_launches = int(lib.c.vx_launch_count() - n0)
        with torch.no_grad():
            for p, q in zip(self.model.parameters(), snap_p):
                p.copy_(q)
            if isinstance(self.opt, VxAdamW):
                for v in self.opt.state_tensors():
                    v.zero_()
            else:
                for st in self.opt.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        torch.cuda.synchronize(self.device)

    # ---- host -> device staging: pinned host batches are copied on a side stream into a staging pair, so the copy of
    # batch i+1 (given as `prefetch=`) overlaps the replay of batch i; labels may arrive as uint8 (7/8 fewer bytes).
    def _stage(self, inputs, labels):
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage_x = torch.empty_like(self._sx)
            self._stage_y = torch.empty(self._sy.shape, dtype=labels.dtype, device=self.device)
            self._stage_ev = torch.cuda.Event()
            self._consumed_ev = None
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))      # once: the blocks may be recycled ones
        if self._stage_y.dtype != labels.dtype:
            self._stage_y = torch.empty(self._sy.shape, dtype=labels.dtype, device=self.device)
        cs = self._copy_stream
        if self._consumed_ev is not None:
            cs.wait_event(self._consumed_ev)      # the staging pair has been read out (NOT the whole step: the copy must overlap it)
        with torch.cuda.stream(cs):
            self._stage_x.copy_(inputs, non_blocking=True)
            self._stage_y.copy_(labels, non_blocking=True)
            self._stage_ev.record(cs)
        self._staged = (inputs, labels)

    def step(self, inputs: torch.Tensor, labels: torch.Tensor, sync: bool = False, prefetch=None):
        """`prefetch=(next_inputs, next_labels)`: start copying the next host batch while this step runs."""
        if self.device.type == "cuda":
            with torch.cuda.device(self.device):      # the library launches on the current device
                return self._step(inputs, labels, sync, prefetch)
        return self._step(inputs, labels, sync, prefetch)

    def _step(self, inputs, labels, sync, prefetch):
        if self.use_graph:
            if self._graph is None:
                self._capture(inputs, labels.long() if labels.dtype != torch.int64 else labels)
            staged = getattr(self, "_staged", None)
            if staged is not None and staged[0] is inputs and staged[1] is labels:
                torch.cuda.current_stream(self.device).wait_event(self._stage_ev)
                self._sx.copy_(self._stage_x)
                self._sy.copy_(self._stage_y)                          # uint8 -> int64 on the device
                self._consumed_ev = torch.cuda.Event()
                self._consumed_ev.record(torch.cuda.current_stream(self.device))
                self._staged = None
            else:
                self._sx.copy_(inputs, non_blocking=True)
                if labels.dtype != torch.int64 and not labels.is_cuda:
                    self._sy.copy_(labels.to(self.device, non_blocking=True))
                else:
                    self._sy.copy_(labels, non_blocking=True)
            self._graph.replay()
            if self._graph_b is not None:
                dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
                self._graph_b.replay()
            if prefetch is not None:
                self._stage(*prefetch)
            loss = self._sloss if sync else self._sloss.clone()      # the static tensor is overwritten by the next replay
        else:
            loss = self._step_eager(inputs.to(self.device, non_blocking=True), labels.to(self.device, non_blocking=True).long())
        return float(loss.item()) if sync else loss
