"""World-size-2 checks of the multi-GPU host logic on the CPU `gloo` backend (no CUDA involved):

* GradBuckets (veloxseg_b200/train.py): bucketed asynchronous all-reduce launched from the grad-ready hooks gives
  every rank the mean of the per-rank gradients, including a parameter that receives no gradient on one step;
* sliding_window_predict (veloxseg_b200/inference.py): windows sharded round-robin over ranks + one all-reduce of the
  partial logit sums equals the single-process result on every rank (ragged volume that needs padding, a volume with
  fewer windows than ranks, batch > 1).

* sliding_window_labels: the sharded-IO variant (1/world of the host volume per rank + all-gather, reduce-scatter of the
  partial sums in slabs, arg-max per slab, label gather) gives rank 0 the arg-max of the single-process blend.

The predictor here is a plain torch function: the product model needs the sm_100a library and is covered by -m gpu.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)


def _small_net():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Conv3d(2, 6, 3, padding=1), torch.nn.GELU(), torch.nn.Conv3d(6, 3, 1))


def _rank_batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(2, 2, 6, 6, 6, generator=g)


def _grad_worker(rank, port, out):
    from veloxseg_b200.train import GradBuckets
    _init(rank, port)
    try:
        net = _small_net()
        unused = torch.nn.Parameter(torch.ones(5))            # never touched by the loss: its bucket slot stays zero
        params = list(net.parameters()) + [unused]
        gb = GradBuckets(params, bucket_bytes=64)               # tiny buckets -> several all-reduces in flight
        assert gb.world == WORLD and len(gb.buckets) >= 2
        res = []
        for step in range(2):                                   # second step checks zero()/re-arm
            gb.zero()
            loss = net(_rank_batch(rank) * (step + 1)).square().mean()
            loss.backward()
            gb.finish()
            res.append(torch.cat([p.grad.flatten().clone() for p in params]))
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def test_grad_buckets_allreduce_mean(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    got = torch.load(out)
    for step in range(2):
        ref = None
        for r in range(WORLD):
            net = _small_net()
            net(_rank_batch(r) * (step + 1)).square().mean().backward()
            g = torch.cat([p.grad.flatten() for p in net.parameters()] + [torch.zeros(5)])
            ref = g if ref is None else ref + g
        ref = ref / WORLD
        assert torch.allclose(got[step], ref, rtol=1e-5, atol=1e-7), (step, float((got[step] - ref).abs().max()))


def _predictor():
    torch.manual_seed(11)
    conv = torch.nn.Conv3d(2, 3, 3, padding=1)
    for p in conv.parameters():
        p.requires_grad_(False)
    return lambda w: [conv(w)]          # list output, like Net.forward on a train-mode model (inference_petct.py:46-51)


CASES = [  # (volume shape, roi)
    ((1, 2, 20, 17, 9), (8, 8, 8)),     # ragged: clamped last windows on every axis
    ((1, 2, 6, 8, 8), (8, 8, 8)),       # needs symmetric padding; a single window (< WORLD)
    ((2, 2, 12, 8, 10), (8, 8, 4)),     # batch of 2 volumes
]


def _sw_worker(rank, port, out):
    from veloxseg_b200.inference import sliding_window_predict
    _init(rank, port)
    try:
        pred = _predictor()
        res = []
        for shape, roi in CASES:
            x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
            res.append(sliding_window_predict(x, pred, roi, sw_batch_size=2, overlap=0.25))
        torch.save(res, out + f".{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_sliding_window_matches_single_process(tmp_path):
    from veloxseg_b200.inference import sliding_window_predict
    out = str(tmp_path / "sw.pt")
    mp.spawn(_sw_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    pred = _predictor()
    for ci, (shape, roi) in enumerate(CASES):
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
        ref = sliding_window_predict(x, pred, roi, sw_batch_size=2, overlap=0.25, shard=False)
        assert ref.shape == (shape[0], 3) + tuple(shape[2:])
        for r in range(WORLD):
            got = torch.load(out + f".{r}")[ci]
            assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6), (ci, r, float((got - ref).abs().max()))


LABEL_CASES = CASES[:2] + [((1, 2, 21, 16, 12), (8, 8, 4))]      # odd first axis: slab padding plane; 3 x 3 x 4 windows


def _labels_worker(rank, port, out):
    from veloxseg_b200.inference import sliding_window_labels
    _init(rank, port)
    try:
        pred = _predictor()
        res = []
        for shape, roi in LABEL_CASES:
            x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
            res.append(sliding_window_labels(x, pred, roi, "cpu", sw_batch_size=2, overlap=0.25))
        torch.save(res, out + f".{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_io_labels_match_single_process_argmax(tmp_path):
    from veloxseg_b200.inference import sliding_window_labels, sliding_window_predict
    out = str(tmp_path / "lab.pt")
    mp.spawn(_labels_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    pred = _predictor()
    got0, got1 = torch.load(out + ".0"), torch.load(out + ".1")
    for ci, (shape, roi) in enumerate(LABEL_CASES):
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
        ref = sliding_window_predict(x, pred, roi, sw_batch_size=2, overlap=0.25, shard=False)[0]
        single = sliding_window_labels(x, pred, roi, "cpu", sw_batch_size=2, overlap=0.25)     # no process group here
        assert single.dtype == torch.uint8 and torch.equal(single, ref.argmax(0).to(torch.uint8))    # same sum order: bit-exact
        assert got1[ci] is None and got0[ci].shape == tuple(shape[2:])
        top2 = ref.topk(2, dim=0).values
        decided = (top2[0] - top2[1]) > 1e-5             # the sharded sum folds windows in another order: near-ties may flip
        assert torch.equal(got0[ci][decided], single[decided]), ci
        assert float((got0[ci] != single).float().mean()) < 1e-3
