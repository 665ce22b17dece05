"""Data-parallel training step on 2 GPUs (skipped on a single-GPU box): NCCL, one rank per GPU.  Every loss term is a
batch mean and every norm is per-sample, so two ranks with 2 patches each must take the same optimisation step as one rank
with the 4 patches.  Three variants of the exchange: `graph` (default: forward, backward, NCCL all-reduce and AdamW in ONE
captured graph), `split` (VX_DP_GRAPH=split: two graphs around an eager all-reduce) and `eager` (GradBuckets: bucketed
all-reduces launched from the grad-ready hooks while the model's forked streams are still producing gradients)."""
import os
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 2, 64, 64, 64, generator=g)
    y = (torch.rand(4, 1, 64, 64, 64, generator=g) > 0.9).long()
    return x, y


def _model():
    from tests import _golden as G
    from veloxseg_b200.configs import MODEL_CONFIGS
    from veloxseg_b200.nn import VeloxSeg
    torch.manual_seed(3)
    m = VeloxSeg(**MODEL_CONFIGS["tiny"])
    G.zero_dropout(m)
    return m


def _worker(rank, port, out, mode):
    import torch.distributed as dist
    os.environ["VX_DP_GRAPH"] = "split" if mode == "split" else "one"
    from veloxseg_b200.train import TrainStep
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=2, device_id=torch.device("cuda", rank))
    try:
        x, y = _batch()
        ts = TrainStep(_model(), 2, f"cuda:{rank}", use_graph=mode != "eager", bucket_bytes=256 << 10)
        losses = [ts.step(x[2 * rank:2 * rank + 2], y[2 * rank:2 * rank + 2], sync=True) for _ in range(3)]
        flat = torch.cat([p.detach().flatten() for p in ts.model.parameters()]).cpu()
        torch.save((losses, flat), out + f".{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["graph", "split", "eager"])
def test_two_rank_step_equals_single_rank_full_batch(tmp_path, mode):
    import torch.multiprocessing as mp
    from veloxseg_b200.train import TrainStep
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(_free_port(), out, mode), nprocs=2, join=True)
    x, y = _batch()
    ts = TrainStep(_model(), 2, "cuda:0", use_graph=True)
    ref_losses = [ts.step(x, y, sync=True) for _ in range(3)]
    ref = torch.cat([p.detach().flatten() for p in ts.model.parameters()]).cpu()
    (l0, p0), (l1, p1) = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(p0, p1)                                     # ranks stay bit-identical
    assert float((p0 - ref).norm()) <= 2e-4 * float(ref.norm()), float((p0 - ref).norm() / ref.norm())
    for a, b, c in zip(l0, l1, ref_losses):                         # full-batch loss = mean of the two half-batch losses
        assert abs(0.5 * (a + b) - c) <= 2e-3 * abs(c), (l0, l1, ref_losses)
