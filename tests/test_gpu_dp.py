"""Data-parallel training step on 2 GPUs (skipped on a single-GPU box): NCCL, one rank per GPU.  Every loss term is a
batch mean and every norm is per-sample, so two ranks with 2 patches each must take the same optimisation step as one rank
with the 4 patches.  Three variants of the exchange: `graph` (default: forward, backward, NCCL all-reduce and AdamW in ONE
captured graph), `split` (VX_DP_GRAPH=split: two graphs around an eager all-reduce) and `eager` (GradBuckets: bucketed
all-reduces launched from the grad-ready hooks while the model's forked streams are still producing gradients).
The ranks are launched the way bench.py's are (torch.distributed.run, rendezvous on 127.0.0.1) with a hard time limit, so
a wedged collective fails the test in minutes instead of hanging the suite."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["graph", "split", "eager"])
def test_two_rank_step_equals_single_rank_full_batch(tmp_path, mode):
    from tests.dp_worker import batch, model
    from veloxseg_b200.train import TrainStep
    out = str(tmp_path / "dp.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py"), mode, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    x, y = batch()
    ts = TrainStep(model(), 2, "cuda:0", use_graph=True)
    ref_losses = [ts.step(x, y, sync=True) for _ in range(3)]
    ref = torch.cat([p.detach().flatten() for p in ts.model.parameters()]).cpu()
    (l0, p0), (l1, p1) = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(p0, p1)                                     # ranks stay bit-identical
    assert float((p0 - ref).norm()) <= 2e-4 * float(ref.norm()), float((p0 - ref).norm() / ref.norm())
    for a, b, c in zip(l0, l1, ref_losses):                         # full-batch loss = mean of the two half-batch losses
        assert abs(0.5 * (a + b) - c) <= 2e-3 * abs(c), (l0, l1, ref_losses)
