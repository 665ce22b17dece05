"""One rank of the 2-GPU data-parallel check (launched by tests/test_gpu_dp.py through torch.distributed.run):
    python -m torch.distributed.run --nproc-per-node 2 ... tests/dp_worker.py <mode> <out_prefix>
mode: graph (one captured graph incl. the NCCL all-reduce) | split (two graphs around an eager all-reduce) | eager (GradBuckets)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def batch():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 2, 64, 64, 64, generator=g)
    y = (torch.rand(4, 1, 64, 64, 64, generator=g) > 0.9).long()
    return x, y


def model():
    from tests import _golden as G
    from veloxseg_b200.configs import MODEL_CONFIGS
    from veloxseg_b200.nn import VeloxSeg
    torch.manual_seed(3)
    m = VeloxSeg(**MODEL_CONFIGS["tiny"])
    G.zero_dropout(m)
    return m


def main():
    mode, out = sys.argv[1], sys.argv[2]
    os.environ["VX_DP_GRAPH"] = "split" if mode == "split" else "one"
    from veloxseg_b200.train import TrainStep
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        x, y = batch()
        n = x.shape[0] // world
        ts = TrainStep(model(), 2, f"cuda:{local}", use_graph=mode != "eager", bucket_bytes=256 << 10)
        losses = [ts.step(x[n * rank:n * rank + n], y[n * rank:n * rank + n], sync=True) for _ in range(3)]
        flat = torch.cat([p.detach().flatten() for p in ts.model.parameters()]).cpu()
        torch.save((losses, flat), out + f".{rank}")
        print(f"rank {rank} mode {mode} losses {losses}", flush=True)
        ts.close()           # the step graph holds the captured NCCL all-reduce: release it before the process group goes
        del ts
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
