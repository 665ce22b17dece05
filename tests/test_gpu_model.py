"""Whole-model parity on the GPU: veloxseg_b200.nn.VeloxSeg (custom ops) against vectors produced by the unmodified
reference for the three reference configs (+ a miniature), fp32: eval logits, train-mode outputs, loss and
per-parameter gradients within the north_star tolerance (1e-3 relative; atol for structurally-zero gradients)."""
import pytest
import torch

from tests import _golden as G
from veloxseg_b200.configs import MODEL_CONFIGS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(name):
    from veloxseg_b200.nn import VeloxSeg
    torch.manual_seed(G.MODEL_SEED)
    m = VeloxSeg(**MODEL_CONFIGS[name])
    G.zero_dropout(m)
    return m.to(DEV)


@pytest.mark.parametrize("name", ["tiny", "autopetii", "hecktor2022", "brats2021"])
def test_whole_model_vs_reference(name):
    """The mode bench.py times: every kernel of the model is libveloxseg's (no cuDNN / cuBLAS call on the path), fp32-accurate
    (SIMT fp32 or tcgen05 3xTF32).  north_star tolerance: 1e-3 on outputs and per-parameter gradients."""
    gtol = 1e-3
    fx = G.load(f"model_{name}.pt")
    cfg = MODEL_CONFIGS[name]
    m = _model(name)
    x = G.model_input(cfg, fx["B"]).to(DEV)
    m.eval()
    with torch.no_grad():
        G.check_sample(m(x), fx["eval"], rtol=gtol, what="eval logits")
    m.train()
    outs = m(x)
    assert len(outs) == len(fx["train_outputs"])          # [seg x4, rcs, gram_s, gram_t x M]
    for i, (o, r) in enumerate(zip(outs, fx["train_outputs"])):
        G.check_sample(o, r, rtol=1e-3, atol=1e-7, what=f"train output {i}")
    loss = sum((o * c.to(DEV)).sum() for o, c in zip(outs, G.cotangents(outs)))
    assert abs(float(loss.detach()) - fx["loss"]) < gtol * max(1.0, abs(fx["loss"]))
    loss.backward()
    bad = []
    for k, p in m.named_parameters():
        try:
            G.check_sample(p.grad if p.grad is not None else torch.zeros_like(p), fx["grads"][k], rtol=gtol, atol=1e-5, what=k)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, bad[:10]


def test_dropout_train_mode_runs_and_differs():
    """Train mode with the config's dropout rates: finite outputs, different masks per step, eval unaffected."""
    from veloxseg_b200.nn import VeloxSeg
    torch.manual_seed(0)
    m = VeloxSeg(**MODEL_CONFIGS["tiny"]).to(DEV)
    x = torch.randn(2, 2, 64, 64, 64, device=DEV)
    m.train()
    a = m(x)[0]
    b = m(x)[0]
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    m.eval()
    with torch.no_grad():
        assert torch.equal(m(x), m(x))


def test_train_step_graph_matches_eager_and_redraws_dropout():
    """TrainStep: the CUDA-graph replay is the same computation as the eager step (dropout off), capture leaves the
    model untouched, and with dropout on successive replays draw different masks."""
    from veloxseg_b200.nn import VeloxSeg
    from veloxseg_b200.train import TrainStep
    cfg = MODEL_CONFIGS["tiny"]
    x = torch.randn(2, 2, 64, 64, 64)
    y = (torch.rand(2, 1, 64, 64, 64) > 0.9).long()
    losses = {}
    for mode in (False, True):
        torch.manual_seed(3)
        m = VeloxSeg(**cfg)
        G.zero_dropout(m)
        ts = TrainStep(m, 2, DEV, use_graph=mode)
        losses[mode] = [ts.step(x, y, sync=True) for _ in range(3)]
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses[False], losses[True])
    assert losses[True][2] < losses[True][0]            # it trains
    torch.manual_seed(3)
    ts = TrainStep(VeloxSeg(**cfg), 2, DEV, lr=0.0, weight_decay=0.0, use_graph=True)     # dropout on, no parameter change
    l = [ts.step(x, y, sync=True) for _ in range(3)]
    assert len({round(v, 6) for v in l}) == 3, l        # same weights, same batch, different masks


def test_parallel_branches_match_serial():
    """Forked decoder streams (eager and graph-captured) compute the same loss and gradients as the serial schedule."""
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS["tiny"]
    x = torch.randn(2, 2, 64, 64, 64, device=DEV)
    res = {}
    for par in (False, True):
        torch.manual_seed(3)
        m = VeloxSeg(**cfg)
        G.zero_dropout(m)
        m = m.to(DEV).train()
        m.parallel_branches = par
        m.encoder.pipeline_branches = par      # conv encoder one level behind the transformer encoder on a side stream
        for rep in range(3):       # repeat: a race would show up as run-to-run differences
            m.zero_grad(set_to_none=True)
            outs = m(x)
            loss = sum((o * o).mean() for o in outs)
            loss.backward()
            torch.cuda.synchronize()
            res[(par, rep)] = (float(loss), torch.cat([p.grad.flatten() for p in m.parameters()]).clone())
    ref_loss, ref_g = res[(False, 0)]
    for k, (l, g) in res.items():
        assert abs(l - ref_loss) <= 1e-5 * abs(ref_loss), (k, l, ref_loss)
        assert float((g - ref_g).norm()) <= 1e-4 * float(ref_g.norm()), k
