"""Sliding-window inference on the GPU: the product path (CUDA-graph predictor over the sm_100a eval forward, windows
accumulated on the device) against the oracle's CPU eval forward applied window by window with the restated MONAI
blending (constant importance map, clamped last window, divide by count)."""
import pytest
import torch

from tests import _golden as G
from veloxseg_b200.configs import MODEL_CONFIGS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_sliding_window_vs_oracle():
    from oracle import veloxseg_oracle as O
    from veloxseg_b200.inference import GraphedPredictor, sliding_window_predict, window_starts
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS["tiny"]
    roi = cfg["input_size"]
    torch.manual_seed(G.MODEL_SEED)
    m = VeloxSeg(**cfg)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    spec = O.ModelSpec(cfg)
    vol = torch.randn(1, 2, 72, 64, 100, generator=torch.Generator().manual_seed(1))      # ragged on two axes
    starts = window_starts(vol.shape[2:], roi, 0.25)
    assert len(starts) == 2 * 1 * 2 and starts[-1] == (8, 0, 36)
    ref = torch.zeros(1, cfg["n_classes"], *vol.shape[2:])
    cnt = torch.zeros(1, 1, *vol.shape[2:])
    for a, b, c in starts:
        win = vol[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]]
        ref[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += O.forward(win, p, spec, training=False)
        cnt[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += 1
    ref /= cnt
    m = m.to(DEV).eval()
    for sw in (1, 3):           # 3 does not divide 4 windows: exercises the padded final batch of the graphed predictor
        pred = GraphedPredictor(m, sw, 2, roi, DEV)
        out = sliding_window_predict(vol.to(DEV), pred, roi, sw_batch_size=sw, overlap=0.25).cpu()
        assert out.shape == ref.shape
        err = float((out - ref).norm() / ref.norm())
        assert err < 1e-3, (sw, err)
        assert float((out.argmax(1) != ref.argmax(1)).float().mean()) < 2e-3
