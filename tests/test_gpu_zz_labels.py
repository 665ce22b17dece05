"""sliding_window_labels on one GPU (host volume in, host uint8 labels out; per-axis count tables) against the arg-max of
sliding_window_predict through the same CUDA-graph predictor: same windows, same summation order -> identical labels.
(The world-size-2 exchange of this path is covered on gloo in tests/test_dist_gloo.py.)"""
import pytest
import torch

from tests import _golden as G
from veloxseg_b200.configs import MODEL_CONFIGS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_labels_equal_argmax_of_predict():
    from veloxseg_b200.inference import GraphedPredictor, sliding_window_labels, sliding_window_predict
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS["tiny"]
    roi = cfg["input_size"]
    torch.manual_seed(G.MODEL_SEED)
    m = VeloxSeg(**cfg).to(DEV).eval()
    pred = GraphedPredictor(m, 2, 2, roi, DEV)
    # ragged on two axes (two rows of windows: the upload is staged in two slabs and the first label planes travel back
    # before the last windows run); smaller than the roi on one axis (padding: unstaged path); three rows of windows
    for shape in [(1, 2, 72, 64, 100), (1, 2, 60, 64, 64), (1, 2, 160, 64, 80)]:
        vol = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).pin_memory()
        ref = sliding_window_predict(vol.to(DEV), pred, roi, sw_batch_size=2, overlap=0.25).argmax(1)[0].to(torch.uint8).cpu()
        out_host = torch.empty(shape[2:], dtype=torch.uint8).pin_memory()
        got = sliding_window_labels(vol, pred, roi, DEV, sw_batch_size=2, overlap=0.25, out_host=out_host)
        assert got is out_host and got.shape == ref.shape
        assert torch.equal(got, ref)
        assert 0 < int(got.sum()) < got.numel()                  # both classes present: the comparison is not vacuous
        assert torch.equal(sliding_window_labels(vol, pred, roi, DEV, sw_batch_size=2, overlap=0.25), ref)      # result allocated by the call
