"""veloxseg_b200.patch on the real reference checkout (this build container only: /root/reference is absent on the GPU
box, where the test skips).  The reference's own VeloxSeg.py / Encoder.py / Decoder.py wiring, patched, must build a
model whose hot-path components are the sm_100a modules and whose state_dict is key-for-key and value-for-value the
one the unpatched reference builds under the same seed (so reference checkpoints load both ways)."""
import json
import os
import sys

import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref_modules():
    added = [os.path.join(HERE, "golden", "monai_shim"), REF]      # MONAI is not installed: 4-symbol stand-in
    for p in added:
        sys.path.insert(0, p)
    try:
        import model.VeloxSeg as ref_vs
        yield ref_vs
    finally:
        for p in added:
            sys.path.remove(p)


@pytest.mark.parametrize("cfg_file", ["models_config_autopetii.json", "models_config_brats2021.json"])
def test_patched_reference_builds_identical_state(ref_modules, cfg_file):
    from veloxseg_b200 import nn as vnn
    from veloxseg_b200.patch import fuse_mixers, patch_reference, unpatch_reference
    cfg = json.load(open(os.path.join(REF, "config", cfg_file)))["VeloxSeg"]      # the reference's config file, unchanged
    torch.manual_seed(12345)
    plain = ref_modules.VeloxSeg(**cfg)
    undo = patch_reference()
    try:
        torch.manual_seed(12345)
        fast = fuse_mixers(ref_modules.VeloxSeg(**cfg))
    finally:
        unpatch_reference(undo)
    a, b = plain.state_dict(), fast.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    kinds = {type(m) for m in fast.modules()}
    assert {vnn.JLC, vnn.DownConv, vnn.UpConv, vnn.Transformer_BasicLayer, vnn.ModalMixer} <= kinds
    ref_kinds = {type(m).__module__ for m in fast.modules() if type(m).__name__ in ("JLC", "Transformer_BasicLayer")}
    assert ref_kinds == {"veloxseg_b200.nn"}
    fast.load_state_dict(plain.state_dict())          # a reference checkpoint loads into the patched model
    # after unpatching, the reference builds its own classes again
    torch.manual_seed(12345)
    again = ref_modules.VeloxSeg(**cfg)
    assert not any(isinstance(m, vnn.JLC) for m in again.modules())
