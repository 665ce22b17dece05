import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "emu: developer check of kernel indexing on the CPU shim in tools/emu "
                                       "(opt-in: VX_EMU=1); not a product path")


def pytest_collection_modifyitems(config, items):
    import torch
    if not torch.cuda.is_available():         # a plain `pytest` on a GPU-less box skips the parity tests instead of failing them
        no_gpu = pytest.mark.skip(reason="needs a CUDA device (B200)")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(no_gpu)
    if os.environ.get("VX_EMU") == "1":
        return
    skip = pytest.mark.skip(reason="CPU-shim kernel checks are opt-in (VX_EMU=1)")
    for it in items:
        if "emu" in it.keywords:
            it.add_marker(skip)
