"""Drop-in switch for an UNMODIFIED checkout of the reference: keep its model/VeloxSeg.py, model/Encoder.py and
model/Decoder.py wiring and its config/*.json, and make the components they instantiate the sm_100a ones.

    import sys; sys.path.insert(0, "/path/to/VeloxSeg")          # the reference checkout (needs its own deps: monai)
    from veloxseg_b200.patch import patch_reference, fuse_mixers
    patch_reference()                                            # before the first VeloxSeg(...) is constructed
    from model.VeloxSeg import VeloxSeg
    net = fuse_mixers(VeloxSeg(**cfg["VeloxSeg"])).cuda()        # optional second step: fused 1x1+IN mixers

`patch_reference` rebinds the names the wiring modules imported from model/components (Encoder.py:7-10,
Decoder.py:6-8): JLCLayer, DownConv, UpConv, Transformer_BasicLayer, LayerNorm, get_pram_matrix, PixelShuffle.  The
replacements have the same constructor signatures and parameter registration order, so `state_dict()` keys and the
seed-for-seed He initialisation are unchanged (tests/test_oracle_golden.py::test_state_dict_matches_reference).
`fuse_mixers` swaps the `nn.Sequential(Conv3d 1x1, InstanceNorm3d)` adapters (`encoder.attn2conv_i`,
`rc_decoders.*.enc2rc_i`; Encoder.py:334-337, Decoder.py:54-57) for `ModalMixer`, which is the same Sequential (same
keys) with the fused forward; called as the reference calls it (`mixer(torch.cat(streams, 1))`) it still works, the
concatenation is then just not saved.
"""
from __future__ import annotations

import importlib

import torch.nn as nn

from . import nn as vnn

_WIRING = {
    "model.Encoder": {"JLCLayer": vnn.JLCLayer, "DownConv": vnn.DownConv, "Transformer_BasicLayer": vnn.Transformer_BasicLayer,
                      "LayerNorm": vnn.LayerNorm},
    "model.Decoder": {"JLCLayer": vnn.JLCLayer, "UpConv": vnn.UpConv, "get_pram_matrix": vnn.get_pram_matrix,
                      "PixelShuffle": vnn.PixelShuffle},
    "model.VeloxSeg": {"LayerNorm": vnn.LayerNorm},
}


def patch_reference(package: str = "model") -> dict:
    """Rebind the component names inside the reference's wiring modules.  Returns {module: {name: original}} so that
    a caller can undo it."""
    undo = {}
    for mod_name, names in _WIRING.items():
        mod = importlib.import_module(mod_name.replace("model", package, 1))
        undo[mod.__name__] = {}
        for name, repl in names.items():
            if hasattr(mod, name):
                undo[mod.__name__][name] = getattr(mod, name)
                setattr(mod, name, repl)
    return undo


def unpatch_reference(undo: dict) -> None:
    for mod_name, names in undo.items():
        mod = importlib.import_module(mod_name)
        for name, orig in names.items():
            setattr(mod, name, orig)


def _is_mixer(m: nn.Module) -> bool:
    return (isinstance(m, nn.Sequential) and not isinstance(m, vnn.ModalMixer) and len(m) == 2
            and isinstance(m[0], nn.Conv3d) and m[0].kernel_size == (1, 1, 1) and isinstance(m[1], nn.InstanceNorm3d))


def fuse_mixers(model: nn.Module) -> nn.Module:
    """Replace every Sequential(Conv3d 1x1, InstanceNorm3d) child by a ModalMixer holding the SAME parameters."""
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if _is_mixer(child):
                mix = vnn.ModalMixer(child[0].in_channels, child[0].out_channels)
                mix[0].weight, mix[0].bias = child[0].weight, child[0].bias
                setattr(parent, name, mix)
    return model
