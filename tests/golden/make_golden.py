"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on this container's CPU.

    python tests/golden/make_golden.py            # needs /root/reference; MONAI is replaced by tests/golden/monai_shim

The reference ships no golden vectors for the hot path, so the fixtures are minted here from
seeded synthetic inputs.  /root/reference does not exist on the GPU box: tests read only the files written here.

Fixtures
  ops_small.pt     per-module runs of the reference classes (JLC, mixer Sequential, Paired_Windows_TransformerBlock,
                   PatchMerging, DownConv/UpConv, get_pram_matrix) at small shapes: inputs, parameters, outputs,
                   input- and parameter-gradients for a seeded cotangent.  Dropout p = 0 everywhere.
  tables.pt        integer facts per (config, level): window pyramids, channels_qk / channels_v, the
                   relative_position_index buffer, and the token order of window_gathering_3d run on a voxel-index
                   ramp (bit-exact partition / pooling order).
  model_<cfg>.pt   whole-model runs (seed 12345 He init, input randn under seed 777): state_dict checksums, eval
                   logits and train-mode outputs (strided samples + moments), and per-parameter gradient samples for
                   loss = sum_i <out_i, R_i> with R_i regenerated from seeds.  Inputs / cotangents are NOT stored
                   (they are re-drawn from the same CPU generator seeds by tests/_golden.py).
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "monai_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from model.VeloxSeg import VeloxSeg  # noqa: E402  (reference)
from model.components.conv_blocks import JLC, DownConv, UpConv  # noqa: E402
from model.components.PWA import Paired_Windows_TransformerBlock  # noqa: E402
from model.components.attention_utils import PatchMerging  # noqa: E402
from model.components.common_function import get_pram_matrix  # noqa: E402
import torch.nn as nn  # noqa: E402

from tests._golden import (MODEL_SEED, cotangents, model_input, sample_tensor, sha_int32, state_checksums, zero_dropout)  # noqa: E402
from veloxseg_b200.configs import MODEL_CONFIGS  # noqa: E402

torch.set_num_threads(8)


def run_module(mod, inputs, seed):
    """inputs: list of tensors (or a list-of-list for the PWA block). Returns dict with outputs and grads."""
    g = torch.Generator().manual_seed(seed)
    flat_in = [t.clone().requires_grad_(True) for t in inputs]
    out = mod(flat_in) if getattr(mod, "_takes_list", False) else mod(*flat_in)
    outs = list(out) if isinstance(out, (list, tuple)) else [out]
    cots = [torch.randn(o.shape, generator=g) for o in outs]
    params = [p for p in mod.parameters()]
    grads = torch.autograd.grad(outs, flat_in + params, cots, allow_unused=True)
    names = [n for n, _ in mod.named_parameters()]
    return dict(inputs=[t.detach() for t in inputs], state={k: v.detach().clone() for k, v in mod.state_dict().items()},
                outputs=[o.detach() for o in outs], cot_seed=seed,
                input_grads=[x.detach() for x in grads[:len(flat_in)]],
                param_grads={n: (x.detach() if x is not None else None) for n, x in zip(names, grads[len(flat_in):])})


def ops_small():
    fx = {}
    torch.manual_seed(1)
    for tag, (C, groups, e, shape, B) in {"jlc_c8": (8, 2, 3, (5, 6, 8), 2), "jlc_c16": (16, 2, 2, (4, 4, 7), 1),
                                          "jlc_c32": (32, 2, 2, (3, 3, 3), 2)}.items():
        m = JLC(C, [1, 3, 5], groups, e, dropout=0.0, spatial_dim=3)
        for p in m.parameters():     # biases are zero-initialised by the reference; exercise them
            if p.dim() == 1:
                nn.init.normal_(p, std=0.1)
        fx[tag] = dict(run_module(m, [torch.randn(B, C, *shape)], 11), cfg=dict(C=C, groups=groups, e=e))
    # modal mixer: Sequential(Conv3d 1x1, IN) on cat(streams) + addend (Encoder.py:334-351)
    for tag, (chs, Co, shape, B) in {"mixer_2x16": ((16, 16), 16, (5, 6, 7), 2), "mixer_1x8": ((8,), 24, (4, 4, 4), 1)}.items():
        seq = nn.Sequential(nn.Conv3d(sum(chs), Co, 1, 1), nn.InstanceNorm3d(Co))
        streams = [torch.randn(B, c, *shape) for c in chs]
        addend = torch.randn(B, Co, *shape)

        class Wrap(nn.Module):
            def __init__(self, s):
                super().__init__()
                self.s = s

            def forward(self, addend, *st):
                return addend + self.s(torch.cat(st, dim=1))
        w = Wrap(seq)
        r = run_module(w, [addend] + streams, 12)
        r["state"] = {k[2:]: v for k, v in r["state"].items()}
        r["param_grads"] = {k[2:]: v for k, v in r["param_grads"].items()}
        fx[tag] = dict(r, cfg=dict(chs=chs, Co=Co))
    # PWA block
    for tag, (size, C, mb, ms, heads, mdh, M, e, B) in {
            "pwa_6c8": ((6, 6, 6), 8, [3, 3, 3], [1, 1, 1], 1, 4, 2, 2, 1),
            "pwa_884": ((8, 8, 4), 16, [4, 4, 2], [1, 1, 1], 2, 4, 2, 2, 1),
            "pwa_12m1": ((12, 12, 12), 8, [3, 3, 3], [1, 1, 1], 1, 4, 1, 1, 1)}.items():
        blk = Paired_Windows_TransformerBlock(list(size), [C] * M, mb, ms, 2, heads, mdh, attn_drop=0.0, proj_drop=0.0,
                                              ffn_expansion_ratio=e)
        for n, p in blk.named_parameters():
            if p.dim() == 1 and "norm" not in n:
                nn.init.normal_(p, std=0.1)
            elif "norm" in n:
                nn.init.normal_(p, mean=1.0 if n.endswith("weight") else 0.0, std=0.2)
            elif "table" in n:
                nn.init.normal_(p, std=0.5)
        blk._takes_list = True
        r = run_module(blk, [torch.randn(B, C, *size) for _ in range(M)], 13)
        a = blk.attn
        fx[tag] = dict(r, cfg=dict(size=size, C=C, mb=mb, ms=ms, heads=heads, mdh=mdh, M=M, e=e),
                       geo=dict(bws=a.big_window_size, sws=a.small_window_size, cqk=a.channels_qk, cv=a.channels_v,
                                n=a.n_hwd, heads=heads, nb=a.num_bswin))
    pm = PatchMerging(8, dim=3)
    nn.init.normal_(pm.norm.weight, 1.0, 0.2)
    nn.init.normal_(pm.norm.bias, 0.0, 0.2)
    fx["patch_merging"] = run_module(pm, [torch.randn(2, 8, 4, 6, 8)], 14)
    fx["down_conv"] = run_module(DownConv(3, 8, patch_size=2), [torch.randn(2, 3, 8, 8, 6)], 15)
    fx["up_conv"] = run_module(UpConv(8, 4), [torch.randn(2, 8, 3, 4, 5)], 16)

    class Gram(nn.Module):
        def forward(self, x):
            return get_pram_matrix(x)
    fx["gram"] = run_module(Gram(), [torch.randn(2, 16, 7, 9, 11)], 17)
    torch.save(fx, os.path.join(HERE, "ops_small.pt"))
    print("ops_small.pt", os.path.getsize(os.path.join(HERE, "ops_small.pt")) // 1024, "KiB")


def tables():
    out = {}
    for name, cfg in MODEL_CONFIGS.items():
        torch.manual_seed(0)
        m = VeloxSeg(**cfg)
        levels = []
        for i, layer in enumerate(m.encoder.encoder_attn.layers):
            a = layer.blocks[0].attn
            h, w, d = a.input_size
            # voxel-index ramp through the reference's gather: tokens hold the LARGEST voxel index of their small
            # window (max-pool), i.e. the partition/pooling geometry, bit-exact
            ramp = torch.arange(h * w * d, dtype=torch.float32).view(1, 1, h, w, d).repeat(1, a.channels_qk, 1, 1, 1)
            ramp = ramp + torch.arange(a.channels_qk, dtype=torch.float32).view(1, -1, 1, 1, 1) * (h * w * d)
            tok, Ns, n = a.window_gathering(ramp)
            levels.append(dict(input_size=list(a.input_size), big=a.big_window_size, small=a.small_window_size,
                               channels_qk=a.channels_qk, channels_v=a.channels_v, n=list(n), Ns=[list(x) for x in Ns],
                               rel_index_sha=sha_int32(a.position_embedding.relative_position_index),
                               rel_index=(a.position_embedding.relative_position_index.clone()
                                          if a.position_embedding.relative_position_index.numel() <= 4096 else None),
                               gather_ramp_sha=sha_int32(tok),
                               gather_ramp=tok.to(torch.int32) if name == "tiny" else None))
        out[name] = levels
    torch.save(out, os.path.join(HERE, "tables.pt"))
    print("tables.pt", os.path.getsize(os.path.join(HERE, "tables.pt")) // 1024, "KiB")


def contiguous_in_grads(model):
    """Makes the gradient arriving at each InstanceNorm3d contiguous (the SDKT Gram einsum hands channels-last-strided
    gradients to the JLC InstanceNorms).  Precaution only: re-running the generator with and without the hook gives
    identical fixtures on this torch build (relative difference 0.0 on the miniature config), so the hook does not change
    the reference's results; the reference code itself is untouched."""
    def fwd_hook(_m, _inp, out):
        if out.requires_grad:
            out.register_hook(lambda g: g.contiguous())
    for mod in model.modules():
        if isinstance(mod, nn.InstanceNorm3d):
            mod.register_forward_hook(fwd_hook)


def whole_model(name, B):
    cfg = MODEL_CONFIGS[name]
    torch.manual_seed(MODEL_SEED)
    m = VeloxSeg(**cfg)
    zero_dropout(m)
    contiguous_in_grads(m)
    fx = dict(config=name, B=B, state=state_checksums(m.state_dict()), n_params=sum(p.numel() for p in m.parameters()))
    x = model_input(cfg, B)
    m.eval()
    with torch.no_grad():
        y = m(x)
    fx["eval"] = sample_tensor(y)
    m.train()
    outs = m(x)
    cots = cotangents(outs)
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    loss.backward()
    fx["train_outputs"] = [sample_tensor(o.detach()) for o in outs]
    fx["loss"] = float(loss)
    fx["grads"] = {n: sample_tensor(p.grad) for n, p in m.named_parameters()}
    path = os.path.join(HERE, f"model_{name}.pt")
    torch.save(fx, path)
    print(os.path.basename(path), os.path.getsize(path) // 1024, "KiB", "loss", fx["loss"], "eval absmax", fx["eval"]["absmax"])


if __name__ == "__main__":
    ops_small()
    tables()
    for name, B in (("tiny", 2), ("autopetii", 1), ("hecktor2022", 1), ("brats2021", 1)):
        whole_model(name, B)
    json.dump({"torch": torch.__version__, "reference": "JinPLu/VeloxSeg @ /root/reference (read-only mount)"},
              open(os.path.join(HERE, "PROVENANCE.json"), "w"))
