"""Shared recipe between tests/golden/make_golden.py (writer, runs the reference) and the tests (readers).

Inputs and cotangents are never stored: both sides re-draw them from torch's CPU generator with fixed seeds (same torch
build in the build container and on the GPU box), and only strided samples + moments of the results are kept.
"""
import hashlib
import os

import torch

MODEL_SEED = 12345          # utils/seed.py of the reference seeds everything with 12345
INPUT_SEED = 777
COT_SEED = 4242
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_SAMPLE = 257


def model_input(cfg: dict, B: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(INPUT_SEED)
    return torch.randn(B, sum(cfg["in_ch"]), *cfg["input_size"], generator=g)


def cotangents(outs):
    g = torch.Generator().manual_seed(COT_SEED)
    return [torch.randn(o.shape, generator=g) / o.numel() ** 0.5 for o in outs]


def sample_index(numel: int) -> torch.Tensor:
    if numel <= N_SAMPLE:
        return torch.arange(numel)
    return torch.linspace(0, numel - 1, N_SAMPLE).round().long()


def sample_tensor(t: torch.Tensor) -> dict:
    f = t.detach().reshape(-1).cpu()
    d = f.double()
    return dict(shape=list(t.shape), sample=f[sample_index(f.numel())].clone(), norm=float(d.norm()), sum=float(d.sum()),
                absmax=float(d.abs().max()) if f.numel() else 0.0)


def check_sample(t: torch.Tensor, ref: dict, rtol: float, atol: float = 0.0, what: str = ""):
    """||t - ref|| judged on the stored strided sample (relative to the sample's own norm), plus the full-tensor norm."""
    assert list(t.shape) == ref["shape"], (what, list(t.shape), ref["shape"])
    f = t.detach().reshape(-1).cpu()
    s = f[sample_index(f.numel())].double()
    r = ref["sample"].double()
    scale = max(ref["norm"] / max(f.numel(), 1) ** 0.5, 1e-30)          # rms of the reference tensor
    err = float((s - r).norm()) / r.numel() ** 0.5
    assert err <= rtol * scale + atol, f"{what}: sample rms err {err:.3e} > {rtol:g} * rms {scale:.3e} + {atol:g}"
    n = float(f.double().norm())
    assert abs(n - ref["norm"]) <= 2 * rtol * ref["norm"] + atol * f.numel() ** 0.5, (what, "norm", n, ref["norm"])


def state_checksums(sd) -> dict:
    return {k: dict(shape=list(v.shape), dtype=str(v.dtype), sum=float(v.double().sum()), abs=float(v.double().abs().sum()))
            for k, v in sd.items()}


def zero_dropout(module: torch.nn.Module):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        for attr in ("dropout", "attn_drop", "proj_drop"):
            if isinstance(getattr(m, attr, None), float):
                setattr(m, attr, 0.0)


def sha_int32(t: torch.Tensor) -> str:
    return hashlib.sha256(t.to(torch.int32).contiguous().numpy().tobytes()).hexdigest()


def load(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name), weights_only=False)
