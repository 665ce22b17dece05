"""Shared helpers for the parity tests."""
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / max(||b||, tiny)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def close(a, b, rtol=1e-3, atol=1e-6) -> bool:
    """north_star tolerance for fp32: ||a-b|| <= rtol*||b|| + atol*sqrt(numel) (the atol term covers the tensors whose
    true gradient is exactly zero, SURVEY.md section 7 hard part 3)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm()) <= rtol * float(b.norm()) + atol * (b.numel() ** 0.5)


def jlc_params(C, groups, e, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    cg = C // groups
    def r(*s, k=1.0):
        return torch.randn(*s, generator=g) * k * scale
    return [r(C, cg, 1, 1, 1, k=0.5), r(C, k=0.1), r(C, cg, 3, 3, 3, k=0.1), r(C, k=0.1), r(C, cg, 5, 5, 5, k=0.05), r(C, k=0.1),
            r(e * C, C, k=0.25), r(e * C, k=0.1), r(C, e * C, k=0.15), r(C, k=0.1)]


def jlc_param_dict(params, prefix=""):
    names = ["spatial_convs.0.0.weight", "spatial_convs.0.0.bias", "spatial_convs.1.0.weight", "spatial_convs.1.0.bias",
             "spatial_convs.2.0.weight", "spatial_convs.2.0.bias", "channel_conv.1.weight", "channel_conv.1.bias",
             "channel_conv.3.weight", "channel_conv.3.bias"]
    return {prefix + n: p for n, p in zip(names, params)}


PWA_PARAM_NAMES = ["attn.input_norms.{m}.weight", "attn.input_norms.{m}.bias", "attn.qkv_proj.{m}.0.weight",
                   "attn.qkv_proj.{m}.0.bias", "attn.qkv_proj.{m}.1.weight", "attn.qkv_proj.{m}.1.bias",
                   "attn.qkv_proj.{m}.2.weight", "attn.qkv_proj.{m}.2.bias", "attn.mix_channels.{m}.weight",
                   "attn.mix_channels.{m}.bias", "norms.{m}.weight", "norms.{m}.bias", "ffns.{m}.linear1.weight",
                   "ffns.{m}.linear1.bias", "ffns.{m}.linear2.weight", "ffns.{m}.linear2.bias"]


def pwa_params(M, C, geo, e, seed=0):
    """Random PWA block parameters: (flat list in ABI order, dict under the reference's names, table, index)."""
    from oracle import veloxseg_oracle as O
    g = torch.Generator().manual_seed(seed)
    cqk, cv = geo["cqk"], geo["cv"]
    def r(*s, k=1.0, off=0.0):
        return torch.randn(*s, generator=g) * k + off
    flat, d = [], {}
    for m in range(M):
        ps = [r(C, k=0.2, off=1.0), r(C, k=0.1), r(cqk, C, k=0.3), r(cqk, k=0.1), r(cqk, C, k=0.3), r(cqk, k=0.1),
              r(cv, C, k=0.3), r(cv, k=0.1), r(C, cv, k=0.2), r(C, k=0.1), r(C, k=0.2, off=1.0), r(C, k=0.1),
              r(e * C, C, k=0.25), r(e * C, k=0.1), r(C, e * C, k=0.15), r(C, k=0.1)]
        flat += ps
        for name, p in zip(PWA_PARAM_NAMES, ps):
            d[name.format(m=m)] = p
    n = geo["n"]
    rows = (2 * n[0] - 1) * (2 * n[1] - 1) * (2 * n[2] - 1)
    table = r(rows, geo["heads"], k=0.5)
    index = O.relative_position_index(n)
    d["attn.position_embedding.relative_position_bias_table"] = table
    d["attn.position_embedding.relative_position_index"] = index
    return flat, d, table, index
