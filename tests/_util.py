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
