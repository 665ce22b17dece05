"""torch custom ops `torch.ops.veloxseg.*` over the C ABI in include/veloxseg_abi.h, plus their autograd glue.

Layering
  *_raw(lib, stream, ...)   pointer marshalling only: allocate outputs with torch, fill the POD descriptor, call
                            the C entry point.  Device-agnostic on purpose (tools/emu drives the same code with a
                            test-only CPU build of the kernels); nothing here computes anything.
  torch.ops.veloxseg.*      registered for the CUDA dispatch key only.  A CPU tensor reaches no kernel and raises
                            NotImplementedError from the dispatcher; a missing libveloxseg_sm100.so raises at first
                            use.  There is no fallback implementation.
  autograd.Function         saves what backward needs and calls the *_bwd ops.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import (GramDesc, InormDesc, JlcDesc, LnpwDesc, MixerDesc, PwaDesc, PwaSaved, ResizeDesc, SdktLossDesc, SegLossDesc,
                   VX_MAX_MODAL, VX_MAX_SCALES)

Tensor = torch.Tensor
_f32 = torch.float32


_seed_offset = {}


_PRECISION = "fp32"


def set_precision(mode: str) -> None:
    """`"fp32"` (default): every kernel fp32-accurate (tensor-core contractions 3xTF32).  `"bf16"`: north_star's second
    precision mode -- the tensor-core contractions (1x1 convs of levels 1-2 with their weight gradients, dense 3x3x3 convs
    forward / data / weight gradient) round operands and stored activations to bfloat16 and issue one product with fp32
    accumulation; statistics, norms, softmax, losses and the optimiser stay fp32 (VX_OPT_PRECISION, include/veloxseg_abi.h)."""
    global _PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _lib.get_lib().set_option(14, 1 if mode == "bf16" else 0)
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def seed_offset_tensor(device) -> Optional[Tensor]:
    """Device-resident uint64 (stored as int64) that every dropout kernel adds to its per-call seed.  `advance_seed`
    bumps it with a kernel, so a CUDA graph that captured one training step draws fresh masks on every replay."""
    device = torch.device(device)
    if device.type != "cuda":
        return None
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _seed_offset:
        _seed_offset[key] = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", key))
    return _seed_offset[key]


def seed_offset_ptr(device):
    t = seed_offset_tensor(device)
    return t.data_ptr() if t is not None else None


def advance_seed(device):
    """Call once per training step (capturable: a single in-place add on the current stream)."""
    t = seed_offset_tensor(device)
    if t is not None:
        t.add_(0x632BE59BD9B4E019)


def _stream(t: Tensor) -> int:
    """Current stream of the tensor's device.  The library launches on the CURRENT CUDA device, so the tensor's device has to
    be the current one: a mismatch is reported here instead of surfacing as an invalid-resource-handle launch error."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"veloxseg_b200: tensors live on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()} -- wrap the call in `with torch.cuda.device({t.device.index}):` "
                           "(TrainStep.step and GraphedPredictor do this themselves)")
    return torch.cuda.current_stream(t.device).cuda_stream


def _ws(lib, op: str, desc, like: Tensor) -> Optional[Tensor]:
    n = lib.workspace(op, desc)
    if n == 0:
        return None
    return torch.empty(n, dtype=torch.uint8, device=like.device)


def _chk(t: Tensor, name: str) -> Tensor:
    if t.dtype != _f32:
        raise TypeError(f"veloxseg: {name} must be float32, got {t.dtype}")
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------
# JLC
# ----------------------------------------------------------------------------------------------------
def jlc_desc(x: Tensor, groups: int, expansion: int, drop_p: float, training: bool, seed: int) -> JlcDesc:
    B, Cc, D, H, W = x.shape
    return JlcDesc(B, Cc, D, H, W, groups, expansion, 1e-5, float(drop_p), int(bool(training)), int(seed),
                   seed_offset_ptr(x.device) if training else None)


def jlc_fwd_raw(lib, stream, x, params: Sequence[Tensor], groups, expansion, drop_p=0.0, training=False, seed=0):
    """params = (w1,b1,w3,b3,w5,b5,fw1,fb1,fw2,fb2).  Returns [y, z, o, hpre, stats]."""
    x = _chk(x, "x")
    params = [_chk(p, "param") for p in params]
    d = jlc_desc(x, groups, expansion, drop_p, training, seed)
    B, Cc = x.shape[:2]
    S = x[0, 0].numel()
    y = torch.empty_like(x)
    z = torch.empty((3,) + tuple(x.shape), dtype=_f32, device=x.device)
    o = torch.empty_like(x)
    hpre = torch.empty((B, expansion * Cc) + tuple(x.shape[2:]), dtype=_f32, device=x.device)
    stats = torch.empty((4, B * Cc, 2), dtype=_f32, device=x.device)
    ws = _ws(lib, "jlc", d, x)
    lib.call_ws("vx_jlc_fwd", d, [x] + params, [y, z, o, hpre, stats], ws, stream)
    return [y, z, o, hpre, stats]


def jlc_bwd_raw(lib, stream, dy, x, z, o, hpre, stats, params: Sequence[Tensor], groups, expansion, drop_p=0.0,
                training=False, seed=0):
    """Returns [dx, dw1,db1,dw3,db3,dw5,db5,dfw1,dfb1,dfw2,dfb2]."""
    dy = _chk(dy, "dy")
    w1, b1, w3, b3, w5, b5, fw1, fb1, fw2, fb2 = params
    d = jlc_desc(x, groups, expansion, drop_p, training, seed)
    outs = [torch.empty_like(x)] + [torch.empty_like(p) for p in params]
    ws = _ws(lib, "jlc", d, x)
    lib.call_ws("vx_jlc_bwd", d, [dy, x, z, o, hpre, stats, w1, w3, w5, fw1, fb1, fw2], outs, ws, stream)
    return outs


# ----------------------------------------------------------------------------------------------------
# modal mixer / instance norm
# ----------------------------------------------------------------------------------------------------
def mixer_desc(streams: Sequence[Tensor], c_out: int, has_addend: bool) -> MixerDesc:
    if len(streams) > VX_MAX_MODAL:
        raise ValueError(f"veloxseg: at most {VX_MAX_MODAL} streams")
    ch = (C.c_int32 * VX_MAX_MODAL)(*([int(s.shape[1]) for s in streams] + [0] * (VX_MAX_MODAL - len(streams))))
    return MixerDesc(streams[0].shape[0], streams[0][0, 0].numel(), len(streams), ch, c_out, int(has_addend), 1e-5)


def mixer_fwd_raw(lib, stream, streams: Sequence[Tensor], W, b, addend: Optional[Tensor]):
    """Returns [y, t, stats]."""
    streams = [_chk(s, "stream") for s in streams]
    W, b = _chk(W, "W"), _chk(b, "b")
    c_out = W.shape[0]
    d = mixer_desc(streams, c_out, addend is not None)
    shp = (streams[0].shape[0], c_out) + tuple(streams[0].shape[2:])
    y = torch.empty(shp, dtype=_f32, device=W.device)
    t = torch.empty(shp, dtype=_f32, device=W.device)
    stats = torch.empty((shp[0] * c_out, 2), dtype=_f32, device=W.device)
    ins = list(streams) + [W, b] + ([_chk(addend, "addend")] if addend is not None else [None])
    lib.call_ws("vx_mixer_fwd", d, ins, [y, t, stats], None, stream)
    return [y, t, stats]


def mixer_bwd_raw(lib, stream, dy, streams: Sequence[Tensor], W, t, stats):
    """Returns [da_0.., dW, db]."""
    dy = _chk(dy, "dy")
    d = mixer_desc(streams, W.shape[0], False)
    outs = [torch.empty_like(s) for s in streams] + [torch.empty_like(W), torch.empty(W.shape[0], dtype=_f32, device=W.device)]
    ws = _ws(lib, "mixer", d, dy)
    lib.call_ws("vx_mixer_bwd", d, [dy] + list(streams) + [W, t, stats], outs, ws, stream)
    return outs


def inorm_fwd_raw(lib, stream, x, addend: Optional[Tensor]):
    x = _chk(x, "x")
    rows = x.shape[0] * x.shape[1]
    d = InormDesc(rows, x[0, 0].numel(), 1e-5, int(addend is not None))
    y = torch.empty_like(x)
    stats = torch.empty((rows, 2), dtype=_f32, device=x.device)
    lib.call("vx_inorm_fwd", d, [x, _chk(addend, "addend") if addend is not None else None], [y, stats], stream)
    return [y, stats]


def inorm_bwd_raw(lib, stream, dy, x, stats, with_bias_grad: bool = False):
    dy = _chk(dy, "dy")
    C_ = x.shape[1]
    d = InormDesc(x.shape[0] * C_, x[0, 0].numel(), 1e-5, 0, C_ if with_bias_grad else 0)
    dx = torch.empty_like(x)
    if with_bias_grad:
        db = torch.empty((C_,), dtype=_f32, device=x.device)
        lib.call("vx_inorm_bwd", d, [dy, x, stats], [dx, db], stream)
        return dx, db
    lib.call("vx_inorm_bwd", d, [dy, x, stats], [dx], stream)
    return dx


# ----------------------------------------------------------------------------------------------------
# SDKT
# ----------------------------------------------------------------------------------------------------
def gram_fwd_raw(lib, stream, x):
    x = _chk(x, "x")
    B, Cc = x.shape[:2]
    d = GramDesc(B, Cc, x[0, 0].numel())
    G = torch.empty((B, Cc, Cc), dtype=_f32, device=x.device)
    ws = _ws(lib, "gram", d, x)
    lib.call_ws("vx_gram_fwd", d, [x], [G], ws, stream)
    return G


def gram_bwd_raw(lib, stream, dG, x):
    dG = _chk(dG, "dG")
    B, Cc = x.shape[:2]
    d = GramDesc(B, Cc, x[0, 0].numel())
    dx = torch.empty_like(x)
    lib.call("vx_gram_bwd", d, [dG, x], [dx], stream)
    return dx


def sdkt_loss_fwd_raw(lib, stream, gs, gts: Sequence[Tensor]):
    gs = _chk(gs, "G_s")
    gts = [_chk(g, "G_t") for g in gts]
    d = SdktLossDesc(gs.numel(), len(gts))
    loss = torch.empty((), dtype=_f32, device=gs.device)
    lib.call("vx_sdkt_loss_fwd", d, [gs] + gts, [loss], stream)
    return loss


def sdkt_loss_bwd_raw(lib, stream, dloss, gs, gts: Sequence[Tensor]):
    d = SdktLossDesc(gs.numel(), len(gts))
    outs = [torch.empty_like(gs)] + [torch.empty_like(g) for g in gts]
    lib.call("vx_sdkt_loss_bwd", d, [_chk(dloss, "dloss"), gs] + list(gts), outs, stream)
    return outs


# ----------------------------------------------------------------------------------------------------
# deep-supervision segmentation loss (CE + Dice per output, weighted)
# ----------------------------------------------------------------------------------------------------
def segloss_desc(logits: Sequence[Tensor], weights: Sequence[float]) -> SegLossDesc:
    B, Cc = logits[0].shape[:2]
    d = SegLossDesc(len(logits), B, Cc, logits[0][0, 0].numel())
    for i, w in enumerate(weights):
        d.weights[i] = float(w)
    return d


def segloss_fwd_raw(lib, stream, logits: Sequence[Tensor], labels: Tensor, weights: Sequence[float]):
    logits = [_chk(t, "logits") for t in logits]
    if labels.dtype != torch.int64:
        raise TypeError(f"veloxseg: labels must be int64, got {labels.dtype}")
    labels = labels.contiguous()
    if any(t.shape != logits[0].shape for t in logits) or labels.numel() != logits[0].numel() // logits[0].shape[1]:
        raise ValueError("veloxseg: seg_loss needs equally shaped logits and one label per voxel")
    d = segloss_desc(logits, weights)
    loss = torch.empty((), dtype=_f32, device=labels.device)
    sums = torch.empty((d.n_out, d.B, 1 + 3 * d.C), dtype=_f32, device=labels.device)
    lib.call_ws("vx_segloss_fwd", d, list(logits) + [labels], [loss, sums], _ws(lib, "segloss", d, labels), stream)
    return loss, sums


def segloss_bwd_raw(lib, stream, dloss, logits: Sequence[Tensor], labels: Tensor, sums: Tensor, weights: Sequence[float]):
    d = segloss_desc(logits, weights)
    outs = [torch.empty_like(t) for t in logits]
    lib.call("vx_segloss_bwd", d, [_chk(dloss, "dloss")] + list(logits) + [labels, sums], outs, stream)
    return outs


# ----------------------------------------------------------------------------------------------------
# LayerNorm(channels_first) + 1x1 (PatchMerging tail)
# ----------------------------------------------------------------------------------------------------
def lnpw_fwd_raw(lib, stream, x, ln_w, ln_b, W):
    x = _chk(x, "x")
    B, Ci = x.shape[:2]
    Co = W.shape[0]
    S = x[0, 0].numel()
    d = LnpwDesc(B, Ci, Co, S, 1e-6)
    y = torch.empty((B, Co) + tuple(x.shape[2:]), dtype=_f32, device=x.device)
    xhat = torch.empty_like(x)
    rstd = torch.empty((B, S), dtype=_f32, device=x.device)
    lib.call_ws("vx_lnpw_fwd", d, [x, _chk(ln_w, "ln_w"), _chk(ln_b, "ln_b"), _chk(W, "W")], [y, xhat, rstd], None, stream)
    return [y, xhat, rstd]


def lnpw_bwd_raw(lib, stream, dy, xhat, rstd, ln_w, ln_b, W):
    dy = _chk(dy, "dy")
    B, Ci = xhat.shape[:2]
    d = LnpwDesc(B, Ci, W.shape[0], xhat[0, 0].numel(), 1e-6)
    outs = [torch.empty_like(xhat), torch.empty_like(ln_w), torch.empty_like(ln_b), torch.empty_like(W)]
    ws = _ws(lib, "lnpw", d, dy)
    lib.call_ws("vx_lnpw_bwd", d, [dy, xhat, rstd, ln_w, W, ln_b], outs, ws, stream)
    return outs


# ----------------------------------------------------------------------------------------------------
# trilinear resize (align_corners=True)
# ----------------------------------------------------------------------------------------------------
def resize_fwd_raw(lib, stream, x, size: Sequence[int]):
    x = _chk(x, "x")
    B, Cc, d, h, w = x.shape
    D, H, W = [int(v) for v in size]
    desc = ResizeDesc(B * Cc, d, h, w, D, H, W)
    y = torch.empty((B, Cc, D, H, W), dtype=_f32, device=x.device)
    lib.call("vx_resize_trilinear_fwd", desc, [x], [y], stream)
    return y


def resize_bwd_raw(lib, stream, dy, size: Sequence[int]):
    """dy (B, C, D, H, W) -> dx (B, C, *size)."""
    dy = _chk(dy, "dy")
    B, Cc, D, H, W = dy.shape
    d, h, w = [int(v) for v in size]
    desc = ResizeDesc(B * Cc, d, h, w, D, H, W)
    dx = torch.empty((B, Cc, d, h, w), dtype=_f32, device=dy.device)
    ws = _ws(lib, "resize", desc, dy)
    lib.call_ws("vx_resize_trilinear_bwd", desc, [dy], [dx], ws, stream)
    return dx


# ----------------------------------------------------------------------------------------------------
# PWA block
# ----------------------------------------------------------------------------------------------------
PWA_PARAMS_PER_MODALITY = 16   # ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wmix, bmix, ln2_w, ln2_b, w1, b1, w2, b2


def pwa_desc(xs: Sequence[Tensor], geo: dict, ffn_expansion: int, attn_drop=0.0, proj_drop=0.0, training=False,
             seed=0) -> PwaDesc:
    B, Cc, D, H, W = xs[0].shape
    nb = len(geo["bws"])
    if nb > VX_MAX_SCALES or len(xs) > VX_MAX_MODAL:
        raise ValueError("veloxseg: too many window scales / modalities")
    big = ((C.c_int32 * 3) * VX_MAX_SCALES)()
    small = ((C.c_int32 * 3) * VX_MAX_SCALES)()
    for j in range(nb):
        for a in range(3):
            big[j][a] = int(geo["bws"][j][a])
            small[j][a] = int(geo["sws"][j][a])
    return PwaDesc(B, len(xs), Cc, D, H, W, int(geo["heads"]), nb, big, small, int(geo["cqk"]), int(geo["cv"]),
                   int(ffn_expansion), 1e-6, float(attn_drop), float(proj_drop), int(bool(training)), int(seed),
                   seed_offset_ptr(xs[0].device) if training else None)


def pwa_saved_sizes(lib, d: PwaDesc) -> List[int]:
    lay = PwaSaved()
    lib.check(lib.c.vx_pwa_saved_layout(C.byref(d), C.byref(lay)), "vx_pwa_saved_layout")
    return [int(lay.saved_bytes[i]) for i in range(lay.n_saved)]


def pwa_block_fwd_raw(lib, stream, xs: Sequence[Tensor], params: Sequence[Tensor], table, index, geo,
                      ffn_expansion, attn_drop=0.0, proj_drop=0.0, training=False, seed=0):
    """params: M * 16 tensors (PWA_PARAMS_PER_MODALITY order).  Returns (zs, saved)."""
    xs = [_chk(x, "x") for x in xs]
    params = [_chk(p, "param") for p in params]
    d = pwa_desc(xs, geo, ffn_expansion, attn_drop, proj_drop, training, seed)
    zs = [torch.empty_like(x) for x in xs]
    saved = [torch.empty(max(n, 4), dtype=torch.uint8, device=xs[0].device) for n in pwa_saved_sizes(lib, d)]
    ws = _ws(lib, "pwa", d, xs[0])
    lib.call_ws("vx_pwa_block_fwd", d, xs + params + [_chk(table, "table"), index.contiguous()], zs + saved, ws, stream)
    return zs, saved


def pwa_block_bwd_raw(lib, stream, dzs, xs, params, table, index, saved, geo, ffn_expansion, attn_drop=0.0,
                      proj_drop=0.0, training=False, seed=0):
    """Returns (dxs, dparams, dtable)."""
    dzs = [_chk(g, "dz") for g in dzs]
    d = pwa_desc(xs, geo, ffn_expansion, attn_drop, proj_drop, training, seed)
    dxs = [torch.empty_like(x) for x in xs]
    dparams = [torch.empty_like(p) for p in params]
    dtable = torch.empty_like(table)
    ws = _ws(lib, "pwa", d, xs[0])
    lib.call_ws("vx_pwa_block_bwd", d, dzs + list(xs) + list(params) + [table, index] + list(saved),
                dxs + dparams + [dtable], ws, stream)
    return dxs, dparams, dtable


def pwa_gather_raw(lib, stream, x, geo):
    """Integer window partition + max-pool: tokens (B, heads, Ns, l, c) and arg-max voxel index (int32)."""
    x = _chk(x, "x")
    B, Ct, D, H, W = x.shape
    heads, nb = geo["heads"], len(geo["bws"])
    c = Ct // (heads * nb)
    n = geo["n"]
    l = n[0] * n[1] * n[2]
    Ns = sum((D // bw[0]) * (H // bw[1]) * (W // bw[2]) for bw in geo["bws"])
    g2 = dict(geo)
    g2["cqk"], g2["cv"] = Ct, Ct
    d = pwa_desc([x], g2, 1)
    tok = torch.empty((B, heads, Ns, l, c), dtype=_f32, device=x.device)
    arg = torch.empty((B, heads, Ns, l, c), dtype=torch.int32, device=x.device)
    lib.check(lib.c.vx_pwa_gather(C.byref(d), Ct, x.data_ptr(), tok.data_ptr(), arg.data_ptr(), stream), "vx_pwa_gather")
    return tok, arg


# ----------------------------------------------------------------------------------------------------
# torch.library registration (CUDA key only) and autograd
# ----------------------------------------------------------------------------------------------------
_L = torch.library.Library("veloxseg", "DEF")
_L.define("jlc_fwd(Tensor x, Tensor[] params, int groups, int expansion, float drop_p, bool training, int seed) -> Tensor[]")
_L.define("jlc_bwd(Tensor dy, Tensor x, Tensor z, Tensor o, Tensor hpre, Tensor stats, Tensor[] params, int groups, "
          "int expansion, float drop_p, bool training, int seed) -> Tensor[]")
_L.define("mixer_fwd(Tensor[] streams, Tensor W, Tensor b, Tensor? addend) -> Tensor[]")
_L.define("mixer_bwd(Tensor dy, Tensor[] streams, Tensor W, Tensor t, Tensor stats) -> Tensor[]")
_L.define("inorm_fwd(Tensor x, Tensor? addend) -> Tensor[]")
_L.define("inorm_bwd(Tensor dy, Tensor x, Tensor stats) -> Tensor")
_L.define("gram_fwd(Tensor x) -> Tensor")
_L.define("gram_bwd(Tensor dG, Tensor x) -> Tensor")
_L.define("sdkt_loss_fwd(Tensor gs, Tensor[] gts) -> Tensor")
_L.define("sdkt_loss_bwd(Tensor dloss, Tensor gs, Tensor[] gts) -> Tensor[]")
_L.define("segloss_fwd(Tensor[] logits, Tensor labels, float[] weights) -> Tensor[]")
_L.define("segloss_bwd(Tensor dloss, Tensor[] logits, Tensor labels, Tensor sums, float[] weights) -> Tensor[]")
_L.define("resize_fwd(Tensor x, int[] size) -> Tensor")
_L.define("resize_bwd(Tensor dy, int[] size) -> Tensor")
_L.define("lnpw_fwd(Tensor x, Tensor ln_w, Tensor ln_b, Tensor W) -> Tensor[]")
_L.define("lnpw_bwd(Tensor dy, Tensor xhat, Tensor rstd, Tensor ln_w, Tensor ln_b, Tensor W) -> Tensor[]")


def _cuda(fn):
    def run(*a, **k):
        first = a[0][0] if isinstance(a[0], (list, tuple)) else a[0]
        if first.device.index != torch.cuda.current_device():       # device guard (the library launches on the current device)
            with torch.cuda.device(first.device):
                return fn(_lib.get_lib(), _stream(first), *a, **k)
        return fn(_lib.get_lib(), _stream(first), *a, **k)
    return run


_L.impl("jlc_fwd", _cuda(jlc_fwd_raw), "CUDA")
_L.impl("jlc_bwd", _cuda(jlc_bwd_raw), "CUDA")
_L.impl("mixer_fwd", _cuda(mixer_fwd_raw), "CUDA")
_L.impl("mixer_bwd", _cuda(mixer_bwd_raw), "CUDA")
_L.impl("inorm_fwd", _cuda(inorm_fwd_raw), "CUDA")
_L.impl("inorm_bwd", _cuda(inorm_bwd_raw), "CUDA")
_L.impl("gram_fwd", _cuda(gram_fwd_raw), "CUDA")
_L.impl("gram_bwd", _cuda(gram_bwd_raw), "CUDA")
_L.impl("sdkt_loss_fwd", _cuda(sdkt_loss_fwd_raw), "CUDA")
_L.impl("sdkt_loss_bwd", _cuda(sdkt_loss_bwd_raw), "CUDA")
_L.impl("segloss_fwd", _cuda(segloss_fwd_raw), "CUDA")
_L.impl("segloss_bwd", _cuda(segloss_bwd_raw), "CUDA")
_L.impl("resize_fwd", _cuda(resize_fwd_raw), "CUDA")
_L.impl("resize_bwd", _cuda(resize_bwd_raw), "CUDA")
_L.impl("lnpw_fwd", _cuda(lnpw_fwd_raw), "CUDA")
_L.impl("lnpw_bwd", _cuda(lnpw_bwd_raw), "CUDA")

_vx = torch.ops.veloxseg
_seed_counter = [0]


def next_seed() -> int:
    """Dropout key for one op call: torch's CPU generator state is not consumed; a process-local counter mixed
    with torch.initial_seed() keeps runs reproducible under torch.manual_seed."""
    _seed_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0]) & 0x7FFFFFFFFFFFFFFF


class _JLC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, groups, expansion, drop_p, training, seed, *params):
        y, z, o, hpre, stats = _vx.jlc_fwd(x, list(params), groups, expansion, drop_p, training, seed)
        ctx.save_for_backward(x, z, o, hpre, stats, *params)
        ctx.cfg = (groups, expansion, drop_p, training, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, z, o, hpre, stats, *params = ctx.saved_tensors
        g = _vx.jlc_bwd(dy.contiguous(), x, z, o, hpre, stats, list(params), *ctx.cfg)
        return (g[0], None, None, None, None, None) + tuple(g[1:])


def jlc(x, params: Sequence[Tensor], groups: int, expansion: int, drop_p: float = 0.0, training: bool = False) -> Tensor:
    seed = next_seed() if (training and drop_p > 0) else 0
    return _JLC.apply(x.contiguous(), groups, expansion, float(drop_p), bool(training), seed, *params)


class _Mixer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, b, addend, *streams):
        y, t, stats = _vx.mixer_fwd(list(streams), W, b, addend)
        ctx.save_for_backward(W, t, stats, *streams)
        ctx.has_addend = addend is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        W, t, stats, *streams = ctx.saved_tensors
        dy = dy.contiguous()
        g = _vx.mixer_bwd(dy, list(streams), W, t, stats)
        M = len(streams)
        return (g[M], g[M + 1], dy if ctx.has_addend else None) + tuple(g[:M])


def modal_mixer(streams: Sequence[Tensor], W: Tensor, b: Tensor, addend: Optional[Tensor] = None) -> Tensor:
    """[addend +] IN(W . cat(streams) + b);  W may be (Co, K) or (Co, K, 1, 1, 1)."""
    return _Mixer.apply(W.reshape(W.shape[0], -1), b, addend, *[s.contiguous() for s in streams])


class _INorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, addend):
        y, stats = _vx.inorm_fwd(x, addend)
        ctx.save_for_backward(x, stats)
        ctx.has_addend = addend is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        dy = dy.contiguous()
        return _vx.inorm_bwd(dy, x, stats), (dy if ctx.has_addend else None)


def instance_norm(x: Tensor, addend: Optional[Tensor] = None) -> Tensor:
    return _INorm.apply(x.contiguous(), addend)


class _INormBias(torch.autograd.Function):
    """IN(z + bias[c]) [+ addend] where z is a bias-free convolution output: the per-channel bias cancels in the value of an
    affine-less InstanceNorm, so it is not added; its gradient (sum over batch and voxels of dz, analytically zero --
    SURVEY.md section 7.3) comes out of the norm's backward kernel instead of a separate reduction over dz."""

    @staticmethod
    def forward(ctx, z, bias, addend):
        lib, st = _lib.get_lib(), _stream(z)
        y, stats = inorm_fwd_raw(lib, st, z, addend)
        ctx.save_for_backward(z, stats)
        ctx.has_addend = addend is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        z, stats = ctx.saved_tensors
        dy = dy.contiguous()
        dz, db = inorm_bwd_raw(_lib.get_lib(), _stream(z), dy, z, stats, with_bias_grad=True)
        return dz, db, (dy if ctx.has_addend else None)


def instance_norm_biased(z: Tensor, bias: Tensor, addend: Optional[Tensor] = None) -> Tensor:
    """InstanceNorm3d(affine=False)(z + bias[None, :, None, None, None]) [+ addend] for a bias-free conv output z."""
    return _INormBias.apply(z.contiguous(), bias, addend)


class _Gram(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _vx.gram_fwd(x)

    @staticmethod
    def backward(ctx, dG):
        (x,) = ctx.saved_tensors
        return _vx.gram_bwd(dG.contiguous(), x)


def gram(x: Tensor) -> Tensor:
    return _Gram.apply(x.contiguous())


class _SdktLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gs, *gts):
        ctx.save_for_backward(gs, *gts)
        return _vx.sdkt_loss_fwd(gs, list(gts))

    @staticmethod
    def backward(ctx, dloss):
        gs, *gts = ctx.saved_tensors
        return tuple(_vx.sdkt_loss_bwd(dloss.contiguous(), gs, list(gts)))


def sdkt_loss(gs: Tensor, gts: Sequence[Tensor]) -> Tensor:
    return _SdktLoss.apply(gs.contiguous(), *[g.contiguous() for g in gts])


class _SegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, weights, *logits):
        loss, sums = _vx.segloss_fwd(list(logits), labels, list(weights))
        ctx.save_for_backward(labels, sums, *logits)
        ctx.weights = list(weights)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        labels, sums, *logits = ctx.saved_tensors
        return (None, None) + tuple(_vx.segloss_bwd(dloss.contiguous(), logits, labels, sums, ctx.weights))


def seg_loss(logits: Sequence[Tensor], labels: Tensor, weights: Sequence[float]) -> Tensor:
    """sum_i w_i (CrossEntropy(logits_i, y) + Dice(logits_i, y)); labels (B, 1, *spatial) or (B, *spatial) int64."""
    return _SegLoss.apply(labels.contiguous(), tuple(float(w) for w in weights), *[t.contiguous() for t in logits])


class _LnPw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, W):
        y, xhat, rstd = _vx.lnpw_fwd(x, ln_w, ln_b, W)
        ctx.save_for_backward(xhat, rstd, ln_w, ln_b, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        xhat, rstd, ln_w, ln_b, W = ctx.saved_tensors
        return tuple(_vx.lnpw_bwd(dy.contiguous(), xhat, rstd, ln_w, ln_b, W))


def ln_pointwise(x: Tensor, ln_w: Tensor, ln_b: Tensor, W: Tensor) -> Tensor:
    """W . LayerNorm_channels_first(x)   (no bias);  W (Co, Ci) or (Co, Ci, 1, 1, 1)."""
    Wm = W.reshape(W.shape[0], -1)
    return _LnPw.apply(x.contiguous(), ln_w, ln_b, Wm)


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        ctx.in_size = tuple(x.shape[2:])
        return _vx.resize_fwd(x, list(size))

    @staticmethod
    def backward(ctx, dy):
        return _vx.resize_bwd(dy.contiguous(), list(ctx.in_size)), None


def resize_trilinear(x: Tensor, size: Sequence[int]) -> Tensor:
    """F.interpolate(x, size=size, mode='trilinear', align_corners=True).  Same-size requests are the identity
    (source coordinate == destination index exactly), so the input is returned as is."""
    size = tuple(int(v) for v in size)
    if tuple(x.shape[2:]) == size:
        return x
    return _Resize.apply(x.contiguous(), size)


class _PwaBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, table, index, M, *tensors):
        xs, params = list(tensors[:M]), list(tensors[M:])
        lib = _lib.get_lib()
        for x in xs:
            if not x.is_cuda:
                raise NotImplementedError("veloxseg::pwa_block has no CPU implementation")
        geo, ffn_e, attn_drop, proj_drop, training, seed = cfg
        zs, saved = pwa_block_fwd_raw(lib, _stream(xs[0]), xs, params, table, index, geo, ffn_e, attn_drop, proj_drop,
                                      training, seed)
        ctx.save_for_backward(table, index, *xs, *params, *saved)
        ctx.cfg, ctx.M, ctx.nparam = cfg, M, len(params)
        return tuple(zs)

    @staticmethod
    def backward(ctx, *dzs):
        t = ctx.saved_tensors
        table, index = t[0], t[1]
        M, P = ctx.M, ctx.nparam
        xs, params, saved = list(t[2:2 + M]), list(t[2 + M:2 + M + P]), list(t[2 + M + P:])
        geo, ffn_e, attn_drop, proj_drop, training, seed = ctx.cfg
        lib = _lib.get_lib()
        dxs, dparams, dtable = pwa_block_bwd_raw(lib, _stream(xs[0]), [g.contiguous() for g in dzs], xs, params, table,
                                                 index, saved, geo, ffn_e, attn_drop, proj_drop, training, seed)
        return (None, dtable, None, None) + tuple(dxs) + tuple(dparams)


def pwa_block(xs: Sequence[Tensor], params: Sequence[Tensor], table: Tensor, index: Tensor, geo: dict,
              ffn_expansion: int, attn_drop: float = 0.0, proj_drop: float = 0.0, training: bool = False) -> List[Tensor]:
    """One Paired_Windows_TransformerBlock (attention + FFN) over M modality streams."""
    seed = next_seed() if (training and (attn_drop > 0 or proj_drop > 0)) else 0
    cfg = (geo, int(ffn_expansion), float(attn_drop), float(proj_drop), bool(training), seed)
    return list(_PwaBlock.apply(cfg, table, index, len(xs), *[x.contiguous() for x in xs], *params))


# ----------------------------------------------------------------------------------------------------
# Network-input stem: PatchEmbed (Conv3d k = s = patch) on a channel range of the input, no data gradient
# ----------------------------------------------------------------------------------------------------
def patch_embed_fwd_raw(lib, stream, x: Tensor, c_off: int, weight: Tensor, bias: Optional[Tensor]):
    from ._lib import PatchEmbedDesc
    x, weight = _chk(x, "x"), _chk(weight, "weight")
    B, Ct, D, H, W = x.shape
    Co, Ci, p = weight.shape[0], weight.shape[1], weight.shape[2]
    d = PatchEmbedDesc(B, Ct, int(c_off), Ci, Co, p, D, H, W)
    y = torch.empty((B, Co, D // p, H // p, W // p), dtype=_f32, device=x.device)
    lib.call("vx_patch_embed_fwd", d, [x, weight, _chk(bias, "bias") if bias is not None else None], [y], stream)
    return y


def patch_embed_bwd_raw(lib, stream, dy: Tensor, x: Tensor, c_off: int, weight_shape):
    from ._lib import PatchEmbedDesc
    B, Ct, D, H, W = x.shape
    Co, Ci, p = weight_shape[0], weight_shape[1], weight_shape[2]
    d = PatchEmbedDesc(B, Ct, int(c_off), Ci, Co, p, D, H, W)
    dw = torch.empty(tuple(weight_shape), dtype=_f32, device=x.device)
    db = torch.empty((Co,), dtype=_f32, device=x.device)
    lib.call("vx_patch_embed_bwd", d, [_chk(dy, "dy"), x], [dw, db], stream)
    return dw, db


class _PatchEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c_off, weight, bias):
        ctx.c_off, ctx.has_bias = int(c_off), bias is not None
        ctx.save_for_backward(x, weight)
        return patch_embed_fwd_raw(_lib.get_lib(), _stream(x), x, c_off, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dw, db = patch_embed_bwd_raw(_lib.get_lib(), _stream(x), dy.contiguous(), x, ctx.c_off, weight.shape)
        return None, None, dw, (db if ctx.has_bias else None)


def patch_embed(x: Tensor, c_off: int, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """Conv3d(weight, bias, stride = kernel = patch) applied to channels [c_off, c_off + weight.shape[1]) of the network
    input `x` (B, C_total, D, H, W).  `x` must not require grad (it is the network input)."""
    if x.requires_grad:
        raise RuntimeError("veloxseg: patch_embed has no data gradient (network input); use the torch convolution")
    return _PatchEmbed.apply(x.contiguous(), int(c_off), weight, bias)


# ----------------------------------------------------------------------------------------------------
# bias + 3-D pixel shuffle behind a bias-free convolution (out_conv1 / reconstruction out_conv)
# ----------------------------------------------------------------------------------------------------
def pixel_shuffle_fwd_raw(lib, stream, z: Tensor, bias: Optional[Tensor], scale: int):
    from ._lib import PixelShuffleDesc
    z = _chk(z, "z")
    B, Cs, d, h, w = z.shape
    C_ = Cs // scale ** 3
    desc = PixelShuffleDesc(B, C_, scale, d, h, w)
    y = torch.empty((B, C_, d * scale, h * scale, w * scale), dtype=_f32, device=z.device)
    lib.call("vx_pixel_shuffle_fwd", desc, [z, _chk(bias, "bias") if bias is not None else None], [y], stream)
    return y


def pixel_shuffle_bwd_raw(lib, stream, dy: Tensor, scale: int, with_bias: bool):
    from ._lib import PixelShuffleDesc
    dy = _chk(dy, "dy")
    B, C_, D, H, W = dy.shape
    desc = PixelShuffleDesc(B, C_, scale, D // scale, H // scale, W // scale)
    dz = torch.empty((B, C_ * scale ** 3, D // scale, H // scale, W // scale), dtype=_f32, device=dy.device)
    db = torch.empty((C_ * scale ** 3,), dtype=_f32, device=dy.device) if with_bias else None
    lib.call("vx_pixel_shuffle_bwd", desc, [dy], [dz, db], stream)
    return dz, db


class _PixelShuffleBias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bias, scale):
        ctx.scale, ctx.has_bias = int(scale), bias is not None
        return pixel_shuffle_fwd_raw(_lib.get_lib(), _stream(z), z, bias, int(scale))

    @staticmethod
    def backward(ctx, dy):
        dz, db = pixel_shuffle_bwd_raw(_lib.get_lib(), _stream(dy), dy.contiguous(), ctx.scale, ctx.has_bias)
        return dz, db, None


def pixel_shuffle_bias(z: Tensor, bias: Optional[Tensor], scale: int) -> Tensor:
    """PixelShuffle3d(scale)(z + bias[None, :, None, None, None]) in one pass (z: output of a bias-free convolution)."""
    return _PixelShuffleBias.apply(z.contiguous(), bias, int(scale))


# ----------------------------------------------------------------------------------------------------
# Convolutions of the glue layers through vx_conv_fwd / vx_conv_bwd:
#   dense 3x3x3, 16 input channels (+ fused bias / PixelShuffle(4))   Decoder.py:73-76,150-153   tcgen05 3xTF32 (fp32-accurate)
#   strided k = 2p-1 / transposed k = s = 2                           conv_blocks.py:10-17,31-35  fp32 SIMT
# ----------------------------------------------------------------------------------------------------
def conv_desc(x_shape, c_out: int, kernel: int, stride: int, pad: int, transposed: bool, shuffle: int = 0):
    from ._lib import ConvDesc
    B, Ci, D, H, W = x_shape
    return ConvDesc(B, Ci, int(c_out), D, H, W, int(kernel), int(stride), int(pad), int(bool(transposed)), int(shuffle))


def conv_out_shape(d) -> tuple:
    if d.transposed:
        return (d.B, d.C_out, d.D * d.stride, d.H * d.stride, d.W * d.stride)
    o = [(n + 2 * d.pad - d.kernel) // d.stride + 1 for n in (d.D, d.H, d.W)]
    if d.shuffle:
        return (d.B, d.C_out // d.shuffle ** 3, o[0] * d.shuffle, o[1] * d.shuffle, o[2] * d.shuffle)
    return (d.B, d.C_out, o[0], o[1], o[2])


def conv_fwd_raw(lib, stream, x: Tensor, w: Tensor, bias: Optional[Tensor], kernel: int, stride: int, pad: int, transposed: bool = False,
                 shuffle: int = 0) -> Tensor:
    x, w = _chk(x, "x"), _chk(w, "w")
    c_out = w.shape[1] if transposed else w.shape[0]
    d = conv_desc(x.shape, c_out, kernel, stride, pad, transposed, shuffle)
    y = torch.empty(conv_out_shape(d), dtype=_f32, device=x.device)
    ws = _ws(lib, "conv", d, x)
    lib.call_ws("vx_conv_fwd", d, [x, w, _chk(bias, "bias") if bias is not None else None], [y], ws, stream)
    return y


def conv_bwd_raw(lib, stream, dy: Tensor, x: Tensor, w: Tensor, kernel: int, stride: int, pad: int, transposed: bool = False,
                 shuffle: int = 0, need_dx: bool = True, need_db: bool = True):
    """Returns (dx or None, dw, db or None)."""
    dy, x, w = _chk(dy, "dy"), _chk(x, "x"), _chk(w, "w")
    c_out = w.shape[1] if transposed else w.shape[0]
    d = conv_desc(x.shape, c_out, kernel, stride, pad, transposed, shuffle)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    db = torch.empty((c_out,), dtype=_f32, device=x.device) if need_db else None
    ws = _ws(lib, "conv", d, x)
    lib.call_ws("vx_conv_bwd", d, [dy, x, w], [dx, dw, db], ws, stream)
    return dx, dw, db


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, cfg):
        ctx.save_for_backward(x, w)
        ctx.cfg, ctx.has_bias = cfg, bias is not None
        return conv_fwd_raw(_lib.get_lib(), _stream(x), x, w, bias, *cfg)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = conv_bwd_raw(_lib.get_lib(), _stream(x), dy.contiguous(), x, w, *ctx.cfg, need_dx=ctx.needs_input_grad[0],
                                  need_db=ctx.has_bias)
        return dx, dw, db, None


def conv3d(x: Tensor, w: Tensor, bias: Optional[Tensor], kernel: int, stride: int, pad: int, transposed: bool = False,
           shuffle: int = 0) -> Tensor:
    """Conv3d / ConvTranspose3d of the VeloxSeg glue layers on libveloxseg kernels (cubic kernel; see vx_conv_desc)."""
    if not x.is_cuda:
        raise NotImplementedError("veloxseg::conv3d has no CPU implementation")
    return _Conv.apply(x.contiguous(), w.contiguous(), bias, (int(kernel), int(stride), int(pad), bool(transposed), int(shuffle)))
