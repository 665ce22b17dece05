"""CPU oracle for the VeloxSeg hot path (TEST INFRASTRUCTURE — never imported by veloxseg_b200/).

A functional, torch-fp32 restatement of the reference algorithm: every function takes plain tensors
(or a `params` dict keyed by the reference's own state_dict names) and re-derives the result with
explicit index arithmetic where the reference leans on einops / max_pool3d / interpolate.  It is
the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) and is
never the thing that is shipped or measured as the product.

Parity status: PINNED against the unmodified reference run in the build container
(tests/golden/make_golden.py imports /root/reference through tests/golden/monai_shim and writes
tests/golden/*.pt; tests/test_oracle_golden.py re-checks this file against those fixtures).  The
reference itself ships no golden vectors, and the MONAI-owned pieces
(sliding_window_inference, DiceLoss) are restated from the published MONAI 1.5.0 algorithm — those
two are "parity unpinned" (monai is not installable here).

Citations are path:line into /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------------------------
def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.InstanceNorm3d(affine=False, track_running_stats=False): biased variance over the spatial
    axes per (b, c).  model/components/common_function.py:62-66."""
    dims = tuple(range(2, x.dim()))
    mean = x.mean(dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dims, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def gelu(x: Tensor) -> Tensor:
    """exact-erf GELU (nn.GELU() default).  common_function.py:93-94."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def layer_norm_cf(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-6) -> Tensor:
    """channels_first LayerNorm: per-voxel mean / biased var over C.  attention_utils.py:32-43."""
    u = x.mean(1, keepdim=True)
    s = ((x - u) ** 2).mean(1, keepdim=True)
    xh = (x - u) / torch.sqrt(s + eps)
    shape = (1, -1) + (1,) * (x.dim() - 2)
    return weight.view(shape) * xh + bias.view(shape)


def pointwise(x: Tensor, weight: Tensor, bias: Tensor | None) -> Tensor:
    """1x1x1 convolution as a channel contraction (weight (Co, Ci, 1,1,1) or (Co, Ci))."""
    w = weight.reshape(weight.shape[0], weight.shape[1])
    y = torch.einsum("oc,bc...->bo...", w, x)
    if bias is not None:
        y = y + bias.view((1, -1) + (1,) * (x.dim() - 2))
    return y


# ----------------------------------------------------------------------------------------------
# JLC  (conv_blocks.py:41-75)
# ----------------------------------------------------------------------------------------------
def jlc(x: Tensor, p: Dict[str, Tensor], prefix: str, groups: int, kernel_sizes=(1, 3, 5)) -> Tensor:
    """o = x + sum_k GELU(IN(gconv_k(x)+b_k));  y = o + W2 GELU(W1 IN(o) + b1) + b2   (dropout p = 0).
    conv_blocks.py:50-75; parameter names `spatial_convs.{i}.0.*`, `channel_conv.{1,3}.*`."""
    o = x
    for i, k in enumerate(kernel_sizes):
        z = F.conv3d(x, p[f"{prefix}spatial_convs.{i}.0.weight"], p[f"{prefix}spatial_convs.{i}.0.bias"],
                     padding=k // 2, groups=groups)
        o = o + gelu(instance_norm(z))
    h = gelu(pointwise(instance_norm(o), p[f"{prefix}channel_conv.1.weight"], p[f"{prefix}channel_conv.1.bias"]))
    return o + pointwise(h, p[f"{prefix}channel_conv.3.weight"], p[f"{prefix}channel_conv.3.bias"])


def jlc_groups(channels: int, level: int, min_dim_group=(4, 8, 8, 16)) -> int:
    """groups = C / min_dim_group[level].  Encoder.py:58, Decoder.py:66,145."""
    return channels // min_dim_group[level]


# ----------------------------------------------------------------------------------------------
# modal mixer  (Encoder.py:334-337, 344-351) and the identical RC adapters (Decoder.py:54-57)
# ----------------------------------------------------------------------------------------------
def modal_mixer(streams: Sequence[Tensor], weight: Tensor, bias: Tensor, addend: Tensor | None = None) -> Tensor:
    """IN(W . cat(streams, dim=1) + b) [+ addend]."""
    y = instance_norm(pointwise(torch.cat(list(streams), dim=1), weight, bias))
    return y if addend is None else addend + y


# ----------------------------------------------------------------------------------------------
# PWA integer geometry  (PWA.py:56-85, attention_utils.py:89-117)
# ----------------------------------------------------------------------------------------------
def window_pyramid(input_size: Sequence[int], min_big: Sequence[int], min_small: Sequence[int],
                   scale_factor: int, num_heads: int, min_dim_head: int, in_channels: int):
    """Returns (big_windows, small_windows, channels_qk, channels_v).  PWA.py:56-85: scales are added
    while ANY axis of the big window still fits the input."""
    bws, sws = [], []
    bw, sw = list(min_big), list(min_small)
    while any(b <= s for b, s in zip(bw, input_size)):
        bws.append(list(bw))
        sws.append(list(sw))
        bw = [b * scale_factor for b in bw]
        sw = [s * scale_factor for s in sw]
    need = len(bws) * num_heads * min_dim_head
    return bws, sws, need, int(math.ceil(in_channels / need)) * need


def relative_position_index(n: Sequence[int]) -> Tensor:
    """Swin-style (l, l) int64 index into the ((2n0-1)(2n1-1)(2n2-1), heads) table.
    attention_utils.py:89-117."""
    n0, n1, n2 = n
    idx = torch.empty(n0 * n1 * n2, n0 * n1 * n2, dtype=torch.int64)
    coords = [(a, b, c) for a in range(n0) for b in range(n1) for c in range(n2)]
    for i, (a, b, c) in enumerate(coords):
        for j, (a2, b2, c2) in enumerate(coords):
            idx[i, j] = ((a - a2 + n0 - 1) * (2 * n1 - 1) * (2 * n2 - 1)
                         + (b - b2 + n1 - 1) * (2 * n2 - 1) + (c - c2 + n2 - 1))
    return idx


def gather_tokens(x: Tensor, heads: int, bws, sws) -> Tuple[Tensor, List[List[int]], List[int]]:
    """window_gathering_3d, PWA.py:106-140, by explicit indexing.
    x (B, nb*heads*c, H, W, D) -> tokens (B, heads, Ns, l, c); channel split is (scale, head, c);
    per scale j: small windows of size sw_j are max-pooled, big windows of bw_j are enumerated
    row-major, tokens inside a window row-major over (bw/sw)."""
    B, Ct, H, W, D = x.shape
    nb = len(bws)
    c = Ct // (nb * heads)
    xs = x.view(B, nb, heads, c, H, W, D)
    outs, Ns, n_ref = [], [], None
    for j in range(nb):
        bh, bw_, bd = bws[j]
        sh, sw_, sd = sws[j]
        Nh, Nw, Nd = H // bh, W // bw_, D // bd
        nh, nw, nd = bh // sh, bw_ // sw_, bd // sd
        xi = xs[:, j]                                               # (B, heads, c, H, W, D)
        xi = xi.reshape(B, heads, c, Nh, nh, sh, Nw, nw, sw_, Nd, nd, sd)
        xi = xi.amax(dim=(5, 8, 11))                                # pool the small window
        xi = xi.permute(0, 1, 3, 5, 7, 4, 6, 8, 2)                  # B, head, Nh,Nw,Nd, nh,nw,nd, c
        outs.append(xi.reshape(B, heads, Nh * Nw * Nd, nh * nw * nd, c))
        Ns.append([Nh, Nw, Nd])
        assert n_ref is None or n_ref == [nh, nw, nd]
        n_ref = [nh, nw, nd]
    return torch.cat(outs, dim=2), Ns, n_ref


def _lerp_matrix(n: int, s: int, dtype, device=None) -> Tensor:
    """(n*s, n) matrix of 1-D linear interpolation with align_corners=True
    (src = dst * (n-1)/(n*s-1)); identity when s == 1.  Matches F.interpolate, PWA.py:190."""
    m = torch.zeros(n * s, n, dtype=dtype)
    if n * s == 1:
        m[0, 0] = 1
        return m.to(device) if device is not None else m
    scale = (n - 1) / (n * s - 1) if n * s > 1 else 0.0
    for d in range(n * s):
        src = d * scale
        i0 = min(int(math.floor(src)), n - 1)
        i1 = min(i0 + 1, n - 1)
        w1 = src - i0
        m[d, i0] += 1.0 - w1
        m[d, i1] += w1
    return m.to(device) if device is not None else m


def scatter_tokens(tok: Tensor, heads: int, bws, sws, Ns, n) -> Tensor:
    """window_scattering_3d, PWA.py:177-200: per scale, each big window's (nh,nw,nd) token grid is
    up-sampled independently (trilinear, align_corners=True, factor sw_j), windows are re-tiled and the
    channel axis re-assembled as (scale, head, c)."""
    B, h, _, l, c = tok.shape
    nh, nw, nd = n
    outs, idx = [], 0
    for j in range(len(bws)):
        Nh, Nw, Nd = Ns[j]
        N = Nh * Nw * Nd
        sh, sw_, sd = sws[j]
        t = tok[:, :, idx:idx + N].reshape(B, h, Nh, Nw, Nd, nh, nw, nd, c)
        Mh, Mw, Md = (_lerp_matrix(nh, sh, tok.dtype, tok.device), _lerp_matrix(nw, sw_, tok.dtype, tok.device),
                      _lerp_matrix(nd, sd, tok.dtype, tok.device))
        t = torch.einsum("bhxyzijkc,pi,qj,rk->bhcxpyqzr", t, Mh, Mw, Md)
        outs.append(t.reshape(B, h * c, Nh * nh * sh, Nw * nw * sw_, Nd * nd * sd))
        idx += N
    return torch.cat(outs, dim=1)


# ----------------------------------------------------------------------------------------------
# PWA attention / block  (PWA.py:308-379, 433-439)
# ----------------------------------------------------------------------------------------------
def pwa_geometry(input_size, C, min_big, min_small, scale_factor, heads, min_dim_head):
    bws, sws, cqk, cv = window_pyramid(input_size, min_big, min_small, scale_factor, heads, min_dim_head, C)
    n = [min_big[i] // min_small[i] for i in range(3)]
    return dict(bws=bws, sws=sws, cqk=cqk, cv=cv, n=n, heads=heads, nb=len(bws))


def pwa_attention(xs: Sequence[Tensor], p: Dict[str, Tensor], prefix: str, geo) -> List[Tensor]:
    """MultiModal_Paired_Windows_Attention.forward with dropout p=0.  PWA.py:329-379.
    Returns out_m = x_m + Mix_m(scatter(attn))  (the block adds x_m once more, PWA.py:436)."""
    M = len(xs)
    heads, bws, sws = geo["heads"], geo["bws"], geo["sws"]
    qs, ks, vs = [], [], []
    Ns = n = None
    for m in range(M):
        xh = layer_norm_cf(xs[m], p[f"{prefix}input_norms.{m}.weight"], p[f"{prefix}input_norms.{m}.bias"])
        q = pointwise(xh, p[f"{prefix}qkv_proj.{m}.0.weight"], p.get(f"{prefix}qkv_proj.{m}.0.bias"))
        k = pointwise(xh, p[f"{prefix}qkv_proj.{m}.1.weight"], p.get(f"{prefix}qkv_proj.{m}.1.bias"))
        v = pointwise(xh, p[f"{prefix}qkv_proj.{m}.2.weight"], p.get(f"{prefix}qkv_proj.{m}.2.bias"))
        q, Ns, n = gather_tokens(q, heads, bws, sws)
        k, _, _ = gather_tokens(k, heads, bws, sws)
        v, _, _ = gather_tokens(v, heads, bws, sws)
        qs.append(q), ks.append(k), vs.append(v)
    l = qs[0].shape[-2]
    q, k, v = torch.cat(qs, -2), torch.cat(ks, -2), torch.cat(vs, -2)     # token index = m*l + t
    c = q.shape[-1]
    s = torch.einsum("bhnic,bhnjc->bhnij", q, k) / (c ** 0.5)
    table = p[f"{prefix}position_embedding.relative_position_bias_table"]
    index = p[f"{prefix}position_embedding.relative_position_index"]
    bias = table[index[:l, :l].reshape(-1)].reshape(l, l, heads).permute(2, 0, 1)      # (h, l, l)
    s = s + bias.repeat(1, M, M)[None, :, None]              # same tile for every modality pair, PWA.py:316-320
    w = torch.softmax(s, dim=-1)
    a = torch.einsum("bhnij,bhnjc->bhnic", w, v)
    outs = []
    for m in range(M):
        am = scatter_tokens(a[:, :, :, m * l:(m + 1) * l], heads, bws, sws, Ns, n)
        outs.append(xs[m] + pointwise(am, p[f"{prefix}mix_channels.{m}.weight"], p[f"{prefix}mix_channels.{m}.bias"]))
    return outs


def pwa_block(xs: Sequence[Tensor], p: Dict[str, Tensor], prefix: str, geo) -> List[Tensor]:
    """Paired_Windows_TransformerBlock.forward: y = x + attn(x) (= 2x + ...), z = y + FFN(LN(y)).
    PWA.py:433-439; FFN attention_utils.py:65-71."""
    att = pwa_attention(xs, p, prefix + "attn.", geo)
    outs = []
    for m, x in enumerate(xs):
        y = x + att[m]
        t = layer_norm_cf(y, p[f"{prefix}norms.{m}.weight"], p[f"{prefix}norms.{m}.bias"])
        t = gelu(pointwise(t, p[f"{prefix}ffns.{m}.linear1.weight"], p[f"{prefix}ffns.{m}.linear1.bias"]))
        outs.append(y + pointwise(t, p[f"{prefix}ffns.{m}.linear2.weight"], p[f"{prefix}ffns.{m}.linear2.bias"]))
    return outs


def patch_merging(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """2x2x2 space-to-depth (offset order 000,001,...,111, offset-major channels) -> LN(8C) -> 1x1 8C->2C
    without bias.  attention_utils.py:144-168."""
    parts = [x[:, :, a::2, b::2, c::2] for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    t = layer_norm_cf(torch.cat(parts, 1), p[f"{prefix}norm.weight"], p[f"{prefix}norm.bias"])
    return pointwise(t, p[f"{prefix}reduction.weight"], None)


# ----------------------------------------------------------------------------------------------
# SDKT  (common_function.py:8-14, utils/loss.py:58-64)
# ----------------------------------------------------------------------------------------------
def gram(x: Tensor) -> Tensor:
    """G[b,m,n] = sum_s x[b,m,s] x[b,n,s] / (C*S)."""
    B, C = x.shape[:2]
    f = x.reshape(B, C, -1)
    return torch.bmm(f, f.transpose(1, 2)) / (C * f.shape[-1])


def sdkt_loss(g_student: Tensor, g_teachers: Sequence[Tensor]) -> Tensor:
    """sum_m MSE(G_s, G_t^m) / M, teacher NOT detached."""
    tot = 0.0
    for g in g_teachers:
        tot = tot + ((g_student - g) ** 2).mean()
    return tot / len(g_teachers)


# ----------------------------------------------------------------------------------------------
# glue convs (stay library-backed in the product too)
# ----------------------------------------------------------------------------------------------
def down_conv(x, p, prefix, patch):
    """DownConv: k=2p-1, s=p, pad=p-1, then IN.  conv_blocks.py:4-21."""
    return instance_norm(F.conv3d(x, p[f"{prefix}down.weight"], p[f"{prefix}down.bias"], stride=patch, padding=patch - 1))


def up_conv(x, p, prefix):
    """UpConv: ConvTranspose3d k=s=2, then IN.  conv_blocks.py:23-39."""
    return instance_norm(F.conv_transpose3d(x, p[f"{prefix}up.weight"], p[f"{prefix}up.bias"], stride=2))


def pixel_shuffle3d(x: Tensor, s: int) -> Tensor:
    """'b (c s1 s2 s3) d h w -> b c (d s1) (h s2) (w s3)'.  superpixel.py:15."""
    B, Cs, D, H, W = x.shape
    c = Cs // (s ** 3)
    return x.view(B, c, s, s, s, D, H, W).permute(0, 1, 5, 2, 6, 3, 7, 4).reshape(B, c, D * s, H * s, W * s)


# ----------------------------------------------------------------------------------------------
# whole model  (VeloxSeg.py:186-226, Encoder.py:190-204,339-367, Decoder.py:78-94,160-179)
# ----------------------------------------------------------------------------------------------
class ModelSpec:
    """The integer facts of one VeloxSeg config (config/models_config_*.json "VeloxSeg" block)."""

    def __init__(self, cfg: dict):
        d = dict(n_classes=2, base_ch=16, conv_depths=[1, 1, 1, 1], kernel_sizes=[1, 3, 5],
                 min_dim_group=[4, 8, 8, 16], conv_expansion_factor=[3, 3, 2, 2], attn_base_ch=16,
                 depths=[2, 2, 2, 2], min_big_window_sizes=[[3, 3, 3], [6, 6, 6], [3, 3, 3], [3, 3, 3]],
                 min_small_window_sizes=[[1, 1, 1]] * 4, min_dim_head=[4, 8, 8, 16], scale_factors=[2, 2, 2, 2],
                 num_heads=[1, 2, 2, 4], ffn_expansion_ratio=[3, 3, 2, 2], deep_supervision=True)   # VeloxSeg.py:64-94
        d.update(cfg)
        self.__dict__.update(d)
        self.M = len(self.in_ch)
        self.level_size = []
        size = [s // self.patch_size for s in self.input_size]
        for _ in range(4):
            self.level_size.append(list(size))
            size = [s // 2 for s in size]
        self.geo = [pwa_geometry(self.level_size[i], self.attn_base_ch * 2 ** i, self.min_big_window_sizes[i],
                                 self.min_small_window_sizes[i], self.scale_factors[i], self.num_heads[i],
                                 self.min_dim_head[i]) for i in range(4)]


def _jlc_layer(x, p, prefix, depth, C, level, spec):
    for d in range(depth):
        x = jlc(x, p, f"{prefix}{d}.", jlc_groups(C, level, spec.min_dim_group), tuple(spec.kernel_sizes))
    return x


def encoder(x: Tensor, p, spec: ModelSpec):
    """Returns (attn[level][m], enc[level])."""
    xs = list(torch.split(x, list(spec.in_ch), dim=1))     # torch.chunk(M) for equal in_ch; Encoder.py:192
    cur = [F.conv3d(xs[m], p[f"encoder.encoder_attn.patch_embeds.{m}.proj.weight"],
                    p[f"encoder.encoder_attn.patch_embeds.{m}.proj.bias"], stride=spec.patch_size)
           for m in range(spec.M)]
    attn = []
    for i in range(4):
        for d in range(spec.depths[i]):
            cur = pwa_block(cur, p, f"encoder.encoder_attn.layers.{i}.blocks.{d}.", spec.geo[i])
        attn.append(cur)
        if i < 3:
            cur = [patch_merging(cur[m], p, f"encoder.encoder_attn.layers.{i}.downs.{m}.") for m in range(spec.M)]
    enc, t = [], x
    for i in range(4):
        C = spec.base_ch * 2 ** i
        d = down_conv(t, p, f"encoder.encoder_conv.down{i + 1}.", spec.patch_size if i == 0 else 2)
        t = modal_mixer(attn[i], p[f"encoder.attn2conv_{i + 1}.0.weight"], p[f"encoder.attn2conv_{i + 1}.0.bias"], d)
        t = _jlc_layer(t, p, f"encoder.encoder_conv.layer{i + 1}.", spec.conv_depths[i], C, i, spec)
        enc.append(t)
    return attn, enc


def _decode(feats, p, prefix, spec):
    up = feats[3]
    ups = {}
    for lvl in (3, 2, 1):
        C = spec.base_ch * 2 ** (lvl - 1)
        up = _jlc_layer(feats[lvl - 1] + up_conv(up, p, f"{prefix}layer_up{lvl}."), p, f"{prefix}layer{lvl}.",
                        spec.conv_depths[lvl - 1], C, lvl - 1, spec)
        ups[lvl] = up
    return ups


def forward(x: Tensor, p, spec: ModelSpec, training: bool):
    attn, enc = encoder(x, p, spec)
    ups = _decode(enc, p, "decoder.", spec)
    out = pixel_shuffle3d(F.conv3d(ups[1], p["decoder.out_conv1.0.weight"], p["decoder.out_conv1.0.bias"], padding=1),
                          spec.patch_size)
    if not training:
        return out
    preds = [out]
    if spec.deep_supervision:
        preds += [pointwise(ups[2], p["decoder.out_conv2.weight"], p["decoder.out_conv2.bias"]),
                  pointwise(ups[3], p["decoder.out_conv3.weight"], p["decoder.out_conv3.bias"]),
                  pointwise(enc[3], p["decoder.out_conv4.weight"], p["decoder.out_conv4.bias"])]
    preds = [F.interpolate(q, size=tuple(spec.input_size), mode="trilinear", align_corners=True) for q in preds]
    rcs, grams_t = [], []
    for m in range(spec.M):
        pre = f"rc_decoders.{m}."
        feats = [modal_mixer([attn[i][m], enc[i]], p[f"{pre}enc2rc_{i + 1}.0.weight"], p[f"{pre}enc2rc_{i + 1}.0.bias"])
                 for i in range(4)]
        u = _decode(feats, p, pre, spec)
        rcs.append(pixel_shuffle3d(F.conv3d(u[1], p[f"{pre}out_conv.0.weight"], p[f"{pre}out_conv.0.bias"], padding=1),
                                   spec.patch_size))
        grams_t.append(gram(u[1]))
    return preds + [torch.cat(rcs, 1)] + [gram(ups[1])] + grams_t


# ----------------------------------------------------------------------------------------------
# loss  (utils/loss.py:30-66, utils/runtime.py:125-174; MONAI DiceLoss restated — parity unpinned)
# ----------------------------------------------------------------------------------------------
def dice_loss(logits: Tensor, target: Tensor) -> Tensor:
    """monai DiceLoss(include_background=False, to_onehot_y=True, softmax=True), defaults
    smooth_nr = smooth_dr = 1e-5, reduction mean over (B, C-1)."""
    C = logits.shape[1]
    prob = torch.softmax(logits, 1)[:, 1:]
    onehot = F.one_hot(target.squeeze(1).long(), C).movedim(-1, 1).to(prob.dtype)[:, 1:]
    dims = tuple(range(2, logits.dim()))
    inter = (prob * onehot).sum(dims)
    den = prob.sum(dims) + onehot.sum(dims)
    return (1.0 - (2.0 * inter + 1e-5) / (den + 1e-5)).mean()


def total_loss(outputs: Sequence[Tensor], labels: Tensor, recon_target: Tensor, M: int,
               deep_weights=(1, 1, 1, 1), rc_w=0.5, feat_w=2.0) -> Tensor:
    n_seg = len(outputs) - (2 + M)
    w = [float(v) for v in deep_weights]
    w = [1.0 / n_seg] * n_seg if len(w) != n_seg else [v / sum(w) for v in w]
    seg = 0.0
    for wi, o in zip(w, outputs[:n_seg]):
        seg = seg + wi * (F.cross_entropy(o, labels.squeeze(1).long()) + dice_loss(o, labels))
    rc = ((outputs[n_seg] - recon_target) ** 2).mean()
    return seg + rc_w * rc + feat_w * sdkt_loss(outputs[n_seg + 1], outputs[n_seg + 2:n_seg + 2 + M])


def dice_metric(pred: Tensor, gt: Tensor) -> float:
    """2|A∩B| / (|A|+|B|+1e-5) on binary masks.  utils/metric/metrics.py:93-94."""
    pred, gt = pred.bool(), gt.bool()
    return float(2.0 * (pred & gt).sum() / (pred.sum() + gt.sum() + 1e-5))


# ----------------------------------------------------------------------------------------------
# sliding window (MONAI 1.5.0 sliding_window_inference restated; utils/inference_runtime.py:4-19)
# ----------------------------------------------------------------------------------------------
def sw_scan_interval(image: Sequence[int], roi: Sequence[int], overlap: float) -> List[int]:
    return [int(r) if r == i else max(int(r * (1 - overlap)), 1) for i, r in zip(image, roi)]


def sw_slices(image: Sequence[int], roi: Sequence[int], overlap: float) -> List[Tuple[int, ...]]:
    """Start corners of every window, first spatial axis slowest (meshgrid 'ij'); last window of an
    axis is clamped back so it ends at the image border."""
    interval = sw_scan_interval(image, roi, overlap)
    starts = []
    for i, r, s in zip(image, roi, interval):
        num = int(math.ceil(i / s))
        scan = next((d for d in range(num) if d * s + r >= i), None)
        scan = (scan + 1) if scan is not None else 1
        ax = []
        for d in range(scan):
            st = d * s
            st -= max(st + r - i, 0)
            ax.append(st)
        starts.append(ax)
    return [(a, b, c) for a in starts[0] for b in starts[1] for c in starts[2]]


def sliding_window(x: Tensor, roi: Sequence[int], predictor, overlap: float = 0.25, sw_batch: int = 2) -> Tensor:
    """mode='constant' importance map (ones), constant-0 padding up to roi, output / count."""
    B = x.shape[0]
    size = list(x.shape[2:])
    pads = [max(r - s, 0) for r, s in zip(roi, size)]
    lo = [pd // 2 for pd in pads]
    if any(pads):
        pad = []
        for ax in (2, 1, 0):
            pad += [lo[ax], pads[ax] - lo[ax]]
        x = F.pad(x, pad)
    image = list(x.shape[2:])
    corners = sw_slices(image, roi, overlap)
    nwin = len(corners)
    out = count = None
    for g in range(0, B * nwin, sw_batch):
        ids = range(g, min(g + sw_batch, B * nwin))
        win = torch.cat([x[i // nwin:i // nwin + 1, :, corners[i % nwin][0]:corners[i % nwin][0] + roi[0],
                           corners[i % nwin][1]:corners[i % nwin][1] + roi[1],
                           corners[i % nwin][2]:corners[i % nwin][2] + roi[2]] for i in ids])
        y = predictor(win)
        if out is None:
            out = torch.zeros((B, y.shape[1]) + tuple(image), dtype=y.dtype)
            count = torch.zeros((1, 1) + tuple(image), dtype=y.dtype)
            for (a, b, c) in corners:
                count[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += 1
        for k, i in enumerate(ids):
            a, b, c = corners[i % nwin]
            out[i // nwin, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]] += y[k]
    out = out / count
    return out[:, :, lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]]
