"""Host-side mirror of the reference's nn.Module surface for the hot path.

Every class keeps the reference's name, constructor signature, attribute names and registration order, so that
  * `state_dict()` keys / shapes / dtypes are identical (reference checkpoints load both ways), and
  * constructing under the same torch seed consumes the RNG identically (same He-init weights), because parameters are
    held by the same torch containers (`nn.Conv3d` etc.) created in the same order.
The `forward` bodies do not run those containers: they hand the parameters to `torch.ops.veloxseg.*`
(libveloxseg_sm100.so).  There is no PyTorch fallback in here: on a CPU tensor the ops raise.

Reference files (path:line into the reference repository):
  JLC / JLCLayer / DownConv / UpConv      model/components/conv_blocks.py:4-84
  LayerNorm / FFN / PositionalEmbedding / PatchMerging   model/components/attention_utils.py:11-168
  Paired_Windows_Attention / MultiModal_* / TransformerBlock / BasicLayer   model/components/PWA.py:10-511
  get_pram_matrix                          model/components/common_function.py:8-14
  PixelShuffle                             model/components/superpixel.py:4-17
  Encoder / Seg_Decoder / RC_Decoder / VeloxSeg   model/Encoder.py, model/Decoder.py, model/VeloxSeg.py
Nothing on the model path is a library (cuDNN / cuBLAS) call: the SURVEY.md section 8f "next" rows -- strided DownConv,
ConvTranspose UpConv, PatchEmbed stems, dense 3x3x3 output convs, 1x1 heads -- run on libveloxseg kernels too (ops.conv3d).
"""
from __future__ import annotations

from math import ceil
from typing import List, Sequence

import torch
import torch.nn as nn

from . import ops

__all__ = ["JLC", "JLCLayer", "DownConv", "UpConv", "LayerNorm", "FFN", "PositionalEmbedding", "PatchMerging",
           "MultiModal_Paired_Windows_Attention", "Paired_Windows_TransformerBlock", "Transformer_BasicLayer",
           "ModalMixer", "PixelShuffle", "get_pram_matrix", "Conv_Encoder", "Transformer_Encoder", "Encoder",
           "Seg_Decoder", "RC_Decoder", "VeloxSeg", "InitWeights_He", "PatchEmbed"]


def _w2(conv: nn.Module) -> torch.Tensor:
    """(Co, Ci, 1, 1, 1) -> (Co, Ci) view of a 1x1x1 conv weight."""
    return conv.weight.reshape(conv.weight.shape[0], -1)


def _bias(conv: nn.Module) -> torch.Tensor:
    if conv.bias is not None:
        return conv.bias
    return torch.zeros(conv.weight.shape[0], dtype=conv.weight.dtype, device=conv.weight.device)


# ----------------------------------------------------------------------------------------------------
# conv_blocks.py
# ----------------------------------------------------------------------------------------------------
class DownConv(nn.Module):
    """Strided conv (k = 2p-1, s = p, pad = p-1) + InstanceNorm.  conv_blocks.py:4-21.
    `forward(x, addend)` fuses the `+ attn_i` of Encoder.py:351-361 into the norm kernel."""

    def __init__(self, in_channels, out_channels, patch_size=2, groups=1, use_norm=True, dim=3):
        super().__init__()
        assert dim == 3, "veloxseg_b200 implements the 3-D path (spatial_dim=3 in every VeloxSeg config)"
        self.down = nn.Conv3d(in_channels, out_channels, kernel_size=2 * patch_size - 1, stride=patch_size,
                              padding=patch_size - 1, groups=groups)
        self.norm = nn.InstanceNorm3d(out_channels) if use_norm else nn.Identity()

    def forward(self, x, addend=None):
        c = self.down
        if c.groups != 1:
            raise NotImplementedError("veloxseg_b200.DownConv: groups = 1 (every VeloxSeg config)")
        k, s, p = c.kernel_size[0], c.stride[0], c.padding[0]
        if isinstance(self.norm, nn.Identity):
            y = ops.conv3d(x, c.weight, c.bias, k, s, p)
            return y if addend is None else y + addend
        # the bias cancels inside the affine-less norm: the convolution runs without it (no bias add, no bias-gradient
        # reduction over the conv output) and its (analytically zero) gradient comes from the norm's backward kernel
        z = ops.conv3d(x, c.weight, None, k, s, p)
        return ops.instance_norm_biased(z, c.bias, addend) if c.bias is not None else ops.instance_norm(z, addend)


class UpConv(nn.Module):
    """ConvTranspose3d (k = s = up_rate) + InstanceNorm.  conv_blocks.py:23-39."""

    def __init__(self, in_channels, out_channels, up_rate=2, groups=1, dim=3):
        super().__init__()
        assert dim == 3
        self.up = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=up_rate, stride=up_rate, groups=groups)
        self.norm = nn.InstanceNorm3d(out_channels)

    def forward(self, x, addend=None):
        c = self.up
        if c.groups != 1 or c.kernel_size != c.stride or any(c.padding) or any(c.output_padding):
            raise NotImplementedError("veloxseg_b200.UpConv: kernel = stride, no padding, groups = 1 (every VeloxSeg config)")
        z = ops.conv3d(x, c.weight, None, c.kernel_size[0], c.stride[0], 0, transposed=True)
        return ops.instance_norm_biased(z, c.bias, addend) if c.bias is not None else ops.instance_norm(z, addend)


class JLC(nn.Module):
    """o = x + sum_k GELU(IN(gconv_k(x))); y = o + Drop(W2 GELU(W1 IN(o)))   conv_blocks.py:41-75.
    (argument name `epansion_factor` is the reference's spelling.)"""

    def __init__(self, in_channels, kernel_sizes=[1, 3, 5], groups=1, epansion_factor=4, norm_type="IN",
                 activation='gelu', dropout=0.0, spatial_dim=3):
        super().__init__()
        if list(kernel_sizes) != [1, 3, 5] or norm_type != "IN" or activation.lower() != "gelu" or spatial_dim != 3:
            raise NotImplementedError("veloxseg_b200.JLC implements kernel_sizes [1,3,5] / IN / gelu / 3-D "
                                      "(the only setting used by the VeloxSeg configs)")
        self.spatial_convs = nn.ModuleList([
            nn.Sequential(nn.Conv3d(in_channels, in_channels, k, padding=k // 2, groups=groups),
                          nn.InstanceNorm3d(in_channels), nn.GELU())
            for k in kernel_sizes])
        self.channel_conv = nn.Sequential(
            nn.InstanceNorm3d(in_channels),
            nn.Conv3d(in_channels, in_channels * epansion_factor, 1, 1, 0),
            nn.GELU(),
            nn.Conv3d(in_channels * epansion_factor, in_channels, 1, 1, 0),
            nn.Dropout(dropout))
        self.groups, self.expansion, self.dropout = groups, epansion_factor, float(dropout)

    def forward(self, x):
        sc, cc = self.spatial_convs, self.channel_conv
        params = [sc[0][0].weight, sc[0][0].bias, sc[1][0].weight, sc[1][0].bias, sc[2][0].weight, sc[2][0].bias,
                  _w2(cc[1]), cc[1].bias, _w2(cc[3]), cc[3].bias]
        return ops.jlc(x, params, self.groups, self.expansion, self.dropout, self.training)


def JLCLayer(in_channels, depth=1, kernel_sizes=[1, 3, 5], groups=1, epansion_factor=4, activation='gelu',
             dropout=0.0, spatial_dim=3):
    """conv_blocks.py:77-84: nn.Sequential of `depth` JLC blocks."""
    return nn.Sequential(*[JLC(in_channels, kernel_sizes=kernel_sizes, groups=groups, epansion_factor=epansion_factor,
                               activation=activation, dropout=dropout, spatial_dim=spatial_dim)
                           for _ in range(depth)])


# ----------------------------------------------------------------------------------------------------
# attention_utils.py
# ----------------------------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """Parameter holder of the channels_first LayerNorm (attention_utils.py:11-43); the normalisation itself runs
    inside the fused ops that own one (pwa_block, ln_pointwise)."""

    def __init__(self, normalized_shape: int, eps: float = 1e-6, data_format: str = "channels_last", dim: int = 2):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.eps, self.data_format, self.normalized_shape, self.dim = eps, data_format, (normalized_shape,), dim


class FFN(nn.Module):
    """Parameter holder of the 1x1-conv FFN (attention_utils.py:45-71)."""

    def __init__(self, in_channels: int, groups: int = 1, expansion_ratio: int = 4, dropout_rate: float = 0.0,
                 act: str = "GELU", dim: int = 3):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if groups != 1 or str(act).upper() != "GELU":
            raise NotImplementedError("veloxseg_b200.FFN: groups=1 / GELU only")
        self.linear1 = nn.Conv3d(in_channels, in_channels * expansion_ratio, 1, 1, 0, groups=groups)
        self.linear2 = nn.Conv3d(in_channels * expansion_ratio, in_channels, 1, 1, 0, groups=groups)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = self.drop1
        self.expansion_ratio = expansion_ratio


class PositionalEmbedding(nn.Module):
    """Swin-style relative position table + int64 index buffer.  attention_utils.py:73-125."""

    def __init__(self, dim: int, num_heads: int, window_size: Sequence[int]):
        super().__init__()
        assert dim == 3
        self.dim, self.num_heads, self.window_size = dim, num_heads, window_size
        n0, n1, n2 = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * n0 - 1) * (2 * n1 - 1) * (2 * n2 - 1), num_heads))
        a = torch.arange(n0).view(n0, 1, 1).expand(n0, n1, n2).reshape(-1)
        b = torch.arange(n1).view(1, n1, 1).expand(n0, n1, n2).reshape(-1)
        c = torch.arange(n2).view(1, 1, n2).expand(n0, n1, n2).reshape(-1)
        index = ((a[:, None] - a[None, :] + n0 - 1) * ((2 * n1 - 1) * (2 * n2 - 1))
                 + (b[:, None] - b[None, :] + n1 - 1) * (2 * n2 - 1) + (c[:, None] - c[None, :] + n2 - 1))
        self.register_buffer("relative_position_index", index.contiguous())
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class PatchMerging(nn.Module):
    """2x2x2 space-to-depth (offset order 000..111) -> LN(8C) -> 1x1 8C->2C without bias.  attention_utils.py:127-168."""

    def __init__(self, in_ch: int, norm_layer=LayerNorm, dim: int = 3):
        super().__init__()
        assert dim == 3
        self.in_ch, self.dim, self.mid_ch = in_ch, dim, in_ch * 8
        self.reduction = nn.Conv3d(self.mid_ch, 2 * in_ch, 1, 1, 0, bias=False)
        self.norm = norm_layer(self.mid_ch, data_format="channels_first", dim=dim)

    def faeture_sample(self, x):   # (sic) reference method name
        """cat([x[:, :, i::2, j::2, k::2] for i, j, k in 000..111], 1) of the reference as ONE strided copy (the
        reference's 8 slices + cat cost 1 + 16 kernels per call and direction under autograd)."""
        B, C, D, H, W = x.shape
        if (D | H | W) & 1:
            return torch.cat([x[:, :, i::2, j::2, k::2] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dim=1)
        v = x.view(B, C, D // 2, 2, H // 2, 2, W // 2, 2)                 # (b, c, d, i, h, j, w, k)
        return v.permute(0, 3, 5, 7, 1, 2, 4, 6).reshape(B, 8 * C, D // 2, H // 2, W // 2)

    def forward(self, x):
        return ops.ln_pointwise(self.faeture_sample(x), self.norm.weight, self.norm.bias, _w2(self.reduction))


# ----------------------------------------------------------------------------------------------------
# PWA.py
# ----------------------------------------------------------------------------------------------------
class MultiModal_Paired_Windows_Attention(nn.Module):
    """Parameter holder + integer window geometry of the paired-window attention (PWA.py:10-104, 246-306).
    The arithmetic (LN, Q/K/V, gather, attention, scatter, mix, residual) runs in `ops.pwa_block`, driven by the
    enclosing Paired_Windows_TransformerBlock."""

    def __init__(self, input_size, in_channels, min_big_window_size=[3, 3, 3], min_small_window_size=[1, 1, 1],
                 scale_factor=2, num_heads=1, min_dim_head=4, qkv_bias=True, attn_drop=0.1, proj_drop=0.1,
                 norm_layer=LayerNorm, dim=3, use_pos_embed=True):
        super().__init__()
        assert dim == 3 and use_pos_embed and num_heads > 0
        self.input_size = list(input_size)
        self.mid_channels = max(in_channels)
        self.min_big_window_size, self.min_small_window_size = list(min_big_window_size), list(min_small_window_size)
        self.scale_factor, self.num_heads, self.min_dim_head, self.dim = scale_factor, num_heads, min_dim_head, dim
        self.channels_v = self.mid_channels
        self.big_window_size, self.small_window_size = self.get_window_sizes()
        self.n_hwd = [self.min_big_window_size[i] // self.min_small_window_size[i] for i in range(3)]
        self.num_bswin = len(self.big_window_size)
        self.num_heads_bswin = num_heads * self.num_bswin
        # registration order follows the reference: position_embedding, softmax, dropout_weight, then the lists
        self.position_embedding = PositionalEmbedding(dim=3, num_heads=num_heads, window_size=self.n_hwd)
        self.softmax = nn.Softmax(dim=-1)
        self.dropout_weight = nn.Dropout(attn_drop)
        self.in_channels, self.num_modalities = list(in_channels), len(in_channels)
        norms, qkv, mix, drops = [], [], [], []
        for m in range(self.num_modalities):
            norms.append(norm_layer(self.in_channels[m], data_format='channels_first', dim=3))
            qkv.append(nn.ModuleList([nn.Conv3d(self.in_channels[m], self.channels_qk, kernel_size=1, bias=qkv_bias),
                                      nn.Conv3d(self.in_channels[m], self.channels_qk, kernel_size=1, bias=qkv_bias),
                                      nn.Conv3d(self.in_channels[m], self.channels_v, kernel_size=1, bias=qkv_bias)]))
            mix.append(nn.Conv3d(self.channels_v, self.in_channels[m], kernel_size=1))
            drops.append(nn.Dropout(proj_drop))
        self.input_norms, self.qkv_proj = nn.ModuleList(norms), nn.ModuleList(qkv)
        self.mix_channels, self.dropout_attns = nn.ModuleList(mix), nn.ModuleList(drops)
        self.attn_drop, self.proj_drop = float(attn_drop), float(proj_drop)
        self.geo = dict(bws=self.big_window_size, sws=self.small_window_size, cqk=self.channels_qk, cv=self.channels_v,
                        n=self.n_hwd, heads=num_heads, nb=self.num_bswin)

    def get_window_sizes(self):
        """PWA.py:56-85 (integer, bit-exact): scales are added while ANY axis of the big window fits the input."""
        bws, sws = [], []
        bw, sw = list(self.min_big_window_size), list(self.min_small_window_size)
        while any(b <= s for b, s in zip(bw, self.input_size)):
            bws.append(list(bw))
            sws.append(list(sw))
            bw = [b * self.scale_factor for b in bw]
            sw = [s * self.scale_factor for s in sw]
        need = len(bws) * self.num_heads * self.min_dim_head
        self.channels_qk = need
        self.channels_v = ceil(self.channels_v / need) * need
        return bws, sws


class Paired_Windows_TransformerBlock(nn.Module):
    """y = 2x + Drop(Mix(attn)); z = y + Drop(W2 Drop(GELU(W1 LN(y))))   PWA.py:382-439 (double residual kept)."""

    def __init__(self, input_size, in_channels, min_big_window_size=[3, 3, 3], min_small_window_size=[1, 1, 1],
                 scale_factor=2, num_heads=1, min_dim_head=4, attn_drop=0.1, proj_drop=0.1, drop_path=0.0,
                 ffn_expansion_ratio=4, act_layer="GELU", norm_layer=LayerNorm, qkv_bias=True, dim=3):
        super().__init__()
        if drop_path and drop_path > 0.0:
            raise NotImplementedError("drop_path > 0 is never used by the VeloxSeg configs")
        if len(set(in_channels)) != 1:
            raise NotImplementedError("veloxseg_b200: all modality streams of a PWA level carry the same channel count")
        self.input_size, self.in_channels, self.num_modalities = input_size, list(in_channels), len(in_channels)
        self.attn = MultiModal_Paired_Windows_Attention(
            input_size=input_size, in_channels=in_channels, min_big_window_size=min_big_window_size,
            min_small_window_size=min_small_window_size, scale_factor=scale_factor, num_heads=num_heads,
            min_dim_head=min_dim_head, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=proj_drop,
            norm_layer=norm_layer, dim=dim, use_pos_embed=True)
        self.drop_path = nn.Identity()
        self.ffns, self.norms = nn.ModuleList(), nn.ModuleList()
        for m in range(self.num_modalities):
            self.ffns.append(FFN(in_channels[m], expansion_ratio=ffn_expansion_ratio, dropout_rate=proj_drop,
                                 act=act_layer, dim=dim))
            self.norms.append(norm_layer(in_channels[m], data_format='channels_first', dim=dim))
        self.ffn_expansion_ratio = ffn_expansion_ratio

    def _params(self) -> List[torch.Tensor]:
        a, out = self.attn, []
        for m in range(self.num_modalities):
            q, k, v = a.qkv_proj[m]
            out += [a.input_norms[m].weight, a.input_norms[m].bias, _w2(q), _bias(q), _w2(k), _bias(k), _w2(v), _bias(v),
                    _w2(a.mix_channels[m]), a.mix_channels[m].bias, self.norms[m].weight, self.norms[m].bias,
                    _w2(self.ffns[m].linear1), self.ffns[m].linear1.bias, _w2(self.ffns[m].linear2),
                    self.ffns[m].linear2.bias]
        return out

    def forward(self, xs: Sequence[torch.Tensor]):
        a = self.attn
        assert len(xs) == self.num_modalities, \
            f"The number of modalities should be {self.num_modalities}, but got {len(xs)}"
        return ops.pwa_block(list(xs), self._params(), a.position_embedding.relative_position_bias_table,
                             a.position_embedding.relative_position_index, a.geo, self.ffn_expansion_ratio,
                             a.attn_drop, a.proj_drop, self.training)


class Transformer_BasicLayer(nn.Module):
    """`depth` PWA blocks then per-modality PatchMerging.  PWA.py:444-511.  forward(list) -> (list, list | None)."""

    def __init__(self, input_size, in_channels, depth=2, min_big_window_size=[3, 3, 3],
                 min_small_window_size=[1, 1, 1], scale_factor=2, num_heads=1, min_dim_head=4, attn_drop=0.1,
                 proj_drop=0.1, drop_path=0, ffn_expansion_ratio=4, act_layer="GELU", norm_layer=LayerNorm,
                 qkv_bias=True, do_downsample=True, dim=3):
        super().__init__()
        self.num_modalities = len(in_channels)
        self.blocks = nn.ModuleList([
            Paired_Windows_TransformerBlock(
                input_size=input_size, in_channels=in_channels, min_big_window_size=min_big_window_size,
                min_small_window_size=min_small_window_size, scale_factor=scale_factor, num_heads=num_heads,
                min_dim_head=min_dim_head, attn_drop=attn_drop, proj_drop=proj_drop,
                drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                ffn_expansion_ratio=ffn_expansion_ratio, act_layer=act_layer, norm_layer=norm_layer,
                qkv_bias=qkv_bias, dim=dim)
            for i in range(depth)])
        self.downs = None
        if do_downsample:
            self.downs = nn.ModuleList([PatchMerging(in_ch=in_channels[m], norm_layer=norm_layer, dim=dim)
                                        for m in range(self.num_modalities)])

    def attn_forward(self, xs):
        for blk in self.blocks:
            xs = blk(xs)
        return xs

    def down_forward(self, xs):
        if self.downs is None:
            return None
        return [self.downs[m](xs[m]) for m in range(self.num_modalities)]

    def forward(self, xs):
        xs = self.attn_forward(xs)
        return xs, self.down_forward(xs)


# ----------------------------------------------------------------------------------------------------
# common_function.py / superpixel.py
# ----------------------------------------------------------------------------------------------------
def get_pram_matrix(x: torch.Tensor) -> torch.Tensor:
    """SDKT Gram: einsum('bmhwd,bnhwd->bmn') / (C*H*W*D).  common_function.py:8-14."""
    if x.dim() != 5:
        raise NotImplementedError("veloxseg_b200.get_pram_matrix: 5-D (B,C,H,W,D) input")
    return ops.gram(x)


class PixelShuffle(nn.Module):
    """'b (c s1 s2 s3) d h w -> b c (d s1) (h s2) (w s3)'.  superpixel.py:15 (a pure layout op: view + permute)."""

    def __init__(self, scale, spatial_dim=3):
        super().__init__()
        if spatial_dim != 3:
            raise ValueError("spatial_dim should be 3")
        self.scale, self.spatial_dim = scale, spatial_dim

    def forward(self, x):
        s = self.scale
        B, Cs, D, H, W = x.shape
        c = Cs // s ** 3
        return x.view(B, c, s, s, s, D, H, W).permute(0, 1, 5, 2, 6, 3, 7, 4).reshape(B, c, D * s, H * s, W * s)


def conv_shuffle(seq: nn.Sequential, x):
    """`Sequential(Conv3d k3 p1, PixelShuffle)` (decoder.out_conv1 / reconstruction out_conv, Decoder.py:73-76,150-153).
    16 input channels, scale 4 (every VeloxSeg config): ONE tcgen05 kernel -- convolution, bias and shuffle store; the
    (B, 64 n, D, H, W) intermediate is never written.  Other shapes: convolution kernel + fused bias/shuffle kernel.
    Parameters and state_dict keys are those of the Sequential."""
    conv, ps = seq[0], seq[1]
    if conv.kernel_size != (3, 3, 3) or conv.stride != (1, 1, 1) or conv.padding != (1, 1, 1) or conv.groups != 1:
        raise NotImplementedError("veloxseg_b200.conv_shuffle: Conv3d k3 s1 p1 (Decoder.py:73-76,150-153)")
    if conv.in_channels == 16 and ps.scale == 4 and conv.out_channels % 64 == 0 and (conv.out_channels <= 128 or conv.out_channels % 128 == 0):
        return ops.conv3d(x, conv.weight, conv.bias, 3, 1, 1, shuffle=4)
    return ops.pixel_shuffle_bias(ops.conv3d(x, conv.weight, None, 3, 1, 1), conv.bias, ps.scale)


def conv1x1(conv: nn.Conv3d, x):
    """Deep-supervision heads `out_conv2..4` (Conv3d 1x1 with bias, Decoder.py:155-158)."""
    return ops.conv3d(x, conv.weight, conv.bias, 1, 1, 0)


class ModalMixer(nn.Sequential):
    """`nn.Sequential(Conv3d 1x1, InstanceNorm3d)` of Encoder.py:334-337 / Decoder.py:54-57 with a fused forward:
    keeps the Sequential's state_dict keys (`0.weight`, `0.bias`); accepts either the concatenated tensor (the
    reference's call form) or the list of streams plus the tensor the result is added to."""

    def __init__(self, in_channels, out_channels):
        super().__init__(nn.Conv3d(in_channels, out_channels, 1, 1, 0), nn.InstanceNorm3d(out_channels))

    def forward(self, x, addend=None):
        streams = list(x) if isinstance(x, (list, tuple)) else [x]
        return ops.modal_mixer(streams, _w2(self[0]), self[0].bias, addend)


class PatchEmbed(nn.Module):
    """monai.networks.blocks.PatchEmbed as used by Encoder.py:150-156 (patch_norm=False): Conv3d k = s = patch held
    as `.proj`.  MONAI is an un-vendored dependency of the reference (monai==1.5.0, requirements.txt:3)."""

    def __init__(self, patch_size=2, in_chans=1, embed_dim=48, norm_layer=None, spatial_dims=3):
        super().__init__()
        if norm_layer is not None or spatial_dims != 3:
            raise NotImplementedError("veloxseg_b200.PatchEmbed: patch_norm=False, 3-D")
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        p = self.proj
        return ops.conv3d(x, p.weight, p.bias, p.kernel_size[0], p.stride[0], 0)


class InitWeights_He:
    """kaiming_normal_(a=neg_slope) on every conv weight, zero biases.  components/initialization.py:3-14."""

    def __init__(self, neg_slope=1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)
        elif isinstance(module, (nn.BatchNorm2d, nn.BatchNorm3d, nn.GroupNorm, nn.LayerNorm)):
            nn.init.constant_(module.weight, 1)
            nn.init.constant_(module.bias, 0)


# ----------------------------------------------------------------------------------------------------
# Encoder.py
# ----------------------------------------------------------------------------------------------------
class Conv_Encoder(nn.Module):
    """down1..4 + layer1..4 holders (Encoder.py:13-66); Encoder.forward drives them directly."""

    def __init__(self, patch_size=4, in_ch=1, base_ch=16, depths=[1, 1, 1, 1], kernel_sizes=[1, 3, 5],
                 min_dim_group=[4, 8, 8, 16], expansion_factor=[3, 3, 2, 2], dropout=0.0, spatial_dim=3):
        super().__init__()
        self.down1 = DownConv(in_ch, base_ch, patch_size=patch_size, dim=spatial_dim)
        self.down2 = DownConv(base_ch, base_ch * 2, patch_size=2, dim=spatial_dim)
        self.down3 = DownConv(base_ch * 2, base_ch * 4, patch_size=2, dim=spatial_dim)
        self.down4 = DownConv(base_ch * 4, base_ch * 8, patch_size=2, dim=spatial_dim)
        groups = [base_ch * 2 ** i // min_dim_group[i] for i in range(4)]
        for i in range(4):
            setattr(self, f"layer{i + 1}", JLCLayer(base_ch * 2 ** i, depths[i], kernel_sizes, groups[i],
                                                    expansion_factor[i], dropout=dropout, spatial_dim=spatial_dim))

    def forward(self, x):
        enc1 = self.layer1(self.down1(x))
        enc2 = self.layer2(self.down2(enc1))
        enc3 = self.layer3(self.down3(enc2))
        enc4 = self.layer4(self.down4(enc3))
        return enc1, enc2, enc3, enc4


class Transformer_Encoder(nn.Module):
    """Per-modality PatchEmbed -> 4 x (PWA block(s) -> PatchMerging).  Encoder.py:88-204."""

    def __init__(self, input_size, patch_size, in_channels, embed_dim=16, depths=[2, 2, 2, 2],
                 min_big_window_sizes=[[3, 3, 3], [6, 6, 6], [3, 3, 3], [3, 3, 3]],
                 min_small_window_sizes=[[1, 1, 1]] * 4, scale_factors=[2, 2, 2, 2], num_heads=[1, 2, 2, 4],
                 min_dim_head=[4, 8, 8, 16], ffn_expansion_ratio=[3, 3, 2, 2], attn_drop=0.1, proj_drop=0.1,
                 drop_path=0, act_layer="GELU", norm_layer=LayerNorm, patch_norm=False, qkv_bias=True, spatial_dim=3):
        super().__init__()
        self.in_channels, self.num_modalities, self.num_layers = in_channels, len(in_channels), len(depths)
        self.patch_size = patch_size
        self.patch_embeds = nn.ModuleList([
            PatchEmbed(patch_size=patch_size, in_chans=in_channels[m], embed_dim=embed_dim,
                       norm_layer=norm_layer if patch_norm else None, spatial_dims=spatial_dim)
            for m in range(self.num_modalities)])
        self.pos_drop = nn.Dropout(p=proj_drop)
        dpr = [x.item() for x in torch.linspace(0, drop_path, sum(depths))]
        self.layers = nn.ModuleList()
        size = [int(s) // patch_size for s in input_size]
        for i in range(self.num_layers):
            self.layers.append(Transformer_BasicLayer(
                input_size=list(size), in_channels=[int(embed_dim * 2 ** i)] * self.num_modalities, depth=depths[i],
                min_big_window_size=min_big_window_sizes[i], min_small_window_size=min_small_window_sizes[i],
                scale_factor=scale_factors[i], num_heads=num_heads[i], min_dim_head=min_dim_head[i],
                attn_drop=attn_drop, proj_drop=proj_drop, drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])],
                ffn_expansion_ratio=ffn_expansion_ratio[i], act_layer=act_layer, norm_layer=norm_layer,
                qkv_bias=qkv_bias, do_downsample=i < self.num_layers - 1, dim=spatial_dim))
            size = [s // 2 for s in size]

    def embed(self, xs):
        if not xs.requires_grad:
            # read each modality's channels where they lie in the input (no slice copies, no layout conversion, and no data
            # gradient: this is the network input)
            out, off = [], 0
            for m in range(self.num_modalities):
                proj = self.patch_embeds[m].proj
                out.append(self.pos_drop(ops.patch_embed(xs, off, proj.weight, proj.bias)))
                off += proj.weight.shape[1]
            return out
        xs = torch.chunk(xs, self.num_modalities, dim=1)
        xs = [self.pos_drop(self.patch_embeds[m](xs[m])) for m in range(self.num_modalities)]
        return [x.contiguous() for x in xs]

    def forward(self, xs):
        outs, cur = [], self.embed(xs)
        for i in range(self.num_layers):
            attn, cur = self.layers[i](cur)
            outs.append(attn)
        return tuple(outs)


class Encoder(nn.Module):
    """Dual-branch encoder with the modal mixer between the branches.  Encoder.py:207-367."""

    def __init__(self, input_size, patch_size, in_ch, base_ch=16, conv_depths=[1, 1, 1, 1], kernel_sizes=[1, 3, 5],
                 min_dim_group=[4, 8, 8, 16], conv_expansion_factor=[4, 4, 4, 4], attn_base_ch=16,
                 depths=[2, 2, 2, 2], min_big_window_sizes=[[3, 3, 3], [6, 6, 6], [3, 3, 3], [3, 3, 3]],
                 min_small_window_sizes=[[1, 1, 1]] * 4, min_dim_head=[4, 8, 8, 16], scale_factors=[2, 2, 2, 2],
                 num_heads=[1, 2, 4, 8], attn_drop=0.1, proj_drop=0.1, drop_path=0, ffn_expansion_ratio=[4, 4, 4, 4],
                 act_layer="GELU", norm_layer=LayerNorm, patch_norm=False, qkv_bias=True, conv_drop=0.0,
                 spatial_dim=3):
        super().__init__()
        self.in_channels, self.num_modalities = in_ch, len(in_ch)
        self.encoder_attn = Transformer_Encoder(
            input_size=input_size, patch_size=patch_size, in_channels=in_ch, embed_dim=attn_base_ch, depths=depths,
            min_big_window_sizes=min_big_window_sizes, min_small_window_sizes=min_small_window_sizes,
            scale_factors=scale_factors, num_heads=num_heads, min_dim_head=min_dim_head, attn_drop=attn_drop,
            proj_drop=proj_drop, drop_path=drop_path, ffn_expansion_ratio=ffn_expansion_ratio, act_layer=act_layer,
            norm_layer=norm_layer, patch_norm=patch_norm, qkv_bias=qkv_bias, spatial_dim=spatial_dim)
        self.encoder_conv = Conv_Encoder(
            patch_size=patch_size, in_ch=sum(in_ch), base_ch=base_ch, depths=conv_depths, kernel_sizes=kernel_sizes,
            min_dim_group=min_dim_group, expansion_factor=conv_expansion_factor, dropout=conv_drop,
            spatial_dim=spatial_dim)
        for i in range(4):
            setattr(self, f"attn2conv_{i + 1}",
                    ModalMixer(attn_base_ch * 2 ** i * self.num_modalities, base_ch * 2 ** i))

    # The conv branch at level i needs only attn_i; the transformer branch at level i+1 needs only attn_i too.  On CUDA the
    # conv branch therefore runs on a side stream, one level behind the transformer branch (and the same in backward:
    # autograd replays every op on the stream of its forward).  Levels 2-4 are small kernels that leave most of the 148
    # SMs idle, so the two branches overlap almost for free.  Under graph capture the fork/join become graph edges.
    pipeline_branches = True

    def _conv_level(self, i, attn, t):
        # x_i = IN(down_i(.)) + IN(W . cat_m(attn_i[m]) + b): the mixer kernel reads the M streams in place (no cat)
        ec = self.encoder_conv
        mixed = getattr(self, f"attn2conv_{i + 1}")(attn, addend=getattr(ec, f"down{i + 1}")(t))
        return getattr(ec, f"layer{i + 1}")(mixed)

    def forward(self, x):
        encs, t = [], x
        if self.pipeline_branches and x.is_cuda:
            main = torch.cuda.current_stream(x.device)
            if getattr(self, "_side_dev", None) != x.device:
                self._side, self._side_dev = torch.cuda.Stream(device=x.device), x.device
            side = self._side
            side.wait_stream(main)
            x.record_stream(side)
            attns, cur = [], self.encoder_attn.embed(x)
            for i in range(4):
                attn, cur = self.encoder_attn.layers[i](cur)
                attns.append(attn)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for a in attn:
                        a.record_stream(side)
                    t = self._conv_level(i, attn, t)
                    encs.append(t)
            main.wait_stream(side)
            for e in encs:
                e.record_stream(main)
            attns = tuple(attns)
        else:
            attns = self.encoder_attn(x)
            for i in range(4):
                t = self._conv_level(i, attns[i], t)
                encs.append(t)
        if self.training:
            return [list(a) for a in attns], encs
        return tuple(encs)


# ----------------------------------------------------------------------------------------------------
# Decoder.py
# ----------------------------------------------------------------------------------------------------
def _jlc_groups(ch, min_dim_group):
    return [ch * 2 ** i // min_dim_group[i] for i in range(4)]


class RC_Decoder(nn.Module):
    """Reconstruction teacher.  Decoder.py:11-94.  The adapters read (attn_i[m], enc_i) as two streams (no concat)."""

    def __init__(self, in_channel, enc_channel, dec_channel, patch_size, depths=[1, 1, 1, 1], kernel_sizes=[1, 3, 5],
                 min_dim_group=[4, 8, 8, 16], expansion_factor=[3, 3, 2, 2], spatial_dim=3, dropout=0.0):
        super().__init__()
        self.enc2rc_4 = ModalMixer(enc_channel * 8, dec_channel * 8)
        self.enc2rc_3 = ModalMixer(enc_channel * 4, dec_channel * 4)
        self.enc2rc_2 = ModalMixer(enc_channel * 2, dec_channel * 2)
        self.enc2rc_1 = ModalMixer(enc_channel, dec_channel)
        self.layer_up3 = UpConv(dec_channel * 8, dec_channel * 4, up_rate=2, dim=spatial_dim)
        self.layer_up2 = UpConv(dec_channel * 4, dec_channel * 2, up_rate=2, dim=spatial_dim)
        self.layer_up1 = UpConv(dec_channel * 2, dec_channel, up_rate=2, dim=spatial_dim)
        g = _jlc_groups(dec_channel, min_dim_group)
        self.layer1 = JLCLayer(dec_channel, depths[0], kernel_sizes, g[0], expansion_factor[0], dropout=dropout, spatial_dim=spatial_dim)
        self.layer2 = JLCLayer(dec_channel * 2, depths[1], kernel_sizes, g[1], expansion_factor[1], dropout=dropout, spatial_dim=spatial_dim)
        self.layer3 = JLCLayer(dec_channel * 4, depths[2], kernel_sizes, g[2], expansion_factor[2], dropout=dropout, spatial_dim=spatial_dim)
        self.out_conv = nn.Sequential(
            nn.Conv3d(dec_channel, (patch_size ** 3) * in_channel, kernel_size=3, stride=1, padding=1),
            PixelShuffle(scale=patch_size, spatial_dim=spatial_dim))

    def forward(self, enc1, enc2, enc3, enc4):
        """enc_i: concatenated tensor (reference call form) or the (attn_i[m], enc_i) pair."""
        e4, e3, e2, e1 = self.enc2rc_4(enc4), self.enc2rc_3(enc3), self.enc2rc_2(enc2), self.enc2rc_1(enc1)
        up3 = self.layer3(self.layer_up3(e4, addend=e3))
        up2 = self.layer2(self.layer_up2(up3, addend=e2))
        up1 = self.layer1(self.layer_up1(up2, addend=e1))
        if self.training:
            return conv_shuffle(self.out_conv, up1), get_pram_matrix(up1)
        return conv_shuffle(self.out_conv, up1)


class Seg_Decoder(nn.Module):
    """Segmentation student.  Decoder.py:97-179."""

    def __init__(self, patch_size, base_ch=32, out_ch=2, depths=[1, 1, 1, 1], kernel_sizes=[1, 3, 5],
                 min_dim_group=[4, 8, 8, 16], expansion_factor=[3, 3, 2, 2], dropout=0.0, deep_supervision=False,
                 spatial_dim=3):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.layer_up3 = UpConv(base_ch * 8, base_ch * 4, up_rate=2, dim=spatial_dim)
        self.layer_up2 = UpConv(base_ch * 4, base_ch * 2, up_rate=2, dim=spatial_dim)
        self.layer_up1 = UpConv(base_ch * 2, base_ch, up_rate=2, dim=spatial_dim)
        g = _jlc_groups(base_ch, min_dim_group)
        self.layer1 = JLCLayer(base_ch, depths[0], kernel_sizes, g[0], expansion_factor[0], dropout=dropout, spatial_dim=spatial_dim)
        self.layer2 = JLCLayer(base_ch * 2, depths[1], kernel_sizes, g[1], expansion_factor[1], dropout=dropout, spatial_dim=spatial_dim)
        self.layer3 = JLCLayer(base_ch * 4, depths[2], kernel_sizes, g[2], expansion_factor[2], dropout=dropout, spatial_dim=spatial_dim)
        self.out_conv1 = nn.Sequential(
            nn.Conv3d(base_ch, (patch_size ** 3) * out_ch, kernel_size=3, stride=1, padding=1),
            PixelShuffle(scale=patch_size, spatial_dim=spatial_dim))
        if deep_supervision:
            self.out_conv2 = nn.Conv3d(base_ch * 2, out_ch, 1, 1)
            self.out_conv3 = nn.Conv3d(base_ch * 4, out_ch, 1, 1)
            self.out_conv4 = nn.Conv3d(base_ch * 8, out_ch, 1, 1)

    def forward(self, enc1, enc2, enc3, enc4, head_post=None):
        """`head_post` (training, deep supervision, CUDA): callable applied to every head output (VeloxSeg passes its
        `scale_prediction`).  The three low-resolution heads and their post-processing then run on a forked stream as soon
        as their input exists, beside the rest of the decoder, instead of queueing behind `out_conv1` on the main stream
        (~140 us of the step's critical path); under graph capture the fork is a parallel branch."""
        side = None
        if head_post is not None and self.training and self.deep_supervision and enc4.is_cuda and getattr(self, "fork_heads", True):
            main = torch.cuda.current_stream(enc4.device)
            if getattr(self, "_head_stream_dev", None) != enc4.device:
                self._head_stream, self._head_stream_dev = torch.cuda.Stream(device=enc4.device), enc4.device
            side = self._head_stream
            heads = {}

            def head(i, conv, feat):
                side.wait_stream(main)                    # `feat` has just been produced on main
                feat.record_stream(side)
                with torch.cuda.stream(side):
                    heads[i] = head_post(conv1x1(conv, feat))

            head(4, self.out_conv4, enc4)
        up3 = self.layer3(self.layer_up3(enc4, addend=enc3))
        if side is not None:
            head(3, self.out_conv3, up3)
        up2 = self.layer2(self.layer_up2(up3, addend=enc2))
        if side is not None:
            head(2, self.out_conv2, up2)
        up1 = self.layer1(self.layer_up1(up2, addend=enc1))
        out = conv_shuffle(self.out_conv1, up1)
        if self.training:
            pram = get_pram_matrix(up1)
            if side is not None:
                out = head_post(out)
                main.wait_stream(side)
                for t in heads.values():
                    t.record_stream(main)
                return [out, heads[2], heads[3], heads[4]], pram
            if self.deep_supervision:
                outs = [out, conv1x1(self.out_conv2, up2), conv1x1(self.out_conv3, up3), conv1x1(self.out_conv4, enc4)]
            else:
                outs = [out]
            return ([head_post(o) for o in outs] if head_post is not None else outs), pram
        return out


# ----------------------------------------------------------------------------------------------------
# VeloxSeg.py
# ----------------------------------------------------------------------------------------------------
class VeloxSeg(nn.Module):
    """Same constructor and train/eval output contract as model/VeloxSeg.py:64-226.
    train: [seg x4 (resized to input), rcs, gram_student, gram_teacher x M];  eval: logits."""

    def __init__(self, input_size, patch_size, in_ch, n_classes=2, base_ch=16, conv_depths=[1, 1, 1, 1],
                 kernel_sizes=[1, 3, 5], min_dim_group=[4, 8, 8, 16], conv_expansion_factor=[3, 3, 2, 2],
                 attn_base_ch=16, depths=[2, 2, 2, 2],
                 min_big_window_sizes=[[3, 3, 3], [6, 6, 6], [3, 3, 3], [3, 3, 3]],
                 min_small_window_sizes=[[1, 1, 1], [1, 1, 1], [1, 1, 1], [1, 1, 1]], min_dim_head=[4, 8, 8, 16],
                 scale_factors=[2, 2, 2, 2], num_heads=[1, 2, 2, 4], attn_drop=0.1, proj_drop=0.1, drop_path=0,
                 ffn_expansion_ratio=[3, 3, 2, 2], act_layer="GELU", norm_layer=LayerNorm, patch_norm=False,
                 qkv_bias=True, conv_drop=0.0, deep_supervision=True, spatial_dim=3):
        super().__init__()
        self.size, self.spatial_dim, self.patch_size = input_size, spatial_dim, patch_size
        self.in_ch, self.n_classes, self.num_modalities = in_ch, n_classes, len(in_ch)
        self.encoder = Encoder(
            input_size=input_size, patch_size=patch_size, in_ch=in_ch, base_ch=base_ch, conv_depths=conv_depths,
            kernel_sizes=kernel_sizes, min_dim_group=min_dim_group, conv_expansion_factor=conv_expansion_factor,
            attn_base_ch=attn_base_ch, depths=depths, min_big_window_sizes=min_big_window_sizes,
            min_small_window_sizes=min_small_window_sizes, min_dim_head=min_dim_head, scale_factors=scale_factors,
            num_heads=num_heads, attn_drop=attn_drop, proj_drop=proj_drop, drop_path=drop_path,
            ffn_expansion_ratio=ffn_expansion_ratio, act_layer=act_layer, norm_layer=norm_layer,
            patch_norm=patch_norm, qkv_bias=qkv_bias, conv_drop=conv_drop, spatial_dim=spatial_dim)
        self.decoder = Seg_Decoder(
            patch_size=patch_size, base_ch=base_ch, out_ch=n_classes, depths=conv_depths, kernel_sizes=kernel_sizes,
            min_dim_group=min_dim_group, expansion_factor=conv_expansion_factor, dropout=conv_drop,
            deep_supervision=deep_supervision, spatial_dim=spatial_dim)
        self.rc_decoders = nn.ModuleList([
            RC_Decoder(in_channel=in_ch[i], enc_channel=attn_base_ch + base_ch, dec_channel=base_ch,
                       patch_size=patch_size, depths=conv_depths, kernel_sizes=kernel_sizes,
                       min_dim_group=min_dim_group, expansion_factor=conv_expansion_factor, spatial_dim=spatial_dim,
                       dropout=conv_drop)
            for i in range(len(in_ch))])
        self.init_weights()

    def init_weights(self):
        self.apply(InitWeights_He(neg_slope=1e-2))

    def scale_prediction(self, pred):
        return ops.resize_trilinear(pred, tuple(self.size))

    # Independent branches of the training graph (segmentation student, one reconstruction teacher per modality) are
    # issued on forked CUDA streams: their kernels are small (most occupy a fraction of the 148 SMs), so running the
    # branches side by side -- in forward and, because autograd replays every node on its forward stream, in backward
    # too -- fills the machine.  Under CUDA-graph capture the forks become parallel graph branches.
    parallel_branches = True

    def _side_streams(self, device):
        if getattr(self, "_streams_dev", None) != device:
            self._streams = [torch.cuda.Stream(device=device) for _ in range(self.num_modalities)]
            self._streams_dev = device
        return self._streams

    def forward(self, x):
        if self.training:
            attns, encs = self.encoder(x)
            # encoder / decoder boundary for a two-phase backward (train.TrainStep, data-parallel overlap): with `split_boundary`
            # the decoders consume detached aliases of the encoder outputs (same storage, new autograd leaves), so phase 1
            # differentiates the loss down to `_boundary` only and phase 2 feeds those gradients into `_boundary_src`
            if getattr(self, "split_boundary", False):
                self._boundary_src = [a for lvl in attns for a in lvl] + list(encs)
                attns = [[a.detach().requires_grad_(True) for a in lvl] for lvl in attns]
                encs = [e.detach().requires_grad_(True) for e in encs]
                self._boundary = [a for lvl in attns for a in lvl] + list(encs)
            fork = self.parallel_branches and x.is_cuda
            rcs, rc_prams = [None] * self.num_modalities, [None] * self.num_modalities
            if fork:
                main = torch.cuda.current_stream(x.device)
                streams = self._side_streams(x.device)
                for m, st in enumerate(streams):
                    st.wait_stream(main)
                    with torch.cuda.stream(st):
                        feats = [[attns[i][m], encs[i]] for i in range(4)]
                        for a, e in feats:          # produced on `main`, consumed here: tell the caching allocator
                            a.record_stream(st)
                            e.record_stream(st)
                        rcs[m], rc_prams[m] = self.rc_decoders[m](*feats)
            self.decoder.fork_heads = fork
            pred, dec_pram = self.decoder(*encs, head_post=self.scale_prediction)      # heads resized inside (forked stream)
            if fork:
                for m, st in enumerate(streams):
                    main.wait_stream(st)
                    rcs[m].record_stream(main)
                    rc_prams[m].record_stream(main)
            else:
                for m in range(self.num_modalities):
                    rcs[m], rc_prams[m] = self.rc_decoders[m](*[[attns[i][m], encs[i]] for i in range(4)])
            return pred + [torch.cat(rcs, dim=1)] + [dec_pram] + rc_prams
        return self.decoder(*self.encoder(x))
