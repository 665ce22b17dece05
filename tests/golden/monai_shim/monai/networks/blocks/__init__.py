import torch.nn as nn


class PatchEmbed(nn.Module):
    """monai.networks.blocks.PatchEmbed restated for the configs in use: Conv3d(k = s = patch) held
    as `.proj` (patch_norm=False in every VeloxSeg config, so the norm branch is never built)."""

    def __init__(self, patch_size=2, in_chans=1, embed_dim=48, norm_layer=None, spatial_dims=3):
        super().__init__()
        assert norm_layer is None and spatial_dims == 3
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x)
