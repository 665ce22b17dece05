"""The three VeloxSeg model configurations and the training constants of the reference, as data.

Values restate config/models_config_{autopetii,hecktor2022,brats2021}.json ("VeloxSeg" block, :233/:238/:233) and
config/train_config_bs4.json of the reference; `load_model_config(path)` reads a reference JSON file unchanged, so the
reference's own config files remain a drop-in (`VeloxSeg(**load_model_config(path))`).
"""
from __future__ import annotations

import copy
import json

_COMMON = dict(
    patch_size=4, base_ch=16, conv_depths=[1, 1, 1, 1], kernel_sizes=[1, 3, 5], min_dim_group=[4, 8, 8, 16],
    conv_expansion_factor=[3, 3, 2, 2], attn_base_ch=16, depths=[1, 1, 1, 1],
    min_small_window_sizes=[[1, 1, 1], [1, 1, 1], [1, 1, 1], [1, 1, 1]], min_dim_head=[4, 8, 8, 16],
    ffn_expansion_ratio=[3, 3, 2, 2], proj_drop=0.1, conv_drop=0.1, spatial_dim=3)

MODEL_CONFIGS = {
    "autopetii": dict(_COMMON, input_size=[96, 96, 96], in_ch=[1, 1], n_classes=2, num_heads=[1, 2, 2, 4],
                      min_big_window_sizes=[[3, 3, 3], [6, 6, 6], [3, 3, 3], [3, 3, 3]]),
    # no num_heads key in the reference JSON: the constructor default [1, 2, 2, 4] applies
    "hecktor2022": dict(_COMMON, input_size=[128, 128, 64], in_ch=[1, 1], n_classes=2,
                        min_big_window_sizes=[[4, 4, 2], [8, 8, 4], [4, 4, 2], [4, 4, 2]]),
    "brats2021": dict(_COMMON, input_size=[96, 96, 96], in_ch=[4], n_classes=4, num_heads=[1, 2, 2, 4],
                      min_big_window_sizes=[[3, 3, 3], [6, 6, 6], [3, 3, 3], [3, 3, 3]]),
    # test-only miniature (not a reference config): 64^3 input, 8 base channels, keeps every code path
    # (3 window scales at level 1, 2 modalities, all four levels) at fixture-friendly size
    "tiny": dict(_COMMON, input_size=[64, 64, 64], in_ch=[1, 1], n_classes=2, base_ch=8, attn_base_ch=8,
                 num_heads=[1, 2, 2, 4], min_big_window_sizes=[[4, 4, 4], [4, 4, 4], [4, 4, 4], [2, 2, 2]]),
}

# config/train_config_bs4.json: batch_size 2 x RandCropByPosNegLabeld(num_samples=2) = 4 patches / step
TRAIN = dict(patches_per_step=4, deep_Loss_weight=[1, 1, 1, 1], RC_Loss_weight=0.5, Feature_Loss_weight=2.0,
             lr=2.5e-4, weight_decay=0.01, sw_overlap=0.25, sw_batch=2)


def model_config(name: str) -> dict:
    return copy.deepcopy(MODEL_CONFIGS[name])


def load_model_config(path: str, model_name: str = "VeloxSeg") -> dict:
    with open(path) as f:
        return json.load(f)[model_name]
