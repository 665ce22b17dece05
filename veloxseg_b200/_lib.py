"""ctypes binding of libveloxseg_sm100.so (include/veloxseg_abi.h).

The product loads exactly one library — the in-tree sm_100a build — and raises if it is missing: there is no
CPU or PyTorch fallback behind these ops.
"""
from __future__ import annotations

import ctypes as C
import os

VX_MAX_MODAL = 4
VX_MAX_SCALES = 6

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libveloxseg_sm100.so")


class JlcDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("groups", C.c_int32), ("expansion", C.c_int32), ("eps", C.c_float), ("drop_p", C.c_float),
                ("training", C.c_int32), ("seed", C.c_uint64), ("seed_offset", C.c_void_p)]


class MixerDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("S", C.c_int32), ("n_streams", C.c_int32), ("stream_ch", C.c_int32 * VX_MAX_MODAL),
                ("C_out", C.c_int32), ("has_addend", C.c_int32), ("eps", C.c_float)]


class InormDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("S", C.c_int32), ("eps", C.c_float), ("has_addend", C.c_int32),
                ("bias_channels", C.c_int32)]


class PwaDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("M", C.c_int32), ("C", C.c_int32), ("D", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("heads", C.c_int32), ("n_scales", C.c_int32),
                ("big", (C.c_int32 * 3) * VX_MAX_SCALES), ("small", (C.c_int32 * 3) * VX_MAX_SCALES),
                ("c_qk", C.c_int32), ("c_v", C.c_int32), ("ffn_expansion", C.c_int32), ("ln_eps", C.c_float),
                ("attn_drop", C.c_float), ("proj_drop", C.c_float), ("training", C.c_int32), ("seed", C.c_uint64),
                ("seed_offset", C.c_void_p)]


class PwaSaved(C.Structure):
    _fields_ = [("n_saved", C.c_int32), ("saved_bytes", C.c_size_t * 24)]


class GramDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("S", C.c_int32)]


class SdktLossDesc(C.Structure):
    _fields_ = [("n_elem", C.c_int32), ("n_teachers", C.c_int32)]


class ResizeDesc(C.Structure):
    _fields_ = [("planes", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("D", C.c_int32),
                ("H", C.c_int32), ("W", C.c_int32)]


class SegLossDesc(C.Structure):
    _fields_ = [("n_out", C.c_int32), ("B", C.c_int32), ("C", C.c_int32), ("S", C.c_int32), ("weights", C.c_float * 8)]


class LnpwDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C_in", C.c_int32), ("C_out", C.c_int32), ("S", C.c_int32), ("eps", C.c_float)]


class PatchEmbedDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C_in_total", C.c_int32), ("c_in_off", C.c_int32), ("C_in", C.c_int32),
                ("C_out", C.c_int32), ("patch", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class PixelShuffleDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("scale", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C_in", C.c_int32), ("C_out", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("kernel", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("transposed", C.c_int32), ("shuffle", C.c_int32)]


class AdamwDesc(C.Structure):
    _fields_ = [("n_chunks", C.c_int32), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_scale", C.c_float), ("hyper_on_device", C.c_int32)]


# every symbol include/veloxseg_abi.h declares
SYMBOLS = [
    "vx_version", "vx_last_error_string", "vx_launch_count", "vx_set_option", "vx_profile_enable", "vx_profile_reset",
    "vx_profile_report", "vx_profile_timeline",
    "vx_jlc_workspace", "vx_jlc_fwd", "vx_jlc_bwd",
    "vx_mixer_workspace", "vx_mixer_fwd", "vx_mixer_bwd",
    "vx_inorm_fwd", "vx_inorm_bwd",
    "vx_pwa_saved_layout", "vx_pwa_workspace", "vx_pwa_block_fwd", "vx_pwa_block_bwd", "vx_pwa_gather",
    "vx_gram_workspace", "vx_gram_fwd", "vx_gram_bwd",
    "vx_sdkt_loss_fwd", "vx_sdkt_loss_bwd",
    "vx_lnpw_workspace", "vx_lnpw_fwd", "vx_lnpw_bwd",
    "vx_resize_workspace", "vx_resize_trilinear_fwd", "vx_resize_trilinear_bwd",
    "vx_segloss_workspace", "vx_segloss_fwd", "vx_segloss_bwd",
    "vx_patch_embed_fwd", "vx_patch_embed_bwd", "vx_pixel_shuffle_fwd", "vx_pixel_shuffle_bwd", "vx_adamw_step",
    "vx_conv_workspace", "vx_conv_fwd", "vx_conv_bwd", "vx_conv3_trace", "vx_microbench", "vx_copy_block_async",
]

_WS_OPS = {"jlc", "mixer", "pwa_block", "gram_fwd", "lnpw", "segloss"}


def ptr_array(tensors):
    """void*[] from a list of tensors / None / raw ints."""
    vals = []
    for t in tensors:
        if t is None:
            vals.append(None)
        elif isinstance(t, int):
            vals.append(t)
        else:
            vals.append(t.data_ptr())
    return (C.c_void_p * max(len(vals), 1))(*vals)


class VxLib:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise RuntimeError(
                f"veloxseg_b200: native library {path} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no fallback path.")
        self.path = path
        self.c = C.CDLL(path)
        self.c.vx_last_error_string.restype = C.c_char_p
        self.c.vx_version.restype = C.c_int
        self.c.vx_launch_count.restype = C.c_uint64
        self.c.vx_profile_report.restype = C.c_size_t
        self.c.vx_profile_report.argtypes = [C.c_char_p, C.c_size_t]
        self.c.vx_profile_timeline.restype = C.c_size_t
        self.c.vx_profile_timeline.argtypes = [C.c_char_p, C.c_size_t]
        for name in ("vx_jlc_workspace", "vx_mixer_workspace", "vx_pwa_workspace", "vx_gram_workspace",
                     "vx_lnpw_workspace", "vx_resize_workspace", "vx_segloss_workspace", "vx_conv_workspace"):
            getattr(self.c, name).restype = C.c_size_t
            getattr(self.c, name).argtypes = [C.c_void_p]
        vp, sz = C.c_void_p, C.c_size_t
        for name in ("vx_jlc_fwd", "vx_jlc_bwd", "vx_mixer_fwd", "vx_mixer_bwd", "vx_pwa_block_fwd", "vx_pwa_block_bwd",
                     "vx_gram_fwd", "vx_lnpw_fwd", "vx_lnpw_bwd", "vx_resize_trilinear_bwd", "vx_segloss_fwd", "vx_conv_fwd", "vx_conv_bwd"):
            f = getattr(self.c, name)
            f.restype = C.c_int
            f.argtypes = [vp, vp, vp, vp, sz, vp]
        for name in ("vx_inorm_fwd", "vx_inorm_bwd", "vx_gram_bwd", "vx_sdkt_loss_fwd", "vx_sdkt_loss_bwd",
                     "vx_resize_trilinear_fwd", "vx_segloss_bwd", "vx_patch_embed_fwd", "vx_patch_embed_bwd", "vx_pixel_shuffle_fwd", "vx_pixel_shuffle_bwd", "vx_adamw_step"):
            f = getattr(self.c, name)
            f.restype = C.c_int
            f.argtypes = [vp, vp, vp, vp]
        self.c.vx_microbench.restype = C.c_int
        self.c.vx_microbench.argtypes = [C.c_int, C.c_int, vp, vp]
        self.c.vx_copy_block_async.restype = C.c_int
        self.c.vx_copy_block_async.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp]
        self.c.vx_conv3_trace.restype = C.c_int
        self.c.vx_conv3_trace.argtypes = [vp, C.c_int]
        if "VX_ATTN_TC" in os.environ:          # A/B switch of the tcgen05 attention forward (VX_OPT_ATTN_TC)
            self.c.vx_set_option(17, int(os.environ["VX_ATTN_TC"]))
        if "VX_JLC_SMALL_THREADS" in os.environ:
            self.c.vx_set_option(18, int(os.environ["VX_JLC_SMALL_THREADS"]))
        if "VX_FFN_TC" in os.environ:
            self.c.vx_set_option(19, int(os.environ["VX_FFN_TC"]))
        if "VX_PDL" in os.environ:              # A/B switch of the programmatic-dependent-launch path (VX_OPT_PDL)
            self.c.vx_set_option(16, int(os.environ["VX_PDL"]))
        self.c.vx_pwa_saved_layout.restype = C.c_int
        self.c.vx_pwa_saved_layout.argtypes = [vp, vp]
        self.c.vx_pwa_gather.restype = C.c_int
        self.c.vx_pwa_gather.argtypes = [vp, C.c_int32, vp, vp, vp, vp]

    def profile(self, on: bool):
        self.c.vx_profile_reset()
        self.c.vx_profile_enable(int(on))

    def profile_report(self):
        """[(scope, kernel, launches, total_ms, algorithmic_bytes, algorithmic_flops)] since the last profile(True)."""
        n = self.c.vx_profile_report(None, 0)
        buf = C.create_string_buffer(int(n) + 16)
        self.c.vx_profile_report(buf, len(buf))
        rows = []
        for line in buf.value.decode().splitlines():
            scope, kern, cnt, ms, nbytes, nflops = line.rsplit("|", 5)
            rows.append((scope, kern, int(cnt), float(ms), float(nbytes), float(nflops)))
        return rows

    def profile_timeline(self):
        """[(scope, kernel, start_us, dur_us)] of every launch recorded since the last profile(True)."""
        n = self.c.vx_profile_timeline(None, 0)
        buf = C.create_string_buffer(int(n) + 16)
        self.c.vx_profile_timeline(buf, len(buf))
        rows = []
        for line in buf.value.decode().splitlines():
            scope, kern, st, du = line.rsplit("|", 3)
            rows.append((scope, kern, float(st), float(du)))
        return rows

    def set_option(self, option: int, value: int):
        self.check(self.c.vx_set_option(int(option), int(value)), "vx_set_option")

    def last_error(self) -> str:
        s = self.c.vx_last_error_string()
        return s.decode() if s else ""

    def check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"veloxseg_b200: {what} failed (status {rc}): {self.last_error()}")

    def workspace(self, op: str, desc) -> int:
        return int(getattr(self.c, f"vx_{op}_workspace")(C.byref(desc)))

    def call_ws(self, name: str, desc, ins, outs, ws, stream: int):
        """fwd/bwd entry points that take a workspace."""
        rc = getattr(self.c, name)(C.byref(desc), ptr_array(ins), ptr_array(outs),
                                   ws.data_ptr() if ws is not None else None,
                                   ws.numel() * ws.element_size() if ws is not None else 0, stream)
        self.check(rc, name)

    def call(self, name: str, desc, ins, outs, stream: int):
        rc = getattr(self.c, name)(C.byref(desc), ptr_array(ins), ptr_array(outs), stream)
        self.check(rc, name)


_lib = None


def get_lib() -> VxLib:
    global _lib
    if _lib is None:
        _lib = VxLib(LIB_PATH)
    return _lib
