"""veloxseg_b200 — the VeloxSeg JLC / modal-mixer / PWA / SDKT hot path on NVIDIA B200 (sm_100a).

`veloxseg_b200.ops`   torch custom ops over the C ABI (include/veloxseg_abi.h, libveloxseg_sm100.so)
`veloxseg_b200.nn`    the reference's nn.Module surface (same names / signatures / state_dict keys)
Importing the package does not load the native library; the first op call does, and fails loudly if it is missing.
"""
__version__ = "0.1.0"


def set_precision(mode: str) -> None:
    """"fp32" (default) or "bf16" numerics for the tensor-core kernels; see veloxseg_b200.ops.set_precision."""
    from . import ops
    ops.set_precision(mode)
