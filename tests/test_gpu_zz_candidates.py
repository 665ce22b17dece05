"""Parity of the CANDIDATE kernels that are off by default (VX_OPT_JLC_CONV_TC: tcgen05 implicit-GEMM JLC convolutions,
jlc_tc.cu).  They were written without GPU access and have only been checked on the CPU shim, so this file is opt-in
(VX_CANDIDATES=1) and sorts last: it must not stand between the product path's parity tests and a green run."""
import os

import pytest
import torch

from tests._util import close, jlc_param_dict, jlc_params, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("VX_CANDIDATES") != "1", reason="opt-in: VX_CANDIDATES=1")]
DEV = "cuda:0"


@pytest.mark.parametrize("C,shape,B", [(16, (24, 24, 24), 1), (16, (24, 24, 24), 4), (16, (32, 32, 16), 2), (16, (9, 7, 12), 1),
                                       (32, (12, 12, 12), 4), (32, (16, 16, 8), 1), (32, (5, 6, 8), 2)])     # C = 32: 8 channels per group
def test_jlc_conv_tensor_core_vs_simt_and_oracle(C, shape, B):
    from oracle import veloxseg_oracle as O
    from veloxseg_b200 import _lib, ops
    lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
    groups, e = 4, 3
    torch.manual_seed(4)
    x = torch.randn(B, C, *shape)
    params = jlc_params(C, groups, e, seed=6)
    xd, pd = x.to(DEV), [p.to(DEV) for p in params]
    lib.set_option(8, 0)              # never the small-volume kernels (the small cases)
    try:
        y0, z0, o0, h0, s0 = ops.jlc_fwd_raw(lib, st, xd, pd, groups, e)
        lib.set_option(11, 1)
        y1, z1, o1, h1, s1 = ops.jlc_fwd_raw(lib, st, xd, pd, groups, e)
        torch.cuda.synchronize()
        assert rel_err(z1, z0) < 2e-6 and not torch.equal(z1, z0), rel_err(z1, z0)
        assert rel_err(s1, s0) < 2e-5 and rel_err(y1, y0) < 2e-5
        xr = x.clone().requires_grad_(True)
        pr = [p.clone().requires_grad_(True) for p in params]
        yr = O.jlc(xr, jlc_param_dict(pr), "", groups)
        assert rel_err(y1.cpu(), yr) < 1e-4
        dy = torch.randn_like(yr)
        grads = torch.autograd.grad(yr, [xr] + pr, dy)
        got = ops.jlc_bwd_raw(lib, st, dy.to(DEV), xd, z1, o1, h1, s1, pd, groups, e)
        torch.cuda.synchronize()
        for i, (g, r) in enumerate(zip(got, grads)):
            assert close(g.cpu(), r, rtol=1e-3, atol=2e-5), (i, rel_err(g.cpu(), r))
    finally:
        lib.set_option(11, 0)
        lib.set_option(8, 512)
