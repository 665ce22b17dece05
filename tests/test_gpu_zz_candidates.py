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


@pytest.mark.parametrize("Cout,shape,B", [(128, (24, 24, 24), 4), (64, (24, 24, 24), 2), (256, (24, 24, 24), 1), (128, (32, 32, 16), 1),
                                          (64, (5, 6, 8), 2)])
def test_dense_conv_tensor_core_vs_library(Cout, shape, B):
    """conv_dense_tc.cu (out_conv1 / reconstruction out_conv on tcgen05, tf32) against conv3d in fp64 on tf32-rounded
    operands (tight) and on the fp32 operands (tf32 tolerance), forward; backward goes through the library either way."""
    import torch.nn.functional as F
    from veloxseg_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, 16, *shape, generator=g).to(DEV)
    w = (torch.randn(Cout, 16, 3, 3, 3, generator=g) * 0.05).to(DEV)

    def tf32(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    ops.dense_conv_tc_enable(True)
    try:
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        z = ops.dense_conv3(xr, wr)
        torch.cuda.synchronize()
        ref_t = F.conv3d(tf32(x).double(), tf32(w).double(), None, 1, 1)
        ref = F.conv3d(x.double(), w.double(), None, 1, 1)
        assert rel_err(z, ref_t) < 1e-5, rel_err(z, ref_t)
        assert rel_err(z, ref) < 2e-3, rel_err(z, ref)
        dz = torch.randn_like(z)
        z.backward(dz)
        x2, w2 = x.double().requires_grad_(True), w.double().requires_grad_(True)
        F.conv3d(x2, w2, None, 1, 1).backward(dz.double())
        assert rel_err(xr.grad, x2.grad) < 3e-3 and rel_err(wr.grad, w2.grad) < 3e-3
        if Cout % 64 == 0:      # conv + bias + PixelShuffle(4) in one kernel, backward through the inverse shuffle + library conv
            from veloxseg_b200.nn import PixelShuffle
            bias = torch.randn(Cout, generator=g).to(DEV)
            xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
            y = ops.dense_conv3(xr, wr, br, 4)
            yref = PixelShuffle(4, 3)(z.detach() + bias.view(1, -1, 1, 1, 1))
            assert y.shape == yref.shape and torch.equal(y, yref)
            dy = torch.randn_like(y)
            y.backward(dy)
            x2, w2, b2 = x.double().requires_grad_(True), w.double().requires_grad_(True), bias.double().requires_grad_(True)
            PixelShuffle(4, 3)(F.conv3d(x2, w2, b2, 1, 1)).backward(dy.double())
            assert rel_err(xr.grad, x2.grad) < 3e-3 and rel_err(wr.grad, w2.grad) < 3e-3 and rel_err(br.grad, b2.grad) < 1e-4
    finally:
        ops.dense_conv_tc_enable(False)
