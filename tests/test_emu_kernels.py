"""Developer check (opt-in, VX_EMU=1): the kernel sources compiled for the CPU shim in tools/emu, compared with the
oracle.  Catches indexing mistakes without a GPU.  The real parity tests are the -m gpu ones."""
import os

import pytest
import torch

from tests._util import close, jlc_param_dict, jlc_params, rel_err

pytestmark = pytest.mark.emu


@pytest.fixture(scope="module")
def emu():
    from veloxseg_b200._lib import VxLib
    from veloxseg_b200.csrc.build import build_emu
    return VxLib(build_emu())


def _oracle():
    from oracle import veloxseg_oracle as O
    return O


@pytest.mark.parametrize("C,groups,e,shape,B", [(8, 2, 3, (5, 6, 8), 2), (16, 2, 2, (3, 4, 7), 1), (32, 2, 2, (3, 3, 3), 2),
                                                  (8, 2, 3, (9, 7, 12), 1), (16, 2, 2, (9, 8, 12), 1), (8, 2, 2, (9, 9, 7), 1),
                                                  (8, 2, 2, (12, 14, 14), 1),       # two combine chunks per row (S > 2048)
                                                  (16, 2, 2, (7, 12, 13), 1)])      # S = 1092, C = 16: the fused tcgen05 FFN (ragged last tile)
def test_jlc(emu, C, groups, e, shape, B):
    from veloxseg_b200 import ops
    O = _oracle()
    torch.manual_seed(1)
    x = torch.randn(B, C, *shape)
    params = jlc_params(C, groups, e, seed=3)
    y, z, o, hpre, stats = ops.jlc_fwd_raw(emu, 0, x, params, groups, e)
    xr = x.clone().requires_grad_(True)
    pr = [p.clone().requires_grad_(True) for p in params]
    yr = O.jlc(xr, jlc_param_dict(pr), "", groups)
    assert rel_err(y, yr) < 2e-5, rel_err(y, yr)
    dy = torch.randn_like(y)
    grads = torch.autograd.grad(yr, [xr] + pr, dy)
    got = ops.jlc_bwd_raw(emu, 0, dy, x, z, o, hpre, stats, params, groups, e)
    # the conv-bias gradients are structurally zero (the bias cancels inside the InstanceNorm): pure accumulation noise, which
    # grows with the voxel count -- the absolute tolerance admits it
    atol = 2e-5 if x[0, 0].numel() <= 1024 else 6e-5
    for i, (g, r) in enumerate(zip(got, grads)):
        assert close(g, r, rtol=2e-4, atol=atol), (i, rel_err(g, r), float(r.norm()))


def test_jlc_dropout_consistency(emu):
    """train-mode dropout: backward must use the mask forward used (finite-difference-free check: linearity in dy)."""
    from veloxseg_b200 import ops
    torch.manual_seed(2)
    C, groups, e = 8, 2, 2
    x = torch.randn(1, C, 4, 4, 8)
    params = jlc_params(C, groups, e, seed=5)
    y0, z, o, hpre, stats = ops.jlc_fwd_raw(emu, 0, x, params, groups, e, 0.0, False, 0)
    y1, _, o1, *_ = ops.jlc_fwd_raw(emu, 0, x, params, groups, e, 0.5, True, 1234)
    y2, *_ = ops.jlc_fwd_raw(emu, 0, x, params, groups, e, 0.5, True, 1234)
    assert torch.allclose(y1, y2, rtol=1e-5, atol=1e-6)
    # y = o + mask*2*(W2 h + b2): elements are either o (dropped) or o + 2*(y0 - o)
    d0, d1 = (y0 - o), (y1 - o1)
    dropped = d1 == 0
    frac = dropped.float().mean().item()
    assert 0.35 < frac < 0.65, frac
    assert torch.allclose(d1[~dropped], 2 * d0[~dropped], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("chs,Co,S,B,addend", [((16, 16), 16, (5, 6, 7), 2, True), ((8,), 24, (4, 4, 4), 1, False),
                                                ((32, 16, 8), 20, (3, 5, 9), 2, True)])
def test_mixer(emu, chs, Co, S, B, addend):
    from veloxseg_b200 import ops
    O = _oracle()
    torch.manual_seed(0)
    streams = [torch.randn(B, c, *S) for c in chs]
    W = torch.randn(Co, sum(chs)) * 0.2
    b = torch.randn(Co) * 0.1
    add = torch.randn(B, Co, *S) if addend else None
    y, t, stats = ops.mixer_fwd_raw(emu, 0, streams, W, b, add)
    sr = [s.clone().requires_grad_(True) for s in streams]
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.modal_mixer(sr, Wr, br, add)
    assert rel_err(y, yr) < 1e-5
    dy = torch.randn_like(y)
    grads = torch.autograd.grad(yr, sr + [Wr, br], dy)
    got = ops.mixer_bwd_raw(emu, 0, dy, streams, W, t, stats)
    for i, (g, r) in enumerate(zip(got, grads)):
        assert close(g, r, rtol=2e-4, atol=2e-5), (i, rel_err(g, r))


def test_inorm_gram_sdkt_lnpw(emu):
    from veloxseg_b200 import ops
    O = _oracle()
    torch.manual_seed(0)
    x = torch.randn(2, 6, 5, 6, 7) * 2 + 0.5
    add = torch.randn_like(x)
    y, stats = ops.inorm_fwd_raw(emu, 0, x, add)
    xr = x.clone().requires_grad_(True)
    yr = O.instance_norm(xr) + add
    assert rel_err(y, yr) < 1e-5
    dy = torch.randn_like(y)
    assert close(ops.inorm_bwd_raw(emu, 0, dy, x, stats), torch.autograd.grad(yr, xr, dy)[0], rtol=1e-4)

    f = torch.randn(2, 16, 7, 9, 11)
    G = ops.gram_fwd_raw(emu, 0, f)
    fr = f.clone().requires_grad_(True)
    Gr = O.gram(fr)
    assert rel_err(G, Gr) < 1e-5
    dG = torch.randn_like(G)
    assert close(ops.gram_bwd_raw(emu, 0, dG, f), torch.autograd.grad(Gr, fr, dG)[0], rtol=1e-4)
    f12 = torch.randn(1, 12, 4, 4, 4)
    assert rel_err(ops.gram_fwd_raw(emu, 0, f12), O.gram(f12)) < 1e-5

    gs = torch.randn(2, 16, 16).requires_grad_(True)
    gts = [torch.randn(2, 16, 16).requires_grad_(True) for _ in range(2)]
    L = ops.sdkt_loss_fwd_raw(emu, 0, gs.detach(), [g.detach() for g in gts])
    Lr = O.sdkt_loss(gs, gts)
    assert abs(float(L) - float(Lr)) < 1e-5 * abs(float(Lr))
    gr = torch.autograd.grad(Lr, [gs] + gts, torch.tensor(1.7))
    got = ops.sdkt_loss_bwd_raw(emu, 0, torch.tensor(1.7), gs.detach(), [g.detach() for g in gts])
    for g, r in zip(got, gr):
        assert close(g, r, rtol=1e-5)

    x = torch.randn(2, 24, 3, 4, 5)
    lw, lb, W = torch.randn(24) * 0.3 + 1, torch.randn(24) * 0.2, torch.randn(10, 24) * 0.2
    y, xhat, rstd = ops.lnpw_fwd_raw(emu, 0, x, lw, lb, W)
    xr, lwr, lbr, Wr = [t.clone().requires_grad_(True) for t in (x, lw, lb, W)]
    yr = O.pointwise(O.layer_norm_cf(xr, lwr, lbr), Wr, None)
    assert rel_err(y, yr) < 1e-5
    dy = torch.randn_like(y)
    gr = torch.autograd.grad(yr, [xr, lwr, lbr, Wr], dy)
    got = ops.lnpw_bwd_raw(emu, 0, dy, xhat, rstd, lw, lb, W)
    for i, (g, r) in enumerate(zip(got, gr)):
        assert close(g, r, rtol=2e-4, atol=2e-5), (i, rel_err(g, r))


PWA_CASES = [
    # size, C, min_big, min_small, heads, min_dim_head, M, e, B
    ((6, 6, 6), 8, [3, 3, 3], [1, 1, 1], 1, 4, 2, 2, 1),
    ((8, 8, 4), 16, [4, 4, 2], [1, 1, 1], 2, 4, 2, 2, 1),
    ((12, 12, 12), 16, [3, 3, 3], [1, 1, 1], 1, 4, 1, 1, 1),
    ((8, 8, 8), 8, [4, 4, 4], [1, 1, 1], 1, 4, 2, 2, 1),      # l = 64: the vectorised bias-gradient path of level 2
    ((8, 8, 12), 8, [4, 4, 6], [1, 1, 1], 1, 8, 2, 2, 1),     # L = 192, 8 channels per head: the tcgen05 attention forward
    ((8, 8, 16), 16, [4, 4, 8], [1, 1, 1], 1, 4, 2, 2, 1),    # S = 1024, C = 16: the fused tcgen05 FFN, two modalities per launch
]


@pytest.mark.parametrize("size,C,mb,ms,heads,mdh,M,e,B", PWA_CASES)
def test_pwa_gather_bit_exact(emu, size, C, mb, ms, heads, mdh, M, e, B):
    from veloxseg_b200 import ops
    O = _oracle()
    geo = O.pwa_geometry(size, C, mb, ms, 2, heads, mdh)
    torch.manual_seed(0)
    x = torch.randn(B, geo["cv"], *size)
    tok, arg = ops.pwa_gather_raw(emu, 0, x, geo)
    ref, _, _ = O.gather_tokens(x, heads, geo["bws"], geo["sws"])
    assert torch.equal(tok, ref)
    # arg-max voxel indices must point at the values that were gathered
    flat = x.reshape(B, geo["cv"], -1)
    cper = geo["cv"] // (heads * len(geo["bws"]))
    Ns_off = 0
    for j, bw in enumerate(geo["bws"]):
        Nj = (size[0] // bw[0]) * (size[1] // bw[1]) * (size[2] // bw[2])
        for h in range(heads):
            for c in range(cper):
                ch = (j * heads + h) * cper + c
                a = arg[:, h, Ns_off:Ns_off + Nj, :, c].reshape(B, -1).long()
                v = torch.gather(flat[:, ch], 1, a)
                assert torch.equal(v, tok[:, h, Ns_off:Ns_off + Nj, :, c].reshape(B, -1))
        Ns_off += Nj


@pytest.mark.parametrize("size,C,mb,ms,heads,mdh,M,e,B", PWA_CASES)
def test_pwa_block(emu, size, C, mb, ms, heads, mdh, M, e, B):
    from tests._util import pwa_params
    from veloxseg_b200 import ops
    O = _oracle()
    geo = O.pwa_geometry(size, C, mb, ms, 2, heads, mdh)
    torch.manual_seed(1)
    xs = [torch.randn(B, C, *size) for _ in range(M)]
    flat, pd, table, index = pwa_params(M, C, geo, e, seed=2)
    zs, saved = ops.pwa_block_fwd_raw(emu, 0, xs, flat, table, index, geo, e)
    xr = [x.clone().requires_grad_(True) for x in xs]
    pr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in pd.items()}
    zr = O.pwa_block(xr, pr, "", geo)
    for m in range(M):
        assert rel_err(zs[m], zr[m]) < 2e-5, (m, rel_err(zs[m], zr[m]))
    dzs = [torch.randn_like(z) for z in zs]
    from tests._util import PWA_PARAM_NAMES
    plist = [pr[n.format(m=m)] for m in range(M) for n in PWA_PARAM_NAMES]
    tab = pr["attn.position_embedding.relative_position_bias_table"]
    grads = torch.autograd.grad(zr, xr + plist + [tab], dzs)
    dxs, dps, dtable = ops.pwa_block_bwd_raw(emu, 0, dzs, xs, flat, table, index, saved, geo, e)
    got = list(dxs) + list(dps) + [dtable]
    names = [f"dx{m}" for m in range(M)] + [n.format(m=m) for m in range(M) for n in PWA_PARAM_NAMES] + ["table"]
    # with a single modality the key bias shifts every score of a row equally: softmax-invariant, true gradient 0 -- what the
    # kernels return there is round-off of the chain above it (same rule as tests/test_gpu_ops.py::test_pwa_block_levels)
    zero = {"attn.qkv_proj.0.1.bias"} if M == 1 else set()
    bad = [(n, rel_err(g, r)) for n, g, r in zip(names, got, grads) if n not in zero and not close(g, r, rtol=3e-4, atol=2e-5)]
    assert not bad, bad
    for n, g in zip(names, got):
        if n in zero:
            assert float(g.abs().max()) <= 1e-3 * float(grads[names.index("attn.qkv_proj.0.1.weight")].abs().max()), n


@pytest.mark.parametrize("src,dst", [((3, 3, 3), (12, 12, 12)), ((2, 5, 4), (8, 10, 8)), ((6, 6, 6), (12, 12, 12)), ((1, 3, 2), (4, 6, 8)),
                                     ((3, 4, 5), (32, 32, 32))])      # last: long rows -> the row-staged adjoint pass
def test_resize_trilinear(emu, src, dst):
    import torch.nn.functional as F
    from veloxseg_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 3, *src)
    y = ops.resize_fwd_raw(emu, 0, x, dst)
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, size=dst, mode="trilinear", align_corners=True)
    assert rel_err(y, yr) < 1e-6
    dy = torch.randn_like(yr)
    assert close(ops.resize_bwd_raw(emu, 0, dy, src), torch.autograd.grad(yr, xr, dy)[0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("C,n_out,B,S", [(2, 4, 2, (5, 6, 7)), (4, 3, 1, (4, 9, 5)), (3, 1, 3, (3, 3, 3))])
def test_segloss(emu, C, n_out, B, S):
    """CE + Dice over several deep outputs in one pass vs torch cross_entropy + the restated MONAI DiceLoss."""
    import torch.nn.functional as F
    from veloxseg_b200 import ops
    from veloxseg_b200.loss import dice_loss
    torch.manual_seed(4)
    logits = [torch.randn(B, C, *S) * 2 for _ in range(n_out)]
    labels = torch.randint(0, C, (B, 1, *S))
    w = [0.4, 0.3, 0.2, 0.1][:n_out]
    loss, sums = ops.segloss_fwd_raw(emu, 0, logits, labels, w)
    lr = [t.clone().requires_grad_(True) for t in logits]
    ref = sum(wi * (F.cross_entropy(o, labels.squeeze(1)) + dice_loss(o, labels)) for wi, o in zip(w, lr))
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref)), (float(loss), float(ref))
    up = torch.tensor(0.7)
    grads = torch.autograd.grad(ref, lr, up)
    got = ops.segloss_bwd_raw(emu, 0, up, logits, labels, sums, w)
    for g, r in zip(got, grads):
        assert close(g, r, rtol=2e-4, atol=1e-8), rel_err(g, r)


def test_wgrad_tensor_core_layout(emu):
    """pw_wgrad_tc.cu with the MMA replaced by its software model: staging layout (K-major core matrices, hi/lo split),
    chunking over (batch, voxels), the all-ones bias row and the TMEM read-back, through the ops that use it."""
    emu.set_option(9, 8)          # VX_OPT_WGRAD_TC_MIN_S: route every S % 4 == 0 problem to the tensor-core kernel
    try:
        test_jlc(emu, 8, 2, 3, (5, 6, 8), 2)
        test_jlc(emu, 16, 2, 2, (4, 4, 7), 1)
        test_mixer(emu, (16, 16), 16, (4, 6, 7), 2, True)
        test_mixer(emu, (32, 16, 8), 20, (3, 5, 8), 2, True)
        test_pwa_block(emu, *PWA_CASES[1])
    finally:
        emu.set_option(9, 512)


def test_inorm_register_cached(emu):
    """Rows of >= 1024 voxels (S % 4 == 0) take the register-cached InstanceNorm kernels."""
    from veloxseg_b200 import ops
    O = _oracle()
    torch.manual_seed(0)
    x = torch.randn(2, 3, 8, 16, 12) * 2 + 0.5
    add = torch.randn_like(x)
    y, stats = ops.inorm_fwd_raw(emu, 0, x, add)
    xr = x.clone().requires_grad_(True)
    yr = O.instance_norm(xr) + add
    assert rel_err(y, yr) < 1e-5
    dy = torch.randn_like(y)
    gref = torch.autograd.grad(yr, xr, dy)[0]
    assert close(ops.inorm_bwd_raw(emu, 0, dy, x, stats), gref, rtol=1e-4)
    # bias-gradient output: per-channel sums of dx (analytically zero; must match the sum of what the kernel wrote)
    dx, db = ops.inorm_bwd_raw(emu, 0, dy, x, stats, with_bias_grad=True)
    assert close(dx, gref, rtol=1e-4)
    assert torch.allclose(db, dx.sum(dim=(0, 2, 3, 4)), atol=1e-4)
    xs = torch.randn(2, 5, 3, 4, 5)
    ys, sts = ops.inorm_fwd_raw(emu, 0, xs, None)
    dys = torch.randn_like(ys)
    dxs, dbs = ops.inorm_bwd_raw(emu, 0, dys, xs, sts, with_bias_grad=True)
    assert torch.allclose(dbs, dxs.sum(dim=(0, 2, 3, 4)), atol=1e-4)


def test_lnpw_wide_channels(emu):
    """PatchMerging shapes of levels 3-4 (few voxels, >= 64 channels) take the channel-split LayerNorm kernel."""
    from veloxseg_b200 import ops
    O = _oracle()
    torch.manual_seed(3)
    x = torch.randn(2, 72, 3, 4, 5)
    lw, lb, W = torch.randn(72) * 0.3 + 1, torch.randn(72) * 0.2, torch.randn(10, 72) * 0.2
    y, xhat, rstd = ops.lnpw_fwd_raw(emu, 0, x, lw, lb, W)
    xr, lwr, lbr, Wr = [t.clone().requires_grad_(True) for t in (x, lw, lb, W)]
    yr = O.pointwise(O.layer_norm_cf(xr, lwr, lbr), Wr, None)
    assert rel_err(y, yr) < 1e-5
    dy = torch.randn_like(y)
    gr = torch.autograd.grad(yr, [xr, lwr, lbr, Wr], dy)
    got = ops.lnpw_bwd_raw(emu, 0, dy, xhat, rstd, lw, lb, W)
    for i, (g, r) in enumerate(zip(got, gr)):
        assert close(g, r, rtol=2e-4, atol=2e-5), (i, rel_err(g, r))


@pytest.mark.parametrize("case", [0, 3])
def test_pwa_dropout_backward_matches_forward_masks(emu, case):
    """Train-mode dropout (attention weights + projections): with a fixed seed the block is a deterministic smooth
    function, so the backward pass (which regenerates every mask) must agree with a central finite difference of the
    forward pass.  A mask mismatch between the two directions shows up as an O(1) relative error."""
    from tests._util import pwa_params
    from veloxseg_b200 import ops
    O = _oracle()
    size, C, mb, ms, heads, mdh, M, e, B = PWA_CASES[case]
    geo = O.pwa_geometry(size, C, mb, ms, 2, heads, mdh)
    torch.manual_seed(11)
    xs = [torch.randn(B, C, *size) for _ in range(M)]
    flat, pd, table, index = pwa_params(M, C, geo, e, seed=4)
    p_att, p_proj, seed = 0.3, 0.2, 77

    def fwd(inp):
        zs, saved = ops.pwa_block_fwd_raw(emu, 0, inp, flat, table, index, geo, e, p_att, p_proj, True, seed)
        return zs, saved

    zs, saved = fwd(xs)
    zs2, _ = fwd(xs)
    assert all(torch.equal(a, b) for a, b in zip(zs, zs2))
    dzs = [torch.randn_like(z) for z in zs]
    dxs, dps, dtable = ops.pwa_block_bwd_raw(emu, 0, dzs, xs, flat, table, index, saved, geo, e, p_att, p_proj, True, seed)
    d = [torch.randn_like(x) for x in xs]
    an = sum(float((gx.double() * dd.double()).sum()) for gx, dd in zip(dxs, d))
    # Max-pool kinks and curvature make the finite difference itself O(eps)-inexact and the size of that error depends on the
    # mask realisation (measured on a B200, tools/dbg_pwa_drop.py: 0.2-1.3 % at 1e-3, 0.04-0.7 % at 3e-4, 4 % on an unlucky
    # draw), whereas a mask mismatch is O(1) at every step size: the best of two step sizes must be inside the tolerance.
    gaps = []
    for eps in (1e-3, 3e-4):
        zp, _ = fwd([x + eps * dd for x, dd in zip(xs, d)])
        zm, _ = fwd([x - eps * dd for x, dd in zip(xs, d)])
        fd = sum(float(((a - b).double() / (2 * eps) * g.double()).sum()) for a, b, g in zip(zp, zm, dzs))
        gaps.append(abs(fd - an) / max(abs(fd), abs(an), 1.0))
    assert min(gaps) <= 3e-2, (gaps, an)



@pytest.mark.parametrize("Ct,c_off,Ci,Co,p,size,B", [(2, 1, 1, 16, 4, (8, 8, 12), 2), (4, 0, 4, 16, 4, (8, 4, 8), 1), (3, 1, 2, 20, 2, (4, 6, 6), 2)])
def test_patch_embed(emu, Ct, c_off, Ci, Co, p, size, B):
    """Network-input stem vs torch conv3d (k = s = patch) on the channel slice; weight / bias gradients."""
    import torch.nn.functional as F
    from veloxseg_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(B, Ct, *size)
    w = (torch.randn(Co, Ci, p, p, p) * 0.2).requires_grad_(True)
    b = (torch.randn(Co) * 0.1).requires_grad_(True)
    y = ops.patch_embed_fwd_raw(emu, 0, x, c_off, w.detach(), b.detach())
    yr = F.conv3d(x[:, c_off:c_off + Ci], w, b, stride=p)
    assert rel_err(y, yr) < 1e-5
    dy = torch.randn_like(yr)
    gw, gb = torch.autograd.grad(yr, [w, b], dy)
    dw, db = ops.patch_embed_bwd_raw(emu, 0, dy, x, c_off, w.shape)
    assert close(dw, gw, rtol=1e-4, atol=1e-5) and close(db, gb, rtol=1e-4, atol=1e-5)


def test_adamw_matches_torch(emu):
    """vx_adamw_step vs torch.optim.AdamW over a few tensors and steps (one without a gradient)."""
    import ctypes as C
    from veloxseg_b200._lib import AdamwDesc
    torch.manual_seed(0)
    shapes = [(3, 5), (1500,), (7,), (2, 2, 2)]
    ps = [torch.randn(*s) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    offs, chunks, total = [], [], 0
    for i, p in enumerate(ps):
        offs.append(total)
        chunks += [(i, s) for s in range(0, p.numel(), 1024)]
        total += p.numel()
    m, v, step = torch.zeros(total), torch.zeros(total), torch.zeros(1)
    ch = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2)
    for it in range(3):
        grads = [torch.randn_like(p) if not (i == 2 and it == 1) else None for i, p in enumerate(ps)]
        for r, g in zip(ref, grads):
            r.grad = None if g is None else g.clone()
        opt.step()
        tab = torch.tensor([[p.data_ptr(), g.data_ptr() if g is not None else 0, o, p.numel()] for p, g, o in zip(ps, grads, offs)],
                           dtype=torch.int64)
        d = AdamwDesc(len(chunks), 1e-2, 0.9, 0.999, 1e-8, 0.05)
        emu.call("vx_adamw_step", d, [tab, ch], [m, v, step], 0)
        # torch skips a tensor without gradient entirely (its per-tensor step does not advance); the one-launch kernel
        # shares one step counter, so compare only tensors that had a gradient in every step so far
        for i, (p, r) in enumerate(zip(ps, ref)):
            if i != 2:
                assert close(p, r.detach(), rtol=1e-5, atol=1e-6), (it, i, rel_err(p, r.detach()))
    assert float(step) == 3.0


def test_lnpw_backward_multi_group_ctas(emu):
    """Large B * S: the LayerNorm backward packs several 32-voxel groups per CTA and keeps dgamma / dbeta in registers."""
    from veloxseg_b200 import ops
    O = _oracle()
    torch.manual_seed(5)
    x = torch.randn(8, 8, 16, 16, 16)
    lw, lb, W = torch.randn(8) * 0.3 + 1, torch.randn(8) * 0.2, torch.randn(6, 8) * 0.2
    y, xhat, rstd = ops.lnpw_fwd_raw(emu, 0, x, lw, lb, W)
    xr, lwr, lbr, Wr = [t.clone().requires_grad_(True) for t in (x, lw, lb, W)]
    yr = O.pointwise(O.layer_norm_cf(xr, lwr, lbr), Wr, None)
    dy = torch.randn_like(y)
    gr = torch.autograd.grad(yr, [xr, lwr, lbr, Wr], dy)
    got = ops.lnpw_bwd_raw(emu, 0, dy, xhat, rstd, lw, lb, W)
    for i, (g, r) in enumerate(zip(got, gr)):
        assert close(g, r, rtol=3e-4, atol=2e-5), (i, rel_err(g, r))


@pytest.mark.parametrize("B,C,s,size", [(2, 2, 4, (3, 4, 5)), (1, 3, 2, (4, 4, 6)), (2, 1, 4, (2, 3, 4))])
def test_pixel_shuffle_bias(emu, B, C, s, size):
    """bias + 3-D pixel shuffle vs the reference permutation (superpixel.py:15) and its autograd."""
    from veloxseg_b200 import ops
    from veloxseg_b200.nn import PixelShuffle
    torch.manual_seed(0)
    z = torch.randn(B, C * s ** 3, *size, requires_grad=True)
    bias = torch.randn(C * s ** 3, requires_grad=True)
    yr = PixelShuffle(s)(z + bias[None, :, None, None, None])
    y = ops.pixel_shuffle_fwd_raw(emu, 0, z.detach(), bias.detach(), s)
    assert torch.equal(y, yr.detach())
    dy = torch.randn_like(yr)
    gz, gb = torch.autograd.grad(yr, [z, bias], dy)
    dz, db = ops.pixel_shuffle_bwd_raw(emu, 0, dy, s, True)
    assert torch.equal(dz, gz) and close(db, gb, rtol=1e-5, atol=1e-5)


def _conv_ref(x, w, b, kernel, stride, pad, transposed, shuffle):
    import torch.nn.functional as F
    from veloxseg_b200.nn import PixelShuffle
    if transposed:
        y = F.conv_transpose3d(x, w, b, stride=stride)
    else:
        y = F.conv3d(x, w, b, stride=stride, padding=pad)
    return PixelShuffle(shuffle, 3)(y) if shuffle else y


@pytest.mark.parametrize("Cout,shape,B,shuffle", [(64, (3, 5, 8), 1, 4), (128, (4, 9, 8), 1, 4), (32, (2, 3, 12), 2, 0), (64, (5, 4, 8), 1, 0)])
def test_conv3_tensor_core(emu, Cout, shape, B, shuffle):
    """conv3_tc.cu with the tcgen05 / mbarrier / bulk-copy calls replaced by the descriptor-level software model of
    vx_tc2.cuh: brick staging, tap -> shifted-descriptor arithmetic, weight images and their ring, TMEM column bookkeeping,
    fused bias + PixelShuffle store; the im2col weight gradient (inverse shuffle as register transpose, partial tiles, fixed-
    order fold, bias column) and the plane-wise col2im data gradient -- against conv3d in fp64."""
    from veloxseg_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, 16, *shape, generator=g)
    w = torch.randn(Cout, 16, 3, 3, 3, generator=g) * 0.05
    b = torch.randn(Cout, generator=g)
    y = ops.conv_fwd_raw(emu, 0, x, w, b, 3, 1, 1, False, shuffle)
    x2, w2, b2 = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = _conv_ref(x2, w2, b2, 3, 1, 1, False, shuffle)
    assert y.shape == yr.shape and rel_err(y, yr) < 5e-6, rel_err(y, yr)
    dy = torch.randn(y.shape, generator=g)
    yr.backward(dy.double())
    dx, dw, db = ops.conv_bwd_raw(emu, 0, dy, x, w, 3, 1, 1, False, shuffle)
    assert rel_err(dw, w2.grad) < 5e-6, rel_err(dw, w2.grad)
    assert rel_err(db, b2.grad) < 5e-6, rel_err(db, b2.grad)
    assert rel_err(dx, x2.grad) < 5e-6, rel_err(dx, x2.grad)


@pytest.mark.parametrize("Ci,Co,k,s,p,tr,shape,B", [(2, 16, 7, 4, 3, False, (8, 12, 16), 2), (8, 16, 3, 2, 1, False, (6, 4, 8), 1),
                                                    (16, 8, 2, 2, 0, True, (3, 2, 5), 2), (12, 20, 3, 2, 1, False, (5, 7, 3), 1),
                                                    (20, 12, 2, 2, 0, True, (1, 3, 3), 1), (24, 3, 1, 1, 0, False, (3, 4, 5), 2),
                                                    (3, 8, 4, 4, 0, False, (8, 8, 12), 1), (40, 24, 3, 2, 1, False, (4, 4, 4), 1),
                                                    (8, 12, 3, 1, 1, False, (4, 5, 6), 2)])
def test_conv_strided_transposed(emu, Ci, Co, k, s, p, tr, shape, B):
    """conv_simt.cu: DownConv / UpConv convolutions (strided, scatter and weight-gradient kernels) against torch in fp64."""
    from veloxseg_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, Ci, *shape, generator=g)
    w = torch.randn(*((Ci, Co) if tr else (Co, Ci)), k, k, k, generator=g) * 0.1
    b = torch.randn(Co, generator=g)
    y = ops.conv_fwd_raw(emu, 0, x, w, b, k, s, p, tr)
    x2, w2, b2 = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = _conv_ref(x2, w2, b2, k, s, p, tr, 0)
    assert y.shape == yr.shape and rel_err(y, yr) < 2e-6, rel_err(y, yr)
    dy = torch.randn(y.shape, generator=g)
    yr.backward(dy.double())
    dx, dw, db = ops.conv_bwd_raw(emu, 0, dy, x, w, k, s, p, tr)
    assert rel_err(dx, x2.grad) < 2e-6 and rel_err(dw, w2.grad) < 2e-6 and rel_err(db, b2.grad) < 2e-6
    _, dw1, _ = ops.conv_bwd_raw(emu, 0, dy, x, w, k, s, p, tr, need_dx=False, need_db=False)
    assert rel_err(dw1, w2.grad) < 2e-6
    if tr:      # without a bias the transposed convolution runs as a channel contraction + depth-to-space pass
        y0 = ops.conv_fwd_raw(emu, 0, x, w, None, k, s, p, tr)
        assert rel_err(y0, _conv_ref(x.double(), w.double(), None, k, s, p, tr, 0)) < 2e-6
