"""Parity of the sm_100a ops (through torch.ops.veloxseg -> C ABI) with (a) vectors produced by the unmodified
reference (tests/golden/ops_small.pt) and (b) the CPU oracle at the real per-level shapes of the three configs.
fp32 tolerance (north_star): 1e-3 relative on outputs and per-parameter gradients; integer gather bit-exact."""
import pytest
import torch

from oracle import veloxseg_oracle as O
from tests import _golden as G
from tests._util import PWA_PARAM_NAMES, close, jlc_param_dict, jlc_params, pwa_params, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from veloxseg_b200 import ops as _ops
    from veloxseg_b200 import _lib
    assert _lib.get_lib().c.vx_version() >= 100
    return _ops


FX = G.load("ops_small.pt")
JLC_NAMES = ["spatial_convs.0.0.weight", "spatial_convs.0.0.bias", "spatial_convs.1.0.weight", "spatial_convs.1.0.bias",
             "spatial_convs.2.0.weight", "spatial_convs.2.0.bias", "channel_conv.1.weight", "channel_conv.1.bias",
             "channel_conv.3.weight", "channel_conv.3.bias"]


def _cots(outs, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(o.shape, generator=g) for o in outs]


def _leaf(t):
    return t.detach().to(DEV).clone().requires_grad_(True)


@pytest.mark.parametrize("tag", ["jlc_c8", "jlc_c16", "jlc_c32"])
def test_jlc_golden(ops, tag):
    fx = FX[tag]
    x = _leaf(fx["inputs"][0])
    ps = [_leaf(fx["state"][n]) for n in JLC_NAMES]
    call = [p.reshape(p.shape[0], -1) if n.startswith("channel_conv") and n.endswith("weight") else p for n, p in zip(JLC_NAMES, ps)]
    y = ops.jlc(x, call, fx["cfg"]["groups"], fx["cfg"]["e"])
    assert rel_err(y, fx["outputs"][0]) < 1e-4
    grads = torch.autograd.grad(y, [x] + ps, _cots([y], fx["cot_seed"])[0].to(DEV))
    assert close(grads[0], fx["input_grads"][0], rtol=1e-3, atol=1e-6)
    for n, g in zip(JLC_NAMES, grads[1:]):
        assert close(g, fx["param_grads"][n], rtol=1e-3, atol=2e-5), (n, rel_err(g, fx["param_grads"][n]))


@pytest.mark.parametrize("tag", ["mixer_2x16", "mixer_1x8"])
def test_mixer_golden(ops, tag):
    fx = FX[tag]
    ins = [_leaf(t) for t in fx["inputs"]]
    W, b = _leaf(fx["state"]["0.weight"]), _leaf(fx["state"]["0.bias"])
    y = ops.modal_mixer(ins[1:], W, b, ins[0])
    assert rel_err(y, fx["outputs"][0]) < 1e-4
    grads = torch.autograd.grad(y, ins + [W, b], _cots([y], fx["cot_seed"])[0].to(DEV))
    for g, r in zip(grads[:len(ins)], fx["input_grads"]):
        assert close(g, r, rtol=1e-3, atol=1e-6)
    assert close(grads[-2], fx["param_grads"]["0.weight"], rtol=1e-3, atol=2e-5)
    assert close(grads[-1], fx["param_grads"]["0.bias"], rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize("tag", ["pwa_6c8", "pwa_884", "pwa_12m1"])
def test_pwa_block_golden(ops, tag):
    fx = FX[tag]
    c, geo = fx["cfg"], fx["geo"]
    M = c["M"]
    xs = [_leaf(t) for t in fx["inputs"]]
    st = fx["state"]
    names = [n.format(m=m) for m in range(M) for n in PWA_PARAM_NAMES]
    ps = [_leaf(st[n]) for n in names]
    call = [p.reshape(p.shape[0], -1) if p.dim() == 5 else p for p in ps]
    table = _leaf(st["attn.position_embedding.relative_position_bias_table"])
    index = st["attn.position_embedding.relative_position_index"].to(DEV)
    zs = ops.pwa_block(xs, call, table, index, geo, c["e"])
    for z, r in zip(zs, fx["outputs"]):
        assert rel_err(z, r) < 1e-4
    cots = [t.to(DEV) for t in _cots(zs, fx["cot_seed"])]
    grads = torch.autograd.grad(zs, xs + ps + [table], cots)
    for g, r in zip(grads[:M], fx["input_grads"]):
        assert close(g, r, rtol=1e-3, atol=1e-6), rel_err(g, r)
    for n, g in zip(names + ["attn.position_embedding.relative_position_bias_table"], grads[M:]):
        assert close(g, fx["param_grads"][n], rtol=1e-3, atol=2e-5), (n, rel_err(g, fx["param_grads"][n]))


def test_patch_merging_norms_gram_golden(ops):
    from veloxseg_b200 import nn as vnn
    fx = FX["patch_merging"]
    pm = vnn.PatchMerging(8).to(DEV)
    pm.load_state_dict(fx["state"])
    x = _leaf(fx["inputs"][0])
    y = pm(x)
    assert rel_err(y, fx["outputs"][0]) < 1e-4
    grads = torch.autograd.grad(y, [x] + list(pm.parameters()), _cots([y], fx["cot_seed"])[0].to(DEV))
    assert close(grads[0], fx["input_grads"][0], rtol=1e-3, atol=1e-6)
    for (n, _), g in zip(pm.named_parameters(), grads[1:]):
        assert close(g, fx["param_grads"][n], rtol=1e-3, atol=2e-5), n
    for tag, mod in (("down_conv", vnn.DownConv(3, 8, patch_size=2)), ("up_conv", vnn.UpConv(8, 4))):
        fx = FX[tag]
        mod = mod.to(DEV)
        mod.load_state_dict(fx["state"])
        x = _leaf(fx["inputs"][0])
        y = mod(x)
        assert rel_err(y, fx["outputs"][0]) < 1e-4, tag
        g, = torch.autograd.grad(y, x, _cots([y], fx["cot_seed"])[0].to(DEV))
        assert close(g, fx["input_grads"][0], rtol=1e-3, atol=1e-6), tag
    fx = FX["gram"]
    x = _leaf(fx["inputs"][0])
    Gm = vnn.get_pram_matrix(x)
    assert rel_err(Gm, fx["outputs"][0]) < 1e-5
    g, = torch.autograd.grad(Gm, x, _cots([Gm], fx["cot_seed"])[0].to(DEV))
    assert close(g, fx["input_grads"][0], rtol=1e-3, atol=1e-8)


# ---------------------------------------------------------------------------------------------------------------
# real per-level shapes against the oracle
# ---------------------------------------------------------------------------------------------------------------
LEVELS = {  # name: (spatial, C, groups, e)
    "autopet_L1": ((24, 24, 24), 16, 4, 3), "autopet_L2": ((12, 12, 12), 32, 4, 3), "autopet_L3": ((6, 6, 6), 64, 8, 2),
    "autopet_L4": ((3, 3, 3), 128, 8, 2), "hecktor_L1": ((32, 32, 16), 16, 4, 3), "hecktor_L4": ((4, 4, 2), 128, 8, 2)}


@pytest.mark.parametrize("lvl", list(LEVELS))
@pytest.mark.parametrize("B", [1, 3])
def test_jlc_levels(ops, lvl, B):
    shape, C, groups, e = LEVELS[lvl]
    torch.manual_seed(5)
    x = torch.randn(B, C, *shape)
    params = jlc_params(C, groups, e, seed=7)
    xr = x.clone().requires_grad_(True)
    pr = [p.clone().requires_grad_(True) for p in params]
    yr = O.jlc(xr, jlc_param_dict(pr), "", groups)
    dy = torch.randn_like(yr)
    gr = torch.autograd.grad(yr, [xr] + pr, dy)
    xg, pg = _leaf(x), [_leaf(p) for p in params]
    y = ops.jlc(xg, pg, groups, e)
    assert rel_err(y, yr) < 1e-4, rel_err(y, yr)
    gg = torch.autograd.grad(y, [xg] + pg, dy.to(DEV))
    zero_grad = (2, 4, 6)     # conv biases in front of an affine-less InstanceNorm: the true gradient is exactly 0
    bad = [(i, rel_err(a, b)) for i, (a, b) in enumerate(zip(gg, gr))
           if i not in zero_grad and not close(a, b, rtol=1e-3, atol=2e-5)]
    assert not bad, bad
    for i in zero_grad:       # both sides are rounding noise; bound it by 1e-3 of the same conv's weight-gradient scale
        assert float(gg[i].abs().max()) <= 1e-3 * float(gr[i - 1].abs().max()), (i, float(gg[i].abs().max()))


@pytest.mark.parametrize("shape", [(12, 12, 12), (24, 24, 24)])      # SIMT contraction kernels / tcgen05 kernels
def test_jlc_dropout_mask_consistency(ops, shape):
    """train-mode dropout: same seed -> same mask in forward and backward; elements are either dropped or scaled 1/(1-p)."""
    from veloxseg_b200 import _lib
    torch.manual_seed(2)
    C, groups, e = 16, 4, 3
    x = torch.randn(2, C, *shape, device=DEV)
    params = [p.to(DEV) for p in jlc_params(C, groups, e, seed=5)]
    lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
    y0, z, o, hpre, stats = ops.jlc_fwd_raw(lib, st, x, params, groups, e, 0.0, False, 0)
    y1, _, o1, *_ = ops.jlc_fwd_raw(lib, st, x, params, groups, e, 0.5, True, 1234)
    y2, _, o2, *_ = ops.jlc_fwd_raw(lib, st, x, params, groups, e, 0.5, True, 1234)
    y3, _, o3, *_ = ops.jlc_fwd_raw(lib, st, x, params, groups, e, 0.5, True, 99)
    # same seed -> same mask (values agree to rounding only: the InstanceNorm partial sums use fp32 atomics)
    assert torch.equal(y1 == o1, y2 == o2) and torch.allclose(y1, y2, rtol=1e-4, atol=1e-5)
    assert not torch.equal(y1 == o1, y3 == o3)
    d0, d1 = (y0 - o), (y1 - o1)
    dropped = d1 == 0
    assert 0.45 < dropped.float().mean().item() < 0.55
    assert torch.allclose(d1[~dropped], 2 * d0[~dropped], rtol=1e-3, atol=1e-4)
    # backward uses the same mask: d(sum y)/d(fb2) = kept fraction * 2 per channel
    xg = x.clone().requires_grad_(True)
    pg = [p.clone().requires_grad_(True) for p in params]
    from veloxseg_b200.ops import _JLC
    y = _JLC.apply(xg, groups, e, 0.5, True, 1234, *pg)
    g = torch.autograd.grad(y.sum(), pg[9])[0]
    want = (~dropped).float().sum(dim=(0, 2, 3, 4)) * 2
    assert torch.allclose(g, want, rtol=1e-4)


@pytest.mark.parametrize("chs,Co,shape,B,addend", [((16, 16), 16, (24, 24, 24), 2, True), ((128, 128), 128, (3, 3, 3), 2, True),
                                                    ((64,), 64, (6, 6, 6), 1, True), ((32, 32), 32, (16, 16, 8), 3, False),
                                                    ((16, 16), 16, (32, 32, 16), 1, True)])
def test_mixer_levels(ops, chs, Co, shape, B, addend):
    torch.manual_seed(0)
    streams = [torch.randn(B, c, *shape) for c in chs]
    W, b = torch.randn(Co, sum(chs)) * 0.2, torch.randn(Co) * 0.1
    add = torch.randn(B, Co, *shape) if addend else None
    sr = [s.clone().requires_grad_(True) for s in streams]
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.modal_mixer(sr, Wr, br, add)
    dy = torch.randn_like(yr)
    gr = torch.autograd.grad(yr, sr + [Wr, br], dy)
    sg, Wg, bg = [_leaf(s) for s in streams], _leaf(W), _leaf(b)
    y = ops.modal_mixer(sg, Wg, bg, add.to(DEV) if addend else None)
    assert rel_err(y, yr) < 1e-4
    gg = torch.autograd.grad(y, sg + [Wg, bg], dy.to(DEV))
    for i, (a, r) in enumerate(zip(gg[:-1], gr[:-1])):
        assert close(a, r, rtol=1e-3, atol=2e-5), (i, rel_err(a, r))
    # the bias sits in front of an affine-less InstanceNorm: its true gradient is exactly 0
    assert float(gg[-1].abs().max()) <= 1e-3 * float(gr[-2].abs().max())


PWA_LEVELS = {  # name: (size, C, min_big, heads, min_dim_head, M, e)
    "autopet_L1": ((24, 24, 24), 16, [3, 3, 3], 1, 4, 2, 3), "autopet_L2": ((12, 12, 12), 32, [6, 6, 6], 2, 8, 2, 3),
    "autopet_L3": ((6, 6, 6), 64, [3, 3, 3], 2, 8, 2, 2), "autopet_L4": ((3, 3, 3), 128, [3, 3, 3], 4, 16, 2, 2),
    "hecktor_L1": ((32, 32, 16), 16, [4, 4, 2], 1, 4, 2, 3), "hecktor_L2": ((16, 16, 8), 32, [8, 8, 4], 2, 8, 2, 3),
    "brats_L1": ((24, 24, 24), 16, [3, 3, 3], 1, 4, 1, 3), "brats_L2": ((12, 12, 12), 32, [6, 6, 6], 2, 8, 1, 3)}


@pytest.mark.parametrize("lvl", list(PWA_LEVELS))
def test_pwa_gather_bit_exact(ops, lvl):
    from veloxseg_b200 import _lib
    size, C, mb, heads, mdh, M, e = PWA_LEVELS[lvl]
    geo = O.pwa_geometry(size, C, mb, [1, 1, 1], 2, heads, mdh)
    torch.manual_seed(0)
    for Ct in (geo["cqk"], geo["cv"]):
        x = torch.randn(2, Ct, *size)
        tok, arg = ops.pwa_gather_raw(_lib.get_lib(), torch.cuda.current_stream().cuda_stream, x.to(DEV), geo)
        ref, _, _ = O.gather_tokens(x, heads, geo["bws"], geo["sws"])
        assert torch.equal(tok.cpu(), ref)
        # arg-max indices address exactly the gathered values
        cper = Ct // (heads * len(geo["bws"]))
        flat = x.reshape(2, Ct, -1)
        off = 0
        for j, bw in enumerate(geo["bws"]):
            Nj = (size[0] // bw[0]) * (size[1] // bw[1]) * (size[2] // bw[2])
            for h in range(heads):
                for c in range(cper):
                    a = arg[:, h, off:off + Nj, :, c].reshape(2, -1).long().cpu()
                    assert torch.equal(torch.gather(flat[:, (j * heads + h) * cper + c], 1, a),
                                       ref[:, h, off:off + Nj, :, c].reshape(2, -1))
            off += Nj


@pytest.mark.parametrize("lvl", ["autopet_L1", "autopet_L2", "autopet_L3", "hecktor_L1"])
def test_pwa_dropout_backward_matches_forward_masks(ops, lvl):
    """Train-mode dropout (attention weights + projections): with a fixed seed the block is a deterministic smooth function,
    so the backward pass (which regenerates every mask, with the same hashed words per element) must agree with a
    central finite difference of the forward pass; a mask mismatch is an O(1) relative error."""
    from veloxseg_b200 import _lib
    size, C, mb, heads, mdh, M, e = PWA_LEVELS[lvl]
    geo = O.pwa_geometry(size, C, mb, [1, 1, 1], 2, heads, mdh)
    torch.manual_seed(11)
    lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
    xs = [torch.randn(1, C, *size, device=DEV) for _ in range(M)]
    flat, pd, table, index = pwa_params(M, C, geo, e, seed=4)
    flat, table, index = [p.to(DEV) for p in flat], table.to(DEV), index.to(DEV)
    p_att, p_proj, seed = 0.3, 0.2, 77

    def fwd(inp):
        return ops.pwa_block_fwd_raw(lib, st, inp, flat, table, index, geo, e, p_att, p_proj, True, seed)

    zs, saved = fwd(xs)
    zs2, _ = fwd(xs)
    assert all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(zs, zs2))
    dzs = [torch.randn_like(z) for z in zs]
    dxs, dps, dtable = ops.pwa_block_bwd_raw(lib, st, dzs, xs, flat, table, index, saved, geo, e, p_att, p_proj, True, seed)
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


@pytest.mark.parametrize("lvl", ["autopet_L2", "hecktor_L2"])
def test_pwa_attention_tensor_core_matches_simt(ops, lvl):
    """Windows of L >= 128 tokens with 8 channels per head (level 2: L = 432 / 512) run the attention forward on tcgen05
    (QK^T / PV as 3xTF32 MMAs, P in tensor memory).  Same inputs through the fp32 SIMT kernel (VX_OPT_ATTN_TC = 0): outputs
    agree to fp32 round-off, with attention dropout on as well (both kernels draw the same hashed mask words), and the profile
    shows which kernel ran."""
    from veloxseg_b200 import _lib
    size, C, mb, heads, mdh, M, e = PWA_LEVELS[lvl]
    geo = O.pwa_geometry(size, C, mb, [1, 1, 1], 2, heads, mdh)
    torch.manual_seed(3)
    lib, st = _lib.get_lib(), torch.cuda.current_stream().cuda_stream
    xs = [torch.randn(2, C, *size, device=DEV) for _ in range(M)]
    flat, pd, table, index = pwa_params(M, C, geo, e, seed=6)
    flat, table, index = [p.to(DEV) for p in flat], table.to(DEV), index.to(DEV)
    out = {}
    try:
        for on in (0, 1):
            lib.set_option(17, on)
            lib.profile(True)
            z0, _ = ops.pwa_block_fwd_raw(lib, st, xs, flat, table, index, geo, e)
            torch.cuda.synchronize()
            kernels = {k for _, k, *_ in lib.profile_report()}
            lib.profile(False)
            assert ("pwa_attn_fwd_tc_kernel" in kernels) == bool(on), kernels
            z1, _ = ops.pwa_block_fwd_raw(lib, st, xs, flat, table, index, geo, e, 0.3, 0.0, True, 91)
            out[on] = ([z.clone() for z in z0], [z.clone() for z in z1])
    finally:
        lib.set_option(17, 1)
        lib.profile(False)
    for k in (0, 1):
        for a, b in zip(out[0][k], out[1][k]):
            assert rel_err(a, b) < 5e-6, (k, rel_err(a, b))


@pytest.mark.parametrize("lvl", list(PWA_LEVELS))
def test_pwa_block_levels(ops, lvl):
    size, C, mb, heads, mdh, M, e = PWA_LEVELS[lvl]
    B = 2 if size[0] <= 12 else 1
    geo = O.pwa_geometry(size, C, mb, [1, 1, 1], 2, heads, mdh)
    torch.manual_seed(1)
    xs = [torch.randn(B, C, *size) for _ in range(M)]
    flat, pd, table, index = pwa_params(M, C, geo, e, seed=2)
    xr = [x.clone().requires_grad_(True) for x in xs]
    pr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in pd.items()}
    zr = O.pwa_block(xr, pr, "", geo)
    dzs = [torch.randn_like(z) for z in zr]
    plist = [pr[n.format(m=m)] for m in range(M) for n in PWA_PARAM_NAMES]
    tab = pr["attn.position_embedding.relative_position_bias_table"]
    gr = torch.autograd.grad(zr, xr + plist + [tab], dzs)
    xg, pg, tg = [_leaf(x) for x in xs], [_leaf(p) for p in flat], _leaf(table)
    zs = ops.pwa_block(xg, pg, tg, index.to(DEV), geo, e)
    for m in range(M):
        assert rel_err(zs[m], zr[m]) < 1e-4, (m, rel_err(zs[m], zr[m]))
    gg = torch.autograd.grad(zs, xg + pg + [tg], [d.to(DEV) for d in dzs])
    names = [f"dx{m}" for m in range(M)] + [n.format(m=m) for m in range(M) for n in PWA_PARAM_NAMES] + ["table"]
    # with a single modality the key bias shifts every score of a row equally: softmax-invariant, true gradient 0
    zero = {"attn.qkv_proj.0.1.bias"} if M == 1 else set()
    bad = [(n, rel_err(a, r)) for n, a, r in zip(names, gg, gr) if n not in zero and not close(a, r, rtol=1e-3, atol=2e-5)]
    assert not bad, bad
    for n, a, r in zip(names, gg, gr):
        if n in zero:
            assert float(a.abs().max()) <= 1e-3 * float(gr[names.index("attn.qkv_proj.0.1.weight")].abs().max()), n


def test_gram_sdkt_levels(ops):
    torch.manual_seed(0)
    for shape in [(2, 16, 24, 24, 24), (1, 16, 32, 32, 16), (4, 16, 24, 24, 24)]:
        f = torch.randn(*shape)
        fr = f.clone().requires_grad_(True)
        Gr = O.gram(fr)
        dG = torch.randn_like(Gr)
        fg = _leaf(f)
        Gg = ops.gram(fg)
        assert rel_err(Gg, Gr) < 1e-5
        assert close(torch.autograd.grad(Gg, fg, dG.to(DEV))[0], torch.autograd.grad(Gr, fr, dG)[0], rtol=1e-3, atol=1e-9)
    gs = torch.randn(4, 16, 16)
    gts = [torch.randn(4, 16, 16) for _ in range(2)]
    lr_in = [t.clone().requires_grad_(True) for t in [gs] + gts]
    Lr = O.sdkt_loss(lr_in[0], lr_in[1:])
    lg_in = [_leaf(t) for t in [gs] + gts]
    Lg = ops.sdkt_loss(lg_in[0], lg_in[1:])
    assert abs(float(Lg) - float(Lr)) < 1e-5 * abs(float(Lr))
    for a, r in zip(torch.autograd.grad(Lg * 1.7, lg_in), torch.autograd.grad(Lr * 1.7, lr_in)):
        assert close(a, r, rtol=1e-4)


@pytest.mark.parametrize("src,dst,planes", [((3, 3, 3), (96, 96, 96), (4, 2)), ((6, 6, 6), (96, 96, 96), (1, 4)),
                                             ((12, 12, 12), (96, 96, 96), (2, 2)), ((8, 8, 4), (128, 128, 64), (1, 2)),
                                             ((24, 24, 24), (24, 24, 24), (1, 2))])
def test_resize_trilinear(ops, src, dst, planes):
    """scale_prediction (VeloxSeg.py:177-184): trilinear, align_corners=True; checked against ATen on the CPU."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.randn(*planes, *src)
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, size=dst, mode="trilinear", align_corners=True)
    dy = torch.randn_like(yr)
    gr, = torch.autograd.grad(yr, xr, dy)
    xg = _leaf(x)
    y = ops.resize_trilinear(xg, dst)
    assert rel_err(y, yr) < 1e-5
    g, = torch.autograd.grad(y, xg, dy.to(DEV))
    assert close(g, gr, rtol=1e-4, atol=1e-6), rel_err(g, gr)


@pytest.mark.parametrize("C,shape,B", [(2, (96, 96, 96), 4), (4, (96, 96, 96), 1), (2, (128, 128, 64), 2), (3, (7, 5, 3), 2)])
def test_segloss_levels(ops, C, shape, B):
    """Fused deep-supervision loss (CE + Dice x 4 outputs) at the full-size shapes of the three configs: value and
    gradients against torch cross_entropy + the restated MONAI DiceLoss, through autograd (ops.seg_loss)."""
    import torch.nn.functional as F
    from veloxseg_b200.loss import dice_loss
    torch.manual_seed(11)
    logits = [(torch.randn(B, C, *shape, device=DEV) * 3).requires_grad_(True) for _ in range(4)]
    labels = (torch.rand(B, 1, *shape, device=DEV) * C).long().clamp_(max=C - 1)
    w = [0.25, 0.25, 0.25, 0.25]
    loss = ops.seg_loss(logits, labels, w)
    g = torch.autograd.grad(loss * 1.3, logits)
    lr = [t.detach().double().requires_grad_(True) for t in logits]
    ref = sum(wi * (F.cross_entropy(o, labels.squeeze(1)) + dice_loss(o, labels)) for wi, o in zip(w, lr))
    gr = torch.autograd.grad(ref * 1.3, lr)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref)), (float(loss), float(ref))
    for a, b in zip(g, gr):
        assert close(a, b, rtol=1e-4, atol=1e-12), rel_err(a, b)


@pytest.mark.parametrize("Ct,c_off,Ci,size,B", [(2, 0, 1, (96, 96, 96), 2), (2, 1, 1, (128, 128, 64), 1), (4, 0, 4, (96, 96, 96), 1)])
def test_patch_embed_stem(ops, Ct, c_off, Ci, size, B):
    """Network-input stem (PatchEmbed k = s = 4 on a channel range of the input) vs torch conv3d in fp32."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.randn(B, Ct, *size, device=DEV)
    w = (torch.randn(16, Ci, 4, 4, 4, device=DEV) * 0.2).requires_grad_(True)
    b = (torch.randn(16, device=DEV) * 0.1).requires_grad_(True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        yr = F.conv3d(x[:, c_off:c_off + Ci], w, b, stride=4)
        dy = torch.randn_like(yr)
        gw, gb = torch.autograd.grad(yr, [w, b], dy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    wg, bg = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y = ops.patch_embed(x, c_off, wg, bg)
    assert rel_err(y, yr) < 1e-5
    dw, db = torch.autograd.grad(y, [wg, bg], dy)
    assert close(dw, gw, rtol=1e-3, atol=1e-4) and close(db, gb, rtol=1e-3, atol=1e-4), (rel_err(dw, gw), rel_err(db, gb))


def test_adamw_one_launch_matches_torch(ops):
    """veloxseg_b200.train.VxAdamW vs torch.optim.AdamW on the tiny model's parameter list, 3 steps."""
    from veloxseg_b200.configs import MODEL_CONFIGS
    from veloxseg_b200.nn import VeloxSeg
    from veloxseg_b200.train import VxAdamW
    torch.manual_seed(0)
    m = VeloxSeg(**MODEL_CONFIGS["tiny"]).to(DEV)
    ps = [p for p in m.parameters()]
    ref = [p.detach().clone().requires_grad_(True) for p in ps]
    opt_r = torch.optim.AdamW(ref, lr=2.5e-4, weight_decay=0.01)
    opt = VxAdamW(ps, lr=2.5e-4, weight_decay=0.01)
    for it in range(3):
        for p, r in zip(ps, ref):
            g = torch.randn_like(p) * 0.1
            p.grad, r.grad = g, g.clone()
        opt.step()
        opt_r.step()
    torch.cuda.synchronize()
    worst = max(float((p.detach() - r.detach()).abs().max() / (r.detach().abs().max() + 1e-12)) for p, r in zip(ps, ref))
    assert worst < 1e-5, worst


@pytest.mark.parametrize("B,C,s,size", [(4, 2, 4, (24, 24, 24)), (2, 1, 4, (32, 32, 16)), (1, 4, 4, (24, 24, 24))])
def test_pixel_shuffle_bias(ops, B, C, s, size):
    """bias + 3-D pixel shuffle (out_conv1 / reconstruction out_conv tail) vs the reference permutation and its autograd."""
    from veloxseg_b200.nn import PixelShuffle
    torch.manual_seed(0)
    z = torch.randn(B, C * s ** 3, *size, device=DEV, requires_grad=True)
    bias = torch.randn(C * s ** 3, device=DEV, requires_grad=True)
    yr = PixelShuffle(s)(z + bias[None, :, None, None, None])
    zg, bg = z.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    y = ops.pixel_shuffle_bias(zg, bg, s)
    assert torch.equal(y, yr)
    dy = torch.randn_like(yr)
    gz, gb = torch.autograd.grad(yr, [z, bias], dy)
    dz, db = torch.autograd.grad(y, [zg, bg], dy)
    assert torch.equal(dz, gz) and close(db, gb, rtol=1e-4, atol=1e-3)


# ----------------------------------------------------------------------------------------------------
# convolutions of the glue layers: vx_conv_fwd / vx_conv_bwd
# ----------------------------------------------------------------------------------------------------
def _conv_ref(x, w, b, kernel, stride, pad, transposed, shuffle):
    import torch.nn.functional as F
    from veloxseg_b200.nn import PixelShuffle
    y = F.conv_transpose3d(x, w, b, stride=stride) if transposed else F.conv3d(x, w, b, stride=stride, padding=pad)
    return PixelShuffle(shuffle, 3)(y) if shuffle else y


@pytest.mark.parametrize("Cout,shape,B,shuffle", [(128, (24, 24, 24), 4, 4), (64, (24, 24, 24), 2, 4), (256, (24, 24, 24), 1, 4),
                                                  (128, (32, 32, 16), 1, 4), (64, (5, 6, 8), 2, 4), (32, (7, 9, 12), 1, 0),
                                                  (128, (6, 10, 20), 1, 0)])
def test_conv3_tensor_core(ops, Cout, shape, B, shuffle):
    """decoder.out_conv1 / reconstruction out_conv (Decoder.py:73-76,150-153) on the tcgen05 kernels of conv3_tc.cu against
    conv3d (+ PixelShuffle, superpixel.py:15) in fp64: forward, data gradient, weight and bias gradients at 1e-5 (3xTF32 is
    fp32-accurate; single-pass TF32 would sit at ~1e-3)."""
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, 16, *shape, generator=g).to(DEV)
    w = (torch.randn(Cout, 16, 3, 3, 3, generator=g) * 0.05).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ops.conv3d(xr, wr, br, 3, 1, 1, shuffle=shuffle)
    x2, w2, b2 = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = _conv_ref(x2, w2, b2, 3, 1, 1, False, shuffle)
    assert y.shape == yr.shape and rel_err(y, yr) < 1e-5, rel_err(y, yr)
    dy = torch.randn(y.shape, generator=g).to(DEV)
    y.backward(dy)
    yr.backward(dy.double())
    torch.cuda.synchronize()
    assert rel_err(xr.grad, x2.grad) < 1e-5, rel_err(xr.grad, x2.grad)
    assert rel_err(wr.grad, w2.grad) < 1e-5, rel_err(wr.grad, w2.grad)
    assert rel_err(br.grad, b2.grad) < 1e-5, rel_err(br.grad, b2.grad)
    y2 = ops.conv3d(x, w, b, 3, 1, 1, shuffle=shuffle)          # bit-reproducible (fixed fold order everywhere)
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("Ci,Co,k,s,p,tr,shape,B", [
    (2, 16, 7, 4, 3, False, (96, 96, 96), 4), (4, 16, 7, 4, 3, False, (96, 96, 96), 1), (2, 16, 7, 4, 3, False, (128, 128, 64), 1),
    (16, 32, 3, 2, 1, False, (24, 24, 24), 4), (32, 64, 3, 2, 1, False, (12, 12, 12), 4), (64, 128, 3, 2, 1, False, (6, 6, 6), 4),
    (64, 128, 3, 2, 1, False, (8, 8, 4), 1),
    (128, 64, 2, 2, 0, True, (3, 3, 3), 4), (64, 32, 2, 2, 0, True, (6, 6, 6), 4), (32, 16, 2, 2, 0, True, (12, 12, 12), 4),
    (32, 16, 2, 2, 0, True, (16, 16, 8), 1),
    (32, 2, 1, 1, 0, False, (12, 12, 12), 4), (128, 4, 1, 1, 0, False, (3, 3, 3), 2), (1, 16, 4, 4, 0, False, (32, 32, 32), 1),
    (5, 7, 3, 2, 1, False, (5, 7, 9), 2), (8, 128, 3, 1, 1, False, (16, 16, 16), 2)])
def test_conv_strided_transposed(ops, Ci, Co, k, s, p, tr, shape, B):
    """DownConv.down / UpConv.up (conv_blocks.py:10-17,31-35), the 1x1 deep-supervision heads (Decoder.py:155-158) and a
    PatchEmbed-shaped stem at the shapes of the three configs, against torch in fp64."""
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, Ci, *shape, generator=g).to(DEV)
    w = (torch.randn(*((Ci, Co) if tr else (Co, Ci)), k, k, k, generator=g) * 0.1).to(DEV)
    b = torch.randn(Co, generator=g).to(DEV)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ops.conv3d(xr, wr, br, k, s, p, transposed=tr)
    x2, w2, b2 = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = _conv_ref(x2, w2, b2, k, s, p, tr, 0)
    assert y.shape == yr.shape and rel_err(y, yr) < 2e-6, rel_err(y, yr)
    dy = torch.randn(y.shape, generator=g).to(DEV)
    y.backward(dy)
    yr.backward(dy.double())
    torch.cuda.synchronize()
    assert rel_err(xr.grad, x2.grad) < 2e-6, rel_err(xr.grad, x2.grad)
    assert rel_err(wr.grad, w2.grad) < 1e-5, rel_err(wr.grad, w2.grad)
    assert rel_err(br.grad, b2.grad) < 1e-5, rel_err(br.grad, b2.grad)
