"""The oracle (oracle/veloxseg_oracle.py) pinned against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import veloxseg_oracle as O
from tests import _golden as G
from tests._util import close, rel_err
from veloxseg_b200.configs import MODEL_CONFIGS

FX = G.load("ops_small.pt")


def _cots(outs, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(o.shape, generator=g) for o in outs]


def _check_grads(outs, leaves, names, fx, rtol=2e-4):
    grads = torch.autograd.grad(outs, leaves, _cots(outs, fx["cot_seed"]), allow_unused=True)
    nin = len(fx["input_grads"])
    for i in range(nin):
        assert close(grads[i], fx["input_grads"][i], rtol=rtol, atol=1e-6), ("input", i, rel_err(grads[i], fx["input_grads"][i]))
    for n, g in zip(names, grads[nin:]):
        r = fx["param_grads"][n]
        if r is None:
            continue
        assert close(g, r, rtol=rtol, atol=2e-5), (n, rel_err(g, r), float(r.norm()))


@pytest.mark.parametrize("tag", ["jlc_c8", "jlc_c16", "jlc_c32"])
def test_jlc(tag):
    fx = FX[tag]
    x = fx["inputs"][0].clone().requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in fx["state"].items()}
    y = O.jlc(x, p, "", fx["cfg"]["groups"])
    assert rel_err(y, fx["outputs"][0]) < 1e-5
    _check_grads([y], [x] + list(p.values()), list(p.keys()), fx)


@pytest.mark.parametrize("tag", ["mixer_2x16", "mixer_1x8"])
def test_mixer(tag):
    fx = FX[tag]
    ins = [t.clone().requires_grad_(True) for t in fx["inputs"]]
    W = fx["state"]["0.weight"].clone().requires_grad_(True)
    b = fx["state"]["0.bias"].clone().requires_grad_(True)
    y = O.modal_mixer(ins[1:], W, b, ins[0])
    assert rel_err(y, fx["outputs"][0]) < 1e-5
    _check_grads([y], ins + [W, b], ["0.weight", "0.bias"], fx)


@pytest.mark.parametrize("tag", ["pwa_6c8", "pwa_884", "pwa_12m1"])
def test_pwa_block(tag):
    fx = FX[tag]
    c = fx["cfg"]
    geo = O.pwa_geometry(c["size"], c["C"], c["mb"], c["ms"], 2, c["heads"], c["mdh"])
    assert {k: geo[k] for k in ("bws", "sws", "cqk", "cv", "n", "nb")} == {k: fx["geo"][k] for k in ("bws", "sws", "cqk", "cv", "n", "nb")}
    xs = [t.clone().requires_grad_(True) for t in fx["inputs"]]
    p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in fx["state"].items()}
    zs = O.pwa_block(xs, p, "", geo)
    for z, r in zip(zs, fx["outputs"]):
        assert rel_err(z, r) < 1e-5
    names = [k for k, v in p.items() if v.dtype.is_floating_point]
    _check_grads(zs, xs + [p[k] for k in names], names, fx)


def test_patch_merging_gram_norms():
    fx = FX["patch_merging"]
    y = O.patch_merging(fx["inputs"][0], fx["state"], "")
    assert rel_err(y, fx["outputs"][0]) < 1e-5
    fx = FX["gram"]
    assert rel_err(O.gram(fx["inputs"][0]), fx["outputs"][0]) < 1e-6
    fx = FX["down_conv"]
    assert rel_err(O.down_conv(fx["inputs"][0], fx["state"], "", 2), fx["outputs"][0]) < 1e-5
    fx = FX["up_conv"]
    assert rel_err(O.up_conv(fx["inputs"][0], fx["state"], ""), fx["outputs"][0]) < 1e-5


def _model(name):
    """Our module built on CPU under the reference's seed: identical state_dict, checked against the fixture."""
    from veloxseg_b200.nn import VeloxSeg
    torch.manual_seed(G.MODEL_SEED)
    m = VeloxSeg(**MODEL_CONFIGS[name])
    return m


@pytest.mark.parametrize("name", ["tiny", "autopetii", "hecktor2022", "brats2021"])
def test_state_dict_matches_reference(name):
    fx = G.load(f"model_{name}.pt")
    sd = _model(name).state_dict()
    assert list(sd.keys()) == list(fx["state"].keys())
    got = G.state_checksums(sd)
    for k, r in fx["state"].items():
        assert got[k]["shape"] == r["shape"] and got[k]["dtype"] == r["dtype"], k
        assert got[k]["sum"] == r["sum"] and got[k]["abs"] == r["abs"], k     # same RNG stream -> bit-identical init
    assert sum(p.numel() for p in _model(name).parameters()) == fx["n_params"]


@pytest.mark.parametrize("name", ["tiny", "autopetii", "hecktor2022", "brats2021"])
def test_oracle_whole_model(name):
    fx = G.load(f"model_{name}.pt")
    cfg = MODEL_CONFIGS[name]
    spec = O.ModelSpec(cfg)
    m = _model(name)
    p = {k: v.detach() for k, v in m.state_dict().items()}
    x = G.model_input(cfg, fx["B"])
    with torch.no_grad():
        G.check_sample(O.forward(x, p, spec, training=False), fx["eval"], rtol=1e-4, what="eval logits")
    if name in ("tiny", "autopetii"):        # train-mode outputs and per-parameter gradients
        pr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in p.items()}
        outs = O.forward(x, pr, spec, training=True)
        assert len(outs) == len(fx["train_outputs"])
        for i, (o, r) in enumerate(zip(outs, fx["train_outputs"])):
            G.check_sample(o, r, rtol=1e-4, atol=1e-7, what=f"train output {i}")
        loss = sum((o * c).sum() for o, c in zip(outs, G.cotangents(outs)))
        assert abs(float(loss) - fx["loss"]) < 1e-4 * max(1.0, abs(fx["loss"]))
        names = [k for k, v in pr.items() if v.dtype.is_floating_point]
        grads = torch.autograd.grad(loss, [pr[k] for k in names], allow_unused=True)
        for k, g in zip(names, grads):
            # north_star tolerance (1e-3 relative) + the atol that covers structurally-zero gradients
            G.check_sample(g if g is not None else torch.zeros_like(pr[k]), fx["grads"][k], rtol=1e-3, atol=1e-5, what=k)
