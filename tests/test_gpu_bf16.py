"""bf16 numerics mode (north_star's second precision mode, BASELINE configs[1] "bf16 on 1xB200").

SURVEY.md section 8c, "bf16 caveat": on pure-noise input even the reference's own autocast-bf16 forward moves Dice by 0.3
point, so the Dice check uses STRUCTURED synthetic volumes (smooth blobs: boundary voxels are a small fraction) and compares
Dice(pred, label) of the bf16 mode with Dice(pred, label) of the reference arithmetic (the oracle restatement in fp32 on the
CPU, pinned to the unmodified reference by tests/test_oracle_golden.py).  Dice formula: utils/metric/metrics.py:93-94,
2 |gt & pred| / (|gt| + |pred| + 1e-5), in points (x 100).  Tolerance: 0.1 point (north_star)."""
import pytest
import torch

from tests import _golden as G
from veloxseg_b200.configs import MODEL_CONFIGS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def blob_volumes(n, size, seed):
    """n two-modality volumes with 2-4 smooth ellipsoid lesions each: modality 0 bright inside the lesions, modality 1 a
    weaker, blurred copy; additive low-amplitude smooth noise.  Returns x (n, 2, *size) fp32, y (n, 1, *size) int64."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = size
    zz, yy, xx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing="ij")
    xs, ys = [], []
    for _ in range(n):
        m = torch.zeros(D, H, W)
        for _ in range(int(torch.randint(2, 5, (1,), generator=g))):
            c = torch.rand(3, generator=g) * torch.tensor([D, H, W], dtype=torch.float32) * 0.6 + torch.tensor([D, H, W], dtype=torch.float32) * 0.2
            r = torch.rand(3, generator=g) * 0.12 * min(size) + 0.08 * min(size)
            d2 = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2
            m = torch.maximum(m, (d2 < 1.0).float())
        noise = torch.randn(2, D // 4, H // 4, W // 4, generator=g)
        noise = torch.nn.functional.interpolate(noise[None], size=size, mode="trilinear", align_corners=False)[0]
        soft = torch.nn.functional.avg_pool3d(m[None, None], 5, stride=1, padding=2)[0, 0]
        xs.append(torch.stack([2.5 * m + 0.3 * noise[0] - 0.5, 1.5 * soft + 0.3 * noise[1] - 0.3]))
        ys.append(m.long()[None])
    return torch.stack(xs), torch.stack(ys)


def dice_points(pred, gt):
    """utils/metric/metrics.py:93-94 on the foreground class, in points."""
    pred, gt = pred.bool(), gt.bool()
    return 100.0 * float(2.0 * (pred & gt).sum() / (gt.sum() + pred.sum() + 1e-5))


def _oracle_labels(model, cfg, x):
    from oracle import veloxseg_oracle as O
    p = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        return O.forward(x.cpu(), p, O.ModelSpec(cfg), training=False).argmax(1, keepdim=True)


@pytest.fixture
def precision():
    from veloxseg_b200 import ops
    yield ops.set_precision
    ops.set_precision("fp32")


def test_bf16_dice_trained_tiny_model(precision):
    """A miniature VeloxSeg trained for a few dozen steps on blob volumes (so that its predictions are lesion-shaped, not
    noise): Dice of the bf16-mode prediction on held-out volumes is within 0.1 point of the reference arithmetic's."""
    from veloxseg_b200.nn import VeloxSeg
    from veloxseg_b200.train import TrainStep
    cfg = MODEL_CONFIGS["tiny"]
    torch.manual_seed(G.MODEL_SEED)
    m = VeloxSeg(**cfg)
    G.zero_dropout(m)
    xtr, ytr = blob_volumes(4, cfg["input_size"], seed=11)
    ts = TrainStep(m, 2, DEV, lr=2e-3, use_graph=True)
    xd, yd = xtr.to(DEV), ytr.to(DEV)
    losses = [ts.step(xd, yd, sync=True) for _ in range(80)]
    assert losses[-1] < 0.7 * losses[0], (losses[0], losses[-1])
    xte, yte = blob_volumes(4, cfg["input_size"], seed=12)
    m.eval()
    ref = _oracle_labels(m, cfg, xte)
    d_ref = dice_points(ref == 1, yte == 1)
    res = {}
    for mode in ("fp32", "bf16"):
        precision(mode)
        with torch.no_grad():
            logits = m(xte.to(DEV))
        res[mode] = (dice_points(logits.argmax(1, keepdim=True).cpu() == 1, yte == 1), logits.float().cpu())
    rel = float((res["bf16"][1] - res["fp32"][1]).norm() / res["fp32"][1].norm())
    print(f"dice points: reference arithmetic {d_ref:.3f}, ours fp32 {res['fp32'][0]:.3f}, ours bf16 {res['bf16'][0]:.3f}; "
          f"bf16 vs fp32 logits rel {rel:.2e}; loss {losses[0]:.3f} -> {losses[-1]:.3f}")
    assert d_ref > 30.0, f"the trained miniature does not segment the blobs (Dice {d_ref:.2f}): the check would be vacuous"
    assert abs(res["fp32"][0] - d_ref) <= 0.1
    assert abs(res["bf16"][0] - d_ref) <= 0.1
    assert rel < 5e-2


def test_bf16_train_step_tracks_fp32(precision):
    """Three optimisation steps in bf16 mode stay close to the fp32 ones (same seeds, dropout off): the mode trains."""
    from veloxseg_b200.nn import VeloxSeg
    from veloxseg_b200.train import TrainStep
    cfg = MODEL_CONFIGS["tiny"]
    x, y = blob_volumes(2, cfg["input_size"], seed=5)
    out = {}
    for mode in ("fp32", "bf16"):
        precision(mode)
        torch.manual_seed(G.MODEL_SEED)
        m = VeloxSeg(**cfg)
        G.zero_dropout(m)
        ts = TrainStep(m, 2, DEV, use_graph=False)
        out[mode] = [ts.step(x.to(DEV), y.to(DEV), sync=True) for _ in range(3)]
    for a, b in zip(out["fp32"], out["bf16"]):
        assert abs(a - b) <= 2e-2 * abs(a), out


@pytest.mark.parametrize("name", ["autopetii"])
def test_bf16_dice_reference_config_seeded_init(name, precision):
    """SURVEY.md 8d "Dice-parity volumes": the real AutoPET-II configuration with the seed-12345 He initialisation on one
    structured 96^3 volume; prediction agreement with the reference arithmetic is reported next to the Dice difference."""
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS[name]
    torch.manual_seed(G.MODEL_SEED)
    m = VeloxSeg(**cfg).eval()
    x, y = blob_volumes(1, cfg["input_size"], seed=21)
    ref = _oracle_labels(m, cfg, x)
    m = m.to(DEV)
    d_ref = dice_points(ref == 1, y == 1)
    for mode, tol_agree in (("fp32", 0.9999), ("bf16", 0.995)):
        precision(mode)
        with torch.no_grad():
            lab = m(x.to(DEV)).argmax(1, keepdim=True).cpu()
        d = dice_points(lab == 1, y == 1)
        agree = float((lab == ref).float().mean())
        print(f"{name} {mode}: dice {d:.3f} vs reference arithmetic {d_ref:.3f}; label agreement {agree:.5f}")
        assert abs(d - d_ref) <= 0.1
        assert agree >= tol_agree
