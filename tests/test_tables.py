"""Integer paths, bit-exact: window pyramids, channel counts, relative-position index, gather order (PWA.py:56-140,
attention_utils.py:89-117) against tables minted from the reference; sliding-window slice enumeration properties."""
import pytest
import torch

from oracle import veloxseg_oracle as O
from tests import _golden as G
from veloxseg_b200.configs import MODEL_CONFIGS

TABLES = G.load("tables.pt")


@pytest.mark.parametrize("name", list(MODEL_CONFIGS))
def test_window_tables(name):
    from veloxseg_b200.nn import VeloxSeg
    torch.manual_seed(0)
    m = VeloxSeg(**MODEL_CONFIGS[name])
    spec = O.ModelSpec(MODEL_CONFIGS[name])
    for i, ref in enumerate(TABLES[name]):
        a = m.encoder.encoder_attn.layers[i].blocks[0].attn
        geo = spec.geo[i]
        for got in (dict(big=a.big_window_size, small=a.small_window_size, cqk=a.channels_qk, cv=a.channels_v, n=a.n_hwd),
                    dict(big=geo["bws"], small=geo["sws"], cqk=geo["cqk"], cv=geo["cv"], n=geo["n"])):
            assert got["big"] == ref["big"] and got["small"] == ref["small"]
            assert got["cqk"] == ref["channels_qk"] and got["cv"] == ref["channels_v"] and list(got["n"]) == ref["n"]
        idx = a.position_embedding.relative_position_index
        assert idx.dtype == torch.int64
        assert G.sha_int32(idx) == ref["rel_index_sha"]
        assert G.sha_int32(O.relative_position_index(ref["n"])) == ref["rel_index_sha"]
        if ref["rel_index"] is not None:
            assert torch.equal(idx, ref["rel_index"])
        # gather order on a voxel-index ramp
        h, w, d = ref["input_size"]
        cqk = ref["channels_qk"]
        ramp = torch.arange(h * w * d, dtype=torch.float32).view(1, 1, h, w, d).repeat(1, cqk, 1, 1, 1)
        ramp = ramp + torch.arange(cqk, dtype=torch.float32).view(1, -1, 1, 1, 1) * (h * w * d)
        tok, Ns, n = O.gather_tokens(ramp, a.num_heads, ref["big"], ref["small"])
        assert Ns == ref["Ns"] and n == ref["n"]
        assert G.sha_int32(tok) == ref["gather_ramp_sha"]


def test_sliding_window_slices():
    # MONAI 1.5.0 dense_patch_slices restated (parity unpinned: MONAI is not installable here); properties instead
    for image, roi in [((320, 320, 256), (128, 128, 64)), ((96, 96, 96), (96, 96, 96)), ((100, 130, 70), (96, 96, 96)),
                       ((512, 512, 384), (128, 128, 64)), ((97, 96, 200), (96, 96, 96))]:
        img = [max(i, r) for i, r in zip(image, roi)]
        sl = O.sw_slices(img, roi, 0.25)
        assert sl[0] == (0, 0, 0)
        cover = torch.zeros(img, dtype=torch.int32)
        for a, b, c in sl:
            assert 0 <= a <= img[0] - roi[0] and 0 <= b <= img[1] - roi[1] and 0 <= c <= img[2] - roi[2]
            cover[a:a + roi[0], b:b + roi[1], c:c + roi[2]] += 1
        assert int(cover.min()) >= 1                      # no gaps
        assert sl == sorted(sl)                           # first spatial axis slowest
        assert max(s[0] for s in sl) == img[0] - roi[0]   # last window ends at the border
    assert len(O.sw_slices((320, 320, 256), (128, 128, 64), 0.25)) == 45      # 3 x 3 x 5
    assert len(O.sw_slices((512, 512, 384), (128, 128, 64), 0.25)) == 200
    assert O.sw_scan_interval((320, 320, 256), (128, 128, 64), 0.25) == [96, 96, 48]


def test_count_map_factorises_into_axis_tables():
    # the window list is a Cartesian product of per-axis starts, so the overlap count is an outer product (integer-exact)
    from veloxseg_b200.inference import axis_counts, axis_starts, count_map, window_starts
    for image, roi in [((320, 320, 256), (128, 128, 64)), ((100, 130, 70), (96, 96, 96)), ((97, 96, 200), (96, 96, 96)),
                       ((20, 17, 9), (8, 8, 8))]:
        img = [max(i, r) for i, r in zip(image, roi)]
        per_axis = axis_starts(img, roi, 0.25)
        starts = window_starts(img, roi, 0.25)
        assert starts == O.sw_slices(img, roi, 0.25)
        cx, cy, cz = axis_counts(img, roi, per_axis, "cpu")
        assert torch.equal(cx[:, None, None] * cy[None, :, None] * cz[None, None, :], count_map(img, roi, starts, "cpu")[0, 0])
        assert int(cx.min()) >= 1 and int(cy.min()) >= 1 and int(cz.min()) >= 1


def test_sliding_window_blend_properties():
    # size-independent properties of the restated MONAI blend (constant importance map): an identity predictor gives the
    # volume back (every voxel is the mean of identical copies), also through the symmetric padding of a volume smaller
    # than the roi; a predictor that returns the window index map reproduces it, so every window lands where it was cut.
    from veloxseg_b200.inference import sliding_window_labels, sliding_window_predict
    g = torch.Generator().manual_seed(3)
    for shape, roi in [((1, 2, 20, 17, 9), (8, 8, 8)), ((1, 2, 6, 10, 8), (8, 8, 8)), ((2, 2, 12, 8, 10), (8, 8, 4))]:
        x = torch.randn(*shape, generator=g)
        y = sliding_window_predict(x, lambda w: w, roi, sw_batch_size=3, overlap=0.25, shard=False)
        assert y.shape == x.shape and torch.allclose(y, x, rtol=1e-6, atol=1e-7)
        y2 = sliding_window_predict(x, lambda w: [w * 2.0, w], roi, sw_batch_size=2, overlap=0.25, shard=False)   # list output
        assert torch.allclose(y2, 2.0 * x, rtol=1e-6, atol=1e-7)
    ramp = torch.arange(20 * 17 * 9, dtype=torch.float32).view(1, 1, 20, 17, 9)
    two = torch.cat([ramp, -ramp], 1)                       # class 0 wins where ramp > 0: labels are 0 except at voxel 0 (tie)
    lab = sliding_window_labels(two, lambda w: w, (8, 8, 8), "cpu", sw_batch_size=2, overlap=0.25)
    assert lab.shape == (20, 17, 9) and int(lab.sum()) == 0
    lab = sliding_window_labels(-two, lambda w: w, (8, 8, 8), "cpu", sw_batch_size=2, overlap=0.25)
    assert int(lab.sum()) == 20 * 17 * 9 - 1                # class 1 everywhere but the tie at voxel 0 (first index wins)


def test_upload_schedule_covers_every_window_batch():
    """One-GPU sliding window: the volume is uploaded in (plane slab x row slab) blocks in window order and a batch waits for
    the blocks up to its largest (a + r0, b + r1).  Host logic only: for ragged volumes, every overlap and batch size, the
    blocks a batch has waited for cover every voxel its windows read, and every block is uploaded exactly once."""
    import itertools
    from veloxseg_b200.inference import axis_starts, blocks_ready, upload_blocks
    for size, roi, overlap, sw in itertools.product([(20, 17, 9), (32, 32, 16), (45, 24, 8), (16, 40, 8)], [(16, 16, 8), (12, 8, 8)],
                                                    [0.0, 0.25, 0.5], [1, 3, 4]):
        if any(s < r for s, r in zip(size, roi)):
            continue
        blocks = upload_blocks(size, roi, overlap)
        cover = torch.zeros(size[0], size[1], dtype=torch.int32)
        for x0, x1, y0, y1 in blocks:
            cover[x0:x1, y0:y1] += 1
        assert bool((cover == 1).all()), (size, roi, overlap)             # a partition of the (x, y) plane
        assert [(b[1], b[3]) for b in blocks] == sorted((b[1], b[3]) for b in blocks)      # upload order = key order
        per = axis_starts(size, roi, overlap)
        starts = [(a, b, c) for a in per[0] for b in per[1] for c in per[2]]
        pending = [((b[1], b[3]), b) for b in blocks]
        have = torch.zeros(size[0], size[1], dtype=torch.bool)
        for g in range(0, len(starts), sw):
            ids = starts[g:g + sw]
            need = max((a + roi[0], b + roi[1]) for a, b, _ in ids)
            for _, (x0, x1, y0, y1) in blocks_ready(pending, need):
                have[x0:x1, y0:y1] = True
            for a, b, _ in ids:
                assert bool(have[a:a + roi[0], b:b + roi[1]].all()), (size, roi, overlap, sw, (a, b))
        assert not pending, (size, roi, overlap)      # the last batch has waited for everything
