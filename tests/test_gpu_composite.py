"""-m gpu: K3a masked median and K3b nearest-date mosaic through the C ABI vs the oracle (bit-exact)."""
import datetime as dt

import numpy as np
import pytest

import synthetic as syn
from oracle import composite as ocomp

pytestmark = pytest.mark.gpu


def _check_median(dev, stack, valid, nodata=None):
    from dl_image_segmentation_b200 import ops
    out, mask = ops.median_composite(stack, valid, nodata, device=dev)
    ref = ocomp.median_composite(stack, valid, nodata)
    assert out.dtype.is_floating_point and out.element_size() == 8
    np.testing.assert_array_equal(mask.cpu().numpy(), np.ma.getmaskarray(ref))
    np.testing.assert_array_equal(out.cpu().numpy(), ref.filled(0.0))


def test_median_known_answers(dev):
    # values [5,1,9,7] under five validity patterns, and the no-overflow case
    vals = np.array([5, 1, 9, 7], np.uint16).reshape(4, 1, 1, 1).repeat(2, axis=3)
    from dl_image_segmentation_b200 import ops
    for pat, want in [("1111", 6.0), ("1110", 5.0), ("0101", 4.0), ("0010", 9.0), ("0000", None)]:
        v = np.array([int(c) for c in pat], np.uint8).reshape(4, 1, 1)
        out, mask = ops.median_composite(vals, v, device=dev)
        if want is None:
            assert mask.cpu().numpy().all() and (out.cpu().numpy() == 0.0).all()
        else:
            assert not mask.cpu().numpy().any() and (out.cpu().numpy() == want).all()
    big = np.array([65535, 65534, 3, 4], np.uint16).reshape(4, 1, 1, 1).repeat(2, axis=3)
    out, _ = ops.median_composite(big, np.array([1, 1, 0, 0], np.uint8).reshape(4, 1, 1), device=dev)
    assert (out.cpu().numpy() == 65534.5).all()


@pytest.mark.parametrize("T,B", [(16, 8), (16, 4), (12, 8), (5, 2), (32, 6), (3, 8), (1, 4), (2, 2), (9, 12)])
def test_median_matches_numpy_ma(dev, T, B):
    stack, valid = syn.cfg4_tile(T * 100 + B, T=T, H=48, W=40, B=B)
    _check_median(dev, stack, valid)


def test_median_extreme_values_and_all_masked(dev):
    rng = np.random.default_rng(5)
    stack = rng.choice(np.array([0, 1, 65534, 65535], np.uint16), size=(16, 32, 32, 8))
    valid = (rng.random((16, 32, 32)) > 0.5).astype(np.uint8)
    valid[:, :4, :4] = 0
    valid[:, 4:8, :4] = 1
    _check_median(dev, stack, valid)


def test_median_with_nodata_mask(dev):
    rng = np.random.default_rng(6)
    stack, valid = syn.cfg4_tile(77, T=16, H=32, W=32, B=8)
    nodata = (rng.random(stack.shape) > 0.8).astype(np.uint8)
    _check_median(dev, stack, valid, nodata)


@pytest.mark.parametrize("T,B", [(40, 8), (16, 3), (7, 1)])
def test_median_generic_fallback(dev, T, B):
    stack, valid = syn.cfg4_tile(T + B, T=T, H=16, W=24, B=B)
    _check_median(dev, stack, valid)


def test_median_full_size_properties(dev):
    """cfg4 full size (16,1024,1024,8): size-independent properties instead of a slow CPU oracle."""
    import torch
    from dl_image_segmentation_b200 import ops
    stack, valid = syn.cfg4_tile(0)
    sd, vd = ops.to_device(stack, dev), ops.to_device(valid, dev)
    out, mask = ops.median_composite(sd, vd, device=dev)
    o, m = out.cpu().numpy(), mask.cpu().numpy()
    n = valid.sum(axis=0)
    assert np.array_equal(m[..., 0], n == 0) and m[8:12, 8:12].all()
    assert ((o * 2) == np.floor(o * 2)).all()                       # only x.0 or x.5
    assert (o[n % 2 == 1] == np.floor(o[n % 2 == 1])).all()         # odd count -> an actual sample
    # permutation invariance over time, and agreement with the oracle on a crop
    perm = np.random.default_rng(1).permutation(16)
    pt = torch.as_tensor(perm, device=dev)
    out2, mask2 = ops.median_composite(sd.view(torch.int16)[pt].contiguous().view(sd.dtype), vd[pt].contiguous(), device=dev)
    assert torch.equal(out, out2) and torch.equal(mask, mask2)
    ref = ocomp.median_composite(stack[:, :64, :64], valid[:, :64, :64])
    np.testing.assert_array_equal(o[:64, :64], ref.filled(0.0))
    frac_half = float(((o * 2) % 2 == 1).mean())
    assert 0.15 < frac_half < 0.35                                  # ~25 % end in .5


def _check_mosaic(dev, stacks, valids, days, cfs, **flt):
    from dl_image_segmentation_b200 import ops
    out, mask, src, nel = ops.nearest_date_mosaic(stacks, valids, days, cfs, device=dev, **flt)
    out, mask, src, nel = out.cpu().numpy(), mask.cpu().numpy(), src.cpu().numpy(), nel.cpu().numpy()
    for i in range(len(stacks)):
        ref = ocomp.nearest_date_mosaic(stacks[i], valids[i], days[i], cfs[i], flt["ref_day"], flt.get("min_day"),
                                        flt.get("max_day"), flt.get("max_cf"))
        if ref is None:
            assert nel[i] == 0 and mask[i].all() and not out[i].any()
            continue
        r_out, r_mask, r_src = ref
        assert nel[i] > 0
        np.testing.assert_array_equal(out[i].view(stacks[i].dtype), r_out)
        np.testing.assert_array_equal(mask[i], r_mask)
        np.testing.assert_array_equal(src[i], r_src)


@pytest.mark.parametrize("dtype,B", [(np.uint16, 4), (np.uint16, 8), (np.uint8, 3), (np.uint16, 1), (np.float32, 2)])
def test_mosaic_matches_painters_loop(dev, dtype, B):
    stacks, valids, days, cfs = [], [], [], []
    for chip in range(6):
        s, v = syn.cfg5_chip(chip, T=32, H=40, W=36, B=B)
        stacks.append(s.astype(dtype))
        valids.append(v)
        d, c = syn.cfg5_scene_meta(chip)
        days.append(d)
        cfs.append(c)
    _check_mosaic(dev, stacks, valids, np.stack(days), np.stack(cfs), **syn.CFG5_FILTER)


def test_mosaic_ties_filters_and_none(dev):
    rng = np.random.default_rng(3)
    T, H, W, B = 8, 16, 16, 4
    stacks = [rng.integers(1, 9999, (T, H, W, B), dtype=np.uint16) for _ in range(4)]
    valids = [(rng.random((T, H, W)) > 0.3).astype(np.uint8) for _ in range(4)]
    days = np.array([[10, 20, 20, 30, 30, 40, 40, 50]] * 4, np.int32)       # ties around ref=30 and 25
    cfs = np.array([[0.1, 0.5, 0.2, 0.2, 0.39999, 0.4, 0.1, 0.0]] * 4, np.float32)
    _check_mosaic(dev, stacks, valids, days, cfs, ref_day=25)                 # 20 vs 30 tie -> later index
    _check_mosaic(dev, stacks, valids, days, cfs, ref_day=30, max_cf=0.4)     # strict <
    _check_mosaic(dev, stacks, valids, days, cfs, ref_day=30, min_day=20, max_day=40)   # end exclusive
    _check_mosaic(dev, stacks, valids, days, cfs, ref_day=30, min_day=60)     # nothing eligible -> None
    _check_mosaic(dev, stacks, valids, days, cfs, ref_day=-100000, max_cf=0.05)


def test_reference_entry_points(dev):
    """create_cloudmasked_s2_array / create_img_array_for_tile keep the reference signatures and None rules."""
    import dl_image_segmentation_b200 as pkg
    stack, valid = syn.cfg4_tile(3, T=6, H=20, W=20, B=3)
    dates = [dt.date(2020, 1, 1) + dt.timedelta(days=10 * i) for i in range(6)]
    cf = [0.0, 0.5, 0.1, 0.9, 0.2, 0.3]
    src = pkg.SyntheticSceneSource()
    src.add("tile-a", "sentinel-2:L1C", pkg.SceneStack(stack, valid, dates, cf))
    res = pkg.create_cloudmasked_s2_array("tile-a", dt.date(2020, 1, 11), dt.date(2020, 2, 10), "red green blue",
                                          scene_source=src)
    ref = ocomp.median_composite(stack[1:4], valid[1:4])
    np.testing.assert_array_equal(res.to_masked_array().filled(-1), ref.filled(-1))
    assert pkg.create_cloudmasked_s2_array("tile-a", dt.date(2021, 1, 1), None, scene_source=src) is None
    assert pkg.create_cloudmasked_s2_array("tile-b", scene_source=src) is None
    res = pkg.create_img_array_for_tile("tile-a", "sentinel-2:L1C", dt.date(2020, 1, 25), max_cloud_fraction=0.4,
                                        scene_source=src)
    r_out, r_mask, _ = ocomp.nearest_date_mosaic(stack, valid, [d.toordinal() for d in dates], cf,
                                                 dt.date(2020, 1, 25).toordinal(), None, None, 0.4)
    np.testing.assert_array_equal(res.data.cpu().numpy().view(np.uint16), r_out)
    np.testing.assert_array_equal(res.mask.cpu().numpy()[..., 0], r_mask)
    assert pkg.create_img_array_for_tile("tile-a", "sentinel-2:L1C", dt.date(2020, 1, 25), max_cloud_fraction=0.0,
                                         scene_source=src) is None


@pytest.mark.parametrize("dtype,B", [(np.uint16, 4), (np.uint16, 2), (np.uint8, 4), (np.uint16, 1)])
def test_mosaic_fused_band_statistics(dev, dtype, B):
    """Statistics accumulated inside the mosaic kernel == oracle band_stats over the valid output pixels, and they
    accumulate across calls (configs[4]: one allreduce of these counters gives the dataset mean / std)."""
    import torch

    from dl_image_segmentation_b200 import ops
    from oracle import normalise as onorm
    rng = np.random.default_rng(11)
    n, T, H, W = 5, 9, 24, 20
    stacks = rng.integers(0, np.iinfo(dtype).max + 1, (n, T, H, W, B)).astype(dtype)
    valids = (rng.random((n, T, H, W)) > 0.6).astype(np.uint8)
    valids[0] = 0                                              # a chip without any valid pixel
    days = np.sort(rng.integers(0, 100, (n, T)), axis=1).astype(np.int32)
    cfs = rng.random((n, T)).astype(np.float32)
    acc = torch.zeros((B, 4), dtype=torch.int64, device=dev)
    want = [(0, 0, 0)] * B
    for rep in range(2):
        out, mask, _, nel = ops.nearest_date_mosaic(stacks, valids, days, cfs, 50, 10, 90, 0.7, device=dev, stats_acc=acc)
        o, m = out.cpu().numpy(), mask.cpu().numpy()
        got = onorm.band_stats(o.reshape(-1, B), (~m).reshape(-1).astype(np.uint8))
        want = [(a[0] + b[0], a[1] + b[1], a[2] + b[2]) for a, b in zip(want, got)]
    assert ops.stats_to_python(acc) == want
