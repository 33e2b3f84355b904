"""-m gpu: the encoded-blob parsers and band stacking (A16) at the BASELINE chip shapes, through
the drop-in API, against the oracle's restatement of `_tfrecord_image_translation.py:269-386` and `np.dstack`
(`_descartes_img_chips.py:516`).  The small reference-run fixtures for the same entry points live in
tests/test_reference_parity.py."""
import numpy as np
import pytest
import torch

import synthetic as syn
from oracle import example_proto as oep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def _np(t):
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def _same(got, want):
    g = _np(got)
    assert g.dtype == want.dtype and g.shape == want.shape and np.array_equal(g, want)


def test_gdal_parsers_on_cfg3_lzw_geotiff_records(dev):
    """configs[2] raw-bytes records: 512x512x4 uint16 tiled-LZW GeoTIFF + 512x512 uint8 label blob."""
    import dl_image_segmentation_b200 as pkg
    for i, kw in enumerate((dict(tile=256), dict(tile=None, predictor=2), dict(tile=256, compression="deflate"))):
        img, lab, key = syn.cfg3_chip(40 + i)
        ib, lb = syn.tiff_bytes(img, **kw), syn.tiff_bytes(lab, nodata=255, **kw)
        rec = oep.convert_to_example(ib, lb, 512, 512, 4, 512, 512, key).SerializeToString()
        gi, gt, gid = pkg.parse_encoded_gdal_proto_eager(rec)
        wi, wt, wid = oep.parse_encoded_gdal_proto_eager(rec)
        assert gi.dtype == torch.uint16 and tuple(gt.shape) == (512, 512, 1) and gid == wid == key.encode()
        _same(gi, wi)
        _same(gt, wt)
        assert np.array_equal(wi, img) and np.array_equal(wt[..., 0], lab)
        fi, ft, fid = pkg.parse_encoded_gdal_proto_wrapped(rec)
        wfi, wft, _ = oep.parse_encoded_gdal_proto_wrapped(rec)
        assert fi.dtype == torch.float32 and ft.dtype == torch.float32 and fid == wid
        _same(fi, wfi)
        _same(ft, wft)


def test_eager_parser_checks_the_recorded_shape_and_wrapped_does_not(dev):
    """`_eager` asserts decoded shape == recorded shape (`:377,383-384`); `_wrapped` has no such check (`:332-346`)."""
    import dl_image_segmentation_b200 as pkg
    img, lab, key = syn.cfg3_chip(3, size=64)
    ib, lb = syn.tiff_bytes(img, tile=32), syn.tiff_bytes(lab, tile=32)
    rec = oep.convert_to_example(ib, lb, 64, 60, 4, 64, 64, key).SerializeToString()       # width recorded wrongly
    with pytest.raises(AssertionError):
        pkg.parse_encoded_gdal_proto_eager(rec)
    with pytest.raises(AssertionError):
        oep.parse_encoded_gdal_proto_eager(rec)
    fi, ft, _ = pkg.parse_encoded_gdal_proto_wrapped(rec)
    _same(fi, img.astype(np.float32))
    _same(ft, lab[..., None].astype(np.float32))


def test_rgb_and_gdal_parsers_on_cfg1_png_records(dev):
    """configs[0] raw-bytes records (threaded translator, store_as_array=False): PNG blobs, label comes back (H,W,1)."""
    import dl_image_segmentation_b200 as pkg
    for i in range(3):
        img, lab, key = syn.cfg1_chip(70 + i)
        rec = oep.convert_to_example(syn.png_bytes(img), syn.png_bytes(lab), 256, 256, 3, 256, 256, key).SerializeToString()
        for parser in ("parse_encoded_rgb_img_proto", "parse_encoded_gdal_proto_eager", "parse_encoded_gdal_proto_wrapped"):
            gi, gt, gid = getattr(pkg, parser)(rec)
            wi, wt, wid = getattr(oep, parser)(rec)
            assert tuple(gi.shape) == (256, 256, 3) and tuple(gt.shape) == (256, 256, 1) and gid == wid
            _same(gi, wi)
            _same(gt, wt)
        assert np.array_equal(wi, img.astype(np.float32)) and np.array_equal(wt[..., 0], lab.astype(np.float32))


def test_rgb_parser_and_gdal_parser_disagree_on_palette_pngs_as_the_libraries_do(dev):
    """tf.io.decode_image expands a palette to RGB; GDAL presents one band of indices: each parser keeps the semantics of
    the call it replaces."""
    import dl_image_segmentation_b200 as pkg
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 7, (24, 20)).astype(np.uint8)
    pal = rng.integers(0, 256, (7, 3)).astype(np.uint8)
    blob = syn.png_bytes_flavour(idx, 8, 3, palette=pal)
    lab = syn.png_bytes(idx)
    rec = oep.convert_to_example(blob, lab, 24, 20, 3, 24, 20, "p").SerializeToString()
    gi, _, _ = pkg.parse_encoded_rgb_img_proto(rec)
    _same(gi, pal[idx])
    rec1 = oep.convert_to_example(blob, lab, 24, 20, 1, 24, 20, "p").SerializeToString()
    gi, _, _ = pkg.parse_encoded_gdal_proto_eager(rec1)
    _same(gi, idx[..., None])


def test_stack_products_matches_numpy_dstack_for_mixed_dtypes(dev):
    """A16: per-product overlay mosaic, then np.dstack with NumPy's promotion (u16+u8 -> u16, u16+i16 -> i32, +f32 -> f64 / f32)."""
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _descartes_img_chips as dc
    from oracle import composite as ocomp
    rng = np.random.default_rng(77)
    H, W = 256, 256
    specs = {"a": (np.uint16, 3), "b": (np.uint8, 1), "c": (np.int16, 2), "d": (np.float32, 2)}
    src = dc.SyntheticSceneSource()
    want = {}
    for name, (dt_, nb) in specs.items():
        T = 3
        if np.issubdtype(dt_, np.integer):
            st = rng.integers(np.iinfo(dt_).min, np.iinfo(dt_).max, (T, H, W, nb), endpoint=True).astype(dt_)
        else:
            st = rng.normal(size=(T, H, W, nb)).astype(dt_)
        va = (rng.random((T, H, W)) > 0.4).astype(np.uint8)
        src.add("ctx", name, dc.SceneStack(st, va, [0] * T))
        want[name] = ocomp.nearest_date_mosaic(st, va, [0] * T, [0.0] * T, 0)[0]
    for combo in (("a", "b"), ("a", "c"), ("b", "c"), ("a", "b", "c"), ("a", "d"), ("c", "d"), ("a",)):
        got = pkg.stack_products_for_tile("ctx", list(combo), ["x"] * len(combo), scene_source=src)
        ref = np.dstack([want[n] for n in combo])
        g = got.view(torch.int32).cpu().numpy().view(np.uint32) if got.dtype == torch.uint32 else _np(got)
        assert g.dtype == ref.dtype, (combo, g.dtype, ref.dtype)
        assert np.array_equal(g, ref), combo


def test_uint16_mosaic_filled_and_noncontiguous_date_filter(dev):
    """`filled()` on a uint16 mosaic and a search filter that keeps a non-contiguous set of scenes."""
    import datetime as dt

    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _descartes_img_chips as dc
    from oracle import composite as ocomp
    rng = np.random.default_rng(78)
    T, H, W, B = 6, 40, 36, 3
    st = rng.integers(0, 65536, (T, H, W, B)).astype(np.uint16)
    va = (rng.random((T, H, W)) > 0.5).astype(np.uint8)
    va[:, :3, :3] = 0
    days = [dt.date(2020, 1, 1) + dt.timedelta(days=10 * t) for t in range(T)]
    res = pkg.nearest_date_mosaic(st, va, days, None, dt.date(2020, 1, 25))
    out, mask, _ = ocomp.nearest_date_mosaic(st, va, [d.toordinal() for d in days], [0.0] * T, dt.date(2020, 1, 25).toordinal())
    filled = res.filled(65535)
    assert filled.dtype == torch.uint16
    want = out.copy()
    want[mask] = 65535
    _same(filled, want)
    # median over scenes {0, 2, 3, 5}: dates out of order in the catalogue so that the window keeps a non-contiguous set
    shuffled = [days[0], days[5], days[1], days[2], days[4], days[3]]
    src = dc.SyntheticSceneSource()
    src.add("c", "sentinel-2:L1C", dc.SceneStack(st, va, shuffled))
    lo, hi = dt.date(2020, 1, 1), dt.date(2020, 2, 5)                     # keeps days[0..3] -> catalogue indices 0, 2, 3, 5
    got = pkg.create_cloudmasked_s2_array("c", min_date=lo, max_date=hi, scene_source=src).to_masked_array()
    ref = ocomp.create_cloudmasked_s2_array(shuffled, st, va, None, lo, hi)
    assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(ref)) and np.array_equal(got.filled(0), ref.filled(0))


def _shard_of(records):
    from oracle import tfrecord as otfr
    return b"".join(otfr.frame(r) for r in records)


@pytest.mark.parametrize("parser", ["gdal_eager", "gdal_wrapped", "rgb"])
def test_parse_encoded_shard_equals_the_per_record_parsers(dev, parser):
    """The whole-shard form (one scan + index, one CRC pass, one batched decode) returns, record for record, what the
    per-record parse function returns — and that is checked against the oracle's restatement of the reference parser."""
    import dl_image_segmentation_b200 as pkg
    recs = []
    for i in range(7):
        if parser == "rgb":
            img, lab, key = syn.cfg1_chip(i, size=64 + 8 * (i % 2))          # two shapes in one shard
            ib, lb = syn.png_bytes(img), syn.png_bytes(lab)
            h, w, c = img.shape
        else:
            img, lab, key = syn.cfg3_chip(i, size=96)
            kw = (dict(tile=32), dict(tile=None, predictor=2), dict(tile=32, compression="deflate"))[i % 3]
            ib, lb = syn.tiff_bytes(img, **kw), syn.tiff_bytes(lab, nodata=255, **kw)
            h, w, c = img.shape
        recs.append(oep.convert_to_example(ib, lb, h, w, c, h, w, key).SerializeToString())
    one = getattr(pkg, {"gdal_eager": "parse_encoded_gdal_proto_eager", "gdal_wrapped": "parse_encoded_gdal_proto_wrapped",
                        "rgb": "parse_encoded_rgb_img_proto"}[parser])
    ref = getattr(oep, one.__name__)
    got = pkg.parse_encoded_shard(_shard_of(recs), parser=parser)
    assert len(got) == len(recs)
    for rec, (gi, gt, gid) in zip(recs, got):
        si, st, sid = one(rec)
        wi, wt, wid = ref(rec)
        assert gid == sid == wid
        assert gi.dtype == si.dtype and gt.dtype == st.dtype
        _same(gi, wi)
        _same(gt, wt)
        _same(si, wi)


def test_parse_encoded_shard_errors_are_the_per_record_ones(dev, tmp_path):
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import ops
    from dl_image_segmentation_b200._tfrecord_image_translation import InvalidArgumentError
    img, lab, key = syn.cfg3_chip(1, size=64)
    ib, lb = syn.tiff_bytes(img, tile=32), syn.tiff_bytes(lab, tile=32)
    good = oep.convert_to_example(ib, lb, 64, 64, 4, 64, 64, key).SerializeToString()
    shard = bytearray(_shard_of([good, good]))
    assert pkg.parse_encoded_shard(b"") == []
    path = tmp_path / "s"
    path.write_bytes(bytes(shard))
    assert len(pkg.parse_encoded_shard(str(path))) == 2                       # a path works like the bytes
    shard[len(shard) // 2 + 40] ^= 1                                          # inside the second record's payload
    with pytest.raises(ops.DataLossError):
        pkg.parse_encoded_shard(bytes(shard))
    broken = oep.convert_to_example(ib[:len(ib) // 2], lb, 64, 64, 4, 64, 64, key).SerializeToString()   # truncated TIFF
    with pytest.raises(InvalidArgumentError):
        pkg.parse_encoded_shard(_shard_of([good, broken]))
    wrong = oep.convert_to_example(ib, lb, 64, 60, 4, 64, 64, key).SerializeToString()
    with pytest.raises(AssertionError):
        pkg.parse_encoded_shard(_shard_of([wrong]), parser="gdal_eager")
    assert len(pkg.parse_encoded_shard(_shard_of([wrong]), parser="gdal_wrapped")) == 1
    arr = oep.convert_to_example(img[..., :3].astype(np.uint8), lab, 64, 64, 3, 64, 64, key).SerializeToString()
    with pytest.raises(InvalidArgumentError):                                 # FloatList / array records do not fit the template
        pkg.parse_encoded_shard(_shard_of([oep.convert_to_example(img.astype(np.float32), lab.astype(np.float32), 64, 64, 4, 64, 64,
                                                                  key).SerializeToString()]))
    del arr


def test_iter_parse_encoded_shards_prefetches_and_equals_shard_by_shard(dev, tmp_path):
    """The generator form reads shard k+1 into the other pinned buffer while shard k is decoded: same results as calling
    parse_encoded_shard on each, for paths and bytes, for more shards than staging buffers, and with an empty shard between."""
    import dl_image_segmentation_b200 as pkg
    shards = []
    for s in range(5):
        recs = []
        for i in range(3 + s % 2):
            img, lab, key = syn.cfg3_chip(10 * s + i, size=64)
            ib, lb = syn.tiff_bytes(img, tile=32), syn.tiff_bytes(lab, tile=32, nodata=255)
            recs.append(oep.convert_to_example(ib, lb, 64, 64, 4, 64, 64, key).SerializeToString())
        shards.append(_shard_of(recs))
    shards.insert(2, b"")
    paths = []
    for k, b in enumerate(shards):
        p = tmp_path / ("s%d" % k)
        p.write_bytes(b)
        paths.append(str(p))
    want = [pkg.parse_encoded_shard(b, parser="gdal_wrapped") for b in shards]
    for source in (shards, paths):
        got = list(pkg.iter_parse_encoded_shards(source, parser="gdal_wrapped"))
        assert [len(g) for g in got] == [len(w) for w in want]
        for g, w in zip(got, want):
            for (gi, gt, gid), (wi, wt, wid) in zip(g, w):
                assert gid == wid and torch.equal(gi, wi) and torch.equal(gt, wt)
    assert list(pkg.iter_parse_encoded_shards([])) == []


def test_parse_encoded_shard_mixed_blob_formats(dev):
    """parser='rgb' stands for tf.io.decode_image: PNG of any flavour and JPEG blobs in one shard.  A palette PNG makes the
    in-place planner give way to the gathering one; the .jpg blobs are the ones the TIFF / PNG planner refuses and go
    through the JPEG path; every record must equal the per-record parser's result."""
    import cv2
    import dl_image_segmentation_b200 as pkg
    rng = np.random.default_rng(3)
    recs = []
    for i in range(6):
        img, lab, key = syn.cfg1_chip(i, size=64)
        if i % 3 == 0:                                                       # palette image, 2-bit grey label
            idx = rng.integers(0, 16, (64, 64))
            ib = syn.png_bytes_flavour(idx, 8, 3, palette=rng.integers(0, 256, (16, 3)))
            lb = syn.png_bytes_flavour(rng.integers(0, 4, (64, 64)), 2, 0)
        elif i % 3 == 1:                                                     # baseline JPEG image, PNG label
            ib = cv2.imencode(".jpg", np.ascontiguousarray(img[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes()
            lb = syn.png_bytes(lab)
        else:
            ib, lb = syn.png_bytes(img), syn.png_bytes(lab)
        recs.append(oep.convert_to_example(ib, lb, 64, 64, 3, 64, 64, key).SerializeToString())
    got = pkg.parse_encoded_shard(_shard_of(recs), parser="rgb")
    assert len(got) == len(recs)
    for rec, (gi, gt, gid) in zip(recs, got):
        si, st, sid = pkg.parse_encoded_rgb_img_proto(rec)
        assert gid == sid and gi.dtype == si.dtype and gt.dtype == st.dtype
        assert tuple(gi.shape) == tuple(si.shape) and tuple(gt.shape) == tuple(st.shape)
        assert torch.equal(gi, si) and torch.equal(gt, st)
    # and without the palette record the in-place plan is kept: same answers
    plain = [r for k, r in enumerate(recs) if k % 3 != 0]
    for rec, (gi, gt, gid) in zip(plain, pkg.parse_encoded_shard(_shard_of(plain), parser="rgb")):
        si, st, sid = pkg.parse_encoded_rgb_img_proto(rec)
        assert gid == sid and torch.equal(gi, si) and torch.equal(gt, st)
