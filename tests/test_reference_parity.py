"""Parity pinned on the reference's OWN code.

`tests/golden/ref_*` were written by running the unmodified `/root/reference/dl_segmentation_utils` under the stub
dependencies of `oracle/refstubs` (`tests/golden/gen_golden_reference.py`).  Here:

  * CPU (`-m "not gpu"`): `oracle/` reproduces every fixture byte for byte / value for value, and — wherever
    `/root/reference` exists (the build container) — the reference is run again, live, on the committed chip folders and
    on fresh random inputs, and must agree with both the committed fixtures and `oracle/`.
  * GPU (`-m gpu`): the product (drop-in Python API -> C ABI -> CUDA) reproduces the same fixtures.
"""
import datetime as dt
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import composite as ocomp
from oracle import example_proto as oep
from oracle import refrun
from oracle import tfrecord as otfr
from oracle import translate as otr

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
META = json.load(open(os.path.join(GOLD, "ref_meta.json")))
needs_reference = pytest.mark.skipif(not refrun.available(), reason="/root/reference is not on this machine")

MP_RUNS = ["ref_mp_tif_arrays", "ref_mp_tif_raw_georef", "ref_mp_png_arrays"]
MT_RUNS = ["ref_mt_png_raw", "ref_mt_png_arrays", "ref_mt_png_to_jpg"]
PARSER_RUNS = [k.split("|") for k in sorted(META) if k.startswith("parse_") and k.count("|") == 1]


def _gen():
    spec = importlib.util.spec_from_file_location("gen_golden_reference", os.path.join(GOLD, "gen_golden_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _chips(run):
    return os.path.join(GOLD, "ref_chips_tif" if "tif" in run else "ref_chips_png")


def _shards(d):
    return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))}


def _records(run):
    out = []
    for f, blob in _shards(os.path.join(GOLD, run)).items():
        out += otfr.read_records(blob, True)
    return out


def _mp_kwargs(run):
    m = META[run]
    return dict(file_ext=m["file_ext"], store_as_array=m["store_as_array"], dltile_from_filename=m.get("dltile_from_filename", True))


def _mt_kwargs(run):
    m = META[run]
    return dict(store_as_array=m["store_as_array"], convert_png_to_jpg=m.get("convert_png_to_jpg", False))


def _dates():
    return [dt.date.fromisoformat(s) for s in META["composite_dates"]]


def _d(s):
    return None if s is None else dt.date.fromisoformat(s)


COMP = np.load(os.path.join(GOLD, "ref_composites.npz"))


# ================================================================================================ CPU: oracle == reference
@pytest.mark.parametrize("run", MP_RUNS)
def test_oracle_mp_translator_reproduces_reference_shards(run, tmp_path):
    m = META[run]
    n = otr.images_to_tfrecords("chips", _chips(run), str(tmp_path), m["num_shards"], m["num_proc"], n_jobs=1, **_mp_kwargs(run))
    want = _shards(os.path.join(GOLD, run))
    assert _shards(str(tmp_path)) == want
    assert n == sum(len(otfr.read_records(b, True)) for b in want.values())
    found = len([f for f in os.listdir(os.path.join(_chips(run), "images")) if f.endswith("." + m["file_ext"])])
    assert m["skipped"] == found - n                               # the reference printed SKIPPED for exactly the others


@pytest.mark.parametrize("run", MT_RUNS)
def test_oracle_mt_translator_reproduces_reference_shards(run, tmp_path):
    m = META[run]
    n = otr.images_to_tfrecords_mt("chips", _chips(run), str(tmp_path), m["num_shards"], m["num_threads"], **_mt_kwargs(run))
    assert _shards(str(tmp_path)) == _shards(os.path.join(GOLD, run))
    assert m["skipped"] == 6 - n == 1                               # the RGBA chip: more than 3 bands


@pytest.mark.parametrize("parser,run", PARSER_RUNS)
def test_oracle_parsers_reproduce_reference_outputs(parser, run):
    parsed = np.load(os.path.join(GOLD, "ref_parsed.npz"))
    recs = _records(run)
    assert len(recs) == META["%s|%s" % (parser, run)]
    for k, rec in enumerate(recs):
        key = "%s|%s|%d" % (parser, run, k)
        if key in META.get("parse_errors", {}):
            with pytest.raises(Exception):
                getattr(oep, parser)(rec)
            continue
        img, tgt, ident = getattr(oep, parser)(rec)
        for got, name in ((img, "img"), (tgt, "tgt")):
            want = parsed[key + "|" + name]
            assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), key
        assert ident == parsed[key + "|id"].tobytes()


def test_oracle_convert_to_example_reproduces_reference_bytes():
    arrs = np.load(os.path.join(GOLD, "ref_convert_inputs.npz"))
    n = 0
    for name, hexs in META["convert_to_example"].items():
        if name == "bytes+bytes":
            got = oep.convert_to_example(b"\x89PNG fake image", b"\x89PNG fake label", 7, 8, 3, 7, 8, "key/with|bar")
        else:
            iname, lname = name.split("+")
            img, lab = arrs[iname], arrs[lname]
            ident = _convert_identifier(name)
            got = oep.convert_to_example(img, lab, 5, 6, img.shape[2], 5, 6, ident)
        assert got.SerializeToString(deterministic=True).hex() == hexs, name
        n += 1
    assert n == 8


CONVERT_ORDER = ["u8x3+lab8", "u8x3+lab8_3d", "u16x4+lab8", "i16x2+lab8", "f32x1+lab8", "f64x2+lab16", "u8x3+lab16"]


def _convert_identifier(name):
    return "64:0:10.0:43:-3:%d" % CONVERT_ORDER.index(name)              # as gen_golden_reference.py numbered them


def test_oracle_compositors_reproduce_reference_outputs():
    dates = _dates()
    stack, nodata, cloudfree, cf = COMP["stack"], COMP["nodata"], COMP["cloudfree"], COMP["cf"]
    nd4 = np.repeat(nodata[..., None], stack.shape[-1], -1)
    for tag in ("all", "window", "empty"):
        m = META["median_%s" % tag]
        lo, hi = (None, None) if m is None and tag != "empty" else ((_d(m[0]), _d(m[1])) if m else (dt.date(2021, 1, 1), None))
        res = ocomp.create_cloudmasked_s2_array(dates, stack, cloudfree, nd4, lo, hi)
        if m is None:
            assert res is None
            continue
        assert res.dtype == np.float64
        assert np.array_equal(np.ma.getmaskarray(res), COMP["median_%s_mask" % tag])
        assert np.array_equal(res.filled(0), COMP["median_%s_data" % tag])
    ref_day = _d(META["mosaic_reference_date"]).toordinal()
    days = [d.toordinal() for d in dates]
    for tag in ("plain", "cloud", "window", "both", "none"):
        m = META["mosaic_%s" % tag]
        kw = m if m is not None else dict(max_cloud_fraction=0.0)
        res = ocomp.nearest_date_mosaic(stack, (~nodata).astype(np.uint8), days, cf, ref_day,
                                        None if "min_date" not in kw else _d(kw["min_date"]).toordinal(),
                                        None if "max_date" not in kw else _d(kw["max_date"]).toordinal(), kw.get("max_cloud_fraction"))
        if m is None:
            assert res is None
            continue
        out, mask, _ = res
        assert out.dtype == COMP["mosaic_%s_data" % tag].dtype
        assert np.array_equal(np.repeat(mask[..., None], out.shape[-1], -1), COMP["mosaic_%s_mask" % tag])
        assert np.array_equal(out, COMP["mosaic_%s_data" % tag])
    for tag, which in (("dstack_all", (0, 1, 2)), ("dstack_u16_u8", (0, 1)), ("dstack_u16_i16", (0, 2))):
        arrays = []
        for p in which:
            st, nd = COMP["prod%d_stack" % p], COMP["prod%d_nodata" % p]
            T = st.shape[0]
            out, _, _ = ocomp.nearest_date_mosaic(st, (~nd).astype(np.uint8), [0] * T, [0.0] * T, 0)   # plain overlay: last wins
            arrays.append(out)
        got = ocomp.stack_products(arrays)
        assert str(got.dtype) == META["dstack_dtypes"][tag] and np.array_equal(got, COMP[tag]), tag


# ================================================================================================ CPU: the reference, live
@needs_reference
@pytest.mark.parametrize("run", MP_RUNS + MT_RUNS)
def test_committed_shards_are_what_the_reference_writes(run, tmp_path):
    g = _gen()
    m = META[run]
    if run in MP_RUNS:
        g.ref_images_to_tfrecords_mp(_chips(run), str(tmp_path), "chips", m["num_shards"], m["num_proc"], **_mp_kwargs(run))
    else:
        g.ref_images_to_tfrecords_mt(_chips(run), str(tmp_path), "chips", m["num_shards"], m["num_threads"], **_mt_kwargs(run))
    assert _shards(str(tmp_path)) == _shards(os.path.join(GOLD, run))


@needs_reference
def test_live_reference_vs_oracle_convert_and_parse_random():
    """Random dtypes / shapes through the reference's convert_to_example and back through its parsers, against oracle/."""
    g = _gen()
    rng = np.random.default_rng(424242)
    for it in range(40):
        h, w, c = int(rng.integers(1, 9)), int(rng.integers(1, 9)), int(rng.integers(1, 6))
        idt = [np.uint8, np.uint16, np.int16, np.float32, np.int32][int(rng.integers(0, 5))]
        ldt = [np.uint8, np.uint8, np.uint16][int(rng.integers(0, 3))]
        img = (rng.random((h, w, c)) * 250).astype(idt)
        lab = rng.integers(0, 11, (h, w)).astype(ldt)
        ident = "k%d:%d" % (it, h)
        want = g.ref_convert_to_example(img, lab, h, w, c, h, w, ident)
        got = oep.convert_to_example(img, lab, h, w, c, h, w, ident).SerializeToString(deterministic=True)
        assert got == want, (it, idt, ldt)
        parser = "parse_8bit_array_proto" if (idt == np.uint8 and ldt == np.uint8) else "parse_higher_dtype_array_proto"
        ri, rt, rid = g.ref_parse(parser, want)
        oi, ot, oid = getattr(oep, parser)(want)
        assert ri.dtype == oi.dtype and np.array_equal(ri, oi) and rt.dtype == ot.dtype and np.array_equal(rt, ot) and rid == oid
    # a record without its image payload, a FloatList where bytes are expected: both sides refuse
    feats = oep.convert_to_example(np.zeros((2, 2, 1), np.uint16), np.zeros((2, 2), np.uint8), 2, 2, 1, 2, 2, "x")
    rec = feats.SerializeToString()
    with pytest.raises(Exception):
        g.ref_parse("parse_8bit_array_proto", rec)
    with pytest.raises(Exception):
        oep.parse_8bit_array_proto(rec)


@needs_reference
def test_live_reference_vs_oracle_compositors_random():
    g = _gen()
    for seed in range(25):
        rng = np.random.default_rng(9000 + seed)
        T = int(rng.integers(1, 9))
        dates, cf, stack, nodata, cloudfree = g.make_catalog(rng, T=T, H=6, W=5, tie_days=bool(seed % 2))
        nd4 = np.repeat(nodata[..., None], 3, -1)
        lo = None if seed % 3 == 0 else dt.date(2020, 1, 1) + dt.timedelta(days=int(rng.integers(0, 60)))
        hi = None if seed % 4 == 0 else dt.date(2020, 1, 1) + dt.timedelta(days=int(rng.integers(40, 130)))
        r = g.ref_cloudmasked(dates, cf, stack, nodata, cloudfree, lo, hi)
        o = ocomp.create_cloudmasked_s2_array(dates, stack, cloudfree, nd4, lo, hi)
        assert (r is None) == (o is None)
        if r is not None:
            assert np.array_equal(o.filled(0), r[0]) and np.array_equal(np.ma.getmaskarray(o), r[1])
        refd = dt.date(2020, 1, 1) + dt.timedelta(days=int(rng.integers(0, 120)))
        mcf = None if seed % 5 == 0 else float(rng.random())
        r = g.ref_img_array(dates, cf, stack, nodata, refd, lo, hi, mcf)
        o = ocomp.nearest_date_mosaic(stack, (~nodata).astype(np.uint8), [d.toordinal() for d in dates], cf, refd.toordinal(),
                                      None if lo is None else lo.toordinal(), None if hi is None else hi.toordinal(), mcf)
        assert (r is None) == (o is None), seed
        if r is not None:
            assert np.array_equal(o[0], r[0]) and np.array_equal(np.repeat(o[1][..., None], 3, -1), r[1]), seed


@needs_reference
def test_live_reference_partition_matches_oracle_for_awkward_sizes(tmp_path):
    """The nested linspace of the reference's worker / shard ranges (`_img_to_tf_mp.py:102-108,167-170`) on folder sizes
    that do not divide: shard files written by the reference and by the oracle hold the same records."""
    g = _gen()
    src = tmp_path / "chips"
    g.write_png_folder(str(src), n=13, size=8, seed=7300, with_jpg=False, with_rgba=False)
    for shards, procs in ((6, 3), (4, 1), (6, 6)):
        a, b = tmp_path / ("ref%d_%d" % (shards, procs)), tmp_path / ("orc%d_%d" % (shards, procs))
        g.ref_images_to_tfrecords_mp(str(src), str(a), "p", shards, procs, file_ext="png", store_as_array=True)
        otr.images_to_tfrecords("p", str(src), str(b), shards, procs, file_ext="png", store_as_array=True, n_jobs=1)
        assert _shards(str(a)) == _shards(str(b))


# ================================================================================================ GPU: product == reference
@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def _np(t):
    import torch
    if isinstance(t, np.ndarray):
        return t
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    if t.dtype == torch.uint32:
        return t.view(torch.int32).cpu().numpy().view(np.uint32)
    return t.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("run", MP_RUNS)
def test_gpu_mp_translator_reproduces_reference_shards(dev, run, tmp_path, capsys):
    import dl_image_segmentation_b200 as pkg
    m = META[run]
    pkg.images_to_tfrecords_mp("chips", _chips(run), str(tmp_path), m["num_shards"], m["num_proc"], **_mp_kwargs(run))
    assert _shards(str(tmp_path)) == _shards(os.path.join(GOLD, run))
    assert capsys.readouterr().out.count("SKIPPED: Unexpected eror while decoding") == m["skipped"]


@pytest.mark.gpu
@pytest.mark.parametrize("run", MT_RUNS)
def test_gpu_mt_translator_reproduces_reference_shards(dev, run, tmp_path, capsys):
    import dl_image_segmentation_b200 as pkg
    m = META[run]
    pkg.images_to_tfrecords_mt("chips", _chips(run), str(tmp_path), m["num_shards"], m["num_threads"], **_mt_kwargs(run))
    assert _shards(str(tmp_path)) == _shards(os.path.join(GOLD, run))
    assert capsys.readouterr().out.count("SKIPPED: Unexpected eror while decoding") == m["skipped"]


@pytest.mark.gpu
@pytest.mark.parametrize("parser,run", PARSER_RUNS)
def test_gpu_parsers_reproduce_reference_outputs(dev, parser, run):
    import dl_image_segmentation_b200 as pkg
    parsed = np.load(os.path.join(GOLD, "ref_parsed.npz"))
    for k, rec in enumerate(_records(run)):
        key = "%s|%s|%d" % (parser, run, k)
        if key in META.get("parse_errors", {}):
            with pytest.raises(Exception):
                getattr(pkg, parser)(rec)
            continue
        img, tgt, ident = getattr(pkg, parser)(rec)
        assert img.is_cuda and tgt.is_cuda
        for got, name in ((_np(img), "img"), (_np(tgt), "tgt")):
            want = parsed[key + "|" + name]
            assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), key
        assert ident == parsed[key + "|id"].tobytes()


@pytest.mark.gpu
def test_gpu_convert_to_example_reproduces_reference_bytes(dev):
    import dl_image_segmentation_b200 as pkg
    arrs = np.load(os.path.join(GOLD, "ref_convert_inputs.npz"))
    for name, hexs in META["convert_to_example"].items():
        if name == "bytes+bytes":
            got = pkg.convert_to_example(b"\x89PNG fake image", b"\x89PNG fake label", 7, 8, 3, 7, 8, "key/with|bar")
        else:
            iname, lname = name.split("+")
            img, lab = arrs[iname], arrs[lname]
            got = pkg.convert_to_example(img, lab, 5, 6, img.shape[2], 5, 6, _convert_identifier(name))
        assert got.SerializeToString().hex() == hexs, name


@pytest.mark.gpu
def test_gpu_compositors_reproduce_reference_outputs(dev):
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _descartes_img_chips as dc
    dates = _dates()
    stack, nodata, cloudfree, cf = COMP["stack"], COMP["nodata"], COMP["cloudfree"], COMP["cf"]
    nd4 = np.repeat(nodata[..., None], stack.shape[-1], -1).astype(np.uint8)
    src = dc.SyntheticSceneSource()
    ctx = "64:0:10.0:43:1:2"
    src.add(ctx, "sentinel-2:L1C", dc.SceneStack(stack, cloudfree, dates, cf, nodata_mask=nd4))
    src.add(ctx, "some:product", dc.SceneStack(stack, (~nodata).astype(np.uint8), dates, cf))
    for tag in ("all", "window", "empty"):
        m = META["median_%s" % tag]
        lo, hi = (_d(m[0]), _d(m[1])) if m else (dt.date(2021, 1, 1), None)
        res = pkg.create_cloudmasked_s2_array(ctx, min_date=lo, max_date=hi, scene_source=src)
        if m is None:
            assert res is None
            continue
        ma = res.to_masked_array()
        assert ma.dtype == np.float64
        assert np.array_equal(np.ma.getmaskarray(ma), COMP["median_%s_mask" % tag])
        assert np.array_equal(ma.filled(0), COMP["median_%s_data" % tag])
    refd = _d(META["mosaic_reference_date"])
    for tag in ("plain", "cloud", "window", "both", "none"):
        m = META["mosaic_%s" % tag]
        kw = m if m is not None else dict(max_cloud_fraction=0.0)
        res = pkg.create_img_array_for_tile(ctx, "some:product", refd, min_date=_d(kw.get("min_date")), max_date=_d(kw.get("max_date")),
                                            max_cloud_fraction=kw.get("max_cloud_fraction"), scene_source=src)
        if m is None:
            assert res is None
            continue
        assert np.array_equal(_np(res.mask), COMP["mosaic_%s_mask" % tag])
        got = _np(res.filled(0))
        assert got.dtype == COMP["mosaic_%s_data" % tag].dtype and np.array_equal(got, COMP["mosaic_%s_data" % tag])
    for p in range(3):
        st, nd = COMP["prod%d_stack" % p], COMP["prod%d_nodata" % p]
        src.add(ctx, "prod%d" % p, dc.SceneStack(st, (~nd).astype(np.uint8), [_d(s) for s in META["prod%d" % p]["dates"]]))
    for tag, which in (("dstack_all", (0, 1, 2)), ("dstack_u16_u8", (0, 1)), ("dstack_u16_i16", (0, 2))):
        got = _np(pkg.stack_products_for_tile(ctx, ["prod%d" % p for p in which], [META["prod%d" % p]["bands"] for p in which],
                                              scene_source=src))
        assert str(got.dtype) == META["dstack_dtypes"][tag] and np.array_equal(got, COMP[tag]), tag
    with pytest.raises(ValueError):                                   # a product without scenes: mosaic() of an empty collection
        pkg.stack_products_for_tile(ctx, ["prod0", "nothing:here"], ["red green blue", "x"], scene_source=src)


# ================================================================================================ create_chips_for_tile
CHIPS = np.load(os.path.join(GOLD, "ref_chips_for_tile.npz"))
GDAL_NP = {1: np.uint8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.float32, 7: np.float64}


def _chips_layer():
    return [([CHIPS["layer_%d_%d" % (k, j)] for j in range(nr)], {"cls": c})
            for k, (nr, c) in enumerate(zip(META["chips_layer_rings"], META["chips_layer_cls"]))]


def _chips_dates(which):
    return [dt.date.fromisoformat(s) for s in META["chips_%s_dates" % which]]


def test_oracle_pieces_reproduce_the_reference_create_chips_for_tile():
    """The arrays the unmodified reference handed to GDAL band by band (recording stub) for its three dispatch modes."""
    from oracle import rasterize as orr
    gt = META["chips_jobs"]["median"]["img"]["geotransform"]
    layer = _chips_layer()
    st, nd, cfree, cf = CHIPS["s2_stack"], CHIPS["s2_nodata"], CHIPS["s2_cloudfree"], CHIPS["s2_cf"]
    dates = _chips_dates("s2")
    med = ocomp.create_cloudmasked_s2_array(dates, st, cfree, np.repeat(nd[..., None], 3, -1))
    assert CHIPS["median_img"].dtype == np.float64 and np.array_equal(np.asarray(med.data), CHIPS["median_img"])
    days = [d.toordinal() for d in dates]
    mos = ocomp.nearest_date_mosaic(st, (~nd).astype(np.uint8), days, cf, dt.date(2020, 3, 1).toordinal(),
                                    dt.date(2020, 1, 10).toordinal(), None, 0.6)[0]
    assert np.array_equal(mos, CHIPS["mosaic_img"]) and CHIPS["mosaic_img"].dtype == np.uint16
    a = ocomp.nearest_date_mosaic(st, (~nd).astype(np.uint8), [0] * len(days), [0.0] * len(days), 0)[0]
    cst, cnd = CHIPS["cls_stack"], CHIPS["cls_nodata"]
    b = ocomp.nearest_date_mosaic(cst, (~cnd).astype(np.uint8), [0] * len(cst), [0.0] * len(cst), 0)[0]
    assert np.array_equal(ocomp.stack_products([a, b]), CHIPS["stack_img"])
    for kind, attr in (("median", "cls"), ("stack", "cls"), ("mosaic", None)):
        lab = orr.create_label_array_for_tile(12, 2, gt, layer, attr, 255)
        assert np.array_equal(lab, CHIPS[kind + "_lbl"][..., 0]), kind
        rec = META["chips_jobs"][kind]
        assert rec["lbl"]["nodata"] == [255] and rec["lbl"]["gdal_type"] == 1 and rec["img"]["file"] == "images/12#2#10.0#43#7#11.tif"
        assert rec["img"]["options"] == ["COMPRESS=LZW", "TILED=TRUE", "NUM_THREADS=4"]


@needs_reference
def test_committed_chip_fixtures_are_what_the_reference_writes(tmp_path):
    g = _gen()
    tile = g.Tile()
    scenes = {"median": {"sentinel-2:L1C": ("red green blue", _chips_dates("s2"), CHIPS["s2_cf"], CHIPS["s2_stack"], CHIPS["s2_nodata"],
                                            "sentinel-2:L1C:dlcloud:v1", CHIPS["s2_cloudfree"])}}
    (img_file, lbl_file), created = g.ref_create_chips("median", tile, str(tmp_path), _chips_layer(), scenes["median"])
    assert np.array_equal(np.transpose(created[img_file].data, (1, 2, 0)), CHIPS["median_img"])
    assert np.array_equal(created[lbl_file].data[0], CHIPS["median_lbl"][..., 0])


@pytest.mark.gpu
def test_gpu_create_chips_for_tile_reproduces_the_reference(dev, tmp_path):
    """Drop-in create_chips_for_tile in the reference's three dispatch modes: the GeoTIFFs it writes hold exactly the arrays
    the reference handed to GDAL (dtype per `_numpy_dtype_to_gdal`, tiled LZW, label nodata tag, geotransform, file names)."""
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _descartes_img_chips as dc
    from oracle import imagecodecs as oic

    class Tile:
        key, tilesize, pad, epsg = "12:2:10.0:43:7:11", 12, 2, 32643
        geotrans = tuple(META["chips_jobs"]["median"]["img"]["geotransform"])
    src = dc.SyntheticSceneSource()
    st, nd, cfree, cf = CHIPS["s2_stack"], CHIPS["s2_nodata"], CHIPS["s2_cloudfree"], CHIPS["s2_cf"]
    dates = _chips_dates("s2")
    src.add(Tile, "sentinel-2:L1C", dc.SceneStack(st, cfree, dates, cf, nodata_mask=np.repeat(nd[..., None], 3, -1).astype(np.uint8)))
    src.add(Tile, "airbus:oneatlas:spot:v2", dc.SceneStack(st, (~nd).astype(np.uint8), dates, cf))
    src.add(Tile, "modelout:classes", dc.SceneStack(CHIPS["cls_stack"], (~CHIPS["cls_nodata"]).astype(np.uint8), _chips_dates("cls"), CHIPS["cls_cf"]))
    layer = _chips_layer()
    jobs = {"median": pkg.DLTileJobConfig(Tile, str(tmp_path / "median"), "sentinel-2:L1C", dt.date(2020, 3, 1), layer, max_cloud_fraction=0,
                                          label_attr="cls"),
            "mosaic": pkg.DLTileJobConfig(Tile, str(tmp_path / "mosaic"), "airbus:oneatlas:spot:v2", dt.date(2020, 3, 1), layer,
                                          max_cloud_fraction=0.6, min_date=dt.date(2020, 1, 10), label_attr=None),
            "stack": pkg.DLTileJobConfig(Tile, str(tmp_path / "stack"), ["airbus:oneatlas:spot:v2", "modelout:classes"], dt.date(2020, 3, 1),
                                         layer, label_attr="cls", bands=["red green blue", "class"])}
    for kind, job in jobs.items():
        ret, img_file, lbl_file = pkg.create_chips_for_tile(job, scene_source=src)
        rec = META["chips_jobs"][kind]
        assert ret is job and os.path.relpath(img_file, job.OUTFOLDER) == rec["img"]["file"]
        assert os.path.relpath(lbl_file, job.OUTFOLDER) == rec["lbl"]["file"]
        for role, path in (("img", img_file), ("lbl", lbl_file)):
            blob = open(path, "rb").read()
            got = oic.decode_image(blob, png_as_tf=False)
            want = CHIPS["%s_%s" % (kind, role)]
            assert got.dtype == GDAL_NP[rec[role]["gdal_type"]] == want.dtype and np.array_equal(got, want), (kind, role)
            t = oic.parse_tiff(blob)
            assert t["compression"] == 5 and t["tiled"]
            assert oic.georef_strings(blob) == (str([float(v) for v in rec[role]["geotransform"]]), "EPSG:32643")
            assert t.get("nodata") == (None if rec[role]["nodata"][0] is None else str(rec[role]["nodata"][0]))
    none_job = pkg.DLTileJobConfig(Tile, str(tmp_path / "none"), "airbus:oneatlas:spot:v2", dt.date(2020, 3, 1), layer,
                                   max_cloud_fraction=0.0000001)
    assert pkg.create_chips_for_tile(none_job, scene_source=src) == (none_job, None, None)      # reference :772-773
