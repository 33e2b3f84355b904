#!/usr/bin/env python3
"""Golden fixtures produced by the UNMODIFIED reference itself.  Run from the repo root, in the build container:

    python tests/golden/gen_golden_reference.py            # rewrites tests/golden/ref_*

`/root/reference/dl_segmentation_utils` is imported exactly as it lies on disk through `oracle/refrun.py`, with the stub
modules under `oracle/refstubs/` standing in for tensorflow / rasterio / descarteslabs / geopandas / osgeo (their
arithmetic is delegated to google.protobuf, NumPy, Pillow/libpng, OpenCV/libtiff/libjpeg-turbo — never to `oracle/` or
the product).  What runs is therefore the reference's own control logic: `images_to_tfrecords_mp` / `_mt` (discovery,
seeded shuffle, worker and shard ranges, skip-and-continue, identifiers), `convert_to_example` (type dispatch), the five
`parse_*_proto`, `create_cloudmasked_s2_array`, `create_img_array_for_tile`, `stack_products_for_tile`.

The functions below are also imported by `tests/test_reference_parity.py` to compare `oracle/` with the reference
live on fresh random inputs whenever `/root/reference` is present.
"""
import contextlib
import datetime as dt
import io
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("OPENCV_LOG_LEVEL", "SILENT")

import synthetic as syn  # noqa: E402
from oracle import refrun  # noqa: E402


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()) as so:
        yield so


# ------------------------------------------------------------------------------------------------ chip folders
def write_tif_folder(root, n=7, size=24, bands=4, tile=16, seed=7003, bad_label=3, bad_image=5):
    """GDAL-style tiled LZW GeoTIFF pairs (the format create_chips_for_tile writes); one pair has a label that is not
    a raster at all and one an image cut off in mid-tile: the reference must skip both and carry on."""
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "labels"), exist_ok=True)
    for i in range(n):
        img, lab, key = syn.cfg3_chip(i, seed=seed, size=size, bands=bands)
        fn = key.replace(":", "#") + ".tif"
        ib, lb = syn.tiff_bytes(img, tile=tile), syn.tiff_bytes(lab, tile=tile, nodata=255)
        if i == bad_label:
            lb = b"this is not a raster file\n" * 3
        if i == bad_image:
            ib = ib[:len(ib) // 4]                    # IFD intact (it sits in front), first tile's bytes missing
        open(os.path.join(root, "images", fn), "wb").write(ib)
        open(os.path.join(root, "labels", fn), "wb").write(lb)


def write_png_folder(root, n=6, size=20, seed=7001, with_jpg=True, with_rgba=True):
    """PNG pairs as a user's chips look (Pillow/libpng), plus one .jpg pair and one RGBA chip (4 bands: the threaded
    translator must refuse it, `_img_to_tf_threaded.py:107-112`; the multiprocess one stores it)."""
    import cv2
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "labels"), exist_ok=True)
    for i in range(n):
        img, lab, key = syn.cfg1_chip(i, seed=seed, size=size)
        fn = key.replace(":", "#")
        if with_rgba and i == 2:
            img = np.concatenate([img, 255 - img[:, :, :1]], axis=-1)
        if with_jpg and i == 4:
            ok1, ij = cv2.imencode(".jpg", np.ascontiguousarray(img[:, :, ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90])
            ok2, lj = cv2.imencode(".jpg", lab, [cv2.IMWRITE_JPEG_QUALITY, 100])
            open(os.path.join(root, "images", fn + ".jpg"), "wb").write(ij.tobytes())
            open(os.path.join(root, "labels", fn + ".jpg"), "wb").write(lj.tobytes())
            continue
        open(os.path.join(root, "images", fn + ".png"), "wb").write(syn.png_bytes(img))
        open(os.path.join(root, "labels", fn + ".png"), "wb").write(syn.png_bytes(lab))


# ------------------------------------------------------------------------------------------------ reference runs
def ref_images_to_tfrecords_mp(directory, out_dir, name, num_shards, num_proc, **kw):
    """The reference's images_to_tfrecords_mp, joblib on threads (its workers share nothing), out_dir made up front
    (the reference's own makedirs races between workers, create_training_samples.ipynb cell 76)."""
    import joblib
    ref = refrun.load()
    os.makedirs(out_dir, exist_ok=True)
    with joblib.parallel_config(backend="threading"), quiet() as so:
        ref.images_to_tfrecords_mp(name, directory, out_dir, num_shards, num_proc, **kw)
    return so.getvalue()


def ref_images_to_tfrecords_mt(directory, out_dir, name, num_shards, num_threads, **kw):
    ref = refrun.load()
    os.makedirs(out_dir, exist_ok=True)
    with quiet() as so:
        ref.images_to_tfrecords_mt(name, directory, out_dir, num_shards, num_threads, **kw)
    return so.getvalue()


def read_shard(path):
    """Records of one shard file as bytes, through the stub's reader (CRCs verified)."""
    refrun.load()
    return [bytes(t.numpy()) for t in refrun.stub("tensorflow").read_tfrecords(path)]


def ref_parse(parser_name, record):
    """Run one of the reference's parse_*_proto on a serialized Example -> (img ndarray, target ndarray, identifier bytes)."""
    ref = refrun.load()
    tf = refrun.stub("tensorflow")
    img, tgt, ident = getattr(ref, parser_name)(tf.constant(record))
    as_np = lambda x: np.asarray(x.numpy() if hasattr(x, "numpy") else x)
    return as_np(img), as_np(tgt), bytes(ident.numpy())


def ref_convert_to_example(*args):
    return refrun.load().convert_to_example(*args).SerializeToString()


# ------------------------------------------------------------------------------------------------ compositors
class Ctx:
    """What the compositors need of a DLTile geocontext."""

    def __init__(self, key="64:0:10.0:43:1:2"):
        self.key = key


def make_catalog(rng, T=7, H=12, W=10, bands=("red", "green", "blue"), dtype=np.uint16, day0=dt.date(2020, 1, 1),
                 tie_days=True, valid_frac=0.6):
    """Synthetic scenes: (dates, cloud fractions, (T,H,W,B) stack, (T,H,W) nodata mask, (T,H,W) valid_cloudfree)."""
    days = np.sort(rng.integers(0, 120, T))
    if tie_days and T >= 4:
        days[2] = days[1]                                       # two scenes on one day
        days[-1] = days[0] + 2 * (60 - days[0]) if days[0] < 60 else days[-1]   # equidistant from day 60 where possible
        days = np.sort(days)
    dates = [day0 + dt.timedelta(days=int(d)) for d in days]
    cf = rng.random(T).round(3)
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    hi = 10000 if info is None or info.max > 10000 else info.max
    lo = -200 if info is not None and info.min < 0 else 0
    stack = rng.integers(lo, hi, (T, H, W, len(bands))).astype(dtype)
    nodata = rng.random((T, H, W)) > 0.85
    cloudfree = (rng.random((T, H, W)) < valid_frac).astype(np.uint8)
    cloudfree[:, :2, :2] = 0                                    # a patch no scene sees
    return dates, cf, stack, nodata, cloudfree


def load_catalog(product, dates, cf, stack, nodata, bands, cloud_product=None, cloudfree=None):
    dl = refrun.stub("descarteslabs")
    for t in range(len(dates)):
        dl.catalog.add(product, dl.Scene(dates[t], {b: stack[t, :, :, k] for k, b in enumerate(bands)},
                                         mask=None if nodata is None else nodata[t], cloud_fraction=cf[t]))
        if cloud_product is not None:
            dl.catalog.add(cloud_product, dl.Scene(dates[t], {"valid_cloudfree": cloudfree[t]}))


def ref_cloudmasked(dates, cf, stack, nodata, cloudfree, min_date=None, max_date=None, bands="red green blue"):
    """create_cloudmasked_s2_array on a synthetic catalogue -> (data float64 filled with 0, mask) or None."""
    ref = refrun.load()
    dl = refrun.stub("descarteslabs")
    dl.catalog.clear()
    load_catalog("sentinel-2:L1C", dates, cf, stack, nodata, bands.split(" "), "sentinel-2:L1C:dlcloud:v1", cloudfree)
    res = ref.create_cloudmasked_s2_array(Ctx(), min_date=min_date, max_date=max_date, bands=bands)
    if res is None:
        return None
    return np.asarray(res.filled(0)), np.ma.getmaskarray(res)


def ref_img_array(dates, cf, stack, nodata, reference_date, min_date=None, max_date=None, max_cloud_fraction=None,
                  bands="red green blue", product="some:product"):
    """create_img_array_for_tile on a synthetic catalogue -> (data filled with 0, mask) or None."""
    ref = refrun.load()
    dl = refrun.stub("descarteslabs")
    dl.catalog.clear()
    load_catalog(product, dates, cf, stack, nodata, bands.split(" "))
    res = ref.create_img_array_for_tile(Ctx(), product, reference_date, min_date=min_date, max_date=max_date, bands=bands,
                                        max_cloud_fraction=max_cloud_fraction)
    if res is None:
        return None
    return np.asarray(res.filled(0)), np.ma.getmaskarray(res)


def ref_stack_products(products):
    """products: list of (name, bands str, dates, cf, stack, nodata).  -> np.dstack result (masked pixels filled with 0)."""
    ref = refrun.load()
    dl = refrun.stub("descarteslabs")
    dl.catalog.clear()
    for name, bands, dates, cf, stack, nodata in products:
        load_catalog(name, dates, cf, stack, nodata, bands.split(" "))
    res = ref.stack_products_for_tile(Ctx(), [p[0] for p in products], [p[1] for p in products])
    return np.asarray(np.ma.filled(res, 0))


# ------------------------------------------------------------------------------------------------ create_chips_for_tile
class Tile:
    """What create_chips_for_tile reads of a DLTile."""

    def __init__(self, key="12:2:10.0:43:7:11", tilesize=12, pad=2, geotrans=(499680.0, 10.0, 0.0, 5300360.0, 0.0, -10.0)):
        self.key, self.tilesize, self.pad, self.geotrans = key, tilesize, pad, geotrans
        self.wkt, self.crs, self.epsg = 'PROJCS["WGS 84 / UTM zone 43N"]', "EPSG:32643", 32643


def label_layer(rng, tile, n=5):
    """Random polygons (some with holes) in the tile's map coordinates with a `cls` attribute."""
    S = tile.tilesize + 2 * tile.pad
    gt = tile.geotrans
    layer = []
    for f in range(n):
        k = int(rng.integers(3, 8))
        ang = np.sort(rng.random(k)) * 2 * np.pi
        rad = rng.uniform(0.1 * S, 0.4 * S, k)
        c = rng.uniform(0, S, 2)
        ring = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        ring = np.vstack([ring, ring[:1]])
        rings = [ring] + ([(c + (ring - c) * 0.4)[::-1].copy()] if f % 2 == 0 else [])
        layer.append(([np.stack([gt[0] + r[:, 0] * gt[1], gt[3] + r[:, 1] * gt[5]], 1) for r in rings], {"cls": int(rng.integers(0, 9))}))
    return layer


def ref_create_chips(job_kind, tile, out_dir, layer, scenes, label_attr="cls", label_ndv=255):
    """Run the reference's create_chips_for_tile (:693-800) with the recording GDAL stub.  scenes: what load_catalog takes
    per product.  -> (returned paths, {path: recorded dataset}).  gdal.RasterizeLayer is served by oracle/rasterize.py (GDAL
    itself cannot be had here: that arithmetic stays unpinned), everything else is the reference's own code."""
    from oracle import rasterize as orr
    ref = refrun.load()
    chips = refrun.submodule("_descartes_img_chips")
    dl, gdal, ogr = refrun.stub("descarteslabs"), sys.modules["osgeo.gdal"], sys.modules["osgeo.ogr"]
    dl.catalog.clear()
    for product, (bands, dates, cf, stack, nodata, cloud_product, cloudfree) in scenes.items():
        load_catalog(product, dates, cf, stack, nodata, bands.split(" "), cloud_product, cloudfree)
    ogr.datasets["labels.geojson"] = [layer]
    gdal.created.clear()

    def hook(ds, bands, lyr, burn_values, options):
        assert bands == [1] and "ALL_TOUCHED=TRUE" in options
        attr = [o.split("=")[1] for o in options if o.startswith("ATTRIBUTE=")]
        feats = [([orr.to_pixel_space(r, ds.geotransform) for r in rings], int(a[attr[0]]) if attr else int(burn_values[0])) for rings, a in lyr]
        burnt = orr.rasterize(feats, (ds.ysize, ds.xsize), 0)
        touched = orr.rasterize([(f[0], 1) for f in feats], (ds.ysize, ds.xsize), 0) == 1
        ds.data[0][touched] = burnt[touched]
    gdal.rasterize_hook = hook
    products = list(scenes)
    if job_kind == "stack":
        job = chips.DLTileJobConfig(tile, out_dir, products, dt.date(2020, 3, 1), "labels.geojson", label_attr=label_attr,
                                    bands=[scenes[p][0] for p in products], label_nodata_value=label_ndv)
    elif job_kind == "median":
        job = chips.DLTileJobConfig(tile, out_dir, "sentinel-2:L1C", dt.date(2020, 3, 1), "labels.geojson", max_cloud_fraction=0,
                                    label_attr=label_attr, bands=scenes["sentinel-2:L1C"][0], label_nodata_value=label_ndv)
    else:
        job = chips.DLTileJobConfig(tile, out_dir, products[0], dt.date(2020, 3, 1), "labels.geojson", max_cloud_fraction=0.6,
                                    min_date=dt.date(2020, 1, 10), label_attr=None, bands=scenes[products[0]][0],
                                    label_nodata_value=label_ndv)
    _, img_file, lbl_file = ref.create_chips_for_tile(job)
    return (img_file, lbl_file), dict(gdal.created)


# ------------------------------------------------------------------------------------------------ main
def _copy_shards(src, dst):
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(src)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))


def main():
    import tempfile
    assert refrun.available(), "needs /root/reference"
    tmp = tempfile.mkdtemp()
    meta = {}
    # ---- A: GeoTIFF chips through images_to_tfrecords_mp
    tif_dir = os.path.join(HERE, "ref_chips_tif")
    shutil.rmtree(tif_dir, ignore_errors=True)
    write_tif_folder(tif_dir)
    runs = {"ref_mp_tif_arrays": dict(file_ext="tif", store_as_array=True, dltile_from_filename=True),
            "ref_mp_tif_raw_georef": dict(file_ext="tif", store_as_array=False, dltile_from_filename=False)}
    for name, kw in runs.items():
        out = os.path.join(tmp, name)
        log = ref_images_to_tfrecords_mp(tif_dir, out, "chips", 4, 2, **kw)
        shutil.rmtree(os.path.join(HERE, name), ignore_errors=True)
        _copy_shards(out, os.path.join(HERE, name))
        meta[name] = dict(kw, num_shards=4, num_proc=2, skipped=log.count("SKIPPED"))
    # ---- B: PNG (+ one JPG pair, one RGBA chip) through both translators
    png_dir = os.path.join(HERE, "ref_chips_png")
    shutil.rmtree(png_dir, ignore_errors=True)
    write_png_folder(png_dir)
    runs = {"ref_mt_png_raw": dict(store_as_array=False), "ref_mt_png_arrays": dict(store_as_array=True),
            "ref_mt_png_to_jpg": dict(store_as_array=False, convert_png_to_jpg=True)}
    for name, kw in runs.items():
        out = os.path.join(tmp, name)
        log = ref_images_to_tfrecords_mt(png_dir, out, "chips", 2, 2, **kw)
        shutil.rmtree(os.path.join(HERE, name), ignore_errors=True)
        _copy_shards(out, os.path.join(HERE, name))
        meta[name] = dict(kw, num_shards=2, num_threads=2, skipped=log.count("SKIPPED"))
    out = os.path.join(tmp, "ref_mp_png_arrays")
    log = ref_images_to_tfrecords_mp(png_dir, out, "chips", 2, 1, file_ext="png", store_as_array=True)
    shutil.rmtree(os.path.join(HERE, "ref_mp_png_arrays"), ignore_errors=True)
    _copy_shards(out, os.path.join(HERE, "ref_mp_png_arrays"))
    meta["ref_mp_png_arrays"] = dict(file_ext="png", store_as_array=True, num_shards=2, num_proc=1, skipped=log.count("SKIPPED"))
    # ---- C: the five parsers on those records
    parsed = {}
    for parser, shard_dir in (("parse_8bit_array_proto", "ref_mt_png_arrays"), ("parse_8bit_array_proto", "ref_mp_png_arrays"),
                              ("parse_higher_dtype_array_proto", "ref_mp_tif_arrays"),
                              ("parse_encoded_rgb_img_proto", "ref_mt_png_raw"), ("parse_encoded_rgb_img_proto", "ref_mt_png_to_jpg"),
                              ("parse_encoded_gdal_proto_eager", "ref_mp_tif_raw_georef"),
                              ("parse_encoded_gdal_proto_wrapped", "ref_mp_tif_raw_georef"),
                              ("parse_encoded_gdal_proto_eager", "ref_mt_png_raw"),
                              ("parse_encoded_gdal_proto_wrapped", "ref_mt_png_raw")):
        k = 0
        for f in sorted(os.listdir(os.path.join(HERE, shard_dir))):
            for rec in read_shard(os.path.join(HERE, shard_dir, f)):
                try:
                    img, tgt, ident = ref_parse(parser, rec)
                except Exception as e:                          # e.g. the chip that was stored truncated (raw mode)
                    meta.setdefault("parse_errors", {})["%s|%s|%d" % (parser, shard_dir, k)] = type(e).__name__
                    k += 1
                    continue
                parsed["%s|%s|%d|img" % (parser, shard_dir, k)] = img
                parsed["%s|%s|%d|tgt" % (parser, shard_dir, k)] = tgt
                parsed["%s|%s|%d|id" % (parser, shard_dir, k)] = np.frombuffer(ident, np.uint8)
                k += 1
        meta["%s|%s" % (parser, shard_dir)] = k
    np.savez_compressed(os.path.join(HERE, "ref_parsed.npz"), **parsed)
    # ---- D: convert_to_example type dispatch
    rng = np.random.default_rng(7100)
    cases = {}
    arrs = {"u8x3": rng.integers(0, 256, (5, 6, 3)).astype(np.uint8), "u16x4": rng.integers(0, 65536, (5, 6, 4)).astype(np.uint16),
            "i16x2": rng.integers(-3000, 3000, (5, 6, 2)).astype(np.int16), "f32x1": rng.random((5, 6, 1)).astype(np.float32),
            "f64x2": rng.random((5, 6, 2)) * 1e3, "lab8": rng.integers(0, 10, (5, 6)).astype(np.uint8),
            "lab16": rng.integers(0, 300, (5, 6)).astype(np.uint16), "lab8_3d": rng.integers(0, 10, (5, 6, 1)).astype(np.uint8)}
    for iname, lname in (("u8x3", "lab8"), ("u8x3", "lab8_3d"), ("u16x4", "lab8"), ("i16x2", "lab8"), ("f32x1", "lab8"),
                         ("f64x2", "lab16"), ("u8x3", "lab16")):
        img, lab = arrs[iname], arrs[lname]
        cases["%s+%s" % (iname, lname)] = ref_convert_to_example(img, lab, 5, 6, img.shape[2], 5, 6, "64:0:10.0:43:-3:%d" % len(cases)).hex()
    cases["bytes+bytes"] = ref_convert_to_example(b"\x89PNG fake image", b"\x89PNG fake label", 7, 8, 3, 7, 8, "key/with|bar").hex()
    np.savez_compressed(os.path.join(HERE, "ref_convert_inputs.npz"), **arrs)
    meta["convert_to_example"] = cases
    # ---- E: compositors
    comp = {}
    rng = np.random.default_rng(7200)
    dates, cf, stack, nodata, cloudfree = make_catalog(rng)
    comp["stack"], comp["nodata"], comp["cloudfree"], comp["cf"] = stack, nodata, cloudfree, cf
    meta["composite_dates"] = [d.isoformat() for d in dates]
    for tag, (lo, hi) in {"all": (None, None), "window": (dt.date(2020, 1, 20), dt.date(2020, 4, 1)),
                          "empty": (dt.date(2021, 1, 1), None)}.items():
        r = ref_cloudmasked(dates, cf, stack, nodata, cloudfree, lo, hi)
        meta["median_%s" % tag] = None if r is None else [None if lo is None else lo.isoformat(), None if hi is None else hi.isoformat()]
        if r is not None:
            comp["median_%s_data" % tag], comp["median_%s_mask" % tag] = r
    refd = dt.date(2020, 3, 1)
    for tag, kw in {"plain": {}, "cloud": dict(max_cloud_fraction=0.5), "window": dict(min_date=dt.date(2020, 1, 20), max_date=dt.date(2020, 4, 1)),
                    "both": dict(min_date=dt.date(2020, 1, 20), max_date=dt.date(2020, 4, 1), max_cloud_fraction=0.7),
                    "none": dict(max_cloud_fraction=0.0)}.items():
        r = ref_img_array(dates, cf, stack, nodata, refd, **kw)
        meta["mosaic_%s" % tag] = None if r is None else {k: (v.isoformat() if hasattr(v, "isoformat") else v) for k, v in kw.items()}
        if r is not None:
            comp["mosaic_%s_data" % tag], comp["mosaic_%s_mask" % tag] = r
    meta["mosaic_reference_date"] = refd.isoformat()
    prods = []
    for k, (dtype, bands) in enumerate(((np.uint16, "red green blue"), (np.uint8, "class"), (np.int16, "ndvi evi"))):
        d2, cf2, st2, nd2, _ = make_catalog(rng, T=3, bands=bands.split(" "), dtype=dtype, tie_days=False)
        prods.append(("prod%d" % k, bands, d2, cf2, st2, nd2))
        comp["prod%d_stack" % k], comp["prod%d_nodata" % k] = st2, nd2
        meta["prod%d" % k] = dict(bands=bands, dates=[d.isoformat() for d in d2])
    comp["dstack_all"] = ref_stack_products(prods)
    comp["dstack_u16_u8"] = ref_stack_products(prods[:2])
    comp["dstack_u16_i16"] = ref_stack_products([prods[0], prods[2]])
    meta["dstack_dtypes"] = {k: str(comp[k].dtype) for k in ("dstack_all", "dstack_u16_u8", "dstack_u16_i16")}
    np.savez_compressed(os.path.join(HERE, "ref_composites.npz"), **comp)
    # ---- F: create_chips_for_tile (dispatch, file names, band-by-band writes, dtypes, nodata) with the recording GDAL stub
    rng = np.random.default_rng(7300)
    tile = Tile()
    chipfix = {}
    layer = label_layer(rng, tile)
    for k, (rings, attrs) in enumerate(layer):
        for j, r in enumerate(rings):
            chipfix["layer_%d_%d" % (k, j)] = r
    meta["chips_layer_cls"] = [a["cls"] for _, a in layer]
    meta["chips_layer_rings"] = [len(r) for r, _ in layer]
    S = tile.tilesize + 2 * tile.pad
    d1, cf1, st1, nd1, cfree1 = make_catalog(rng, T=5, H=S, W=S)
    d2, cf2, st2, nd2, _ = make_catalog(rng, T=3, H=S, W=S, bands=["class"], dtype=np.uint8, tie_days=False)
    chipfix.update(s2_stack=st1, s2_nodata=nd1, s2_cloudfree=cfree1, s2_cf=cf1, cls_stack=st2, cls_nodata=nd2, cls_cf=cf2)
    meta["chips_s2_dates"], meta["chips_cls_dates"] = [d.isoformat() for d in d1], [d.isoformat() for d in d2]
    jobs = {"median": {"sentinel-2:L1C": ("red green blue", d1, cf1, st1, nd1, "sentinel-2:L1C:dlcloud:v1", cfree1)},
            "mosaic": {"airbus:oneatlas:spot:v2": ("red green blue", d1, cf1, st1, nd1, None, None)},
            "stack": {"airbus:oneatlas:spot:v2": ("red green blue", d1, cf1, st1, nd1, None, None),
                      "modelout:classes": ("class", d2, cf2, st2, nd2, None, None)}}
    meta["chips_jobs"] = {}
    for kind, scenes in jobs.items():
        (img_file, lbl_file), created = ref_create_chips(kind, tile, os.path.join(tmp, "chips_" + kind), layer, scenes)
        rec = {}
        for role, path in (("img", img_file), ("lbl", lbl_file)):
            ds = created[path]
            chipfix["%s_%s" % (kind, role)] = np.transpose(ds.data, (1, 2, 0))
            rec[role] = dict(file=os.path.relpath(path, os.path.join(tmp, "chips_" + kind)), gdal_type=ds.gdt, options=ds.options,
                             nodata=ds.nodata, geotransform=list(ds.geotransform), driver=ds.driver)
        meta["chips_jobs"][kind] = rec
    np.savez_compressed(os.path.join(HERE, "ref_chips_for_tile.npz"), **chipfix)
    json.dump(meta, open(os.path.join(HERE, "ref_meta.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)
    print("wrote reference-run fixtures:", ", ".join(sorted(k for k in meta if k.startswith("ref_"))))


if __name__ == "__main__":
    main()
