"""CPU tests: pin the oracle against the committed golden vectors (tests/golden, made by gen_golden.py from
libtiff / libpng / zlib / google.protobuf / numpy.ma and the RFC 3720 + TFRecord known answers) and against
the live third-party codecs in this image."""
import io
import json
import os
import random

import numpy as np
import pytest

import synthetic as syn
from oracle import composite as ocomp
from oracle import example_proto as oep
from oracle import imagecodecs as oic
from oracle import normalise as onorm
from oracle import partition as opart
from oracle import tfrecord as otfr

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = json.load(open(os.path.join(G, "vectors.json")))


def _g(name):
    p = os.path.join(G, name)
    return np.load(p) if name.endswith(".npy") else open(p, "rb").read()


def test_crc32c_rfc3720_vectors():
    for v in VEC["crc32c"]:
        data = bytes.fromhex(v["hex"])
        assert otfr.crc32c(data) == v["crc"] == otfr.crc32c_py(data)
        assert otfr.masked_crc32c(data) == v["masked"]
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 4097):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert otfr.crc32c(d) == otfr.crc32c_py(d)


def test_tfrecord_frame_vectors_and_scan():
    for v in VEC["frames"]:
        assert otfr.frame(bytes.fromhex(v["data_hex"])).hex() == v["frame_hex"]
    recs = [b"", b"abc", bytes(range(16)), b"x" * 1000]
    buf = b"".join(otfr.frame(r) for r in recs)
    assert otfr.read_records(buf) == recs
    bad = bytearray(buf)
    bad[-6] ^= 1
    with pytest.raises(otfr.DataLossError):
        otfr.read_records(bytes(bad))
    assert otfr.read_records(bytes(bad), verify=False)[-1] != recs[-1]
    with pytest.raises(otfr.DataLossError):
        otfr.scan(buf[:-1])


def test_example_bytes_match_google_protobuf():
    img8, lab8 = _g("chip8_img.npy"), _g("chip8_lab.npy")
    img16, lab16 = _g("chip16_img.npy"), _g("chip16_lab.npy")
    h, w, c = img8.shape
    s = oep.convert_to_example(img8, lab8, h, w, c, h, w, VEC["example"]["key8"]).SerializeToString()
    assert s == _g("example_bytes.bin")
    h, w, c = img16.shape
    s16 = oep.convert_to_example(img16, lab16, h, w, c, h, w, VEC["example"]["key16"]).SerializeToString()
    assert s16 == _g("example_float.bin")
    # and back
    i, t, ident = oep.parse_8bit_array_proto(_g("example_bytes.bin"))
    assert np.array_equal(i, img8) and np.array_equal(t, lab8) and ident == VEC["example"]["key8"].encode()
    i, t, ident = oep.parse_higher_dtype_array_proto(_g("example_float.bin"))
    assert i.dtype == np.float32 and np.array_equal(i, img16.astype(np.float32)) and np.array_equal(t, lab16.astype(np.float32))
    with pytest.raises(oep.ParseError):
        oep.parse_8bit_array_proto(_g("example_float.bin"))


def test_convert_to_example_type_dispatch():
    img8, lab8 = _g("chip8_img.npy"), _g("chip8_lab.npy")
    f = oep.parse_example(oep.convert_to_example(img8, lab8, 24, 24, 3, 24, 24, "k").SerializeToString())
    assert f["image/image_data"][0] == "bytes" and f["target/target_data"][0] == "bytes"
    # uint8 label next to a uint16 image is forced to FloatList (reference :184-197)
    f = oep.parse_example(oep.convert_to_example(img8.astype(np.uint16), lab8, 24, 24, 3, 24, 24, "k").SerializeToString())
    assert f["image/image_data"][0] == "float" and f["target/target_data"][0] == "float"
    f = oep.parse_example(oep.convert_to_example(b"raw-img", b"raw-lab", 1, 2, 3, 1, 2, "k").SerializeToString())
    assert f["image/image_data"] == ("bytes", [b"raw-img"]) and f["identifier"] == ("bytes", [b"k"])
    assert sorted(f) == sorted(oep.KEYS)
    # payload sizes of SURVEY.md Appendix B (22-byte key)
    key22 = "448:32:10.0:43:-38:349"
    assert len(key22) == 22
    big = oep.convert_to_example(np.zeros((256, 256, 3), np.uint8), np.zeros((256, 256), np.uint8), 256, 256, 3, 256, 256, key22)
    assert len(big.SerializeToString()) == 262381
    bigf = oep.convert_to_example(np.zeros((512, 512, 4), np.uint16), np.zeros((512, 512), np.uint8), 512, 512, 4, 512, 512, key22)
    assert len(bigf.SerializeToString()) == 5243122


def test_tiff_and_png_decode_against_golden_files():
    img, lab = _g("tiff_img.npy"), _g("tiff_lab.npy")
    assert np.array_equal(oic.decode_image(_g("libtiff_cv2_lzw_u16x4.tif")), img)          # libtiff-encoded, strips, predictor 2
    assert np.array_equal(oic.decode_image(_g("libtiff_pil_lzw_u8.tif"))[..., 0], lab)
    assert np.array_equal(oic.decode_image(_g("libtiff_pil_deflate_u8.tif"))[..., 0], lab)
    assert np.array_equal(oic.decode_image(_g("gdalstyle_tiled_lzw_u16x4.tif")), img)      # GDAL layout: 64x64 tiles, predictor 1
    assert np.array_equal(oic.decode_image(_g("gdalstyle_tiled_lzw_label.tif"))[..., 0], lab)
    assert oic.parse_tiff(_g("gdalstyle_tiled_lzw_label.tif"))["nodata"] == "255"
    assert np.array_equal(oic.decode_image(_g("libpng_rgb.png")), _g("png_img.npy"))
    assert np.array_equal(oic.decode_image(_g("libpng_label.png"))[..., 0], _g("png_lab.npy"))
    assert oic.image_shape(_g("libpng_rgb.png")) == (64, 64, 3)
    assert oic.image_shape(_g("gdalstyle_tiled_lzw_u16x4.tif")) == (96, 96, 4)


def test_decoders_against_live_libtiff_and_libpng():
    import cv2
    from PIL import Image
    rng = np.random.default_rng(3)
    img, lab, _ = syn.cfg3_chip(9, size=130)
    for arr in (img, (img >> 4).astype(np.uint8)[..., :3], lab):
        ok, enc = cv2.imencode(".tif", arr if arr.ndim == 2 else arr[..., [2, 1, 0, 3][:arr.shape[2]] if arr.shape[2] == 4 else [2, 1, 0]],
                               [cv2.IMWRITE_TIFF_COMPRESSION, 5])
        assert ok
        got = oic.decode_image(enc.tobytes())
        assert np.array_equal(got if arr.ndim == 3 else got[..., 0], arr)
    # our writers are readable by libtiff / libpng (so the GPU tests' inputs are legitimate files)
    for kw in (dict(tile=64), dict(tile=None, predictor=2), dict(tile=64, compression="deflate"), dict(tile=None, big_endian=True)):
        b = syn.tiff_bytes(lab, **kw)
        assert np.array_equal(np.array(Image.open(io.BytesIO(b))), lab), kw
        assert np.array_equal(oic.decode_image(b)[..., 0], lab)
    noise = rng.integers(0, 256, (70, 50, 3), dtype=np.uint8)
    for ft in ((0,), (1,), (2,), (3,), (4,), (0, 1, 2, 3, 4)):
        b = syn.png_bytes_manual(noise, filter_types=ft, idat_chunk=333)
        assert np.array_equal(np.array(Image.open(io.BytesIO(b))), noise)
        assert np.array_equal(oic.decode_image(b), noise)
    with pytest.raises(oic.DecodeError):
        oic.decode_image(b"nope")
    with pytest.raises(oic.DecodeError):
        oic.decode_image(syn.tiff_bytes(lab, tile=64)[:600])


def _png_flavour_cases():
    rng = np.random.default_rng(77)
    H, W = 37, 53                                             # W * depth not a multiple of 8: rows end mid-byte
    pal = rng.integers(0, 256, (256, 3), dtype=np.uint8)
    cases = []
    for depth in (1, 2, 4, 8):
        idx = rng.integers(0, 1 << depth, (H, W), dtype=np.uint8)
        cases.append(("palette%d" % depth, syn.png_bytes_flavour(idx, depth, 3, palette=pal[:1 << depth])))
        cases.append(("palette%d+tRNS" % depth, syn.png_bytes_flavour(idx, depth, 3, palette=pal[:1 << depth],
                                                                      trns=rng.integers(0, 256, max(1, (1 << depth) // 2), dtype=np.uint8))))
        if depth < 8:
            cases.append(("grey%d" % depth, syn.png_bytes_flavour(idx, depth, 0)))
    cases.append(("short palette", syn.png_bytes_flavour(rng.integers(0, 8, (H, W), dtype=np.uint8), 4, 3, palette=pal[:5])))
    for ct, C in ((0, 1), (2, 3), (4, 2), (6, 4)):
        cases.append(("16-bit ct%d" % ct, syn.png_bytes_flavour(rng.integers(0, 65536, (H, W, C)), 16, ct)))
    return cases


def _png_interlaced_cases():
    """Adam7 files (name, blob, samples as written) of several geometries, including ones where passes are empty."""
    rng = np.random.default_rng(78)
    cases = []
    pal = rng.integers(0, 256, (256, 3), dtype=np.uint8)
    for (H, W) in ((37, 53), (1, 1), (3, 2), (8, 8), (9, 5), (64, 64)):
        for ct, C, depth in ((2, 3, 8), (0, 1, 8), (6, 4, 8), (4, 2, 8), (3, 1, 8), (2, 3, 16), (0, 1, 16)):
            a = rng.integers(0, 1 << depth, (H, W, C))
            cases.append(("adam7 %dx%d ct%d/%d" % (H, W, ct, depth),
                          syn.png_bytes_flavour(a, depth, ct, palette=pal if ct == 3 else None, interlace=True), a, ct, depth))
    return cases


def test_png_adam7_against_pillow():
    from PIL import Image
    for name, blob, a, ct, depth in _png_interlaced_cases():
        got = oic.decode_png(blob, True)
        im = Image.open(io.BytesIO(blob))
        im.load()
        if ct == 3:
            want = np.asarray(im.convert("RGB"))
        elif depth == 16 and ct == 0:
            want = (np.asarray(im).astype(np.uint16) >> 8).astype(np.uint8)[..., None]
        else:
            want = np.asarray(im)
            want = want[..., None] if want.ndim == 2 else want
        assert got.shape == want.shape and np.array_equal(got, want), name
        if depth == 16:
            assert np.array_equal(oic.decode_png(blob, False), a.astype(np.uint16)), name
        elif ct != 3:
            assert np.array_equal(oic.decode_png(blob, False), a.astype(np.uint8)), name
    with pytest.raises(oic.DecodeError):                                  # sub-byte interlaced stays out of scope
        oic.decode_png(syn.png_bytes_flavour(np.zeros((5, 5), np.uint8), 4, 0, interlace=True))


def test_png_flavours_against_pillow_and_libpng():
    """Palette, 1/2/4-bit and 16-bit PNGs under both presentations: tf.image.decode_png's libpng transforms (checked
    against Pillow's, which applies the same expansions, and OpenCV's libpng for palettes) and GDAL's raw view
    (indices / unscaled / uint16; checked against Pillow's raw modes and OpenCV's 16-bit read)."""
    import cv2
    from PIL import Image
    for name, blob in _png_flavour_cases():
        tf_view, gdal_view = oic.decode_png(blob, True), oic.decode_png(blob, False)
        im = Image.open(io.BytesIO(blob))
        im.load()
        if name.startswith("palette") or name == "short palette":
            want = np.asarray(im.convert("RGBA" if "tRNS" in name else "RGB"))
            assert np.array_equal(tf_view, want), name
            assert np.array_equal(gdal_view[..., 0], np.asarray(im)), name            # mode P: the indices
            if "tRNS" not in name:
                assert np.array_equal(cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_COLOR)[..., ::-1], tf_view), name
        elif name.startswith("grey"):
            depth = int(name[4:])
            want = np.asarray(im.convert("L"))                                       # '1' -> 0/255, L;2 -> x85, L;4 -> x17
            assert np.array_equal(tf_view[..., 0], want), name
            assert np.array_equal(gdal_view[..., 0], want // {1: 255, 2: 85, 4: 17}[depth]), name
        else:
            full = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)  # libpng, 16 bits kept
            if full.ndim == 2:
                full = full[..., None]
            if full.shape[-1] >= 3:                                                  # OpenCV hands back B,G,R[,A]
                full = full[..., [2, 1, 0] + ([3] if full.shape[-1] == 4 else [])]
            if gdal_view.shape[-1] == 2:                                             # grey+alpha: OpenCV expands to BGRA
                assert np.array_equal(gdal_view[..., 0], full[..., 0]) and np.array_equal(gdal_view[..., 1], full[..., 3]), name
            else:
                assert np.array_equal(gdal_view, full), name
            assert gdal_view.dtype == np.uint16 and np.array_equal(tf_view, (gdal_view >> 8).astype(np.uint8)), name


def test_png_error_behaviour_follows_libpng_and_zlib():
    """Critical-chunk CRC mismatch is fatal (libpng); an incomplete Huffman set is rejected in the block header (zlib)."""
    import cv2
    good = syn.png_bytes_raw_zlib(1, 1, 1, syn.handmade_dynamic_deflate(None, {0: 2, 65: 2, 66: 2, 256: 2}, [0, 65, 256]))
    assert oic.decode_image(good).tolist() == [[[65]]]
    assert cv2.imdecode(np.frombuffer(good, np.uint8), cv2.IMREAD_UNCHANGED).tolist() == [[65]]
    bad_set = syn.png_bytes_raw_zlib(1, 1, 1, syn.handmade_dynamic_deflate(None, {0: 2, 65: 2, 256: 2}, [0, 65, 256]))
    bad_crc = bytearray(good)
    bad_crc[-16] ^= 1
    for blob in (bad_set, bytes(bad_crc)):
        with pytest.raises(oic.DecodeError):
            oic.decode_image(blob)
        assert cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED) is None      # libpng agrees


def test_median_against_numpy_ma_golden():
    stack, valid = _g("median_stack.npy"), _g("median_valid.npy")
    ref = ocomp.median_composite(stack, valid)
    assert ref.dtype == np.float64
    assert np.array_equal(ref.filled(0.0), _g("median_out.npy"))
    assert np.array_equal(np.ma.getmaskarray(ref), _g("median_mask.npy"))
    for kat in VEC["median_kat"]:
        v = np.array([int(c) for c in kat["valid"]], np.uint8).reshape(4, 1, 1)
        r = ocomp.median_composite(np.array(kat["values"], np.uint16).reshape(4, 1, 1, 1), v)
        if kat["median"] is None:
            assert np.ma.getmaskarray(r).all()
        else:
            assert float(r[0, 0, 0]) == kat["median"]
    r = ocomp.median_composite(np.array([65535, 65534], np.uint16).reshape(2, 1, 1, 1), np.ones((2, 1, 1), np.uint8))
    assert float(r[0, 0, 0]) == 65534.5


def test_nearest_date_rules():
    rng = np.random.default_rng(1)
    T, H, W, B = 6, 5, 4, 2
    stack = rng.integers(1, 100, (T, H, W, B), dtype=np.uint16)
    valid = np.ones((T, H, W), np.uint8)
    days = [10, 20, 20, 30, 40, 50]
    cf = [0.1, 0.2, 0.3, 0.5, 0.1, 0.0]
    out, mask, src = ocomp.nearest_date_mosaic(stack, valid, days, cf, 20)
    assert (src == 2).all()                                   # tie at distance 0 -> later scene
    out, mask, src = ocomp.nearest_date_mosaic(stack, valid, days, cf, 25)
    assert (src == 3).all()                                   # 20,20,30 all at distance 5 -> last of them
    out, mask, src = ocomp.nearest_date_mosaic(stack, valid, days, cf, 25, max_cf=0.5)
    assert (src == 2).all()                                   # strict < drops scene 3
    out, mask, src = ocomp.nearest_date_mosaic(stack, valid, days, cf, 45, min_day=20, max_day=40)
    assert (src == 3).all()                                   # end exclusive: day 40 is out
    assert ocomp.nearest_date_mosaic(stack, valid, days, cf, 25, min_day=60) is None
    valid[3] = 0
    valid[2, 0, 0] = 0
    out, mask, src = ocomp.nearest_date_mosaic(stack, valid, days, cf, 25)
    assert src[0, 0] == 1 and src[1, 1] == 2 and np.array_equal(out[0, 0], stack[1, 0, 0])
    valid[:, 4, 3] = 0
    out, mask, src = ocomp.nearest_date_mosaic(stack, valid, days, cf, 25)
    assert mask[4, 3] and not out[4, 3].any() and src[4, 3] == -1


def test_partition_shuffle_and_names(tmp_path):
    assert opart.tile_key(VEC["identifier"]["path"]) == VEC["identifier"]["key"]
    idx = list(range(20))
    random.seed(12345)
    random.shuffle(idx)
    assert idx == VEC["shuffle20"]
    assert opart.worker_ranges(6000, 12) == [[500 * i, 500 * (i + 1)] for i in range(12)]
    plan = opart.shard_plan(6000, 12, 12)
    assert plan == [(s, 500 * s, 500 * (s + 1)) for s in range(12)]
    # benchmark sizes give worker-count-invariant shard boundaries
    for n, S, gs in ((6000, 24, (1, 2, 3, 4, 6, 8, 12, 24)), (1024, 16, (1, 2, 4, 8, 16))):
        ref = opart.shard_plan(n, S, 1)
        for g in gs:
            assert opart.shard_plan(n, S, g) == ref
    assert opart.shard_plan(103, 12, 4) != opart.shard_plan(103, 12, 1)        # awkward N: boundaries move, as in the reference
    assert opart.shard_plan(1000003, 16, 8) != opart.shard_plan(1000003, 16, 1)
    assert opart.shard_name("train", 2, 10) == "train-00002-of-00010"
    for sub in ("images", "labels"):
        os.makedirs(tmp_path / sub)
        for k in range(7):
            (tmp_path / sub / ("1#2#%d.tif" % k)).write_bytes(b"x")
    a, b = opart.find_image_files(str(tmp_path), "tif")
    assert [os.path.basename(x) for x in a] == [os.path.basename(x) for x in b] and len(a) == 7
    a2, _ = opart.find_image_files(str(tmp_path), "tif")
    assert a == a2


def test_normalise_onehot_and_stats_definitions():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (2, 5, 6, 3), dtype=np.uint8)
    lab = rng.integers(0, 12, (2, 5, 6), dtype=np.uint8)
    lab[0, 0, 0] = 255
    mean, std = np.array([1.5, 100.0, 200.25], np.float32), np.array([2.0, 3.0, 0.5], np.float32)
    x = onorm.normalise(img, mean, std)
    assert x.dtype == np.float32 and x[1, 2, 3, 1] == np.float32((np.float32(img[1, 2, 3, 1]) - mean[1]) / std[1])
    h = onorm.one_hot(lab, 10)
    assert h.shape == (2, 5, 6, 10) and h.dtype == np.float32 and not h[0, 0, 0].any()
    assert (h.sum(-1) == (lab < 10)).all()
    st = onorm.band_stats(img)
    flat = img.reshape(-1, 3).astype(np.int64)
    assert st == [(60, int(flat[:, b].sum()), int((flat[:, b] ** 2).sum())) for b in range(3)]
    m, s = onorm.mean_std_from_stats(st)
    np.testing.assert_allclose(m, flat.mean(0), rtol=1e-6)
    np.testing.assert_allclose(s, flat.std(0), rtol=1e-6)


def test_fixture_lzw_encoder_with_restarts_round_trips_and_is_read_by_libtiff():
    """The fixture encoder the GPU writer is compared with byte for byte (synthetic/csrc/lzwenc.c): with a Clear every R input
    bytes its streams must decode back through the oracle's TIFF 6.0 decoder for any R and any data, restart 0 must be the
    classic stream, and libtiff (through OpenCV) must read a tiled file made of such streams."""
    import cv2
    import synthetic as syn
    from oracle import imagecodecs as oic
    rng = np.random.default_rng(21)
    msgs = [b"", b"x", bytes(3000), bytes(range(256)) * 9, rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
            rng.integers(0, 3, 20000, dtype=np.uint8).tobytes(), rng.integers(0, 256, 1024, dtype=np.uint8).tobytes(),
            rng.integers(0, 256, 1025, dtype=np.uint8).tobytes()]
    for m in msgs:
        assert syn.lzw_encode(m, 0) == syn.lzw_encode(m)
        for restart in (16, 48, 400, 1024):
            enc = syn.lzw_encode(m, restart)
            assert oic.lzw_decode(enc, len(m)) == m, (len(m), restart)
    arr = rng.integers(0, 60000, (300, 500), dtype=np.uint16)
    blob = syn.tiff_bytes(arr, tile=256, lzw_restart=1024)
    assert np.array_equal(cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED), arr)
    assert np.array_equal(oic.decode_image(blob)[..., 0], arr)
