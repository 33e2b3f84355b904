"""JPEG chips (.jpg): oracle pinned against libjpeg-turbo (golden files + live cv2 / Pillow), the host marker walk of
the C ABI against the oracle's, and — on the GPU — the decode kernels against both.  Replaces tf.image.decode_jpeg
behind ImageCoder.decode_jpeg (reference _img_to_tf_threaded.py:36-38,51-56,97-103).  Bar: bit-exact."""
import glob
import io
import os

import numpy as np
import pytest

from oracle import jpegcodec as ojpg
from oracle import jpegenc as ojenc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(G, "jpeg_*.npy")))


def _case(name):
    return open(os.path.join(G, "jpeg_%s.jpg" % name), "rb").read(), np.load(os.path.join(G, "jpeg_%s.npy" % name))


def _smooth(h, w, c, rng):
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / (7.0 + 3 * i)) * np.cos(yy / (5.0 + 2 * i)) for i in range(c)], -1)
    return np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)


def _live_files(seed, sizes, qualities=(100, 75, 30), restarts=(0, 3)):
    """(name, file bytes, pixels as libjpeg-turbo decodes them) for every sampling mode cv2 can write."""
    import cv2
    S = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
         "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
         "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    rng = np.random.default_rng(seed)
    out = []
    for (h, w) in sizes:
        for name, sf in S.items():
            for q in qualities:
                for rst in restarts:
                    img = _smooth(h, w, 3, rng) if (h + w + q) % 2 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
                    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf,
                                                         cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                    assert ok
                    out.append(("%dx%d %s q%d rst%d" % (h, w, name, q, rst), buf.tobytes(),
                                np.ascontiguousarray(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)[..., ::-1])))
        img = _smooth(h, w, 1, rng)[..., 0]
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
        out.append(("%dx%d grey" % (h, w), buf.tobytes(), cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)[..., None]))
    return out


# ------------------------------------------------------------------------------------------------ CPU: the oracle
@pytest.mark.parametrize("name", CASES)
def test_oracle_against_golden_libjpeg_files(name):
    data, want = _case(name)
    assert np.array_equal(ojpg.decode_jpeg(data), want)
    assert ojpg.jpeg_shape(data) == want.shape


def test_oracle_against_live_libjpeg_turbo():
    from PIL import Image
    files = _live_files(1, [(64, 64), (45, 67), (16, 16), (17, 33), (8, 5), (3, 3), (9, 2), (1, 9), (1, 1)], qualities=(100, 50))
    for name, data, want in files:
        assert np.array_equal(ojpg.decode_jpeg(data), want), name
        pil = np.array(Image.open(io.BytesIO(data)))
        assert np.array_equal(pil.reshape(want.shape), want), name      # Pillow's libjpeg agrees with cv2's


def test_oracle_reports_out_of_scope_flavours():
    with pytest.raises(ojpg.Unsupported):
        ojpg.decode_jpeg(open(os.path.join(G, "jpeg_progressive.jpg"), "rb").read())
    with pytest.raises(ojpg.DecodeError):
        ojpg.decode_jpeg(b"\x89PNG\r\n\x1a\n")


# ------------------------------------------------------------------------------------ CPU: host half of the C ABI
def test_host_probe_matches_oracle_parse():
    from dl_image_segmentation_b200 import _codec
    files = [(n,) + _case(n) for n in CASES] + _live_files(2, [(45, 67), (9, 2)], qualities=(75,))
    for name, data, want in files:
        st, info = _codec.probe_jpeg(data)
        fr = ojpg.parse_jpeg(data)
        assert st == 0, name
        assert (info.height, info.width, info.components) == want.shape, name
        assert info.scan_off == fr["scan"] and info.restart_interval == fr["restart"] and bool(info.ycc) == fr["ycc"], name
        for c, comp in enumerate(fr["comps"]):
            assert (info.h[c], info.v[c], info.tq[c], info.td[c], info.ta[c]) == (comp["h"], comp["v"], comp["tq"], comp["td"], comp["ta"])
            assert list(info.qt[comp["tq"]]) == list(fr["qt"][comp["tq"]]), name
            for cls, sel in ((0, comp["td"]), (1, comp["ta"])):
                counts, syms = fr["ht"][(cls, sel)]
                assert list(info.huff_counts[cls][sel]) == counts and list(info.huff_syms[cls][sel])[:len(syms)] == syms
        gen = _codec.probe(data)                                          # the format-agnostic header read
        assert (gen.format, gen.status, gen.height, gen.width, gen.samples) == (_codec.FORMAT_JPEG, 0) + want.shape
    assert _codec.probe_jpeg(open(os.path.join(G, "jpeg_progressive.jpg"), "rb").read())[0] == 3
    assert _codec.probe_jpeg(b"\xff\xd8\xff\xe0\x00")[0] == 1
    assert _codec.probe_jpeg(_case(CASES[0])[0][:200])[0] == 1           # cut inside the tables
    assert _codec.probe_jpeg(b"")[0] == 1
    data = bytearray(_case("420_q90")[0])                                 # same file, frame header claiming 40000 x 40000
    sof = data.index(b"\xff\xc0")
    data[sof + 5:sof + 9] = (40000).to_bytes(2, "big") * 2
    assert _codec.probe_jpeg(bytes(data))[0] == 3                        # refused like tf.image.decode_jpeg ("too large")
    with pytest.raises(ojpg.Unsupported):
        ojpg.parse_jpeg(bytes(data))


def test_host_batch_planner_matches_per_file_calls():
    """b2_jpeg_plan_batch = b2_jpeg_probe + b2_jpeg_sizes per file, compacted, with the bytes gathered (no GPU involved)."""
    import ctypes
    from dl_image_segmentation_b200 import _codec
    from dl_image_segmentation_b200._lib import lib
    prog = open(os.path.join(G, "jpeg_progressive.jpg"), "rb").read()
    blobs = [_case(CASES[0])[0], prog, b"", _case(CASES[1])[0], b"junk", _case(CASES[2])[0]] * 9
    n = len(blobs)
    ptrs = (ctypes.c_void_p * n)()
    sizes = np.zeros(n, np.uint64)
    keep = []
    for i, b in enumerate(blobs):
        ptrs[i], sizes[i], k = _codec._ptr_of(b)
        keep.append(k)
    infos, jobs, status = (_codec.JpegInfo * n)(), np.zeros(n, _codec.JPEG_JOB_DTYPE), np.zeros(n, np.int32)
    plan = _codec.JpegPlan()
    for stage in (None, np.zeros(int(sizes.sum()) + 16 * n, np.uint8)):
        assert lib().b2_jpeg_plan_batch(ptrs, sizes.ctypes.data, n, infos, status.ctypes.data, jobs.ctypes.data,
                                        stage.ctypes.data if stage is not None else None, stage.size if stage is not None else 0,
                                        3, ctypes.byref(plan)) == 0
        assert list(status) == [0, 3, 1, 0, 1, 0] * 9 and plan.n_jobs == 27 and plan.filled == (stage is not None)
    ok = [i for i in range(n) if status[i] == 0]
    src = coef = plane = out = 0
    cc, pb, ob = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    for j, i in enumerate(ok):
        st, want = _codec.probe_jpeg(blobs[i])
        assert bytes(infos[j]) == bytes(want)
        assert tuple(jobs[j]) == (src, coef, plane, out, len(blobs[i]), i)
        assert stage[src:src + len(blobs[i])].tobytes() == blobs[i]
        lib().b2_jpeg_sizes(ctypes.byref(want), ctypes.byref(cc), ctypes.byref(pb), ctypes.byref(ob))
        src, coef = (src + len(blobs[i]) + 15) & ~15, coef + cc.value
        plane, out = (plane + pb.value + 15) & ~15, (out + ob.value + 255) & ~255
    assert (plan.stage_bytes, plan.coef_count, plan.plane_bytes, plan.out_bytes) == (src, coef, plane, out)


def _enc_cases(seed):
    rng = np.random.default_rng(seed)
    out = []
    for (h, w) in [(16, 16), (24, 40), (17, 33), (64, 64), (5, 7), (1, 1), (100, 130)]:
        out += [rng.integers(0, 256, (h, w, 3), dtype=np.uint8), _smooth(h, w, 3, rng), _smooth(h, w, 1, rng)]
    return out


def test_encoder_oracle_is_byte_identical_to_libjpeg_turbo():
    """The whole file, header included, against cv2.imencode (libjpeg-turbo) at several qualities; cv2 writes JFIF
    density 1 x 1 without a unit, TensorFlow 300 x 300 dpi — the only bytes that differ."""
    import cv2
    for img in _enc_cases(7):
        for q in (100, 90, 50, 20):
            ok, buf = cv2.imencode(".jpg", img[..., ::-1] if img.shape[2] == 3 else img[..., 0], [cv2.IMWRITE_JPEG_QUALITY, q])
            got = ojenc.encode_jpeg(img, q, density=(0, 1, 1))
            assert got == buf.tobytes(), (img.shape, q)
            tf_like = ojenc.encode_jpeg(img, q)
            assert tf_like[:13] == got[:13] and tf_like[13:18] == b"\x01\x01\x2c\x01\x2c" and tf_like[18:] == got[18:]
            assert np.array_equal(ojpg.decode_jpeg(tf_like), ojpg.decode_jpeg(got))


ENC_CASES = {"rgb_q100": 100, "rgb_q75": 75, "grey_q100": 100}


def _enc_case(name):
    return np.load(os.path.join(G, "jpegenc_%s_in.npy" % name)), open(os.path.join(G, "jpegenc_%s_out.jpg" % name), "rb").read()


@pytest.mark.parametrize("name", sorted(ENC_CASES))
def test_encoder_oracle_against_golden_libjpeg_files(name):
    img, want = _enc_case(name)
    assert ojenc.encode_jpeg(img, ENC_CASES[name], density=(0, 1, 1)) == want


def test_host_header_matches_oracle_header():
    from dl_image_segmentation_b200 import _codec
    for (h, w, c, q) in [(256, 256, 3, 100), (17, 33, 1, 75), (65535, 1, 3, 1), (40, 48, 3, 50)]:
        for dens in ((1, 300, 300), (0, 1, 1)):
            assert _codec.jpeg_header(h, w, c, q, dens) == ojenc.header(h, w, c, ojenc.quant_tables(q), dens)


# ---------------------------------------------------------------------------------------------- GPU: the kernels
@pytest.mark.gpu
def test_gpu_encode_is_byte_identical_to_libjpeg_turbo_and_oracle(dev):
    import cv2
    import torch
    from dl_image_segmentation_b200 import _codec
    imgs = _enc_cases(8) + [np.random.default_rng(1).integers(0, 256, (256, 256, 3), dtype=np.uint8)]
    for q in (100, 60):
        files = _codec.encode_jpeg_arrays([torch.from_numpy(i).to(dev) for i in imgs], quality=q, density=(0, 1, 1), device=dev)
        for img, f in zip(imgs, files):
            ok, buf = cv2.imencode(".jpg", img[..., ::-1] if img.shape[2] == 3 else img[..., 0], [cv2.IMWRITE_JPEG_QUALITY, q])
            assert f == buf.tobytes(), (img.shape, q)
    files = _codec.encode_jpeg_arrays(imgs[:6], device=dev)                         # host arrays in, TensorFlow's header
    for img, f in zip(imgs, files):
        assert f == ojenc.encode_jpeg(img)
    arrays, status, _ = _codec.decode_jpeg_blobs(files, dev)                        # and back through the decoder
    assert not status.any()
    for a, f in zip(arrays, files):
        assert np.array_equal(a.cpu().numpy(), ojpg.decode_jpeg(f))


@pytest.mark.gpu
def test_gpu_encode_matches_golden_libjpeg_files(dev):
    from dl_image_segmentation_b200 import _codec
    imgs, wants = zip(*[_enc_case(n) for n in sorted(ENC_CASES)])
    for img, want, name in zip(imgs, wants, sorted(ENC_CASES)):
        assert _codec.encode_jpeg_arrays([img], quality=ENC_CASES[name], density=(0, 1, 1), device=dev)[0] == want, name


@pytest.mark.gpu
def test_convert_png_to_jpg_translator(dev, tmp_path):
    """images_to_tfrecords_mt(convert_png_to_jpg=True): records carry encode_jpeg(decode_png(file), quality=100) — bytes
    or decoded pixels — for image and label alike (reference :92-95 applies to both)."""
    import cv2
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _img_to_tf_threaded
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from oracle import example_proto as oep
    from oracle import tfrecord as otfr
    rng = np.random.default_rng(9)
    d = tmp_path / "chips"
    (d / "images").mkdir(parents=True)
    (d / "labels").mkdir()
    want = {}
    for i in range(5):
        key = "64:0:10.0:30:%d:%d" % (i, 2 * i)
        img, lab = _smooth(40, 56, 3, rng), (rng.integers(0, 4, (40, 56)) * 60).astype(np.uint8)
        for sub, arr in (("images", img[..., ::-1]), ("labels", lab)):
            ok, buf = cv2.imencode(".png", np.ascontiguousarray(arr))
            (d / sub / (key.replace(":", "#") + ".png")).write_bytes(buf.tobytes())
        want[key.encode()] = (ojenc.encode_jpeg(img), ojenc.encode_jpeg(lab))
    coder = _img_to_tf_threaded.ImageCoder(device=dev)
    png = (d / "images" / "64#0#10.0#30#1#2.png").read_bytes()
    assert coder.png_to_jpeg(png) == want[b"64:0:10.0:30:1:2"][0]
    for as_array in (False, True):
        out = tmp_path / ("out%d" % as_array)
        out.mkdir()
        pkg.images_to_tfrecords_mt("j", str(d), str(out), 1, num_threads=1, convert_png_to_jpg=True, store_as_array=as_array)
        (shard,) = sorted(os.listdir(out))
        recs = otfr.read_records(open(os.path.join(out, shard), "rb").read())
        assert len(recs) == 5
        for rec in recs:
            if as_array:
                img, tgt, ident = tr.parse_8bit_array_proto(rec, device=dev)
                fi, fl = want[ident]
                assert np.array_equal(img.cpu().numpy(), ojpg.decode_jpeg(fi)) and np.array_equal(tgt.cpu().numpy(), ojpg.decode_jpeg(fl)[..., 0])
            else:
                img_bytes, dims, tgt_bytes, tdims, ident = oep._parse_byteslist_proto(rec)
                assert (bytes(img_bytes), bytes(tgt_bytes)) == want[bytes(ident)] and tuple(int(x) for x in dims) == (40, 56, 3)


@pytest.mark.gpu
def test_gpu_decode_matches_golden_and_oracle(dev):
    from dl_image_segmentation_b200 import _codec
    blobs, wants = zip(*[_case(n) for n in CASES])
    arrays, status, infos = _codec.decode_jpeg_blobs(list(blobs), dev)
    assert list(status) == [0] * len(CASES)
    for name, a, want, data in zip(CASES, arrays, wants, blobs):
        got = a.cpu().numpy()
        assert got.dtype == np.uint8 and got.shape == want.shape, name
        assert np.array_equal(got, want), name                            # libjpeg-turbo's pixels
        assert np.array_equal(got, ojpg.decode_jpeg(data)), name          # the oracle's


@pytest.mark.gpu
def test_gpu_decode_sweep_against_live_libjpeg_turbo(dev):
    from dl_image_segmentation_b200 import _codec
    files = _live_files(3, [(64, 64), (45, 67), (17, 33), (8, 5), (3, 3), (9, 2), (1, 9), (1, 1), (256, 256)], qualities=(100, 40))
    arrays, status, _ = _codec.decode_jpeg_blobs([f[1] for f in files], dev)
    assert not status.any()
    for (name, _, want), a in zip(files, arrays):
        assert np.array_equal(a.cpu().numpy(), want), name


@pytest.mark.gpu
def test_gpu_decode_error_behaviour(dev):
    """Out-of-scope and damaged files are data, not exceptions: a status per chip, the neighbours unaffected."""
    from dl_image_segmentation_b200 import _codec
    good, want = _case("420_q90")
    prog = open(os.path.join(G, "jpeg_progressive.jpg"), "rb").read()
    cut = good[:len(good) - 300]                                          # entropy data truncated
    arrays, status, _ = _codec.decode_jpeg_blobs([good, prog, cut, b"not a jpeg", good], dev)
    assert list(status) == [0, 3, 2, 1, 0]
    assert arrays[1] is None and arrays[2] is None and arrays[3] is None
    assert np.array_equal(arrays[0].cpu().numpy(), want) and np.array_equal(arrays[4].cpu().numpy(), want)
    # the format-agnostic entry point routes .jpg blobs here and everything else to the TIFF / PNG decoders
    import cv2
    ok, png = cv2.imencode(".png", want[..., ::-1])
    (a, b), st = _codec.decode_blobs([png.tobytes(), good], dev, png_as_tf=True)
    assert list(st) == [0, 0] and np.array_equal(a.cpu().numpy(), want) and np.array_equal(b.cpu().numpy(), want)


@pytest.mark.gpu
def test_gpu_decode_of_damaged_files_is_contained(dev):
    """Bit flips and cuts in the entropy-coded segment: every file gets a status, no file disturbs its neighbours, a file
    reported as fine has exactly the oracle's pixels, and whatever the oracle rejects (bad code, bad run, missing
    restart marker) the GPU rejects too."""
    from dl_image_segmentation_b200 import _codec
    rng = np.random.default_rng(11)
    blobs = []
    for name in ("420_q90", "422_q75_rst", "411_q60_rst", "grey_q90"):
        data, _ = _case(name)
        scan = ojpg.parse_jpeg(data)["scan"]
        for _ in range(30):
            b = bytearray(data)
            if rng.random() < 0.25:
                b = b[:int(rng.integers(scan + 1, len(b)))]
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(scan, len(b)))] = int(rng.integers(0, 256))
            blobs.append(bytes(b))
        blobs.append(data)                                                  # an intact neighbour in every group
    arrays, status, _ = _codec.decode_jpeg_blobs(blobs, dev)
    assert set(status.tolist()) <= {0, 2} and (status == 0).sum() >= 4 and (status == 2).sum() >= 10
    for b, a, st in zip(blobs, arrays, status):
        try:
            want = ojpg.decode_jpeg(b)
        except ojpg.DecodeError:
            assert st == 2
            continue
        if st == 0:
            assert np.array_equal(a.cpu().numpy(), want)


@pytest.mark.gpu
def test_mp_translator_and_loaders_on_jpg_chips(dev, tmp_path):
    """images_to_tfrecords_mp(file_ext='jpg') (rasterio -> GDAL's JPEG driver = the same libjpeg) and the single-chip
    loaders of both modules."""
    import cv2
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _img_to_tf_mp, _img_to_tf_threaded
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from oracle import tfrecord as otfr
    rng = np.random.default_rng(6)
    d = tmp_path / "chips"
    (d / "images").mkdir(parents=True)
    (d / "labels").mkdir()
    files = {}
    for i in range(5):
        name = "64#0#10.0#30#%d#%d.jpg" % (i, i + 3)
        img, lab = _smooth(48, 40, 3, rng), (rng.integers(0, 2, (48, 40)) * 255).astype(np.uint8)
        for sub, arr in (("images", img), ("labels", lab)):
            ok, buf = cv2.imencode(".jpg", arr, [cv2.IMWRITE_JPEG_QUALITY, 90])
            (d / sub / name).write_bytes(buf.tobytes())
        files[name[:-4].replace("#", ":").encode()] = ((d / "images" / name).read_bytes(), (d / "labels" / name).read_bytes())
    out = tmp_path / "out"
    out.mkdir()
    pkg.images_to_tfrecords_mp("jpgs", str(d), str(out), 1, num_proc=1, file_ext="jpg", store_as_array=True)
    (shard,) = sorted(os.listdir(out))
    recs = otfr.read_records(open(os.path.join(out, shard), "rb").read())
    assert len(recs) == 5
    for rec in recs:
        img, tgt, ident = tr.parse_8bit_array_proto(rec, device=dev)
        fi, fl = files[ident]
        assert np.array_equal(img.cpu().numpy(), ojpg.decode_jpeg(fi)) and np.array_equal(tgt.cpu().numpy(), ojpg.decode_jpeg(fl)[..., 0])
    path = str(d / "images" / "64#0#10.0#30#2#5.jpg")
    want = ojpg.decode_jpeg(open(path, "rb").read())
    arr, h, w, b, key = _img_to_tf_mp.load_image_rasterio(path, device=dev)
    assert (h, w, b, key) == (48, 40, 3, "64:0:10.0:30:2:5") and np.array_equal(arr.cpu().numpy(), want)
    raw, h, w, b, key = _img_to_tf_mp.load_image_rasterio(path, decode=False, device=dev)
    assert raw == open(path, "rb").read() and (h, w, b) == (48, 40, 3)
    coder = _img_to_tf_threaded.ImageCoder(device=dev)
    arr, h, w, b, key = _img_to_tf_threaded._process_image(path, coder, decode=True)
    assert (h, w, b, key) == (48, 40, 3, "64:0:10.0:30:2:5") and np.array_equal(arr.cpu().numpy(), want)


@pytest.mark.gpu
def test_threaded_translator_on_jpg_chips(dev, tmp_path):
    """images_to_tfrecords_mt over a folder of .jpg chips: raw file bytes + header dims in the records (store_as_array
    False, reference :113-121) or the decoded arrays (True), next to the oracle's worker loop pieces."""
    import cv2
    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from oracle import tfrecord as otfr
    rng = np.random.default_rng(5)
    d = tmp_path / "chips"
    (d / "images").mkdir(parents=True)
    (d / "labels").mkdir()
    want = {}
    for i in range(6):
        key = "64:0:10.0:30:%d:%d" % (i, 7 * i)
        img = _smooth(32, 40, 3, rng)
        lab = (rng.integers(0, 3, (32, 40)) * 100).astype(np.uint8)
        for sub, arr in (("images", img), ("labels", lab)):
            ok, buf = cv2.imencode(".jpg", arr, [cv2.IMWRITE_JPEG_QUALITY, 95])
            (d / sub / (key.replace(":", "#") + ".jpg")).write_bytes(buf.tobytes())
        want[key.encode()] = tuple((d / s / (key.replace(":", "#") + ".jpg")).read_bytes() for s in ("images", "labels"))
    (d / "images" / "64#0#10.0#30#9#9.jpg").write_bytes(b"\xff\xd8\xff garbage")        # skipped, as the reference would
    (d / "labels" / "64#0#10.0#30#9#9.jpg").write_bytes(want[b"64:0:10.0:30:0:0"][1])
    for as_array in (False, True):
        out = tmp_path / ("out%d" % as_array)
        out.mkdir()
        pkg.images_to_tfrecords_mt("jpgs", str(d), str(out), 1, num_threads=1, store_as_array=as_array)
        (shard,) = sorted(os.listdir(out))
        recs = otfr.read_records(open(os.path.join(out, shard), "rb").read())
        assert len(recs) == 6
        for rec in recs:
            if as_array:
                img, tgt, ident = tr.parse_8bit_array_proto(rec, device=dev)
                fi, fl = want[ident]
                assert np.array_equal(img.cpu().numpy(), ojpg.decode_jpeg(fi))
                assert np.array_equal(tgt.cpu().numpy(), ojpg.decode_jpeg(fl)[..., 0])
            else:
                img, tgt, ident = tr.parse_encoded_rgb_img_proto(rec, device=dev)     # tf.io.decode_image on both blobs
                fi, fl = want[ident]
                assert np.array_equal(img.cpu().numpy(), ojpg.decode_jpeg(fi))
                assert np.array_equal(tgt.cpu().numpy(), ojpg.decode_jpeg(fl))
