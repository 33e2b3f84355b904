"""-m gpu: the GeoTIFF chip writer (tile split + TIFF-LZW encode on the GPU, IFD on the host) — the save step of
create_chips_for_tile (reference _descartes_img_chips.py:781-797)."""
import io
import os

import numpy as np
import pytest

import synthetic as syn
from oracle import imagecodecs as oic

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(8)
    yield b""
    yield b"a"
    yield b"ab" * 3
    yield bytes(300000)                                                    # one long run: the table fills again and again
    yield rng.integers(0, 256, 100001, dtype=np.uint8).tobytes()           # incompressible: widest codes, frequent Clears
    yield bytes(range(256)) * 700
    yield rng.integers(0, 4, 262144, dtype=np.uint8).tobytes()
    img, lab, _ = syn.cfg3_chip(3, size=256)
    yield img.tobytes()
    yield lab.tobytes()


@pytest.mark.parametrize("restart", [0, 1024, 16, 400])
def test_lzw_encode_is_byte_identical_with_the_cpu_encoder_and_round_trips(dev, restart):
    """restart 0: the classic stream (Clear only when the table is full); otherwise a Clear every `restart` input bytes —
    the pieces are encoded independently on the device and shifted together; 1024 is what the writer uses."""
    import torch

    from dl_image_segmentation_b200 import _geotiff
    msgs = list(_cases())
    if restart:
        rng = np.random.default_rng(restart)
        msgs += [bytes(restart), bytes(restart + 1), bytes(restart - 1), rng.integers(0, 256, 3 * restart, dtype=np.uint8).tobytes()]
    buf = bytearray()
    for m in msgs:
        buf += m + bytes((-len(m)) % 16)
    raw = torch.from_numpy(np.frombuffer(bytes(buf) or b"\0" * 16, np.uint8).copy()).to(dev)
    got = _geotiff.lzw_encode_tiles(raw, [len(m) for m in msgs], dev, restart=restart)
    for m, g in zip(msgs, got):
        assert g == syn.lzw_encode(m, restart), len(m)
        assert oic.lzw_decode(g, len(m)) == m                              # oracle decoder (TIFF 6.0 section 13)


def test_lzw_restart_rejects_bad_intervals(dev):
    import torch

    from dl_image_segmentation_b200 import _geotiff
    from dl_image_segmentation_b200._lib import B2Error
    raw = torch.zeros(64, dtype=torch.uint8, device=dev)
    for bad in (8, 1000, 2048):
        with pytest.raises(B2Error):
            _geotiff.lzw_encode_tiles(raw, [40], dev, restart=bad)


@pytest.mark.parametrize("shape,dtype,nodata", [((512, 512, 4), np.uint16, None), ((512, 512), np.uint8, 255),
                                                ((300, 500, 3), np.uint8, None), ((70, 33, 8), np.uint16, None),
                                                ((256, 256, 2), np.float64, None), ((100, 257, 1), np.int16, -1)])
def test_geotiff_files_match_the_fixture_writer_and_libtiff_reads_them(dev, shape, dtype, nodata):
    import cv2
    import torch

    from dl_image_segmentation_b200 import _codec, _geotiff
    rng = np.random.default_rng(sum(shape))
    if np.issubdtype(dtype, np.floating):
        arr = (rng.integers(0, 20000, shape) * 0.5).astype(dtype)          # what the median composite produces
    else:
        info = np.iinfo(dtype)
        smooth = (syn.smooth_field(rng, shape[0], shape[1])[..., None] * 4000).astype(np.int64)
        arr = np.clip(smooth + rng.integers(0, 40, shape if len(shape) == 3 else shape + (1,)), info.min, info.max).astype(dtype)
        arr = arr.reshape(shape)
    (blob,) = _geotiff.encode_geotiffs([torch.from_numpy(arr).to(dev) if dtype != np.uint16 else torch.from_numpy(arr.view(np.int16)).to(dev).view(torch.uint16)],
                                       nodata=nodata, device=dev)
    want = syn.tiff_bytes(arr, tile=256, nodata=nodata, lzw_restart=_geotiff.LZW_RESTART)
    assert blob == want                                                    # same bytes as the fixture writer
    a3 = arr if arr.ndim == 3 else arr[:, :, None]
    np.testing.assert_array_equal(oic.decode_image(blob), a3)              # oracle decoder
    (dec,), (st,) = _codec.decode_blobs([blob], device=dev)                # K1 decoder
    assert st == 0
    np.testing.assert_array_equal(dec.cpu().numpy().view(dtype).reshape(a3.shape), a3)
    info = _codec.probe(blob)
    assert _codec.georef_strings(info) == ("[499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0]", "EPSG:32643")
    if nodata is not None:
        assert info.has_nodata == 1 and info.nodata == float(nodata)
    if a3.shape[2] == 1 or (dtype == np.uint8 and a3.shape[2] == 3):       # libtiff through OpenCV: single band or RGB
        cvd = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
        cvd = cvd if cvd.ndim == 3 else cvd[:, :, None]
        if a3.shape[2] == 3:
            cvd = cvd[:, :, ::-1]                                          # OpenCV returns B,G,R
        np.testing.assert_array_equal(cvd, a3)


def test_write_chip_pair_feeds_the_translator(dev, tmp_path):
    """composite -> chip files on disk -> images_to_tfrecords_mp: the loop the reference's notebooks run."""
    import contextlib

    import dl_image_segmentation_b200 as pkg
    from oracle import tfrecord as otfr
    from oracle import example_proto as oep
    keys = []
    for i in range(3):
        stack, valid = syn.cfg4_tile(i, T=5, H=96, W=80, B=4)
        med = pkg.median_composite(stack, valid, device=dev)               # float64 + mask, as np.ma.median
        img = med.filled(0)
        lab = syn.label_field(np.random.default_rng(i), 96, 80)
        key = "64:16:10.0:43:%d:7" % i
        f_img, f_lab = pkg.write_chip_pair(img, lab, str(tmp_path), key, label_ndv=255, device=dev)
        assert os.path.basename(f_img) == key.replace(":", "#") + ".tif"
        keys.append(key)
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("c", str(tmp_path), str(tmp_path / "out"), 1, num_proc=1, file_ext="tif")
    recs = otfr.read_records(open(tmp_path / "out" / "c-00000-of-00001", "rb").read(), verify=True)
    assert sorted(oep.parse_example(r)["identifier"][1][0].decode() for r in recs) == sorted(keys)
    f = oep.parse_example(recs[0])
    assert f["image/channels"][1] == [4] and f["image/image_data"][0] == "float"


def test_lzw_restart_large_batch_spans_several_scratch_groups(dev):
    """More pieces than one 256 MiB scratch group holds (the library encodes the tiles group by group): every stream must
    still be the fixture encoder's, including tiles of different lengths side by side."""
    import torch

    from dl_image_segmentation_b200 import _geotiff
    rng = np.random.default_rng(5)
    base = [rng.integers(0, 256, 524288, dtype=np.uint8).tobytes(), bytes(65536), rng.integers(0, 3, 70001, dtype=np.uint8).tobytes()]
    want = [syn.lzw_encode(m, 1024) for m in base]
    n = 720                                                                # 240 x (512 + 64 + 69) pieces x 1.5 KiB > 256 MiB
    msgs = [base[i % 3] for i in range(n)]
    raw = torch.from_numpy(np.frombuffer(b"".join(m + bytes((-len(m)) % 16) for m in msgs), np.uint8).copy()).to(dev)
    got = _geotiff.lzw_encode_tiles(raw, [len(m) for m in msgs], dev, restart=1024)
    for i, g in enumerate(got):
        assert g == want[i % 3], i


def test_writer_on_a_second_device(dev):
    """The restart encoder opts in to 224 KiB of dynamic shared memory per DEVICE: a process that encodes on GPU 0 and then
    on GPU 1 (one context per GPU, as run_workers does) must not rely on a once-per-process setting."""
    import torch

    from dl_image_segmentation_b200 import _geotiff
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    img, _, _ = syn.cfg3_chip(5, size=256)
    want = syn.tiff_bytes(img, tile=256, lzw_restart=_geotiff.LZW_RESTART)
    for d in (0, 1, 0):
        t = torch.from_numpy(img.view(np.int16)).to("cuda:%d" % d).view(torch.uint16)
        (blob,) = _geotiff.encode_geotiffs([t], device=torch.device("cuda", d))
        assert blob == want, d
