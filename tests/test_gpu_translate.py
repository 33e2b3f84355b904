"""-m gpu: the drop-in translators (images_to_tfrecords_mp / _mt) against the oracle's restatement of the reference
worker loop: same shard files, byte for byte, including skipped chips, empty shards and every identifier flavour."""
import contextlib
import io
import os

import numpy as np
import pytest

import synthetic as syn
from oracle import partition as opart
from oracle import tfrecord as otfr
from oracle import translate as otr

pytestmark = pytest.mark.gpu


def _write_dataset(root, kind, n, size, corrupt=(), mismatch=()):
    os.makedirs(root / "images")
    os.makedirs(root / "labels")
    ext = "png" if kind == "png" else "tif"
    for i in range(n):
        if kind == "png":
            img, lab, key = syn.cfg1_chip(i, size=size)
            a, b = syn.png_bytes(img), syn.png_bytes(lab)
        else:
            img, lab, key = syn.cfg3_chip(i, size=size)
            a, b = syn.tiff_bytes(img, tile=32), syn.tiff_bytes(lab, tile=32, nodata=255)
        name = key.replace(":", "#")
        if i in corrupt:
            a = a[:len(a) // 2]
        (root / "images" / (name + "." + ext)).write_bytes(a)
        lname = name + ("x" if i in mismatch else "")
        (root / "labels" / (lname + "." + ext)).write_bytes(b)
    return ext


def _same_shards(a, b):
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert fa == fb and fa
    for f in fa:
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f
    return fa


@pytest.mark.parametrize("kind,store_as_array", [("png", True), ("png", False), ("tif", True), ("tif", False)])
def test_images_to_tfrecords_mp_matches_reference_worker_loop(dev, tmp_path, kind, store_as_array):
    import dl_image_segmentation_b200 as pkg
    n = 23
    ext = _write_dataset(tmp_path, kind, n, 48 if kind == "png" else 64, corrupt=(5,), mismatch=(11,))
    out_g, out_c = str(tmp_path / "g"), str(tmp_path / "c")
    with contextlib.redirect_stdout(io.StringIO()) as log:
        wrote = pkg.images_to_tfrecords_mp("t", str(tmp_path), out_g, 6, num_proc=3, file_ext=ext, store_as_array=store_as_array)
    want = otr.images_to_tfrecords("t", str(tmp_path), out_c, 6, num_proc=3, file_ext=ext, store_as_array=store_as_array, n_jobs=1)
    files = _same_shards(out_g, out_c)
    assert len(files) == 6 and files[0] == "t-00000-of-00006"
    skipped = 1 + (1 if store_as_array else 0)      # the key mismatch always skips; the truncated file only fails when decoded
    assert sum(wrote) == want
    assert log.getvalue().count("SKIPPED: Unexpected eror while decoding") >= skipped
    total = sum(len(otfr.read_records(open(os.path.join(out_g, f), "rb").read(), verify=True)) for f in files)
    assert total == want and total <= n - 1


def test_two_gpus_in_one_process_write_the_same_shards(dev, tmp_path):
    """the shard files do not depend on how many GPUs wrote them.  num_proc=4 workers on 2 GPUs (two
    concurrent host threads, two workers each) vs the same call confined to one GPU vs the CPU restatement."""
    import torch

    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _translate
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ext = _write_dataset(tmp_path, "png", 41, 48)
    out2, out1, out_c = str(tmp_path / "g2"), str(tmp_path / "g1"), str(tmp_path / "c")
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("t", str(tmp_path), out2, 8, num_proc=4, file_ext=ext)
        assert sorted({d for _, d in _translate.my_workers(4)}) == [0, 1]
        real = torch.cuda.device_count
        torch.cuda.device_count = lambda: 1                                  # same call, one GPU
        try:
            pkg.images_to_tfrecords_mp("t", str(tmp_path), out1, 8, num_proc=4, file_ext=ext)
        finally:
            torch.cuda.device_count = real
    otr.images_to_tfrecords("t", str(tmp_path), out_c, 8, num_proc=4, file_ext=ext, n_jobs=1)
    assert len(_same_shards(out2, out_c)) == 8 and len(_same_shards(out1, out_c)) == 8


def test_georeferenced_identifiers_and_more_shards_than_chips(dev, tmp_path):
    """dltile_from_filename=False (identifier = name|geotransform|crs) and shards that stay empty."""
    import dl_image_segmentation_b200 as pkg
    ext = _write_dataset(tmp_path, "tif", 3, 32)
    out_g, out_c = str(tmp_path / "g"), str(tmp_path / "c")
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("e", str(tmp_path), out_g, 4, num_proc=2, file_ext=ext, dltile_from_filename=False)
    otr.images_to_tfrecords("e", str(tmp_path), out_c, 4, num_proc=2, file_ext=ext, dltile_from_filename=False, n_jobs=1)
    files = _same_shards(out_g, out_c)
    assert len(files) == 4
    ids = []
    from oracle import example_proto as oep
    for f in files:
        for r in otfr.read_records(open(os.path.join(out_g, f), "rb").read(), verify=True):
            ids.append(oep.parse_example(r)["identifier"][1][0].decode())
    assert len(ids) == 3 and all(i.endswith("|[499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0]|EPSG:32643") for i in ids)


@pytest.mark.parametrize("store_as_array", [False, True])
def test_images_to_tfrecords_mt_png(dev, tmp_path, store_as_array):
    import dl_image_segmentation_b200 as pkg
    _write_dataset(tmp_path, "png", 10, 40)
    out_g, out_c = str(tmp_path / "g"), str(tmp_path / "c")
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mt("m", str(tmp_path), out_g, 2, num_threads=2, store_as_array=store_as_array)
    otr.images_to_tfrecords("m", str(tmp_path), out_c, 2, num_proc=2, file_ext="png", store_as_array=store_as_array, n_jobs=1)
    _same_shards(out_g, out_c)


@pytest.mark.parametrize("kind,store_as_array", [("png", True), ("tif", True), ("tif", False)])
def test_pipelined_worker_mixes_clean_and_irregular_batches(dev, tmp_path, kind, store_as_array):
    """The worker pipeline with many small decode batches: clean batches take the vectorised path (one plan call, one build
    launch per batch), a batch with a damaged chip or a key mismatch falls back to the chip-by-chip path, and batches
    straddle shard boundaries — the shard files must still be the oracle's, byte for byte, with the reference's progress
    and SKIPPED lines."""
    from dl_image_segmentation_b200 import _translate
    n = 61
    ext = _write_dataset(tmp_path, kind, n, 40 if kind == "png" else 48, corrupt=(20,), mismatch=(33, 34))
    imgs, lbls = opart.find_image_files(str(tmp_path), ext)
    out_g, out_c = str(tmp_path / "g"), str(tmp_path / "c")
    ranges = _translate.worker_ranges(n, 1)
    key = lambda p, info=None: _translate.tile_key_from_path(p, True)
    with contextlib.redirect_stdout(io.StringIO()) as log:
        wrote = _translate.run_worker(0, ranges, "t", imgs, lbls, out_g, 5, key, store_as_array, progress_every=10,
                                      batch_pairs=8, path_key=lambda p: _translate.tile_key_from_path(p, True))
    want = otr.images_to_tfrecords("t", str(tmp_path), out_c, 5, num_proc=1, file_ext=ext, store_as_array=store_as_array, n_jobs=1)
    _same_shards(out_g, out_c)
    assert wrote == want
    text = log.getvalue()
    assert text.count("SKIPPED: Unexpected eror while decoding") == n - want
    assert text.count("Processed ") == want // 10 and text.count("Wrote ") == 5 + 1
    # and again straight away: the pinned staging sets, write-back buffers and the reader are reused
    with contextlib.redirect_stdout(io.StringIO()):
        assert _translate.run_worker(0, ranges, "t", imgs, lbls, str(tmp_path / "g2"), 5, key, store_as_array, batch_pairs=16,
                                     path_key=lambda p: _translate.tile_key_from_path(p, True)) == want
    _same_shards(str(tmp_path / "g2"), out_c)


@pytest.mark.gpu
def test_large_pieces_written_through_shared_mapping(dev, tmp_path, monkeypatch):
    """Pieces above _MAPPED_WRITE_MIN are written as a pwrite head plus mapped copies of the rest (run_worker.write_back);
    with the threshold lowered every piece of this small job takes that path, over several batches per shard file, and the
    shards must not change by a byte."""
    from dl_image_segmentation_b200 import _translate
    monkeypatch.setattr(_translate, "_MAPPED_WRITE_MIN", 4096)
    n = 40
    ext = _write_dataset(tmp_path, "tif", n, 48)
    imgs, lbls = opart.find_image_files(str(tmp_path), ext)
    out_g, out_c = str(tmp_path / "g"), str(tmp_path / "c")
    key = lambda p, info=None: _translate.tile_key_from_path(p, True)
    with contextlib.redirect_stdout(io.StringIO()):
        wrote = _translate.run_worker(0, _translate.worker_ranges(n, 1), "t", imgs, lbls, out_g, 2, key, True, batch_pairs=6,
                                      path_key=lambda p: _translate.tile_key_from_path(p, True))
    want = otr.images_to_tfrecords("t", str(tmp_path), out_c, 2, num_proc=1, file_ext=ext, store_as_array=True, n_jobs=1)
    assert wrote == want == n
    _same_shards(out_g, out_c)


@pytest.mark.parametrize("store_as_array", [False, True])
@pytest.mark.parametrize("convert", [False, True])
def test_threaded_translator_modes_against_the_oracle_loop(dev, tmp_path, store_as_array, convert):
    """images_to_tfrecords_mt in its four modes on a clean folder (every batch takes the batched path: decode-to-validate
    beside the upload of the files; device-side JPEG encode and file assembly; records from the re-decoded JPEG files) and on
    a clean folder of .jpg chips (JPEG plan on the read-ahead thread): shards byte-identical with the oracle's restatement of
    the reference's worker loop, which decodes and encodes with libjpeg-turbo itself."""
    import cv2
    import dl_image_segmentation_b200 as pkg
    os.environ["B2_ORACLE_JPEG"] = "libjpeg"
    try:
        _write_dataset(tmp_path / "png", "png", 23, 48)
        jd = tmp_path / "jpg"
        os.makedirs(jd / "images")
        os.makedirs(jd / "labels")
        for i in range(23):
            img, lab, key = syn.cfg1_chip(i, size=48)
            for sub, arr in (("images", np.ascontiguousarray(img[..., ::-1])), ("labels", lab)):
                (jd / sub / (key.replace(":", "#") + ".jpg")).write_bytes(cv2.imencode(".jpg", arr, [cv2.IMWRITE_JPEG_QUALITY, 92])[1].tobytes())
        for folder in ("png", "jpg"):
            if folder == "jpg" and convert:
                continue                                                     # convert_png_to_jpg leaves .jpg chips alone: same job
            root = str(tmp_path / folder)
            out_g, out_c = os.path.join(root, "g"), os.path.join(root, "c")
            with contextlib.redirect_stdout(io.StringIO()):
                pkg.images_to_tfrecords_mt("m", root, out_g, 3, num_threads=1, store_as_array=store_as_array, convert_png_to_jpg=convert)
                want = otr.images_to_tfrecords_mt("m", root, out_c, 3, num_threads=1, store_as_array=store_as_array,
                                                  convert_png_to_jpg=convert)
            assert want == 23
            _same_shards(out_g, out_c)
    finally:
        os.environ.pop("B2_ORACLE_JPEG", None)
