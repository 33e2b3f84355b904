"""Whole-job restatements used as the CPU baseline (oracle; test infrastructure only).

``images_to_tfrecords``  = ``process_dataset_mp`` ``_img_to_tf_mp.py:233-275`` with the worker loop
``_process_image_files_mp_worker :78-157`` (per shard: writer, per chip: two ``load_image_rasterio``
calls ``:128-131``, key assert ``:132``, skip on exception ``:133-136``, ``convert_to_example`` ``:138``,
``writer.write(SerializeToString())`` ``:141``) and the same joblib fan-out ``:180``.
``parse_shards``         = ``TFRecordDataset(...).map(parse_fn, 8)`` ``parse_tfrecords.ipynb`` cells 4, 30
followed by the north-star cast / normalise / one-hot.
"""
import os

import numpy as np

from . import example_proto, imagecodecs, normalise, partition, tfrecord


def load_image(path, parse_dltile_filename=True, decode=True):
    """load_image_rasterio (_img_to_tf_mp.py:22-75) -> (data, h, w, bands, tile_key)."""
    with open(path, "rb") as f:
        blob = f.read()
    if decode:
        arr = imagecodecs.decode_image(blob, png_as_tf=False)        # rasterio -> GDAL's PNG driver
        h, w, b = arr.shape
        data = arr
    else:
        h, w, b = imagecodecs.image_shape(blob)
        data = blob
    if parse_dltile_filename:
        return data, h, w, b, partition.tile_key(path, True)
    gt_str, crs_str = imagecodecs.georef_strings(blob)                                  # :49-50
    return data, h, w, b, "|".join((os.path.basename(path), gt_str, crs_str))         # :63-67


def build_record(img_path, lbl_path, dltile_from_filename=True, store_as_array=True):
    ib, ih, iw, ibands, ikey = load_image(img_path, dltile_from_filename, store_as_array)
    lb, lh, lw, _, lkey = load_image(lbl_path, dltile_from_filename, store_as_array)
    assert ikey == lkey
    ex = example_proto.convert_to_example(ib, lb, ih, iw, ibands, lh, lw, ikey)
    return ex.SerializeToString(deterministic=True)


def _worker(proc_index, ranges, name, img_files, lbl_files, out_dir, num_shards, dltile, store_as_array, quiet=True):
    num_proc = len(ranges)
    assert not num_shards % num_proc
    per = num_shards // num_proc
    sr = np.linspace(ranges[proc_index][0], ranges[proc_index][1], per + 1).astype(int)
    written = 0
    for s in range(per):
        shard = proc_index * per + s
        path = os.path.join(out_dir, partition.shard_name(name, shard, num_shards))
        os.makedirs(out_dir, exist_ok=True)
        with tfrecord.TFRecordWriter(path) as w:
            for i in range(int(sr[s]), int(sr[s + 1])):
                try:
                    rec = build_record(img_files[i], lbl_files[i], dltile, store_as_array)
                except Exception as e:                      # :133-136
                    if not quiet:
                        print(e)
                        print("SKIPPED: Unexpected eror while decoding %s." % img_files[i])
                    continue
                w.write(rec)
                written += 1
    return written


def images_to_tfrecords(name, directory, out_directory, num_shards, num_proc=None, dltile_from_filename=True,
                        file_ext="tif", store_as_array=True, n_jobs=None, limit=None):
    """n_jobs = OS processes actually used (the reference uses num_proc of them)."""
    from joblib import Parallel, delayed
    if not num_proc:
        num_proc = num_shards
    imgs, lbls = partition.find_image_files(directory, file_ext)
    if limit is not None:
        imgs, lbls = imgs[:limit], lbls[:limit]
    ranges = partition.worker_ranges(len(imgs), num_proc)
    args = [(p, ranges, name, imgs, lbls, out_directory, num_shards, dltile_from_filename, store_as_array)
            for p in range(len(ranges))]
    res = Parallel(n_jobs=n_jobs or num_proc)(__import__("joblib").delayed(_worker)(*a) for a in args)
    return sum(res)


# ------------------------------------------------------------------ threaded flavour (_img_to_tf_threaded.py)
def process_image_mt(path, parse_dltile_filename=True, png_to_jpg=False, decode=False):
    """_process_image (_img_to_tf_threaded.py:75-121): read; PNG (substring test :72) -> decode_png, or transcode to
    JPEG q=100 then decode_jpeg (:92-100); anything else -> decode_jpeg (:103); assert 3-D and <= 3 bands (:107-112);
    key = file name only (:113-116); return the decoded array or the (possibly transcoded) file bytes (:118-121)."""
    from . import jpegcodec, jpegenc
    with open(path, "rb") as f:
        data = f.read()
    if ".png" in path:
        image = imagecodecs.decode_png(data, as_tf=True)                    # tf.image.decode_png
        if png_to_jpg:
            if image.shape[2] not in (1, 3):
                raise ValueError("encode_jpeg: image must have 1 or 3 channels")
            data = jpegenc.encode_jpeg(image, quality=100, density=(1, 300, 300))   # tf.image.encode_jpeg defaults
            image = jpegcodec.decode_jpeg(data)
    else:
        image = jpegcodec.decode_jpeg(data)
    assert image.ndim == 3
    h, w, b = image.shape
    assert b <= 3
    return (image if decode else data), h, w, b, partition.tile_key(path, parse_dltile_filename)


def build_record_mt(img_path, lbl_path, dltile_from_filename=True, png_to_jpg=False, store_as_array=False):
    ib, ih, iw, ibands, ikey = process_image_mt(img_path, dltile_from_filename, png_to_jpg, store_as_array)
    lb, lh, lw, _, lkey = process_image_mt(lbl_path, dltile_from_filename, png_to_jpg, store_as_array)
    assert ikey == lkey
    return example_proto.convert_to_example(ib, lb, ih, iw, ibands, lh, lw, ikey).SerializeToString(deterministic=True)


def images_to_tfrecords_mt(name, directory, out_directory, num_shards, num_threads=None, dltile_from_filename=True,
                           convert_png_to_jpg=False, store_as_array=False):
    """process_dataset_multithreaded (_img_to_tf_threaded.py:321-350) with the worker loop :136-219, run serially."""
    if not num_threads:
        num_threads = num_shards
    assert not num_shards % num_threads
    imgs, lbls = partition.find_image_files(directory, "png", also_jpg=True)
    ranges = partition.worker_ranges(len(imgs), num_threads)
    per = num_shards // num_threads
    written = 0
    os.makedirs(out_directory, exist_ok=True)
    for t in range(len(ranges)):
        sr = np.linspace(ranges[t][0], ranges[t][1], per + 1).astype(int)
        for s in range(per):
            with tfrecord.TFRecordWriter(os.path.join(out_directory, partition.shard_name(name, t * per + s, num_shards))) as w:
                for i in range(int(sr[s]), int(sr[s + 1])):
                    try:
                        rec = build_record_mt(imgs[i], lbls[i], dltile_from_filename, convert_png_to_jpg, store_as_array)
                    except Exception:                           # :196-199
                        continue
                    w.write(rec)
                    written += 1
    return written


def parse_records_norm_onehot(records, mean, std, num_classes):
    """records: list of Example bytes (uint8 arrays stored as BytesList) -> (N,H,W,C) f32, (N,H,W,K) f32.

    Same values as normalise.normalise / normalise.one_hot (float32 subtract + IEEE divide; label == k), written
    straight into the batch arrays so that the CPU baseline is not handicapped by temporaries."""
    mean = np.asarray(mean, dtype=np.float32)
    std = np.asarray(std, dtype=np.float32)
    imgs = hots = None
    mean_row = std_row = None
    for n, r in enumerate(records):
        img, tgt, _ = example_proto.parse_8bit_array_proto(r)
        if imgs is None:
            imgs = np.empty((len(records),) + img.shape, np.float32)
            hots = np.empty((len(records),) + tgt.shape[:2] + (num_classes,), np.float32)
        h, w, c = img.shape
        if mean_row is None or mean_row.size != w * c:
            mean_row, std_row = np.tile(mean, w), np.tile(std, w)       # long inner loops: NumPy broadcasts over 3 are slow
        o2 = imgs[n].reshape(h, w * c)
        np.subtract(img.reshape(h, w * c), mean_row, out=o2, dtype=np.float32)
        np.divide(o2, std_row, out=o2)
        lab = (tgt[..., 0] if tgt.ndim == 3 else tgt).reshape(-1)
        hn = hots[n].reshape(-1, num_classes)
        hn[:] = 0.0
        ok = np.nonzero(lab < num_classes)[0]
        hn[ok, lab[ok]] = 1.0
    if imgs is None:
        return np.zeros((0,), np.float32), np.zeros((0,), np.float32)
    return imgs, hots


def parse_shard_bytes(buf, mean, std, num_classes, verify=True):
    return parse_records_norm_onehot(tfrecord.read_records(buf, verify), mean, std, num_classes)
