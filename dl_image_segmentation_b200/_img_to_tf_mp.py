"""Chip folders -> sharded TFRecords for any supported raster chip (GeoTIFF LZW/DEFLATE/none, PNG).

B200 drop-in for the reference's multiprocess translator ``dl_segmentation_utils/_img_to_tf_mp.py``: same
function names and arguments; ``num_proc`` now counts GPU workers (one contiguous range of the shuffled file
list and ``num_shards / num_proc`` shards each, exactly the reference's partition ``:102-108,167-170``).
File discovery / shuffle ``:213-226`` is kept (``random.seed(12345)``), with both globs sorted first because
the reference pairs images and labels by position and ``gfile.glob`` order is unspecified.
"""
import glob
import os
import random

from . import _codec, _translate


def load_image_rasterio(img_path, parse_dltile_filename=True, decode=True, device=None):
    """One chip -> (image_data, height, width, bands, tile_key)  (reference :22-75).

    decode=True: (H,W,bands) CUDA tensor in the file's dtype; decode=False: the raw file bytes (header parsed
    only).  Raises on an unreadable / unsupported file, as rasterio would."""
    with open(img_path, "rb") as f:
        image_data = f.read()
    info = _codec.probe(image_data)
    if info.status != 0:
        raise _translate.ChipError("'%s' not recognized as a supported file format." % img_path)
    tile_key = _translate.tile_key_from_path(img_path, parse_dltile_filename)
    if not parse_dltile_filename:
        # reference :63-67: filename | str(geotransform) | str(crs), from the GeoTIFF tags (GDAL defaults for a PNG)
        tile_key = "|".join((os.path.basename(img_path),) + _codec.georef_strings(info))
    if decode:
        (arr,), (st,) = _codec.decode_blobs([image_data], device=device)
        if st != 0:
            raise _translate.ChipError("could not decode %s (codec status %d)" % (img_path, int(st)))
        assert (info.height, info.width, info.samples) == tuple(arr.shape)          # reference :72
        return arr, info.height, info.width, info.samples, tile_key
    return image_data, info.height, info.width, info.samples, tile_key


def _process_image_files_mp_worker(proc_index, ranges, name, img_filenames, lbl_filenames, output_directory,
                                   num_shards, dltile_from_filename, store_as_array, device=None):
    """One worker = one GPU: writes its shards (reference :78-157)."""
    def key_fn(p, info=None):
        if dltile_from_filename:
            return _translate.tile_key_from_path(p, True)
        gt, crs = _codec.georef_strings(info) if info is not None else ("[0.0, 1.0, 0.0, 0.0, 0.0, 1.0]", "None")
        return "|".join((os.path.basename(p), gt, crs))
    path_key = (lambda p: _translate.tile_key_from_path(p, True)) if dltile_from_filename else None
    return _translate.run_worker(proc_index, ranges, name, img_filenames, lbl_filenames, output_directory, num_shards,
                                 key_fn, store_as_array, label="process", progress_every=100, device=device, path_key=path_key,
                                 info_keys=not dltile_from_filename)


def _process_image_files_mp(name, img_files, lbl_files, out_folder, num_shards, num_proc, dltile_from_filename,
                            store_as_array):
    assert len(img_files) == len(lbl_files)
    ranges = _translate.worker_ranges(len(img_files), num_proc)
    # joblib.Parallel(n_jobs=num_proc) over processes in the reference (:180); here one GPU per worker, concurrently
    return _translate.run_workers(len(ranges), lambda proc_idx, dev: _process_image_files_mp_worker(
        proc_idx, ranges, name, img_files, lbl_files, out_folder, num_shards, dltile_from_filename, store_as_array, device=dev))


def _find_image_files(data_dir, file_ext):
    """Paired image / label lists in the seeded-shuffle order (reference :184-230)."""
    print("Determining list of input files and labels from %s." % data_dir)
    filenames = sorted(glob.glob("%s/images/*.%s" % (data_dir, file_ext)))
    labels = sorted(glob.glob("%s/labels/*.%s" % (data_dir, file_ext)))
    shuffled_index = list(range(len(filenames)))
    random.seed(12345)
    random.shuffle(shuffled_index)
    filenames = [filenames[i] for i in shuffled_index]
    labels = [labels[i] for i in shuffled_index]
    print("Found %d %s image files and %d label files inside %s." % (len(filenames), file_ext, len(labels), data_dir))
    return filenames, labels


def process_dataset_mp(name, directory, out_directory, num_shards, num_proc=None, dltile_from_filename=True,
                       file_ext="tif", store_as_array=True):
    """Process a folder of images and label images and save it as TFRecords (reference :233-275)."""
    if not num_proc:
        num_proc = num_shards
    filenames, labels = _find_image_files(directory, file_ext)
    return _process_image_files_mp(name, filenames, labels, out_directory, num_shards, num_proc, dltile_from_filename,
                                   store_as_array)
