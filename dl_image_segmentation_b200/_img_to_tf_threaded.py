"""PNG chip folders -> sharded TFRecords: B200 drop-in for ``dl_segmentation_utils/_img_to_tf_threaded.py``.

Same names and arguments as the reference; ``num_threads`` counts GPU workers.  Kept: ``*.png`` then ``*.jpg``
discovery and the seeded shuffle (``:297-314``), the always-validate rule (3-D, at most 3 bands, ``:105-112``) even
when the raw bytes are stored, the substring ``_is_png`` test (``:72``), skip-and-continue (``:190-199``), progress
every 1000 chips (``:207-210``).  ``.jpg`` chips are decoded and ``convert_png_to_jpg`` transcodes on the GPU with
libjpeg's arithmetic (``csrc/jpeg.cu``, ``csrc/jpeg_enc.cu``); progressive / CMYK / 12-bit JPEGs are skipped with the
reference's message.
"""
import glob
import os
import random
import sys
from datetime import datetime

from . import _codec, _translate


class ImageCoder(object):
    """Helper with the reference's method names (``:16-62``); decoding runs on the GPU."""

    def __init__(self, device=None):
        self.device = device

    def png_to_jpeg(self, image_data):
        # tf.image.encode_jpeg(tf.image.decode_png(image_data), format='', quality=100)  (reference :36-38)
        image = self.decode_png(image_data)
        if image.shape[2] not in (1, 3):
            raise _translate.ChipError("encode_jpeg: image must have 1 or 3 channels, has %d" % int(image.shape[2]))
        return _codec.encode_jpeg_arrays([image], quality=100, device=self.device)[0]

    def decode_jpeg(self, image_data):
        (image,), (st,), _ = _codec.decode_jpeg_blobs([image_data], device=self.device)           # tf.image.decode_jpeg
        if st != 0:
            raise _translate.ChipError("could not decode JPEG (codec status %d)" % int(st))
        return image

    def decode_png(self, image_data):
        (image,), (st,) = _codec.decode_blobs([image_data], device=self.device, png_as_tf=True)   # tf.image.decode_png
        if st != 0:
            raise _translate.ChipError("could not decode PNG (codec status %d)" % int(st))
        assert len(image.shape) == 3
        assert image.shape[2] <= 3
        return image


def _is_png(filename):
    return ".png" in filename


def _validate(info):
    # reference :107-112 — decoded image must be 3-D with <= 3 bands (checked from the header here)
    assert info.samples <= 3


def _process_image(filename, coder, parse_dltile_filename=True, png_to_jpg=False, decode=False):
    """One chip -> (image tensor | raw bytes, height, width, bands, tile_key)  (reference :75-121)."""
    with open(filename, "rb") as f:
        image_data = f.read()
    if _is_png(filename):
        if not png_to_jpg:
            image = coder.decode_png(image_data)
        else:
            print("Converting PNG to JPEG for %s" % filename)
            image_data = coder.png_to_jpeg(image_data)
            image = coder.decode_jpeg(image_data)
    else:
        image = coder.decode_jpeg(image_data)
    assert len(image.shape) == 3
    height, width, bands = (int(x) for x in image.shape)
    assert bands <= 3
    tile_key = _translate.tile_key_from_path(filename, parse_dltile_filename)
    if decode:
        return image, height, width, bands, tile_key
    return image_data, height, width, bands, tile_key


def _process_image_files_worker(coder, thread_index, ranges, name, filenames, labels, out_folder, num_shards,
                                dltile_from_filename, png_to_jpg, store_as_array=False, device=None):
    """One worker = one GPU (reference :136-219)."""
    def key_fn(p, info=None):
        return _translate.tile_key_from_path(p, dltile_from_filename)

    def validate(info):
        if info.format not in (2, _codec.FORMAT_JPEG):
            raise NotImplementedError("only PNG and JPEG chips are handled by the threaded translator on the GPU")
        _validate(info)
    return _translate.run_worker(thread_index, ranges, name, filenames, labels, out_folder, num_shards, key_fn,
                                 store_as_array, label="thread", progress_every=1000, validate=validate, device=device,
                                 png_as_tf=True,                     # this translator decodes with tf.image.decode_png
                                 png_to_jpg=png_to_jpg, path_key=key_fn,
                                 fast_validate=lambda infos: (infos["format"] == 2) & (infos["samples"] <= 3))


def _process_image_files(name, img_files, lbl_files, out_folder, num_shards, num_threads, dltile_from_filename,
                         png_to_jpg, store_as_array):
    assert len(img_files) == len(lbl_files)
    ranges = _translate.worker_ranges(len(img_files), num_threads)
    print("Launching %d threads for spacings: %s" % (num_threads, ranges))
    sys.stdout.flush()
    coder = ImageCoder()
    _translate.run_workers(len(ranges), lambda thread_index, dev: _process_image_files_worker(
        coder, thread_index, ranges, name, img_files, lbl_files, out_folder, num_shards, dltile_from_filename, png_to_jpg,
        store_as_array, device=dev))
    print("%s: Finished writing all %d images in data set." % (datetime.now(), len(img_files)))
    sys.stdout.flush()


def _find_image_files(data_dir):
    """*.png then *.jpg under images/ and labels/, seeded shuffle (reference :268-318)."""
    print("Determining list of input files and labels from %s." % data_dir)
    filenames = sorted(glob.glob("%s/images/*.png" % data_dir))
    labels = sorted(glob.glob("%s/labels/*.png" % data_dir))
    fn_jpg = sorted(glob.glob("%s/images/*.jpg" % data_dir))
    lb_jpg = sorted(glob.glob("%s/labels/*.jpg" % data_dir))
    filenames.extend(fn_jpg)
    labels.extend(lb_jpg)
    shuffled_index = list(range(len(filenames)))
    random.seed(12345)
    random.shuffle(shuffled_index)
    filenames = [filenames[i] for i in shuffled_index]
    labels = [labels[i] for i in shuffled_index]
    print("Found %d image files (of which %d JPGs) and %d label files inside %s." %
          (len(filenames), len(fn_jpg), len(labels), data_dir))
    return filenames, labels


def process_dataset_multithreaded(name, directory, out_directory, num_shards, num_threads=None,
                                  dltile_from_filename=True, convert_png_to_jpg=False, store_as_array=False):
    """Process a folder of PNG chips + label chips and save it as TFRecords (reference :321-350)."""
    if not num_threads:
        num_threads = num_shards
    assert not num_shards % num_threads, ("Num shards must be a multiple of num threads (incl 1*)")
    filenames, labels = _find_image_files(directory)
    _process_image_files(name, filenames, labels, out_directory, num_shards, num_threads, dltile_from_filename,
                         convert_png_to_jpg, store_as_array)
