"""oracle/ — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()``, and ``bench.py``'s
``cpu_baseline`` leg / ``--impl reference`` arm.  The product package
``dl_image_segmentation_b200`` never imports it and has no CPU fallback.

What it restates (reference = harry-gibson/dl_image_segmentation, paths under
``/root/reference/dl_segmentation_utils``):

=====================  ==========================================================================
``partition``          ``_img_to_tf_mp.py:102-122,167-170,213-226`` (seeded shuffle, linspace ranges,
                       shard names); same code in ``_img_to_tf_threaded.py:163-170,236-239,297-314``
``example_proto``      ``_tfrecord_image_translation.py:7-211`` (feature wrappers + ``convert_to_example``)
                       and ``:216-415`` (templates + the five ``parse_*_proto``)
``tfrecord``           ``tf.io.TFRecordWriter`` / ``TFRecordDataset`` call sites
                       (``_img_to_tf_mp.py:119,141,150``; ``parse_tfrecords.ipynb`` cell 4)
``imagecodecs``        ``load_image_rasterio`` (``_img_to_tf_mp.py:22-75``) and ``_process_image``
                       (``_img_to_tf_threaded.py:75-121``): TIFF (LZW / DEFLATE / none) and PNG decode
``jpegcodec``          ``ImageCoder.decode_jpeg`` / ``.jpg`` chips (``_img_to_tf_threaded.py:36-38,51-56,97-103``):
                       baseline JPEG decode with libjpeg's integer arithmetic
``jpegenc``            ``ImageCoder.png_to_jpeg`` (``_img_to_tf_threaded.py:36-38,92-95``): baseline JPEG encode,
                       byte-identical to libjpeg-turbo's files
``composite``          ``_descartes_img_chips.py:562-567`` (np.ma median) and ``:461-469,603-626``
                       (date/cloud filter, stable descending sort, painter's mosaic), ``:516`` dstack
``rasterize``          ``create_label_array_for_tile`` ``_descartes_img_chips.py:633-689`` (GDAL RasterizeLayer, ALL_TOUCHED)
``refrun``             runs the unmodified reference under ``refstubs/`` (fixture generator, live checks)
``normalise``          north-star row A17 (cast, per-band normalise, one-hot, integer band statistics)
``translate``          the whole worker loop ``_img_to_tf_mp.py:78-157`` on top of the pieces above
=====================  ==========================================================================

PARITY STATUS — pinned on the reference's own code: ``oracle/refrun.py`` imports the UNMODIFIED
``/root/reference/dl_segmentation_utils`` under the stub dependencies of ``oracle/refstubs/`` (tensorflow, rasterio,
descarteslabs, geopandas, osgeo are not installable here; the stubs delegate their arithmetic to google.protobuf, NumPy,
Pillow/libpng, OpenCV/libtiff and libjpeg-turbo, never to this package).  ``tests/golden/gen_golden_reference.py`` ran the
reference's translators, ``convert_to_example``, its five parsers, its compositors and ``create_chips_for_tile`` and
committed the outputs (``tests/golden/ref_*``); ``tests/test_reference_parity.py`` checks that every module here reproduces
them byte for byte and re-runs the reference live where ``/root/reference`` exists.  Unpinned remain only the two things
for which no code exists in this image: the Descartes Labs service's search / mosaic semantics (restated from its
documentation) and GDAL's ``RasterizeLayer`` (``rasterize``: restated from GDAL's public algorithm, pinned on the exact
square-meets-polygon set).  Beneath that the codecs stay pinned from the outside: RFC 3720 B.4 CRC-32C vectors;
``google.protobuf``; ``zlib``; libtiff 4.7.1 through ``cv2`` and ``Pillow``; libpng; libjpeg-turbo (decoded pixels bit for
bit, encoded files byte for byte); ``numpy.ma.median`` (``tests/test_oracle_*.py``, ``tests/golden/``).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libb2oracle.so")
_SRC = os.path.join(_HERE, "csrc", "b2oracle.c")


def build(force: bool = False) -> str:
    """Compile oracle/csrc/b2oracle.c with gcc (no GPU, no CUDA)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-msse4.2", "-fPIC", "-shared", "-o", _SO, _SRC])
    return _SO


_lib = None


def clib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        u8p, u64p = ctypes.c_void_p, ctypes.c_void_p
        L.orc_crc32c.restype = ctypes.c_uint32
        L.orc_crc32c.argtypes = [u8p, ctypes.c_size_t]
        L.orc_crc32c_sw.restype = ctypes.c_uint32
        L.orc_crc32c_sw.argtypes = [u8p, ctypes.c_size_t]
        L.orc_mask_crc.restype = ctypes.c_uint32
        L.orc_mask_crc.argtypes = [ctypes.c_uint32]
        L.orc_masked_crc32c.restype = ctypes.c_uint32
        L.orc_masked_crc32c.argtypes = [u8p, ctypes.c_size_t]
        L.orc_tfrecord_frame.restype = None
        L.orc_tfrecord_frame.argtypes = [u8p, ctypes.c_uint64, u8p]
        L.orc_tfrecord_scan.restype = ctypes.c_int64
        L.orc_tfrecord_scan.argtypes = [u8p, ctypes.c_uint64, u64p, u64p, ctypes.c_int64, ctypes.c_int]
        L.orc_lzw_decode.restype = ctypes.c_int64
        L.orc_lzw_decode.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_size_t]
        L.orc_hdiff_undo.restype = None
        L.orc_hdiff_undo.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
        L.orc_png_unfilter.restype = ctypes.c_int
        L.orc_png_unfilter.argtypes = [u8p, u8p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
        _lib = L
    return _lib
