"""`rasterio` stand-in for running the unmodified reference (test infrastructure; see ../README.md).

Provides `MemoryFile(bytes).open()` -> dataset with `.read()`, `.get_transform()`, `.read_crs()`, `.height`, `.width`,
`.count` (reference `_img_to_tf_mp.py:45-53`, `_tfrecord_image_translation.py:320-326,369-381`) on top of OpenCV / Pillow
(libtiff 4.7.1 — the codec GDAL's GTiff driver links — and libpng), with GDAL's band conventions: a palette PNG is ONE
band of indices, 16-bit stays 16-bit, grey+alpha is 2 bands.
"""
from .io import MemoryFile, RasterioIOError  # noqa: F401

__version__ = "0.0-refstub"
