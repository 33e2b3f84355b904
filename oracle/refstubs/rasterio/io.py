"""rasterio.io.MemoryFile over OpenCV / Pillow decoders (test infrastructure; see ../README.md)."""
import io as _io

import numpy as np


class RasterioIOError(OSError):
    pass


class _CRS:
    def __init__(self, epsg):
        self.epsg = epsg

    def __str__(self):
        return "EPSG:%d" % self.epsg

    to_string = __str__


def _tiff_georef(tags):
    gt = [0.0, 1.0, 0.0, 0.0, 0.0, 1.0]                      # GDAL's default geotransform
    crs = None
    if 34264 in tags and len(tags[34264]) == 16:
        m = [float(v) for v in tags[34264]]
        gt = [m[3], m[0], m[1], m[7], m[4], m[5]]
    elif 33550 in tags and 33922 in tags:
        sx, sy = float(tags[33550][0]), float(tags[33550][1])
        i, j, _, x, y, _ = [float(v) for v in tags[33922][:6]]
        gt = [x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy]
    if 34735 in tags:
        d = [int(v) for v in tags[34735]]
        proj = geog = 0
        for k in range(d[3]):
            key, loc, _, val = d[4 * (k + 1):4 * (k + 2)]
            if loc == 0 and key == 3072:
                proj = val
            if loc == 0 and key == 2048:
                geog = val
        code = proj if proj not in (0, 32767) else geog
        if code not in (0, 32767):
            crs = _CRS(code)
    return gt, crs


def _retag_for_opencv(data):
    """OpenCV's TIFF reader only hands back 3 / 4 channels for Photometric=RGB; GDAL writes non-Byte multi-band chips as
    MinIsBlack + (bands-1) ExtraSamples.  Re-label a COPY of the file as RGB (+1 extra sample) — two IFD entries change,
    the compressed tile / strip bytes libtiff decodes stay exactly as they are."""
    import struct
    bo = "<" if data[:2] == b"II" else ">"
    b = bytearray(data)
    (ifd,) = struct.unpack_from(bo + "I", b, 4)
    (n,) = struct.unpack_from(bo + "H", b, ifd)
    ents = [list(struct.unpack_from(bo + "HHI", b, ifd + 2 + 12 * k)) + [ifd + 2 + 12 * k] for k in range(n)]
    tag = {e[0]: e for e in ents}
    spp = struct.unpack_from(bo + "H", b, tag[277][3] + 8)[0] if 277 in tag else 1
    pm = struct.unpack_from(bo + "H", b, tag[262][3] + 8)[0] if 262 in tag else 1
    if spp not in (3, 4) or pm != 1:
        return data
    struct.pack_into(bo + "H", b, tag[262][3] + 8, 2)
    if 338 in tag:
        pos = tag[338][3]
        if spp == 4:
            struct.pack_into(bo + "HHI", b, pos, 338, 3, 1)
            b[pos + 8:pos + 12] = b"\0\0\0\0"
        else:                                                  # RGB without extras: drop the entry
            end = ifd + 2 + 12 * n + 4
            b[pos:end - 12] = b[pos + 12:end]
            struct.pack_into(bo + "H", b, ifd, n - 1)
    return bytes(b)


class _Dataset:
    """Opening parses the header only, as GDALOpen does; pixels are decoded by read()."""

    def __init__(self, data):
        import struct
        self._data = data
        self._gt, self._crs = [0.0, 1.0, 0.0, 0.0, 0.0, 1.0], None
        if data[:8] == b"\x89PNG\r\n\x1a\n":
            if len(data) < 33 or data[12:16] != b"IHDR":
                raise RasterioIOError("not a PNG")
            self.width, self.height = struct.unpack(">II", data[16:24])
            self.count = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[data[25]]
            self._kind = "png"
        elif data[:4] in (b"II*\0", b"MM\0*"):
            try:                                               # Pillow's IFD reader alone: its image modes stop at 4 x 8 bit
                from PIL.TiffImagePlugin import ImageFileDirectory_v2
                fp = _io.BytesIO(data)
                ifd = ImageFileDirectory_v2(fp.read(8))
                fp.seek(ifd.next)
                ifd.load(fp)
                tags = {k: (v if isinstance(v, tuple) else (v,)) for k, v in ifd.items()}
            except Exception as e:
                raise RasterioIOError(str(e))
            self._gt, self._crs = _tiff_georef(tags)
            self.height, self.width, self.count = int(tags[257][0]), int(tags[256][0]), int(tags.get(277, (1,))[0])
            self._kind = "tiff"
        elif data[:3] == b"\xff\xd8\xff":
            from PIL import Image
            im = Image.open(_io.BytesIO(data))                 # header only until load()
            self.width, self.height = im.size
            self.count = len(im.getbands())
            self._kind = "jpeg"
        else:
            raise RasterioIOError("'/vsimem/refstub' not recognized as a supported file format.")

    def _decode(self):
        import cv2
        from PIL import Image
        data = self._data
        if self._kind == "png":
            try:
                im = Image.open(_io.BytesIO(data))
                im.load()
            except Exception as e:
                raise RasterioIOError(str(e))
            if im.mode in ("I;16", "I;16B", "I"):
                arr = np.asarray(im).astype(np.uint16)
            elif im.mode == "1":
                arr = np.asarray(im, dtype=np.uint8)
            else:
                arr = np.asarray(im)                           # 'P' -> palette indices, as GDAL presents them
            if data[24] == 16 and data[25] in (2, 4, 6):       # Pillow narrows 16-bit colour; OpenCV keeps it
                arr = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
                arr = arr[:, :, [2, 1, 0] + ([3] if arr.shape[2] == 4 else [])] if arr.ndim == 3 and arr.shape[2] >= 3 else arr
        elif self._kind == "tiff":
            arr = cv2.imdecode(np.frombuffer(_retag_for_opencv(data), np.uint8), cv2.IMREAD_UNCHANGED)
            if arr is None:
                raise RasterioIOError("TIFF decode failed (libtiff)")
            if arr.ndim == 3 and arr.shape[2] >= 3:            # OpenCV hands back B,G,R[,A]
                arr = arr[:, :, [2, 1, 0] + ([3] if arr.shape[2] == 4 else [])]
        else:
            try:
                im = Image.open(_io.BytesIO(data))
                im.load()
            except Exception as e:
                raise RasterioIOError(str(e))
            arr = np.asarray(im)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        if arr.shape != (self.height, self.width, self.count):
            raise NotImplementedError("refstub rasterio: %r decoded as %r (more than 4 bands?)" % ((self.height, self.width, self.count), arr.shape))
        return np.ascontiguousarray(arr)

    def read(self):
        return np.ascontiguousarray(np.transpose(self._decode(), (2, 0, 1)))      # (bands, rows, cols)

    def get_transform(self):
        return list(self._gt)

    def read_crs(self):
        return self._crs

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class MemoryFile:
    def __init__(self, file_or_bytes=None):
        self._data = bytes(file_or_bytes)

    def open(self):
        return _Dataset(self._data)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
