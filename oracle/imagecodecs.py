"""Chip decode: TIFF (LZW / DEFLATE / none) and PNG (oracle; test infrastructure only).

Stands in for the two decode calls of the reference:
  * ``load_image_rasterio`` ``_img_to_tf_mp.py:43-75`` — ``MemoryFile(bytes).open().read()`` (GDAL ->
    libtiff / libpng) then ``reshape_as_image`` ``(B,H,W)->(H,W,B)`` ``:69``; ``decode=False`` keeps the file
    bytes and only reads height/width/count ``:51-53``.
  * ``_process_image`` ``_img_to_tf_threaded.py:87-121`` — ``tf.image.decode_png`` ``:59``.
GDAL/TF are not installable; the codecs are restated from the public specs (TIFF 6.0 sections 2, 13, 14;
PNG sections 5, 9; RFC 1950/1951 via the real ``zlib``).  Chip format written by the reference:
``_descartes_img_chips.py:781-797`` (GTiff, COMPRESS=LZW, TILED=TRUE -> 256x256 tiles, pixel
interleaved, predictor 1).  Tests cross-check against libtiff (cv2, Pillow) and libpng (Pillow).
"""
import os
import struct
import zlib

import numpy as np

from . import clib

_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8}
_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 6: "b", 7: "B", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q"}


class DecodeError(Exception):
    pass


def parse_tiff(blob: bytes) -> dict:
    """First IFD of a classic TIFF -> dict of the tags the decoder needs."""
    if len(blob) < 8:
        raise DecodeError("short TIFF")
    bo = {b"II": "<", b"MM": ">"}.get(blob[:2])
    if bo is None:
        raise DecodeError("not a TIFF")
    magic, ifd = struct.unpack(bo + "HI", blob[2:8])
    if magic != 42:
        raise DecodeError("BigTIFF / unknown magic %d" % magic)
    (n,) = struct.unpack(bo + "H", blob[ifd:ifd + 2])
    tags = {}
    for i in range(n):
        e = ifd + 2 + 12 * i
        tag, typ, cnt = struct.unpack(bo + "HHI", blob[e:e + 8])
        sz = _TYPE_SIZE.get(typ)
        if sz is None:
            continue
        nb = sz * cnt
        off = e + 8 if nb <= 4 else struct.unpack(bo + "I", blob[e + 8:e + 12])[0]
        raw = blob[off:off + nb]
        if len(raw) < nb:
            raise DecodeError("tag %d data out of range" % tag)
        if typ == 5 or typ == 10:
            vals = struct.unpack(bo + ("II" if typ == 5 else "ii") * cnt, raw)
        elif typ == 2:
            vals = raw
        else:
            vals = struct.unpack(bo + _TYPE_FMT[typ] * cnt, raw)
        tags[tag] = vals
    g = lambda t, d=None: tags[t][0] if t in tags else d
    info = dict(
        byteorder=bo, width=g(256), height=g(257), compression=g(259, 1), photometric=g(262, 1),
        spp=g(277, 1), planar=g(284, 1), predictor=g(317, 1), fill_order=g(266, 1),
        bps=tags.get(258, (1,)), sample_format=tags.get(339, (1,)), tags=tags)
    if info["width"] is None or info["height"] is None:
        raise DecodeError("missing size tags")
    if 322 in tags:
        info.update(tiled=True, block_w=g(322), block_h=g(323), offsets=tags[324], counts=tags[325])
    else:
        rps = min(g(278, info["height"]), info["height"])
        info.update(tiled=False, block_w=info["width"], block_h=rps, offsets=tags[273], counts=tags[279])
    if len(set(info["bps"])) != 1 or len(set(info["sample_format"])) != 1:
        raise DecodeError("mixed per-band sample types")
    info["nodata"] = tags[42113].rstrip(b"\0").decode() if 42113 in tags else None
    return info


def _dtype(bps, fmt):
    try:
        return {(8, 1): np.uint8, (8, 2): np.int8, (16, 1): np.uint16, (16, 2): np.int16, (32, 1): np.uint32,
                (32, 2): np.int32, (32, 3): np.float32, (64, 3): np.float64}[(bps, fmt)]
    except KeyError:
        raise DecodeError("unsupported sample type bps=%d fmt=%d" % (bps, fmt))


def lzw_decode(src: bytes, nbytes: int) -> bytes:
    out = np.zeros(nbytes, dtype=np.uint8)
    s = np.frombuffer(src, dtype=np.uint8)
    n = clib().orc_lzw_decode(s.ctypes.data, s.size, out.ctypes.data, nbytes)
    if n != nbytes:
        raise DecodeError("LZW stream produced %d of %d bytes" % (n, nbytes))
    return out.tobytes()


def decode_tiff(blob: bytes) -> np.ndarray:
    t = parse_tiff(blob)
    W, H, spp = t["width"], t["height"], t["spp"]
    bps, fmt = t["bps"][0], t["sample_format"][0]
    dt = np.dtype(_dtype(bps, fmt))
    bpsamp = dt.itemsize
    if t["fill_order"] != 1:
        raise DecodeError("FillOrder 2 unsupported")
    if t["predictor"] not in (1, 2):
        raise DecodeError("predictor %d unsupported" % t["predictor"])
    planes = spp if t["planar"] == 2 else 1
    spb = 1 if t["planar"] == 2 else spp            # samples per pixel inside one block
    bw, bh = t["block_w"], t["block_h"]
    across = (W + bw - 1) // bw
    down = (H + bh - 1) // bh
    if len(t["offsets"]) < across * down * planes:
        raise DecodeError("too few blocks")
    out = np.zeros((H, W, spp), dtype=dt)
    for p in range(planes):
        for by in range(down):
            for bx in range(across):
                k = (p * down + by) * across + bx
                rows = bh if t["tiled"] else min(bh, H - by * bh)
                nbytes = rows * bw * spb * bpsamp
                raw = blob[t["offsets"][k]:t["offsets"][k] + t["counts"][k]]
                if len(raw) != t["counts"][k]:
                    raise DecodeError("block %d out of file" % k)
                c = t["compression"]
                if c == 5:
                    dec = lzw_decode(raw, nbytes)
                elif c in (8, 32946):
                    try:
                        dec = zlib.decompress(raw)
                    except zlib.error as e:
                        raise DecodeError(str(e))
                    if len(dec) < nbytes:
                        raise DecodeError("short deflate block")
                    dec = dec[:nbytes]
                elif c == 1:
                    if len(raw) < nbytes:
                        raise DecodeError("short raw block")
                    dec = raw[:nbytes]
                else:
                    raise DecodeError("compression %d unsupported" % c)
                arr = np.frombuffer(dec, dtype=dt.newbyteorder(t["byteorder"])).astype(dt).reshape(rows, bw * spb).copy()
                if t["predictor"] == 2:
                    if dt.kind == "f":
                        raise DecodeError("predictor 2 on float")
                    clib().orc_hdiff_undo(arr.ctypes.data, rows, bw * spb, spb, bpsamp)
                arr = arr.reshape(rows, bw, spb)
                y0, x0 = by * bh, bx * bw
                hh, ww = min(rows, H - y0), min(bw, W - x0)
                if t["planar"] == 2:
                    out[y0:y0 + hh, x0:x0 + ww, p] = arr[:hh, :ww, 0]
                else:
                    out[y0:y0 + hh, x0:x0 + ww, :] = arr[:hh, :ww, :]
    return out


_PNG_SIG = b"\x89PNG\r\n\x1a\n"


def parse_png(blob: bytes) -> dict:
    if blob[:8] != _PNG_SIG:
        raise DecodeError("not a PNG")
    p, idat, ihdr, plte, trns = 8, [], None, None, None
    while p + 8 <= len(blob):
        (n,) = struct.unpack(">I", blob[p:p + 4])
        typ = blob[p + 4:p + 8]
        body = blob[p + 8:p + 8 + n]
        if len(body) != n:
            raise DecodeError("truncated chunk")
        if typ in (b"IHDR", b"IDAT", b"PLTE"):
            # libpng (behind tf.image.decode_png and GDAL's PNG driver) treats a CRC mismatch in a critical chunk as fatal
            if blob[p + 8 + n:p + 12 + n] != struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF):
                raise DecodeError("CRC mismatch in %s chunk" % typ.decode())
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"PLTE" and not idat:
            if n % 3 or n > 768:
                raise DecodeError("bad PLTE")
            plte = body
        elif typ == b"tRNS" and not idat:
            trns = body
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        p += 12 + n
    if ihdr is None:
        raise DecodeError("no IHDR")
    w, h, depth, ctype, comp, flt, inter = ihdr
    return dict(width=w, height=h, depth=depth, color_type=ctype, interlace=inter, idat=b"".join(idat), plte=plte, trns=trns)


def decode_png(blob: bytes, as_tf: bool = True) -> np.ndarray:
    """Non-interlaced PNG -> (H,W,C).  8-bit grey / grey+alpha / RGB / RGBA are the samples as stored.  The other
    flavours depend on who decodes (the two reference paths disagree):
      as_tf=True   tf.image.decode_png(dtype=uint8) (_img_to_tf_threaded.py:59; libpng png_set_palette_to_rgb,
                   png_set_expand_gray_1_2_4_to_8, png_set_strip_16): palette -> RGB (RGBA when a tRNS chunk is present),
                   1/2/4-bit grey scaled by 255/(2^bits-1), 16-bit -> the high byte of every sample;
      as_tf=False  rasterio / GDAL PNG driver (_img_to_tf_mp.py:45-48): palette -> the indices as one band, sub-byte
                   grey unscaled, 16-bit -> uint16."""
    t = parse_png(blob)
    depth, ct = t["depth"], t["color_type"]
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}.get(ct)
    ok = ch is not None and (depth == 8 or (depth == 16 and ct != 3) or (depth in (1, 2, 4) and ct in (0, 3)))
    if not ok or t["interlace"] > 1 or (t["interlace"] == 1 and depth < 8):
        raise DecodeError("PNG flavour out of scope (depth %d, colour type %d, interlace %d)" % (depth, ct, t["interlace"]))
    if ct == 3 and t["plte"] is None:
        raise DecodeError("palette image without PLTE")
    h, w = t["height"], t["width"]
    rb = (w * ch * depth + 7) // 8
    bpp = max(1, ch * depth // 8)
    try:
        raw = zlib.decompress(t["idat"])
    except zlib.error as e:
        raise DecodeError(str(e))

    def unfilter(buf, off, hh, rbb):
        if len(buf) < off + hh * (rbb + 1):
            raise DecodeError("short PNG stream")
        src = np.frombuffer(buf, dtype=np.uint8, count=hh * (rbb + 1), offset=off)
        dst = np.zeros(hh * rbb, dtype=np.uint8)
        if clib().orc_png_unfilter(src.ctypes.data, dst.ctypes.data, hh, rbb, bpp) != 0:
            raise DecodeError("bad PNG filter type")
        return dst.reshape(hh, rbb)
    if t["interlace"]:
        # Adam7 (PNG spec section 8.2): seven reduced images, each with its own filtered scanlines, depth >= 8 here
        rows = np.zeros((h, w, bpp), np.uint8)
        off = 0
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            wp, hp = (w - x0 + dx - 1) // dx if w > x0 else 0, (h - y0 + dy - 1) // dy if h > y0 else 0
            if wp == 0 or hp == 0:
                continue
            rows[y0::dy, x0::dx] = unfilter(raw, off, hp, wp * bpp).reshape(hp, wp, bpp)
            off += hp * (wp * bpp + 1)
        rows = rows.reshape(h, rb)
    else:
        rows = unfilter(raw, 0, h, rb)
    if depth == 8 and ct != 3:
        return rows.reshape(h, w, ch)
    if depth == 16:
        be = rows.reshape(h, w, ch, 2)
        if as_tf:
            return be[..., 0].copy()
        return (be[..., 0].astype(np.uint16) << 8) | be[..., 1]
    # packed samples, most significant bits first
    bits = np.unpackbits(rows, axis=1)[:, :w * depth].reshape(h, w, depth)
    v = np.zeros((h, w), np.uint8)
    for k in range(depth):
        v = (v << 1) | bits[..., k]
    if ct == 0:
        scale = {1: 255, 2: 85, 4: 17}[depth] if as_tf else 1
        return (v * np.uint8(scale))[..., None]
    if not as_tf:
        return v[..., None]
    pal = np.zeros((256, 4), np.uint8)
    pal[:, 3] = 255
    pl = np.frombuffer(t["plte"], np.uint8).reshape(-1, 3)
    pal[:len(pl), :3] = pl
    if t["trns"] is not None:
        tr = np.frombuffer(t["trns"], np.uint8)[:256]
        pal[:len(tr), 3] = tr
    return pal[v][..., :4 if t["trns"] is not None else 3].copy()


def decode_image(blob: bytes, png_as_tf: bool = True) -> np.ndarray:
    if blob[:8] == _PNG_SIG:
        return decode_png(blob, png_as_tf)
    if blob[:2] in (b"II", b"MM"):
        return decode_tiff(blob)
    if blob[:3] == b"\xff\xd8\xff":
        if os.environ.get("B2_ORACLE_JPEG") == "libjpeg":     # CPU-baseline timing only: libjpeg-turbo itself (through cv2),
            import cv2                                        # the library behind tf.image.decode_jpeg / GDAL's JPEG driver
            a = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
            if a is None:
                raise DecodeError("libjpeg could not decode the file")
            return np.ascontiguousarray(a[..., ::-1]) if a.ndim == 3 else a[..., None]
        from . import jpegcodec
        try:
            return jpegcodec.decode_jpeg(blob)
        except jpegcodec.DecodeError as e:
            raise DecodeError(str(e))
    raise DecodeError("unknown image format")


def georef_strings(blob: bytes):
    """(str(src.get_transform()), str(src.read_crs())) of load_image_rasterio (_img_to_tf_mp.py:49-50), restated from
    the GeoTIFF tags: affine from ModelTransformation (34264) or ModelPixelScale (33550) + ModelTiepoint (33922) in
    GDAL order [x0, dx, rx, y0, ry, dy]; CRS from the GeoKeyDirectory's ProjectedCSType (3072) / GeographicType (2048)
    EPSG code.  Files without georeferencing (PNG, plain TIFF) give GDAL's default transform and no CRS."""
    gt, crs = [0.0, 1.0, 0.0, 0.0, 0.0, 1.0], "None"
    if blob[:8] != _PNG_SIG:
        tags = parse_tiff(blob)["tags"]
        if 34264 in tags and len(tags[34264]) == 16:
            m = tags[34264]
            gt = [m[3], m[0], m[1], m[7], m[4], m[5]]
        elif 33550 in tags and 33922 in tags and len(tags[33550]) >= 2 and len(tags[33922]) >= 6:
            sx, sy = tags[33550][0], tags[33550][1]
            i, j, _, x, y, _ = tags[33922][:6]
            gt = [x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy]
        if 34735 in tags and len(tags[34735]) >= 4:
            d = tags[34735]
            proj = geog = 0
            for k in range(d[3]):
                if 4 * (k + 2) > len(d):
                    break
                key, loc, _, val = d[4 * (k + 1):4 * (k + 2)]
                if loc == 0 and key == 3072:
                    proj = val
                elif loc == 0 and key == 2048:
                    geog = val
            code = proj if proj not in (0, 32767) else (geog if geog not in (0, 32767) else 0)
            if code:
                crs = "EPSG:%d" % code
    return str([float(v) for v in gt]), crs


def image_shape(blob: bytes):
    """(height, width, bands) from the header only — load_image_rasterio(decode=False), :51-53."""
    if blob[:8] == _PNG_SIG:
        t = parse_png(blob)
        return t["height"], t["width"], {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[t["color_type"]]     # GDAL: a palette image is one band
    if blob[:3] == b"\xff\xd8\xff":
        from . import jpegcodec
        return jpegcodec.jpeg_shape(blob)
    t = parse_tiff(blob)
    return t["height"], t["width"], t["spp"]
