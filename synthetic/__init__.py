"""synthetic/ — workload generators and chip *writers* for tests and bench.py.

Neither product nor oracle: this package only manufactures inputs (there is no network and the
reference ships no data, ``.MISSING_LARGE_BLOBS``).  Workloads follow SURVEY.md section 8(d):
seed = 1000 + config number, ``np.random.default_rng``.

Writers:
  * ``tiff_bytes``  — classic little/big-endian TIFF the way GDAL writes the reference's chips
    (``_descartes_img_chips.py:781-797``: COMPRESS=LZW, TILED=TRUE -> 256x256 tiles, pixel interleaved,
    predictor 1) plus the libtiff/cv2 flavour (strips, predictor 2), DEFLATE, planar=2, GDAL_NODATA.
  * ``png_bytes``   — Pillow (libpng, adaptive filters, dynamic Huffman) and ``png_bytes_manual``
    (chosen filter type per row, chosen IDAT chunking, stored/fixed/dynamic deflate blocks).
"""
import ctypes
import os
import struct
import subprocess
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libb2synth.so")
_SRC = os.path.join(_HERE, "csrc", "lzwenc.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-o", _SO, _SRC])
    return _SO


def _clib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.syn_lzw_encode_restart.restype = ctypes.c_int64
        L.syn_lzw_encode_restart.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
        L.syn_hdiff.restype = None
        L.syn_hdiff.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def lzw_encode(data: bytes, restart: int = 0) -> bytes:
    """TIFF-LZW (MSB first, early change, Clear at 4094 entries); restart > 0: a Clear every `restart` input bytes too."""
    s = np.frombuffer(data, dtype=np.uint8)
    cap = 2 * s.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _clib().syn_lzw_encode_restart(s.ctypes.data, s.size, out.ctypes.data, cap, int(restart))
    assert n >= 0
    return out[:n].tobytes()


# ----------------------------------------------------------------------------- TIFF writer
_SF = {"u": 1, "i": 2, "f": 3}


def tiff_bytes(arr, tile=256, compression="lzw", predictor=1, planar=1, big_endian=False,
               rows_per_strip=None, nodata=None, geo=True, zlevel=6, photometric=None, lzw_restart=0):
    """(H,W,B) or (H,W) array -> TIFF file bytes.  tile=None -> strips."""
    arr = np.asarray(arr)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    H, W, B = arr.shape
    dt = arr.dtype
    bo = ">" if big_endian else "<"
    comp = {"lzw": 5, "deflate": 8, "none": 1}[compression]
    planes = B if planar == 2 else 1
    spb = 1 if planar == 2 else B
    if tile:
        bw = bh = int(tile)
    else:
        bw, bh = W, int(rows_per_strip or max(1, 8192 // max(1, W * spb * dt.itemsize)))
    across, down = (W + bw - 1) // bw, (H + bh - 1) // bh
    blocks = []
    for p in range(planes):
        src = arr[:, :, p:p + 1] if planar == 2 else arr
        for by in range(down):
            for bx in range(across):
                rows = bh if tile else min(bh, H - by * bh)
                blk = np.zeros((rows, bw, spb), dtype=dt)
                y0, x0 = by * bh, bx * bw
                hh, ww = min(rows, H - y0), min(bw, W - x0)
                blk[:hh, :ww] = src[y0:y0 + hh, x0:x0 + ww]
                blk = np.ascontiguousarray(blk)
                if predictor == 2:
                    _clib().syn_hdiff(blk.ctypes.data, rows, bw * spb, spb, dt.itemsize)
                raw = blk.astype(dt.newbyteorder(bo)).tobytes()
                if comp == 5:
                    raw = lzw_encode(raw, lzw_restart)
                elif comp == 8:
                    raw = zlib.compress(raw, zlevel)
                blocks.append(raw)
    # ---- IFD
    entries = []       # (tag, type, count, value-bytes)

    def short(tag, *v):
        entries.append((tag, 3, len(v), struct.pack(bo + "H" * len(v), *v)))

    def long_(tag, *v):
        entries.append((tag, 4, len(v), struct.pack(bo + "I" * len(v), *v)))

    def dbl(tag, *v):
        entries.append((tag, 12, len(v), struct.pack(bo + "d" * len(v), *v)))

    def ascii_(tag, s):
        b = s.encode() + b"\0"
        entries.append((tag, 2, len(b), b))
    short(256, W)
    short(257, H)
    short(258, *([dt.itemsize * 8] * B))
    short(259, comp)
    # GDAL: RGB only for >=3 Byte bands, else MinIsBlack + (B-1) unspecified ExtraSamples
    pm = photometric if photometric is not None else (2 if (B >= 3 and dt == np.uint8) else 1)
    short(262, pm)
    short(277, B)
    short(284, planar)
    if predictor != 1:
        short(317, predictor)
    n_extra = B - (3 if pm == 2 else 1)
    if n_extra > 0:
        short(338, *([0] * n_extra))
    short(339, *([_SF[dt.kind]] * B))
    if geo:                                                     # GeoTIFF tags as GDAL writes them for a UTM DLTile
        dbl(33550, 10.0, 10.0, 0.0)
        dbl(33922, 0.0, 0.0, 0.0, 499980.0, 5300040.0, 0.0)
        short(34735, 1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32643)
    if nodata is not None:
        ascii_(42113, str(nodata))
    n_blocks = len(blocks)
    off_tag, cnt_tag = (324, 325) if tile else (273, 279)
    if tile:
        short(322, bw)
        short(323, bh)
    else:
        short(278, bh)
    long_(off_tag, *([0] * n_blocks))
    long_(cnt_tag, *[len(b) for b in blocks])
    entries.sort(key=lambda e: e[0])
    ifd_off = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_off = ifd_off + ifd_size
    extra = b""
    placed = []
    for tag, typ, cnt, val in entries:
        if len(val) <= 4:
            placed.append((tag, typ, cnt, val.ljust(4, b"\0"), None))
        else:
            if len(extra) % 2:
                extra += b"\0"
            placed.append((tag, typ, cnt, struct.pack(bo + "I", extra_off + len(extra)), (len(extra), len(val))))
            extra += val
    data_off = extra_off + len(extra)
    offs, cur = [], data_off
    for b in blocks:
        offs.append(cur)
        cur += len(b)
    offs_bytes = struct.pack(bo + "I" * n_blocks, *offs)
    extra = bytearray(extra)
    out_entries = b""
    for tag, typ, cnt, val, where in placed:
        if tag == off_tag:
            if where is None:
                val = offs_bytes.ljust(4, b"\0")
            else:
                extra[where[0]:where[0] + where[1]] = offs_bytes
        out_entries += struct.pack(bo + "HHI", tag, typ, cnt) + val
    hdr = (b"MM" if big_endian else b"II") + struct.pack(bo + "HI", 42, ifd_off)
    ifd = struct.pack(bo + "H", len(entries)) + out_entries + struct.pack(bo + "I", 0)
    return hdr + ifd + bytes(extra) + b"".join(blocks)


# ----------------------------------------------------------------------------- PNG writers
def png_bytes(arr, compress_level=6):
    """Pillow / libpng encoder (what a user's PNG chips look like)."""
    import io

    from PIL import Image
    arr = np.asarray(arr)
    if arr.ndim == 3 and arr.shape[2] == 1:
        arr = arr[:, :, 0]
    bio = io.BytesIO()
    Image.fromarray(arr).save(bio, format="PNG", compress_level=compress_level)
    return bio.getvalue()


def _png_chunk(typ, body):
    return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF)


def png_filter_rows(arr, filter_types):
    """Apply PNG filters (type per row, cycled) -> filtered scanlines incl. filter bytes."""
    arr = np.asarray(arr, dtype=np.uint8)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    H, W, C = arr.shape
    rows = arr.reshape(H, W * C).astype(np.int32)
    out = bytearray()
    prev = np.zeros(W * C, dtype=np.int32)
    for y in range(H):
        ft = filter_types[y % len(filter_types)]
        cur = rows[y]
        a = np.concatenate([np.zeros(C, np.int32), cur[:-C]])
        b = prev
        c = np.concatenate([np.zeros(C, np.int32), prev[:-C]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - a
        elif ft == 2:
            f = cur - b
        elif ft == 3:
            f = cur - ((a + b) >> 1)
        else:
            p = a + b - c
            pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
            f = cur - pred
        out.append(ft)
        out += (f & 0xFF).astype(np.uint8).tobytes()
        prev = cur
    return bytes(out)


def png_bytes_manual(arr, filter_types=(0, 1, 2, 3, 4), zlevel=6, idat_chunk=None, strategy=zlib.Z_DEFAULT_STRATEGY):
    arr = np.asarray(arr, dtype=np.uint8)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    H, W, C = arr.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[C]
    co = zlib.compressobj(zlevel, zlib.DEFLATED, 15, 9, strategy)
    z = co.compress(png_filter_rows(arr, filter_types)) + co.flush()
    body = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, ctype, 0, 0, 0))
    step = idat_chunk or max(1, len(z))
    for i in range(0, max(1, len(z)), step):
        body += _png_chunk(b"IDAT", z[i:i + step])
    return body + _png_chunk(b"IEND", b"")


ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))   # x0, y0, dx, dy


def png_bytes_flavour(samples, depth, ctype, palette=None, trns=None, filter_types=(0, 1, 2, 3, 4), zlevel=6, interlace=False):
    """PNG of any flavour.  samples: (H,W) or (H,W,C) integers < 2**depth (palette indices for colour type 3);
    depth 1/2/4 packs them most-significant-bits first, depth 16 stores them big-endian; interlace=True writes the
    seven Adam7 passes (each a reduced image with its own filtered scanlines)."""
    a = np.asarray(samples)
    if a.ndim == 2:
        a = a[:, :, None]
    H, W, C = a.shape
    assert C == {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]

    def scanlines(sub):
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            return b""
        if depth == 16:
            rows = sub.astype(">u2").view(np.uint8).reshape(sub.shape[0], sub.shape[1], 2 * C)
        elif depth == 8:
            rows = sub.astype(np.uint8)
        else:
            bits = ((sub[:, :, 0].astype(np.uint8)[:, :, None] >> np.arange(depth - 1, -1, -1)) & 1).reshape(sub.shape[0], -1)
            rows = np.packbits(bits, axis=1)[:, :, None]                  # (h, ceil(w*depth/8), 1): filter unit = 1 byte
        return bytes(png_filter_rows(rows, filter_types))
    if interlace:
        raw = b"".join(scanlines(a[y0::dy, x0::dx]) for x0, y0, dx, dy in ADAM7)
    else:
        raw = scanlines(a)
    z = zlib.compress(raw, zlevel)
    body = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        body += _png_chunk(b"PLTE", np.asarray(palette, np.uint8).reshape(-1, 3).tobytes())
    if trns is not None:
        body += _png_chunk(b"tRNS", bytes(trns))
    return body + _png_chunk(b"IDAT", z) + _png_chunk(b"IEND", b"")


def png_bytes_raw_zlib(width, height, channels, zstream):
    """A PNG around a caller-supplied zlib stream (tests of malformed / hand-made DEFLATE data)."""
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[channels]
    return (b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, 8, ctype, 0, 0, 0)) +
            _png_chunk(b"IDAT", zstream) + _png_chunk(b"IEND", b""))


def handmade_dynamic_deflate(symbols, lit_lens, data):
    """zlib stream with ONE dynamic-Huffman block whose literal/length code has the given lengths
    ({symbol: bits}; every other symbol unused) and one 1-bit distance code; `data` = symbols to emit (end with 256).
    Nothing checks that the set is complete: that is the point (zlib rejects an incomplete one in the block header)."""
    bits = []

    def put(v, n):                      # n bits, LSB first (header fields, extra bits)
        bits.extend((v >> i) & 1 for i in range(n))

    def put_code(c, n):                 # Huffman code, MSB first
        bits.extend((c >> (n - 1 - i)) & 1 for i in range(n))

    def canonical(lens):
        out, code = {}, 0
        for ln in range(1, 16):
            for s in sorted(k for k, v in lens.items() if v == ln):
                out[s] = (code, ln)
                code += 1
            code <<= 1
        return out
    lens = [0] * 257
    for s_, l_ in lit_lens.items():
        lens[s_] = l_
    seq = lens + [1]                    # + one distance code of one bit
    # run-length code the sequence with symbols 0..15 and 18 only
    rl, i = [], 0
    while i < len(seq):
        if seq[i] == 0:
            j = i
            while j < len(seq) and seq[j] == 0 and j - i < 138:
                j += 1
            if j - i >= 11:
                rl.append((18, j - i - 11))
                i = j
                continue
        rl.append((seq[i], None))
        i += 1
    used = sorted({s_ for s_, _ in rl})
    cl_lens = {s_: 3 for s_ in used}    # a complete code-length code: pad to 8 three-bit codes
    for s_ in range(19):
        if len(cl_lens) == 8:
            break
        cl_lens.setdefault(s_, 3)
    assert len(cl_lens) == 8
    cl = canonical(cl_lens)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    put(1, 1)
    put(2, 2)                           # final block, dynamic
    put(0, 5)                           # HLIT  = 257
    put(0, 5)                           # HDIST = 1
    put(15, 4)                          # HCLEN = 19
    for s_ in order:
        put(cl_lens.get(s_, 0), 3)
    for s_, extra in rl:
        put_code(*cl[s_])
        if s_ == 18:
            put(extra, 7)
    lit = canonical(lit_lens)
    for s_ in data:
        put_code(*lit[s_])
    while len(bits) % 8:
        bits.append(0)
    raw = bytes(sum(b << k for k, b in enumerate(bits[i:i + 8])) for i in range(0, len(bits), 8))
    payload = bytes(s_ for s_ in data if s_ < 256)
    return b"\x78\x01" + raw + struct.pack(">I", zlib.adler32(payload))


# ----------------------------------------------------------------------------- smooth fields / workloads
def smooth_field(rng, h, w, coarse=34):
    """Uniform noise on a coarse grid, bicubic-upsampled to (h, w), rescaled to [0,1]."""
    import cv2
    g = rng.random((coarse, coarse)).astype(np.float32)
    f = cv2.resize(g, (w, h), interpolation=cv2.INTER_CUBIC)
    lo, hi = float(f.min()), float(f.max())
    return (f - lo) / max(hi - lo, 1e-9)


def label_field(rng, h, w, num_classes=10, nodata=255, nodata_frac=0.02):
    f = smooth_field(rng, h, w, coarse=12)
    lab = np.minimum((f * num_classes).astype(np.int32), num_classes - 1).astype(np.uint8)
    nd = smooth_field(rng, h, w, coarse=9)
    thr = np.quantile(nd, 1.0 - nodata_frac)
    lab[nd > thr] = nodata
    return lab


def dltile_key(size, pad, res, i, j):
    return "%d:%d:%s:43:%d:%d" % (size, pad, res, i, j)


def cfg1_chip(index, seed=1001, size=256):
    """256x256x3 u8 image + 256x256 u8 label (10 classes + 2% nodata)."""
    rng = np.random.default_rng([seed, index])
    img = np.stack([smooth_field(rng, size, size) for _ in range(3)], axis=-1)
    img = np.clip(img * 252.0, 0, 252).astype(np.uint8) + rng.integers(0, 4, (size, size, 3), dtype=np.uint8)
    return img, label_field(rng, size, size), dltile_key(size, 2, "1.0", index // 100, index % 100)


def cfg3_chip(index, seed=1003, size=512, bands=4):
    """512x512x4 u16 Sentinel-like (smooth 0..10000 + noise 0..46) + label."""
    rng = np.random.default_rng([seed, index])
    img = np.stack([smooth_field(rng, size, size) for _ in range(bands)], axis=-1)
    img = (img * 9950.0).astype(np.uint16) + rng.integers(0, 47, (size, size, bands), dtype=np.uint16)
    return img, label_field(rng, size, size), dltile_key(size - 64, 32, "10.0", index // 32, index % 32)


def cfg4_tile(index, seed=1004, T=16, H=1024, W=1024, B=8, cloud=0.4):
    """(T,H,W,B) u16 uniform 0..10000 + (T,H,W) u8 validity (~40% cloud, smooth), 4x4 all-cloud patch."""
    rng = np.random.default_rng([seed, index])
    stack = rng.integers(0, 10001, (T, H, W, B), dtype=np.uint16)
    valid = np.empty((T, H, W), dtype=np.uint8)
    for t in range(T):
        f = smooth_field(rng, H, W, coarse=20)
        valid[t] = (f > np.quantile(f[::8, ::8], cloud)).astype(np.uint8)
    valid[:, 8:12, 8:12] = 0
    return stack, valid


def cfg5_chip(index, seed=1005, T=32, H=256, W=256, B=4, valid_frac=0.85):
    rng = np.random.default_rng([seed, index])
    stack = rng.integers(0, 10001, (T, H, W, B), dtype=np.uint16)
    valid = np.empty((T, H, W), dtype=np.uint8)
    for t in range(T):
        f = smooth_field(rng, H, W, coarse=10)
        valid[t] = (f > np.quantile(f[::4, ::4], 1.0 - valid_frac)).astype(np.uint8)
    return stack, valid


def cfg5_scene_meta(chip, T=32, seed=1005):
    """Per-chip scene metadata: ascending scene_day in [0,730), cloud_fraction U[0,1)."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, int(chip)]))
    day = np.sort(rng.integers(0, 730, T)).astype(np.int32)
    cf = rng.random(T).astype(np.float32)
    return day, cf


CFG5_FILTER = dict(ref_day=365, min_day=180, max_day=545, max_cf=0.4)
