"""Baseline JPEG encode (oracle; test infrastructure only).

Stands in for ``tf.image.encode_jpeg(image, format='rgb', quality=100)`` behind ``ImageCoder.png_to_jpeg``
(``_img_to_tf_threaded.py:31-34,92-95``, the ``convert_png_to_jpg`` option).  TensorFlow is not installable; what it
delegates is libjpeg's compressor with default settings: fixed-point RGB -> YCbCr, 2x2 chroma down-sampling with the
alternating 1,2 bias, the accurate integer ("islow") forward DCT, division by 8 x the quantiser with rounding half away
from zero, the standard Huffman tables (ITU-T T.81 K.3-K.6) and a JFIF header.  Pinned BYTE FOR BYTE against
libjpeg-turbo through ``cv2.imencode`` (``tests/test_jpeg.py``); the only bytes that differ from what TensorFlow writes are
the JFIF density fields, which TensorFlow sets to 300 x 300 dpi (``density`` argument, default as TensorFlow).
"""
import numpy as np

from .jpegcodec import ZIGZAG

STD_HUFF = {  # (class << 4 | id): (bits[16], values)  — ITU-T T.81 Annex K.3.3
    0: ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0],
        [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]),
    16: ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125],
        [1, 2, 3, 0, 4, 17, 5, 18, 33, 49, 65, 6, 19, 81, 97, 7, 34, 113, 20, 50, 129, 145, 161, 8, 35, 66, 177, 193,
        21, 82, 209, 240, 36, 51, 98, 114, 130, 9, 10, 22, 23, 24, 25, 26, 37, 38, 39, 40, 41, 42, 52, 53, 54, 55,
        56, 57, 58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104, 105,
        106, 115, 116, 117, 118, 119, 120, 121, 122, 131, 132, 133, 134, 135, 136, 137, 138, 146, 147, 148, 149, 150,
        151, 152, 153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184, 185, 186,
        194, 195, 196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 225, 226, 227, 228,
        229, 230, 231, 232, 233, 234, 241, 242, 243, 244, 245, 246, 247, 248, 249, 250]),
    1: ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0],
        [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]),
    17: ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119],
        [0, 1, 2, 3, 17, 4, 5, 33, 49, 6, 18, 65, 81, 7, 97, 113, 19, 34, 50, 129, 8, 20, 66, 145, 161, 177, 193, 9,
        35, 51, 82, 240, 21, 98, 114, 209, 10, 22, 36, 52, 225, 37, 241, 23, 24, 25, 26, 38, 39, 40, 41, 42, 53, 54,
        55, 56, 57, 58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104,
        105, 106, 115, 116, 117, 118, 119, 120, 121, 122, 130, 131, 132, 133, 134, 135, 136, 137, 138, 146, 147, 148,
        149, 150, 151, 152, 153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184,
        185, 186, 194, 195, 196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 226, 227,
        228, 229, 230, 231, 232, 233, 234, 242, 243, 244, 245, 246, 247, 248, 249, 250]),
}
# Annex K.1 quantisation tables (natural order): luminance, chrominance
STD_QUANT = (
    [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
     80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
     95, 98, 112, 100, 103, 99],
    [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99,
     99, 99] + [99] * 32,
)


def quant_tables(quality: int):
    """jpeg_set_quality(quality, force_baseline=TRUE): two tables in natural order."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return [np.clip((np.array(t, np.int64) * scale + 50) // 100, 1, 255).astype(np.int32) for t in STD_QUANT]


def _codes(bits, vals):
    """symbol -> (code, length), canonical assignment (T.81 Annex C)."""
    out = {}
    code = 0
    k = 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            out[vals[k]] = (code, ln)
            code += 1
            k += 1
        code <<= 1
    return out


def rgb_to_ycc(img):
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16
    return y, cb, cr


def _pad(pl, mh, mw):
    h, w = pl.shape
    H, W = -(-h // mh) * mh, -(-w // mw) * mw
    return np.pad(pl, ((0, H - h), (0, W - w)), mode="edge")


def _h2v2(pl):
    s = pl[0::2, 0::2] + pl[0::2, 1::2] + pl[1::2, 0::2] + pl[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)
    return (s + bias) >> 2


_C = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _fdct_1d(d, first):
    """One pass of the IJG accurate integer forward DCT; d = list of 8 arrays."""
    c = _C
    t0, t7, t1, t6 = d[0] + d[7], d[0] - d[7], d[1] + d[6], d[1] - d[6]
    t2, t5, t3, t4 = d[2] + d[5], d[2] - d[5], d[3] + d[4], d[3] - d[4]
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    sh = 11 if first else 15
    ds = lambda x: (x + (1 << (sh - 1))) >> sh
    out = [None] * 8
    if first:
        out[0], out[4] = (t10 + t11) << 2, (t10 - t11) << 2
    else:
        out[0], out[4] = (t10 + t11 + 2) >> 2, (t10 - t11 + 2) >> 2
    z1 = (t12 + t13) * c["f0_541"]
    out[2] = ds(z1 + t13 * c["f0_765"])
    out[6] = ds(z1 + t12 * (-c["f1_847"]))
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * c["f1_175"]
    t4, t5, t6, t7 = t4 * c["f0_298"], t5 * c["f2_053"], t6 * c["f3_072"], t7 * c["f1_501"]
    z1, z2 = z1 * (-c["f0_899"]), z2 * (-c["f2_562"])
    z3, z4 = z3 * (-c["f1_961"]) + z5, z4 * (-c["f0_390"]) + z5
    out[7], out[5], out[3], out[1] = ds(t4 + z1 + z3), ds(t5 + z2 + z4), ds(t6 + z2 + z3), ds(t7 + z1 + z4)
    return out


def fdct_quant(pl, q):
    """(H,W) plane (multiples of 8) -> (H/8, W/8, 64) quantised coefficients in natural order."""
    h, w = pl.shape
    blk = (pl.astype(np.int64) - 128).reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)     # (by,bx,row,col)
    rows = np.stack(_fdct_1d([blk[..., c] for c in range(8)], True), axis=-1)                  # along each row
    cols = np.stack(_fdct_1d([rows[..., r, :] for r in range(8)], False), axis=-2)             # along each column
    d = q.reshape(8, 8).astype(np.int64) * 8
    a = np.abs(cols)
    out = np.sign(cols) * ((a + (d >> 1)) // d)
    return out.reshape(h // 8, w // 8, 64).astype(np.int32)


class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, code, ln):
        self.acc = (self.acc << ln) | (code & ((1 << ln) - 1))
        self.n += ln
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def _encode_block(bw, blk, pred, dc, ac):
    diff = int(blk[0]) - pred
    t2 = diff - 1 if diff < 0 else diff
    nb = abs(diff).bit_length()
    bw.put(*dc[nb])
    if nb:
        bw.put(t2, nb)
    r = 0
    for k in range(1, 64):
        v = int(blk[ZIGZAG[k]])
        if v == 0:
            r += 1
            continue
        while r > 15:
            bw.put(*ac[0xF0])
            r -= 16
        nb = abs(v).bit_length()
        bw.put(*ac[(r << 4) | nb])
        bw.put(v - 1 if v < 0 else v, nb)
        r = 0
    if r:
        bw.put(*ac[0])
    return int(blk[0])


def header(h, w, nc, qts, density=(1, 300, 300)):
    """SOI .. SOS the way libjpeg's jcmarker.c lays them out for a baseline file with the standard tables."""
    out = bytearray(b"\xff\xd8\xff\xe0\x00\x10JFIF\x00\x01\x01")
    out += bytes([density[0]]) + density[1].to_bytes(2, "big") + density[2].to_bytes(2, "big") + b"\x00\x00"
    for i in range(2 if nc == 3 else 1):
        out += b"\xff\xdb\x00\x43" + bytes([i]) + bytes(int(qts[i][z]) for z in ZIGZAG)
    out += b"\xff\xc0" + (8 + 3 * nc).to_bytes(2, "big") + b"\x08" + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([nc])
    out += b"\x01\x22\x00\x02\x11\x01\x03\x11\x01" if nc == 3 else b"\x01\x11\x00"
    for key in ((0, 16, 1, 17) if nc == 3 else (0, 16)):
        bits, vals = STD_HUFF[key]
        out += b"\xff\xc4" + (19 + len(vals)).to_bytes(2, "big") + bytes([key]) + bytes(bits) + bytes(vals)
    out += b"\xff\xda" + (6 + 2 * nc).to_bytes(2, "big") + bytes([nc])
    out += b"\x01\x00\x02\x11\x03\x11" if nc == 3 else b"\x01\x00"
    return bytes(out + b"\x00\x3f\x00")


def encode_jpeg(img: np.ndarray, quality: int = 100, density=(1, 300, 300)) -> bytes:
    """(H,W,3) RGB or (H,W,1)/(H,W) grey uint8 -> baseline JFIF bytes (4:2:0 for colour), as libjpeg writes them."""
    img = np.asarray(img, np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, nc = img.shape
    qts = quant_tables(quality)
    codes = {k: _codes(*v) for k, v in STD_HUFF.items()}
    bw = _BitWriter()
    if nc == 1:
        cy = fdct_quant(_pad(img[..., 0].astype(np.int64), 8, 8), qts[0])
        pred = 0
        for by in range(cy.shape[0]):
            for bx in range(cy.shape[1]):
                pred = _encode_block(bw, cy[by, bx], pred, codes[0], codes[16])
    else:
        y, cb, cr = rgb_to_ycc(img)
        # libjpeg replicates the right edge at full resolution inside the down-sampler, but the bottom edge in two steps:
        # to an even number of rows before down-sampling, then the last DOWN-SAMPLED row up to the iMCU height (jcprepct.c)
        cy = fdct_quant(_pad(y, 16, 16), qts[0])
        ccb, ccr = (fdct_quant(_pad(_h2v2(_pad(p, 2, 16)), 8, 8), qts[1]) for p in (cb, cr))
        pred = [0, 0, 0]
        for my in range(ccb.shape[0]):
            for mx in range(ccb.shape[1]):
                for dy in (0, 1):
                    for dx in (0, 1):
                        if 2 * my + dy >= -(-h // 8) or 2 * mx + dx >= -(-w // 8):
                            # beyond the component's own block grid libjpeg codes a dummy block (jccoefct.c): the DC of
                            # the block before it, no AC
                            bw.put(*codes[0][0])
                            bw.put(*codes[16][0])
                            continue
                        pred[0] = _encode_block(bw, cy[2 * my + dy, 2 * mx + dx], pred[0], codes[0], codes[16])
                pred[1] = _encode_block(bw, ccb[my, mx], pred[1], codes[1], codes[17])
                pred[2] = _encode_block(bw, ccr[my, mx], pred[2], codes[1], codes[17])
    bw.flush()
    return header(h, w, nc, qts, density) + bytes(bw.out) + b"\xff\xd9"
