"""Baseline JPEG decode (oracle; test infrastructure only).

Stands in for ``tf.image.decode_jpeg(image_data, channels=0)`` behind ``ImageCoder.decode_jpeg``
(``_img_to_tf_threaded.py:36-38,51-56``) and ``_process_image`` for ``.jpg`` chips (``:97-103``).  TensorFlow is not
installable; the arithmetic it delegates is libjpeg's with its default settings (``dct_method`` default -> the
accurate integer IDCT, ``fancy_upscaling=True``), restated here from the published algorithm:
ITU-T T.81 (markers, Huffman coding, zig-zag) and the IJG reference implementation's integer pipeline
(13-bit "islow" inverse DCT, triangle-filter chroma upsampling, 16-bit fixed-point YCbCr -> RGB).
Pinned against libjpeg-turbo through ``cv2.imdecode`` and ``Pillow`` in ``tests/test_oracle_golden.py``.

Scope: 8-bit sequential Huffman (SOF0 / SOF1), one interleaved scan, 1 or 3 components, restart intervals.
Progressive, arithmetic, lossless, 12-bit, CMYK and multi-scan files raise ``Unsupported`` (codec status 3).
"""
import struct

import numpy as np


class DecodeError(Exception):
    pass


class Unsupported(DecodeError):
    pass


ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                   6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45,
                   38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63], dtype=np.int32)


def parse_jpeg(blob: bytes) -> dict:
    """Marker walk up to the first SOS -> frame header, tables, start of the entropy-coded segment."""
    if len(blob) < 4 or blob[0] != 0xFF or blob[1] != 0xD8:
        raise DecodeError("not a JPEG")
    p = 2
    qt = {}
    ht = {}
    frame = None
    restart = 0
    jfif = False
    adobe = None
    while True:
        if p + 4 > len(blob):
            raise DecodeError("truncated before SOS")
        if blob[p] != 0xFF:
            raise DecodeError("marker expected at %d" % p)
        while p < len(blob) and blob[p] == 0xFF:
            p += 1
        m = blob[p]
        p += 1
        if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01:
            continue
        if m == 0xD9:
            raise DecodeError("EOI before SOS")
        (ln,) = struct.unpack(">H", blob[p:p + 2])
        seg = blob[p + 2:p + ln]
        if ln < 2 or len(seg) != ln - 2:
            raise DecodeError("truncated segment")
        p += ln
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                q += 1
                if tq > 3 or pq > 1:
                    raise DecodeError("bad DQT")
                if pq:
                    vals = struct.unpack(">64H", seg[q:q + 128])
                    q += 128
                else:
                    vals = tuple(seg[q:q + 64])
                    q += 64
                if len(vals) != 64:
                    raise DecodeError("short DQT")
                t = np.zeros(64, np.int32)
                t[ZIGZAG] = vals                      # natural order
                qt[tq] = t
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1:q + 17])
                n = sum(counts)
                syms = list(seg[q + 17:q + 17 + n])
                if tc > 1 or th > 3 or len(counts) != 16 or len(syms) != n or n > 256:
                    raise DecodeError("bad DHT")
                ht[(tc, th)] = (counts, syms)
                q += 17 + n
        elif m in (0xC0, 0xC1):
            prec, h, w, nc = struct.unpack(">BHHB", seg[:6])
            comps = []
            for i in range(nc):
                cid, hv, tq = seg[6 + 3 * i:9 + 3 * i]
                comps.append(dict(id=cid, h=hv >> 4, v=hv & 15, tq=tq))
            frame = dict(precision=prec, height=h, width=w, comps=comps)
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise Unsupported("SOF%d (progressive / lossless / arithmetic) not handled" % (m - 0xC0))
        elif m == 0xDD:
            (restart,) = struct.unpack(">H", seg[:2])
        elif m == 0xE0 and seg[:5] == b"JFIF\0":
            jfif = True
        elif m == 0xEE and seg[:5] == b"Adobe" and len(seg) >= 12:
            adobe = seg[11]
        elif m == 0xDA:
            if frame is None:
                raise DecodeError("SOS before SOF")
            ns = seg[0]
            sel = []
            for i in range(ns):
                cs, tt = seg[1 + 2 * i:3 + 2 * i]
                sel.append((cs, tt >> 4, tt & 15))
            ss, se, ahal = seg[1 + 2 * ns:4 + 2 * ns]
            break
    nc = len(frame["comps"])
    if frame["precision"] != 8:
        raise Unsupported("12-bit JPEG")
    if nc not in (1, 3):
        raise Unsupported("%d-component JPEG" % nc)
    if frame["height"] == 0 or frame["width"] == 0:
        raise Unsupported("DNL / empty frame")
    if frame["height"] * frame["width"] * nc >= 1 << 29:
        raise Unsupported("image too large")           # tf.image.decode_jpeg's own limit (jpeg_mem.cc)
    if ns != nc or [s[0] for s in sel] != [c["id"] for c in frame["comps"]]:
        raise Unsupported("multi-scan sequential JPEG")
    if (ss, se, ahal) != (0, 63, 0):
        raise DecodeError("bad spectral selection for a sequential scan")
    for c, (_, td, ta) in zip(frame["comps"], sel):
        c["td"], c["ta"] = td, ta
        if not (1 <= c["h"] <= 4 and 1 <= c["v"] <= 4):
            raise DecodeError("bad sampling factors")
        if c["tq"] not in qt or (0, td) not in ht or (1, ta) not in ht:
            raise DecodeError("missing table")
    if nc == 1:                                       # a single-component scan is never interleaved (T.81 A.2.2)
        frame["comps"][0]["h"] = frame["comps"][0]["v"] = 1
    ids = [c["id"] for c in frame["comps"]]
    # libjpeg's colour-space guess (jdapimin.c default_decompress_parms)
    if nc == 1:
        ycc = False
    elif jfif:
        ycc = True
    elif adobe is not None:
        ycc = adobe != 0
    else:
        ycc = ids != [82, 71, 66]
    if sum(c["h"] * c["v"] for c in frame["comps"]) > 10:
        raise DecodeError("MCU too large")
    frame.update(qt=qt, ht=ht, restart=restart, scan=p, ycc=ycc)
    return frame


def _huff_lookup(counts, syms):
    """code -> (length, symbol) tables per T.81 Annex C, as arrays indexed by length."""
    mincode, maxcode, valptr = [0] * 17, [-1] * 17, [0] * 17
    code = 0
    k = 0
    for ln in range(1, 17):
        valptr[ln] = k
        mincode[ln] = code
        code += counts[ln - 1]
        k += counts[ln - 1]
        maxcode[ln] = code - 1 if counts[ln - 1] else -1
        code <<= 1
    return mincode, maxcode, valptr, syms


class _Bits:
    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0
        self.marker = None

    def _fill(self):
        while self.n <= 24:
            if self.marker is not None or self.p >= len(self.d):
                b = 0                                  # libjpeg feeds zeros after a marker / end of data
            else:
                b = self.d[self.p]
                if b == 0xFF:
                    nb = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nb == 0:
                        self.p += 2
                    else:
                        self.marker = nb
                        b = 0
                else:
                    self.p += 1
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def get(self, k):
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def decode(self, tab):
        mincode, maxcode, valptr, syms = tab
        code = 0
        for ln in range(1, 17):
            code = (code << 1) | self.get(1)
            if maxcode[ln] >= 0 and code <= maxcode[ln] and code >= mincode[ln]:
                return syms[valptr[ln] + code - mincode[ln]]
        raise DecodeError("bad Huffman code")

    def restart(self, expect):
        """Byte-align, consume RSTn."""
        self.acc = self.n = 0
        if self.marker is None:
            while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and self.d[self.p + 1] not in (0, 0xFF)):
                self.p += 1
            if self.p + 1 >= len(self.d):
                raise DecodeError("restart marker missing")
            self.marker = self.d[self.p + 1]
        if self.marker != 0xD0 + expect:
            raise DecodeError("wrong restart marker")
        self.p += 2
        self.marker = None


def _extend(v, s):
    return v - (1 << s) + 1 if s and v < (1 << (s - 1)) else v


def decode_coefficients(blob: bytes, fr: dict):
    """Entropy decode -> per component int32 array (blocks_down, blocks_across, 64) in natural order (quantised)."""
    comps = fr["comps"]
    hmax = max(c["h"] for c in comps)
    vmax = max(c["v"] for c in comps)
    mx = -(-fr["width"] // (8 * hmax))
    my = -(-fr["height"] // (8 * vmax))
    coefs = [np.zeros((my * c["v"], mx * c["h"], 64), np.int32) for c in comps]
    tabs = {k: _huff_lookup(*v) for k, v in fr["ht"].items()}
    br = _Bits(blob, fr["scan"])
    pred = [0] * len(comps)
    ri = fr["restart"]
    nrst = 0
    for m in range(mx * my):
        if ri and m and m % ri == 0:
            br.restart(nrst & 7)
            nrst += 1
            pred = [0] * len(comps)
        y0, x0 = divmod(m, mx)
        for ci, c in enumerate(comps):
            dc, ac = tabs[(0, c["td"])], tabs[(1, c["ta"])]
            for by in range(c["v"]):
                for bx in range(c["h"]):
                    blk = coefs[ci][y0 * c["v"] + by, x0 * c["h"] + bx]
                    s = br.decode(dc)
                    if s > 11:
                        raise DecodeError("bad DC category")
                    pred[ci] += _extend(br.get(s), s)
                    blk[0] = pred[ci]
                    k = 1
                    while k < 64:
                        rs = br.decode(ac)
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        if k > 63:
                            raise DecodeError("AC run past the block")
                        blk[ZIGZAG[k]] = _extend(br.get(s), s)
                        k += 1
    return coefs


_F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(x, shift_in_even, descale):
    """One pass of the IJG accurate integer IDCT over the second-to-last axis' 8 entries given as a list of arrays."""
    f = _F
    z2, z3 = x[2], x[6]
    z1 = (z2 + z3) * f["f0_541"]
    tmp2 = z1 + z3 * (-f["f1_847"])
    tmp3 = z1 + z2 * f["f0_765"]
    z2, z3 = x[0], x[4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[7], x[5], x[3], x[1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * f["f1_175"]
    tmp0 = tmp0 * f["f0_298"]
    tmp1 = tmp1 * f["f2_053"]
    tmp2 = tmp2 * f["f3_072"]
    tmp3 = tmp3 * f["f1_501"]
    z1 = z1 * (-f["f0_899"])
    z2 = z2 * (-f["f2_562"])
    z3 = z3 * (-f["f1_961"]) + z5
    z4 = z4 * (-f["f0_390"]) + z5
    tmp0 = tmp0 + z1 + z3
    tmp1 = tmp1 + z2 + z4
    tmp2 = tmp2 + z2 + z3
    tmp3 = tmp3 + z1 + z4
    rnd = 1 << (descale - 1)
    outs = [tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3]
    return [(o + rnd) >> descale for o in outs]


def idct_blocks(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """(…,64) quantised coefficients -> (…,8,8) uint8 samples: dequantise, columns (descale 11), rows (descale 18),
    +128 and the 10-bit wrap of libjpeg's range-limit table."""
    w = (coef.astype(np.int64) * q.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))
    w = w.astype(np.int32).astype(np.int64)           # 32-bit workspace, arithmetic carried in 64 then wrapped
    cols = _idct_1d([w[..., r, :] for r in range(8)], 13, 11)      # each (…,8): output rows of the column pass
    ws = np.stack(cols, axis=-2)                                    # (…,8,8)
    ws = ws.astype(np.int32).astype(np.int64)
    rows = _idct_1d([ws[..., :, c] for c in range(8)], 13, 18)
    out = np.stack(rows, axis=-1)
    s = ((out & 0x3FF) ^ 512) - 512
    return np.clip(s + 128, 0, 255).astype(np.uint8)


def _plane(samples: np.ndarray) -> np.ndarray:
    by, bx = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def _upsample(pl: np.ndarray, dw: int, dh: int, hr: int, vr: int, W: int, H: int) -> np.ndarray:
    """Component plane (padded) -> (H,W) at full resolution the way libjpeg's jdsample.c picks its method."""
    p = pl[:dh, :dw].astype(np.int32)
    fancy = dw > 2
    if hr == 1 and vr == 1:
        out = p
    elif hr == 2 and vr == 1 and fancy:
        out = _h2_fancy(p)
    elif hr == 2 and vr == 2 and fancy:
        up = np.vstack([p[:1], p[:-1]])
        dn = np.vstack([p[1:], p[-1:]])
        out = np.empty((2 * dh, 2 * dw), np.int32)
        out[0::2] = _h2v2_rows(p * 3 + up)
        out[1::2] = _h2v2_rows(p * 3 + dn)
    elif hr == 1 and vr == 2:                           # no width test for this one in jinit_upsampler
        up = np.vstack([p[:1], p[:-1]])
        dn = np.vstack([p[1:], p[-1:]])
        out = np.empty((2 * dh, dw), np.int32)
        out[0::2] = (p * 3 + up + 1) >> 2
        out[1::2] = (p * 3 + dn + 2) >> 2
    else:
        out = np.repeat(np.repeat(p, vr, axis=0), hr, axis=1)
    return out[:H, :W]


def _h2_fancy(p):
    n = p.shape[1]
    out = np.empty((p.shape[0], 2 * n), np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (p * 3 + left + 1) >> 2
    out[:, 1::2] = (p * 3 + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def _h2v2_rows(cs):
    """cs = 3*near row + far row (column sums); horizontal triangle filter with the 8 / 7 rounding pattern."""
    n = cs.shape[1]
    out = np.empty((cs.shape[0], 2 * n), np.int32)
    left = np.concatenate([cs[:, :1], cs[:, :-1]], axis=1)
    right = np.concatenate([cs[:, 1:], cs[:, -1:]], axis=1)
    out[:, 0::2] = (cs * 3 + left + 8) >> 4
    out[:, 1::2] = (cs * 3 + right + 7) >> 4
    out[:, 0] = (cs[:, 0] * 4 + 8) >> 4
    out[:, -1] = (cs[:, -1] * 4 + 7) >> 4
    return out


def ycc_to_rgb(y, cb, cr):
    """IJG jdcolor.c: 16-bit fixed point tables, results clamped."""
    cb = cb.astype(np.int64) - 128
    cr = cr.astype(np.int64) - 128
    half = 1 << 15
    r = y + ((91881 * cr + half) >> 16)
    b = y + ((116130 * cb + half) >> 16)
    g = y + ((-22554 * cb + half - 46802 * cr) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def jpeg_shape(blob: bytes):
    fr = parse_jpeg(blob)
    return fr["height"], fr["width"], len(fr["comps"])


def decode_jpeg(blob: bytes) -> np.ndarray:
    """-> (H,W,1) grey or (H,W,3) RGB uint8, as ``tf.image.decode_jpeg(channels=0)`` returns it."""
    fr = parse_jpeg(blob)
    comps = fr["comps"]
    H, W = fr["height"], fr["width"]
    hmax = max(c["h"] for c in comps)
    vmax = max(c["v"] for c in comps)
    coefs = decode_coefficients(blob, fr)
    planes = []
    for c, cf in zip(comps, coefs):
        if hmax % c["h"] or vmax % c["v"]:
            raise Unsupported("fractional sampling ratio")
        pl = _plane(idct_blocks(cf, fr["qt"][c["tq"]]))
        dw = -(-W * c["h"] // hmax)
        dh = -(-H * c["v"] // vmax)
        planes.append(_upsample(pl, dw, dh, hmax // c["h"], vmax // c["v"], W, H))
    if len(planes) == 1:
        return planes[0].astype(np.uint8)[..., None]
    if fr["ycc"]:
        return ycc_to_rgb(planes[0].astype(np.int64), planes[1], planes[2])
    return np.stack(planes, axis=-1).astype(np.uint8)
