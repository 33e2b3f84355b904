"""tf.train.Example wire format + convert_to_example / parse_*_proto (oracle; test infrastructure only).

Restates ``_tfrecord_image_translation.py``:
  * ``_int64_feature`` ``:7-16``, ``_float64_feature`` ``:19-35`` (FloatList == float32),
    ``_bytes_feature`` ``:38-52``, ``convert_to_example`` ``:55-211`` (type dispatch ``:160-197``,
    the eight keys ``:199-209``)
  * templates ``:216-225`` / ``:231-241`` and parsers ``_parse_byteslist_proto :244-266``,
    ``parse_8bit_array_proto :296-316``, ``parse_higher_dtype_array_proto :389-415``,
    ``parse_encoded_rgb_img_proto :269-293``, ``parse_encoded_gdal_proto_wrapped :332-346`` /
    ``_eager :349-386``.
The protobuf encoding is restated from the public spec + ``example.proto``/``feature.proto`` field
numbers with no protobuf dependency; tests compare it byte-for-byte with
``google.protobuf``'s deterministic serialisation of a dynamic ``tensorflow.Example`` descriptor.

Map order: the reference calls plain ``SerializeToString()`` (``_img_to_tf_mp.py:141``), whose map
order is back-end dependent; the contract here is sorted keys.
"""
import struct

import numpy as np

KEYS = ("image/image_data", "image/height", "image/width", "image/channels",
        "target/target_data", "target/height", "target/width", "identifier")


# ----------------------------------------------------------------------------- encoding
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_field(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


class Feature:
    """One tf.train.Feature: kind in {'bytes','float','int64'} and a list of values."""

    def __init__(self, kind, values):
        self.kind, self.values = kind, values

    def serialize(self) -> bytes:
        if self.kind == "bytes":
            inner = b"".join(_len_field(1, bytes(v)) for v in self.values)
            return _len_field(1, inner)
        if self.kind == "float":
            arr = np.asarray(self.values, dtype="<f4").ravel()
            inner = _len_field(1, arr.tobytes()) if arr.size else b""
            return _len_field(2, inner)
        if self.kind == "int64":
            packed = b"".join(_varint(int(v)) for v in self.values)
            inner = _len_field(1, packed) if len(self.values) else b""
            return _len_field(3, inner)
        raise ValueError(self.kind)


class Example:
    def __init__(self, features: dict):
        self.features = features

    def SerializeToString(self, deterministic=True) -> bytes:
        entries = b""
        for key in sorted(self.features) if deterministic else self.features:
            entry = _len_field(1, key.encode()) + _len_field(2, self.features[key].serialize())
            entries += _len_field(1, entry)
        return _len_field(1, entries)


def _int64_feature(value):
    # _tfrecord_image_translation.py:12-16
    if isinstance(value, np.ndarray):
        value = value.flatten().tolist()
    elif not isinstance(value, list):
        value = [value]
    return Feature("int64", value)


def _float64_feature(value):
    # :25-35 — tf.train.FloatList stores float32 despite the name
    if isinstance(value, np.ndarray):
        value = value.flatten()
    elif not isinstance(value, list):
        value = [value]
    return Feature("float", value)


def _bytes_feature(value):
    # :41-52
    if isinstance(value, np.ndarray):
        value = [value.tobytes()]
    elif not isinstance(value, list):
        value = [value]
    return Feature("bytes", value)


def convert_to_example(img_data, target_data, img_h, img_w, img_b, target_h, target_w, identifier):
    """Restatement of _tfrecord_image_translation.py:160-211."""
    image_is_bytes = False
    target_is_bytes = False
    if isinstance(img_data, bytes):
        image_is_bytes = True
    elif isinstance(img_data, np.ndarray):
        if img_data.dtype == "uint8":
            image_is_bytes = True
    if isinstance(target_data, bytes):
        target_is_bytes = True
    elif isinstance(target_data, np.ndarray):
        if target_data.dtype == "uint8" and image_is_bytes:
            target_is_bytes = True
    if image_is_bytes and target_is_bytes:
        wi, wt = _bytes_feature(img_data), _bytes_feature(target_data)
    else:
        wi, wt = _float64_feature(img_data), _float64_feature(target_data)
    ident = identifier if isinstance(identifier, bytes) else str(identifier).encode("utf-8")
    return Example({
        "image/image_data": wi,
        "image/height": _int64_feature(int(img_h)),
        "image/width": _int64_feature(int(img_w)),
        "image/channels": _int64_feature(int(img_b)),
        "target/target_data": wt,
        "target/height": _int64_feature(int(target_h)),
        "target/width": _int64_feature(int(target_w)),
        "identifier": _bytes_feature(ident),
    })


# ----------------------------------------------------------------------------- decoding
def _rd_varint(b, p):
    v = 0
    s = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return v & 0xFFFFFFFFFFFFFFFF, p
        if s > 63 + 7:
            raise ValueError("varint too long")


def _fields(b, p, end):
    """Yield (field, wire_type, value | (start,end)) over a message body."""
    while p < end:
        tag, p = _rd_varint(b, p)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, p = _rd_varint(b, p)
            yield f, wt, v
        elif wt == 1:
            yield f, wt, (p, p + 8)
            p += 8
        elif wt == 2:
            n, p = _rd_varint(b, p)
            if p + n > end:
                raise ValueError("truncated field")
            yield f, wt, (p, p + n)
            p += n
        elif wt == 5:
            yield f, wt, (p, p + 4)
            p += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
    if p != end:
        raise ValueError("message overrun")


def _parse_feature(b, s, e):
    kind, values = None, []
    for f, wt, v in _fields(b, s, e):
        if wt != 2 or f not in (1, 2, 3):
            continue
        ls, le = v
        # oneof: a later member replaces an earlier one
        if f == 1:
            kind, values = "bytes", [bytes(b[a:z]) for ff, w2, (a, z) in
                                     ((ff, w2, vv) for ff, w2, vv in _fields(b, ls, le) if w2 == 2) if ff == 1]
        elif f == 2:
            kind = "float"
            chunks = []
            for ff, w2, vv in _fields(b, ls, le):
                if ff != 1:
                    continue
                if w2 == 2:
                    chunks.append(np.frombuffer(b, dtype="<f4", count=(vv[1] - vv[0]) // 4, offset=vv[0]))
                elif w2 == 5:
                    chunks.append(np.frombuffer(b, dtype="<f4", count=1, offset=vv[0]))
            values = np.concatenate(chunks) if chunks else np.zeros(0, "<f4")
        else:
            kind, values = "int64", []
            for ff, w2, vv in _fields(b, ls, le):
                if ff != 1:
                    continue
                if w2 == 2:
                    q = vv[0]
                    while q < vv[1]:
                        x, q = _rd_varint(b, q)
                        values.append(x - (1 << 64) if x >> 63 else x)
                elif w2 == 0:
                    values.append(vv - (1 << 64) if vv >> 63 else vv)
    return kind, values


def parse_example(record: bytes) -> dict:
    """Example bytes -> {key: (kind, values)}; any entry order, unknown fields skipped, last key wins."""
    b = bytes(record)
    out = {}
    for f, wt, v in _fields(b, 0, len(b)):
        if f != 1 or wt != 2:
            continue
        for f2, wt2, v2 in _fields(b, v[0], v[1]):          # Features.feature entries
            if f2 != 1 or wt2 != 2:
                continue
            key, feat = None, (None, [])
            for f3, wt3, v3 in _fields(b, v2[0], v2[1]):
                if f3 == 1 and wt3 == 2:
                    key = b[v3[0]:v3[1]].decode("utf-8")
                elif f3 == 2 and wt3 == 2:
                    feat = _parse_feature(b, v3[0], v3[1])
            if key is not None:
                out[key] = feat
    return out


class ParseError(Exception):
    """tf.io.parse_single_example raises InvalidArgumentError for a missing / mistyped key."""


def _scalar(feats, key, kind):
    if key not in feats or feats[key][0] != kind:
        raise ParseError("Feature: %s (data type: %s) is required but could not be found." % (key, kind))
    vals = feats[key][1]
    if len(vals) != 1:
        raise ParseError("Key: %s. Can't parse serialized Example (expected 1 value, got %d)." % (key, len(vals)))
    return vals[0]


def _parse_byteslist_proto(example_proto):
    # :244-266 with template :216-225
    f = parse_example(example_proto)
    shp = tuple(np.int32(_scalar(f, k, "int64")) for k in ("image/height", "image/width", "image/channels"))
    tshp = tuple(np.int32(_scalar(f, k, "int64")) for k in ("target/height", "target/width"))
    return (_scalar(f, "image/image_data", "bytes"), shp,
            _scalar(f, "target/target_data", "bytes"), tshp, _scalar(f, "identifier", "bytes"))


def parse_8bit_array_proto(example_proto):
    # :296-316
    ib, ishp, tb, tshp, ident = _parse_byteslist_proto(example_proto)
    img = np.frombuffer(ib, dtype=np.uint8)
    assert img.shape[0] == int(ishp[0]) * int(ishp[1]) * int(ishp[2]), "Decoded shape is %r - does not match" % (img.shape,)
    tgt = np.frombuffer(tb, dtype=np.uint8)
    assert tgt.shape[0] == int(tshp[0]) * int(tshp[1])
    return img.reshape([int(x) for x in ishp]), tgt.reshape([int(x) for x in tshp]), ident


def parse_higher_dtype_array_proto(example_proto):
    # :389-415 with template :231-241 (float sequences, allow_missing=True -> empty if absent)
    f = parse_example(example_proto)
    h, w, c = (int(_scalar(f, k, "int64")) for k in ("image/height", "image/width", "image/channels"))
    th, tw = (int(_scalar(f, k, "int64")) for k in ("target/height", "target/width"))

    def seq(key):
        if key not in f:
            return np.zeros(0, np.float32)
        if f[key][0] != "float":
            raise ParseError("Feature: %s data type mismatch" % key)
        return np.asarray(f[key][1], dtype=np.float32)
    img = seq("image/image_data").reshape(h, w, c)
    tgt = seq("target/target_data").reshape(th, tw)
    return img, tgt, _scalar(f, "identifier", "bytes")


def parse_encoded_rgb_img_proto(example_proto):
    # :269-293 — tf.io.decode_image on both blobs (PNG here)
    from .imagecodecs import decode_image
    ib, _, tb, _, ident = _parse_byteslist_proto(example_proto)
    return decode_image(ib), decode_image(tb), ident


def parse_encoded_gdal_proto_eager(example_proto):
    # :349-386 — native dtype, shape asserts
    from .imagecodecs import decode_image
    ib, ishp, tb, tshp, ident = _parse_byteslist_proto(example_proto)
    img = decode_image(ib, png_as_tf=False)                     # rasterio -> GDAL's view of the file
    assert img.shape == tuple(int(x) for x in ishp)
    tgt = decode_image(tb, png_as_tf=False)
    assert tgt.shape[0] == int(tshp[0]) and tgt.shape[1] == int(tshp[1])
    return img, tgt, ident


def parse_encoded_gdal_proto_wrapped(example_proto):
    # :319-346 — always float32; unlike _eager it never compares with the recorded shape
    from .imagecodecs import decode_image
    ib, _, tb, _, ident = _parse_byteslist_proto(example_proto)
    return (decode_image(ib, png_as_tf=False).astype(np.float32), decode_image(tb, png_as_tf=False).astype(np.float32), ident)
