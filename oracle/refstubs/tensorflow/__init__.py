"""Minimal eager `tensorflow` stand-in so that the UNMODIFIED reference can be imported and run (test infrastructure).

Only the calls the reference makes are provided (file:line in /root/reference/dl_segmentation_utils):
  _tfrecord_image_translation.py:16,35,52 (tf.train.*List / Feature), :161-211 (tf.constant, tf.compat.as_bytes,
  tf.train.Example / Features), :216-241 (tf.io.FixedLen*Feature, dtypes), :249-263 (parse_single_example, tf.cast),
  :283,289 (tf.io.decode_image), :306-314 (decode_raw, reshape, squeeze, stack), :345 (tf.numpy_function);
  _img_to_tf_mp.py:43,119,141,150,215-216 (gfile, TFRecordWriter); _img_to_tf_threaded.py:37,51,59 (tf.image.*),
  :246,262 (tf.train.Coordinator).
Arithmetic comes from google.protobuf, NumPy, Pillow and OpenCV — see ../README.md.  Never imported by the product.
"""
import builtins as _builtins
import glob as _glob
import io as _io
import struct as _struct
import sys as _sys
import types as _types

import numpy as _np

__version__ = "0.0-refstub"


# ------------------------------------------------------------------------------------------------ dtypes
class DType:
    def __init__(self, name, np_dtype):
        self.name, self.as_numpy_dtype = name, np_dtype

    def __eq__(self, other):
        if isinstance(other, DType):
            return self.name == other.name
        if isinstance(other, str):
            return self.name == other
        try:
            return self.as_numpy_dtype is not None and _np.dtype(other) == _np.dtype(self.as_numpy_dtype)
        except TypeError:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "tf.%s" % self.name


string = DType("string", None)
uint8 = DType("uint8", _np.uint8)
uint16 = DType("uint16", _np.uint16)
int16 = DType("int16", _np.int16)
int32 = DType("int32", _np.int32)
int64 = DType("int64", _np.int64)
float32 = DType("float32", _np.float32)
float64 = DType("float64", _np.float64)
bool = DType("bool", _np.bool_)
_ALL = [string, uint8, uint16, int16, int32, int64, float32, float64, bool]


def as_dtype(d):
    if isinstance(d, DType):
        return d
    if isinstance(d, str):
        for t in _ALL:
            if t.name == d:
                return t
    nd = _np.dtype(d)
    for t in _ALL:
        if t.as_numpy_dtype is not None and _np.dtype(t.as_numpy_dtype) == nd:
            return t
    raise TypeError("unsupported dtype %r" % (d,))


# ------------------------------------------------------------------------------------------------ tensors
class TensorShape(tuple):
    pass


class EagerTensor:
    """Eager tensor: a NumPy array, or Python bytes for a scalar tf.string."""

    def __init__(self, value, dtype=None):
        if isinstance(value, EagerTensor):
            value = value._v
        if isinstance(value, (bytes, bytearray, _np.bytes_)):
            self._v, self.dtype = bytes(value), string
        else:
            a = _np.asarray(value) if dtype is None else _np.asarray(value, dtype=as_dtype(dtype).as_numpy_dtype)
            if a.dtype == _np.int64 and dtype is None and isinstance(value, int):
                a = a.astype(_np.int32)                      # tf.constant(0) is int32
            self._v, self.dtype = a, as_dtype(a.dtype)

    def numpy(self):
        return self._v

    @property
    def shape(self):
        return TensorShape(()) if isinstance(self._v, bytes) else TensorShape(self._v.shape)

    # scalar arithmetic / comparisons used by the reference's asserts (:307-313, :377-384)
    def _other(self, o):
        return o._v if isinstance(o, EagerTensor) else o

    def __mul__(self, o):
        return EagerTensor(self._v * self._other(o))

    __rmul__ = __mul__

    def __eq__(self, o):
        return EagerTensor(_np.asarray(self._v == self._other(o)))

    def __ne__(self, o):
        return EagerTensor(_np.asarray(self._v != self._other(o)))

    def __hash__(self):
        return id(self)

    def __bool__(self):
        return _builtins.bool(self._v)

    def __int__(self):
        return int(self._v)

    def __index__(self):
        return int(self._v)

    def __len__(self):
        return len(self._v)

    def __getitem__(self, k):
        return EagerTensor(self._v[k])

    def __array__(self, dtype=None, copy=None):
        return _np.asarray(self._v, dtype=dtype)

    def __repr__(self):
        return "<refstub tf.Tensor dtype=%s value=%r>" % (self.dtype.name, self._v if isinstance(self._v, bytes) and len(self._v) < 40 else "...")


Tensor = EagerTensor


def constant(value, dtype=None):
    return EagerTensor(value, dtype)


def convert_to_tensor(value, dtype=None):
    return EagerTensor(value, dtype)


def _np_of(x):
    return x._v if isinstance(x, EagerTensor) else x


def cast(x, dtype):
    dtype = as_dtype(dtype)
    v = _np_of(x)
    if dtype == string:
        return EagerTensor(v)
    return EagerTensor(_np.asarray(v).astype(dtype.as_numpy_dtype))


def reshape(tensor, shape):
    shp = [int(_np_of(s)) for s in (_np_of(shape) if not isinstance(shape, (list, tuple)) else shape)]
    v = _np.asarray(_np_of(tensor))
    want = 1
    for s in shp:
        want *= s
    if want != v.size:
        raise errors.InvalidArgumentError("Input to reshape is a tensor with %d values, but the requested shape has %d" % (v.size, want))
    return EagerTensor(v.reshape(shp))


def squeeze(x):
    return EagerTensor(_np.squeeze(_np.asarray(_np_of(x))))


def stack(values):
    return EagerTensor(_np.stack([_np.asarray(_np_of(v)) for v in values]))


def numpy_function(func, inp, Tout):
    res = func(*[_np_of(x) for x in inp])
    if not isinstance(res, (list, tuple)):
        res = [res]
    out = []
    for r, t in zip(res, Tout):
        r = _np.asarray(r)
        if r.dtype != _np.dtype(as_dtype(t).as_numpy_dtype):
            raise errors.InvalidArgumentError("numpy_function: returned %s, declared %s" % (r.dtype, t))
        out.append(EagerTensor(r))
    return out


# ------------------------------------------------------------------------------------------------ errors
class _Errors:
    class OpError(Exception):
        pass

    class InvalidArgumentError(OpError):
        pass

    class DataLossError(OpError):
        pass

    class NotFoundError(OpError):
        pass


errors = _Errors()


# ------------------------------------------------------------------------------------------------ compat
class _Compat:
    @staticmethod
    def as_bytes(s, encoding="utf-8"):
        if isinstance(s, (bytes, bytearray)):
            return bytes(s)
        if isinstance(s, str):
            return s.encode(encoding)
        raise TypeError("Expected binary or unicode string, got %r" % (s,))


compat = _Compat()


# ------------------------------------------------------------------------------------------------ tf.train
def _example_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto(name="refstub_example.proto", package="tensorflow", syntax="proto3")

    def msg(name):
        m = fdp.message_type.add()
        m.name = name
        return m
    msg("BytesList").field.add(name="value", number=1, type=12, label=3)
    f = msg("FloatList").field.add(name="value", number=1, type=2, label=3)
    f.options.packed = True
    f = msg("Int64List").field.add(name="value", number=1, type=3, label=3)
    f.options.packed = True
    fe = msg("Feature")
    fe.oneof_decl.add(name="kind")
    for i, (n, t) in enumerate([("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List")]):
        fe.field.add(name=n, number=i + 1, type=11, label=1, type_name=".tensorflow." + t, oneof_index=0)
    fs = msg("Features")
    en = fs.nested_type.add(name="FeatureEntry")
    en.options.map_entry = True
    en.field.add(name="key", number=1, type=9, label=1)
    en.field.add(name="value", number=2, type=11, label=1, type_name=".tensorflow.Feature")
    fs.field.add(name="feature", number=1, type=11, label=3, type_name=".tensorflow.Features.FeatureEntry")
    msg("Example").field.add(name="features", number=1, type=11, label=1, type_name=".tensorflow.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow." + n))
            for n in ("BytesList", "FloatList", "Int64List", "Feature", "Features", "Example")}


_PB = _example_classes()


class _Wrapped:
    """Thin wrapper over a protobuf message so that SerializeToString() is deterministic (sorted map keys)."""
    _cls = None

    def __init__(self, **kw):
        kw = {k: (v._m if isinstance(v, _Wrapped) else
                  {kk: vv._m for kk, vv in v.items()} if isinstance(v, dict) else
                  (v.ravel().astype(_np.float64).tolist() if isinstance(v, _np.ndarray) else v))
              for k, v in kw.items()}
        self._m = self._cls(**kw)

    def SerializeToString(self, deterministic=True):
        return self._m.SerializeToString(deterministic=deterministic)

    def __getattr__(self, name):
        return getattr(self._m, name)


def _wrap(name):
    return type(name, (_Wrapped,), {"_cls": _PB[name]})


class Coordinator:
    def join(self, threads):
        for t in threads:
            t.join()


train = _types.SimpleNamespace(BytesList=_wrap("BytesList"), FloatList=_wrap("FloatList"), Int64List=_wrap("Int64List"),
                               Feature=_wrap("Feature"), Features=_wrap("Features"), Example=_wrap("Example"),
                               Coordinator=Coordinator)


# ------------------------------------------------------------------------------------------------ CRC-32C + records
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)
_CRC_NP = _np.array(_CRC_TABLE, dtype=_np.uint32)


def _crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    t = _CRC_TABLE
    for b in data:
        c = (c >> 8) ^ t[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def _masked(data: bytes) -> int:
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


class TFRecordWriter:
    """tf.io.TFRecordWriter with default (uncompressed) options: len | masked crc(len) | data | masked crc(data)."""

    def __init__(self, path, options=None):
        assert options is None
        self._f = open(path, "wb")

    def write(self, record):
        head = _struct.pack("<Q", len(record))
        self._f.write(head + _struct.pack("<I", _masked(head)) + record + _struct.pack("<I", _masked(record)))

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_tfrecords(path):
    """What iterating tf.data.TFRecordDataset(path) yields (parse_tfrecords.ipynb cell 4), CRCs verified."""
    out = []
    with open(path, "rb") as f:
        buf = f.read()
    o = 0
    while o < len(buf):
        (n,) = _struct.unpack_from("<Q", buf, o)
        if _struct.unpack_from("<I", buf, o + 8)[0] != _masked(buf[o:o + 8]):
            raise errors.DataLossError("corrupted record at %d" % o)
        data = buf[o + 12:o + 12 + n]
        if _struct.unpack_from("<I", buf, o + 12 + n)[0] != _masked(data):
            raise errors.DataLossError("corrupted record at %d" % o)
        out.append(EagerTensor(data))
        o += 16 + n
    return out


# ------------------------------------------------------------------------------------------------ tf.io
class FixedLenFeature:
    def __init__(self, shape, dtype, default_value=None):
        self.shape, self.dtype, self.default_value = list(shape), as_dtype(dtype), default_value


class FixedLenSequenceFeature:
    def __init__(self, shape, dtype, allow_missing=False, default_value=None):
        self.shape, self.dtype, self.allow_missing = list(shape), as_dtype(dtype), allow_missing


def _feature_values(feat, dtype, key):
    kind = feat.WhichOneof("kind")
    want = {"string": "bytes_list", "int64": "int64_list", "float32": "float_list"}[dtype.name]
    if kind is None:
        return None
    if kind != want:
        raise errors.InvalidArgumentError("Key: %s.  Data types don't match. Expected type: %s" % (key, dtype.name))
    return list(getattr(feat, kind).value)


def parse_single_example(serialized, features):
    data = _np_of(serialized)
    ex = _PB["Example"]()
    try:
        ex.ParseFromString(bytes(data))
    except Exception as e:
        raise errors.InvalidArgumentError("Could not parse example input: %s" % e)
    out = {}
    for key, spec in features.items():
        feat = ex.features.feature[key] if key in ex.features.feature else None
        vals = None if feat is None else _feature_values(feat, spec.dtype, key)
        if isinstance(spec, FixedLenFeature):
            if vals is None:
                if spec.default_value is None:
                    raise errors.InvalidArgumentError("Feature: %s (data type: %s) is required but could not be found." % (key, spec.dtype.name))
                vals = [spec.default_value]
            n = 1
            for s in spec.shape:
                n *= s
            if len(vals) != n:
                raise errors.InvalidArgumentError("Key: %s.  Can't parse serialized Example: expected %d values, got %d" % (key, n, len(vals)))
            if spec.dtype == string:
                assert spec.shape == []
                out[key] = EagerTensor(vals[0])
            else:
                out[key] = EagerTensor(_np.asarray(vals, dtype=spec.dtype.as_numpy_dtype).reshape(spec.shape))
        else:
            if vals is None:
                if not spec.allow_missing:
                    raise errors.InvalidArgumentError("Feature: %s is required but could not be found." % key)
                vals = []
            assert spec.shape == [] and spec.dtype != string
            out[key] = EagerTensor(_np.asarray(vals, dtype=spec.dtype.as_numpy_dtype))
    return out


def decode_raw(input_bytes, out_type, little_endian=True):
    dt = _np.dtype(as_dtype(out_type).as_numpy_dtype)
    data = bytes(_np_of(input_bytes))
    if len(data) % dt.itemsize:
        raise errors.InvalidArgumentError("Input to DecodeRaw has length %d that is not a multiple of %d" % (len(data), dt.itemsize))
    return EagerTensor(_np.frombuffer(data, dtype=dt.newbyteorder("<" if little_endian else ">")).astype(dt))


class _GFile:
    def __init__(self, name, mode="r"):
        self._f = open(name, mode)

    def read(self, n=-1):
        return self._f.read(n)

    def write(self, b):
        return self._f.write(b)

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _gfile_glob(pattern):
    # TF's local-filesystem GetMatchingPaths returns the matches of each directory level sorted; the reference pairs
    # images and labels by position, which only works when both globs come back in the same (sorted) order
    return sorted(_glob.glob(pattern))


# ------------------------------------------------------------------------------------------------ tf.image
def _pil_open(data):
    from PIL import Image
    im = Image.open(_io.BytesIO(bytes(data)))
    im.load()
    return im


def decode_png(contents, channels=0, dtype=uint8):
    """tf.image.decode_png(channels=0, dtype=uint8): libpng with palette -> RGB, tRNS -> alpha, sub-byte grey scaled to
    8 bits, 16 -> 8 bits by dropping the low byte (png_set_strip_16)."""
    data = bytes(_np_of(contents))
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise errors.InvalidArgumentError("Invalid PNG header")
    assert channels == 0 and as_dtype(dtype) == uint8
    try:
        im = _pil_open(data)
    except Exception as e:
        raise errors.InvalidArgumentError("Invalid PNG data: %s" % e)
    bit_depth, color_type = data[24], data[25]
    has_trns = b"tRNS" in data[:data.find(b"IDAT")] if b"IDAT" in data else False
    if color_type == 3:
        im = im.convert("RGBA" if has_trns else "RGB")
    elif color_type == 0 and bit_depth == 16:
        a = _np.asarray(im)                                    # mode I;16 / I
        arr = (a.astype(_np.uint32) >> 8).astype(_np.uint8)[:, :, None]
        return EagerTensor(arr)
    elif bit_depth == 16:
        raise NotImplementedError("refstub: 16-bit colour PNG")
    elif color_type == 0 and bit_depth < 8:
        a = _np.asarray(im.convert("L")) if im.mode != "1" else _np.asarray(im, dtype=_np.uint8) * 255
        return EagerTensor(_np.ascontiguousarray(a, dtype=_np.uint8)[:, :, None])
    elif has_trns and color_type in (0, 2):
        im = im.convert("LA" if color_type == 0 else "RGBA")
    arr = _np.asarray(im, dtype=_np.uint8)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    return EagerTensor(_np.ascontiguousarray(arr))


def decode_jpeg(contents, channels=0, ratio=1, fancy_upscaling=True, try_recover_truncated=False, acceptable_fraction=1,
                dct_method=""):
    """tf.image.decode_jpeg with its defaults: libjpeg(-turbo), JDCT_ISLOW, fancy upsampling — what Pillow does too."""
    data = bytes(_np_of(contents))
    if data[:3] != b"\xff\xd8\xff":
        raise errors.InvalidArgumentError("Invalid JPEG data or crop window")
    try:
        im = _pil_open(data)
    except Exception as e:
        raise errors.InvalidArgumentError("Invalid JPEG data: %s" % e)
    if im.mode not in ("L", "RGB"):
        im = im.convert("RGB")
    arr = _np.asarray(im, dtype=_np.uint8)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    return EagerTensor(_np.ascontiguousarray(arr))


def encode_jpeg(image, format="", quality=95, progressive=False, optimize_size=False, chroma_downsampling=True,
                density_unit="in", x_density=300, y_density=300, xmp_metadata=""):
    """tf.image.encode_jpeg: libjpeg with the standard tables, 4:2:0 for RGB, JFIF density 300x300 dpi."""
    import cv2
    arr = _np.asarray(_np_of(image))
    assert arr.dtype == _np.uint8 and arr.ndim == 3 and format == "" and not progressive and not optimize_size
    if arr.shape[2] == 1:
        src = arr[:, :, 0]
    elif arr.shape[2] == 3:
        src = arr[:, :, ::-1]
    else:
        raise errors.InvalidArgumentError("image must have 1 or 3 channels")
    ok, buf = cv2.imencode(".jpg", _np.ascontiguousarray(src), [cv2.IMWRITE_JPEG_QUALITY, int(quality)])
    assert ok
    b = bytearray(buf.tobytes())
    assert b[2:4] == b"\xff\xe0" and b[6:11] == b"JFIF\0"
    b[13] = 1 if density_unit == "in" else 2
    b[14:16] = _struct.pack(">H", x_density)
    b[16:18] = _struct.pack(">H", y_density)
    return EagerTensor(bytes(b))


def decode_image(contents, channels=None, dtype=uint8, name=None, expand_animations=True):
    data = bytes(_np_of(contents))
    if data[:8] == b"\x89PNG\r\n\x1a\n":
        return decode_png(data)
    if data[:3] == b"\xff\xd8\xff":
        return decode_jpeg(data)
    raise errors.InvalidArgumentError("Unknown image file format. One of JPEG, PNG, GIF, BMP required.")


image = _types.SimpleNamespace(decode_png=decode_png, decode_jpeg=decode_jpeg, encode_jpeg=encode_jpeg, decode_image=decode_image)
io = _types.SimpleNamespace(
    gfile=_types.SimpleNamespace(GFile=_GFile, glob=_gfile_glob),
    TFRecordWriter=TFRecordWriter, FixedLenFeature=FixedLenFeature, FixedLenSequenceFeature=FixedLenSequenceFeature,
    parse_single_example=parse_single_example, decode_raw=decode_raw, decode_image=decode_image, decode_png=decode_png,
    decode_jpeg=decode_jpeg)
_sys.modules[__name__ + ".io"] = io
