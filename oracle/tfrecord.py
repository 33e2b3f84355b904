"""TFRecord framing + CRC-32C (oracle; test infrastructure only).

Stands in for ``tf.io.TFRecordWriter(path).write(bytes)`` / ``.close()``
(reference ``_img_to_tf_mp.py:119,141,150``; ``_img_to_tf_threaded.py:182,203,212``) and for the
reader behind ``tf.data.TFRecordDataset`` (``parse_tfrecords.ipynb`` cell 4).  TensorFlow is not
installable here; the format is restated from its public description:

    uint64le length | uint32le masked_crc32c(length bytes) | data | uint32le masked_crc32c(data)
    masked(c) = ((c >> 15) | (c << 17)) + 0xa282ead8  (mod 2**32);  CRC-32C per RFC 3720 B.4.
"""
import ctypes
import struct

import numpy as np

from . import clib

_POLY = 0x82F63B78
_TAB = None


def _table():
    global _TAB
    if _TAB is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ _POLY if c & 1 else c >> 1
            t.append(c)
        _TAB = t
    return _TAB


def crc32c_py(data: bytes) -> int:
    """Bit-definition CRC-32C in pure Python (small inputs; pins the C versions)."""
    t = _table()
    c = 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ t[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def _ptr(buf):
    a = np.frombuffer(buf, dtype=np.uint8)
    return a, a.ctypes.data


def crc32c(data) -> int:
    a, p = _ptr(data)
    return int(clib().orc_crc32c(p, a.size))


def mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def masked_crc32c(data) -> int:
    return mask(crc32c(data))


def frame(data: bytes) -> bytes:
    """One framed record, as RecordWriter::WriteRecord emits it (uncompressed)."""
    hdr = struct.pack("<Q", len(data))
    return hdr + struct.pack("<I", masked_crc32c(hdr)) + bytes(data) + struct.pack("<I", masked_crc32c(data))


class DataLossError(Exception):
    """TF raises tf.errors.DataLossError('corrupted record at <offset>') on a CRC mismatch."""


def scan(buf, verify: bool = True):
    """Sequential RecordReader walk. Returns (data_offsets, lengths) as uint64 arrays."""
    a, p = _ptr(buf)
    cap = max(1, a.size // 16 + 1)
    offs = np.zeros(cap, dtype=np.uint64)
    lens = np.zeros(cap, dtype=np.uint64)
    n = clib().orc_tfrecord_scan(p, a.size, offs.ctypes.data, lens.ctypes.data, cap, int(verify))
    if n < 0:
        raise DataLossError("corrupted record #%d" % (-n - 1))
    return offs[:n].copy(), lens[:n].copy()


def read_records(buf, verify: bool = True):
    offs, lens = scan(buf, verify)
    mv = memoryview(buf)
    return [bytes(mv[int(o):int(o) + int(l)]) for o, l in zip(offs, lens)]


class TFRecordWriter:
    """Minimal stand-in for tf.io.TFRecordWriter(path) with default (uncompressed) options."""

    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, record: bytes):
        self._f.write(frame(record))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
