"""Chip decode driver: host header parse (C ABI, no GPU) -> stream / image descriptors -> GPU decode + assembly.

Stands where the reference calls ``rasterio.MemoryFile(bytes).open().read()`` + ``reshape_as_image``
(``_img_to_tf_mp.py:43-75``) or ``tf.image.decode_png`` (``_img_to_tf_threaded.py:59``).  A chip that
cannot be decoded is reported through a per-image status (the reference prints and skips it,
``_img_to_tf_mp.py:133-136``); nothing here decodes on the CPU.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import B2Error, check, get_ctx, lib, ptr


class ImageInfo(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int32), ("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("samples", ctypes.c_int32), ("dtype", ctypes.c_int32), ("compression", ctypes.c_int32),
                ("predictor", ctypes.c_int32), ("planar", ctypes.c_int32), ("big_endian", ctypes.c_int32),
                ("tiled", ctypes.c_int32), ("block_w", ctypes.c_int32), ("block_h", ctypes.c_int32),
                ("blocks_across", ctypes.c_int32), ("blocks_down", ctypes.c_int32), ("n_blocks", ctypes.c_int32),
                ("status", ctypes.c_int32), ("has_nodata", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("nodata", ctypes.c_double), ("block_bytes", ctypes.c_uint64)]


STREAM_DESC_DTYPE = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("dst_len", "<u4"),
                              ("codec", "<i4"), ("image", "<i4")])
IMAGE_DESC_DTYPE = np.dtype([("scratch_off", "<u8"), ("out_off", "<u8"), ("block_bytes", "<u8"), ("format", "<i4"),
                             ("width", "<i4"), ("height", "<i4"), ("samples", "<i4"), ("bytes_per_sample", "<i4"),
                             ("predictor", "<i4"), ("planar", "<i4"), ("big_endian", "<i4"), ("block_w", "<i4"),
                             ("block_h", "<i4"), ("blocks_across", "<i4"), ("blocks_down", "<i4")])

_vp, _i, _u64, _u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
_lib.register_signatures({
    "b2_image_probe": (_i, [_vp, _u64, ctypes.POINTER(ImageInfo)]),
    "b2_image_blocks": (_i, [_vp, _u64, ctypes.POINTER(ImageInfo), _vp, _vp, _vp, _i]),
    "b2_decode_streams": (_i, [_vp, _vp, _vp, _i, _u32, _u32, _vp, _vp, _vp]),
    "b2_assemble_images": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
})

_B2_TO_TORCH = {_lib.B2_U8: torch.uint8, _lib.B2_I8: torch.int8, _lib.B2_U16: torch.uint16, _lib.B2_I16: torch.int16,
                _lib.B2_U32: torch.uint32, _lib.B2_I32: torch.int32, _lib.B2_F32: torch.float32, _lib.B2_F64: torch.float64}
_B2_SIZE = {_lib.B2_U8: 1, _lib.B2_I8: 1, _lib.B2_U16: 2, _lib.B2_I16: 2, _lib.B2_U32: 4, _lib.B2_I32: 4, _lib.B2_F32: 4,
            _lib.B2_F64: 8}


def _host_bytes(blob):
    if isinstance(blob, torch.Tensor):
        blob = blob.detach().cpu().numpy()
    if isinstance(blob, np.ndarray):
        return np.ascontiguousarray(blob).view(np.uint8).reshape(-1)
    return np.frombuffer(blob, dtype=np.uint8)


def probe(blob) -> ImageInfo:
    """Header-only read: height / width / bands / dtype (load_image_rasterio(decode=False), reference :51-53)."""
    a = _host_bytes(blob)
    info = ImageInfo()
    check(lib().b2_image_probe(a.ctypes.data, a.size, ctypes.byref(info)))
    return info


def _align(x, a=256):
    return (int(x) + a - 1) // a * a


def decode_blobs(blobs, device=None, timings=None):
    """Decode a batch of encoded chips on the GPU.

    blobs: list of bytes / uint8 arrays (host).  Returns (arrays, status): arrays[i] is an (H,W,bands) CUDA
    tensor of the file's dtype (None when status[i] != 0).
    """
    ctx = get_ctx(device)
    n = len(blobs)
    hosts = [_host_bytes(b) for b in blobs]
    infos, status = [], np.zeros(n, dtype=np.int32)
    stage_parts, stage_pos = [], 0
    streams = []
    images = np.zeros(n, dtype=IMAGE_DESC_DTYPE)
    scratch_pos = out_pos = 0
    mask = 0
    max_raw = 0
    for i, a in enumerate(hosts):
        info = ImageInfo()
        check(lib().b2_image_probe(a.ctypes.data, a.size, ctypes.byref(info)))
        infos.append(info)
        if info.status != 0:
            status[i] = info.status
            continue
        nb = info.n_blocks
        offs = np.zeros(nb, np.uint64)
        cnts = np.zeros(nb, np.uint64)
        dlen = np.zeros(nb, np.uint64)
        try:
            check(lib().b2_image_blocks(a.ctypes.data, a.size, ctypes.byref(info), offs.ctypes.data, cnts.ctypes.data,
                                        dlen.ctypes.data, nb))
        except B2Error:
            status[i] = 2
            continue
        bs = _B2_SIZE[info.dtype]
        im = images[i]
        im["scratch_off"], im["out_off"], im["block_bytes"] = scratch_pos, out_pos, info.block_bytes
        im["format"], im["width"], im["height"], im["samples"] = info.format, info.width, info.height, info.samples
        im["bytes_per_sample"], im["predictor"], im["planar"], im["big_endian"] = bs, info.predictor, info.planar, info.big_endian
        im["block_w"], im["block_h"] = info.block_w, info.block_h
        im["blocks_across"], im["blocks_down"] = info.blocks_across, info.blocks_down
        if info.format == 2:                                   # PNG: concatenate the IDAT payloads into one zlib stream
            total = 0
            for o, c in zip(offs, cnts):
                stage_parts.append(a[int(o):int(o + c)])
                total += int(c)
            streams.append((stage_pos, scratch_pos, total, int(info.block_bytes), 8, i))
            stage_pos += total
            scratch_pos += _align(info.block_bytes)
            mask |= 2
        else:                                                  # TIFF: the file as is, one stream per tile / strip
            stage_parts.append(a)
            codec = {1: 1, 5: 5, 8: 8, 32946: 8}[info.compression]
            mask |= {1: 4, 5: 1, 8: 2}[codec]
            for k in range(nb):
                streams.append((stage_pos + int(offs[k]), scratch_pos + k * int(info.block_bytes), int(cnts[k]), int(dlen[k]), codec, i))
                if codec == 1:
                    max_raw = max(max_raw, int(dlen[k]))
            stage_pos += a.size
            scratch_pos += _align(nb * int(info.block_bytes))
        out_pos += _align(info.width * info.height * info.samples * bs)
        pad = (-stage_pos) % 16
        if pad:
            stage_parts.append(np.zeros(pad, np.uint8))
            stage_pos += pad
    arrays = [None] * n
    if not streams:
        return arrays, status
    sd = np.array(streams, dtype=STREAM_DESC_DTYPE)
    blob_h = torch.from_numpy(np.concatenate(stage_parts)) if len(stage_parts) > 1 else torch.from_numpy(np.array(stage_parts[0]))
    blob_d = blob_h.to(ctx.device, non_blocking=True)
    sd_d = torch.from_numpy(sd.view(np.uint8).reshape(-1)).to(ctx.device, non_blocking=True)
    im_d = torch.from_numpy(images.view(np.uint8).reshape(-1).copy()).to(ctx.device, non_blocking=True)
    scratch = torch.empty((max(scratch_pos, 16),), dtype=torch.uint8, device=ctx.device)
    out = torch.empty((max(out_pos, 16),), dtype=torch.uint8, device=ctx.device)
    st_d = torch.from_numpy(status.copy()).to(ctx.device)
    if timings is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
    check(lib().b2_decode_streams(ctx.handle, ptr(blob_d), ptr(sd_d), len(sd), mask, max_raw, ptr(scratch), ptr(st_d), ctx.stream()))
    if timings is not None:
        ev[1].record()
    check(lib().b2_assemble_images(ctx.handle, ptr(scratch), ptr(im_d), images.ctypes.data, n, ptr(out), ptr(st_d), ctx.stream()))
    if timings is not None:
        ev[2].record()
        torch.cuda.synchronize()
        timings.update(decode_ms=ev[0].elapsed_time(ev[1]), assemble_ms=ev[1].elapsed_time(ev[2]),
                       compressed_bytes=int(sum(int(s[2]) for s in streams)), decoded_bytes=int(out_pos), streams=len(streams))
    status = st_d.cpu().numpy()
    for i, info in enumerate(infos):
        if status[i] != 0:
            continue
        bs = _B2_SIZE[info.dtype]
        nbytes = info.width * info.height * info.samples * bs
        o = int(images[i]["out_off"])
        arrays[i] = out[o:o + nbytes].view(_B2_TO_TORCH[info.dtype]).view(info.height, info.width, info.samples)
    return arrays, status


def to_float32(t):
    """.astype(np.float32) of a decoded chip (reference :328-329) via the cast kernel (mean 0, std 1)."""
    from . import ops
    if t.dtype == torch.float32:
        return t
    if t.dtype in (torch.uint8, torch.uint16, torch.int16):
        C = t.shape[-1]
        out, _ = ops.normalise_onehot(t.contiguous(), None, np.zeros(C, np.float32), np.ones(C, np.float32), 1)
        return out
    return t.to(torch.float32)
