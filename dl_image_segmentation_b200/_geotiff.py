"""GeoTIFF chip writer on the GPU — the save step of ``create_chips_for_tile`` (reference
``_descartes_img_chips.py:781-797``: GTiff, ``COMPRESS=LZW``, ``TILED=TRUE``, bands written one by one, nodata on the
label band ``:794-795``) with the georeferencing of ``_gdal_dataset_from_geocontext`` (``:804-849``).

The pixels never leave the device uncompressed: ``b2_tile_split`` cuts the (H,W,bands) raster into zero-padded
256x256 pixel-interleaved tiles, ``b2_lzw_encode`` compresses every tile of a whole batch of chips in one launch,
and only the code streams come back to the host, where the IFD (baseline tags, GeoTIFF keys, GDAL_NODATA) is put
around them.  Files are classic little-endian TIFF, PlanarConfiguration 1, Predictor 1 — what GDAL writes for these
options — and read back bit-exactly through libtiff (tests) and through the K1 decoder.
"""
import ctypes
import os
import struct

import numpy as np
import torch

from . import _lib
from ._lib import B2Error, check, get_ctx, lib, ptr

ENC_DESC_DTYPE = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("dst_cap", "<u4")])
_vp, _i = ctypes.c_void_p, ctypes.c_int
_lib.register_signatures({
    "b2_lzw_encode": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "b2_lzw_encode_restart": (_i, [_vp, _vp, _vp, _i, ctypes.c_uint32, _vp, _vp, _vp]),
    "b2_tile_split": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "b2_gather_ranges": (_i, [_vp, _vp, _vp, _vp, _vp, _i, ctypes.c_uint32, _vp, _vp]),
})

# A Clear code every LZW_RESTART input bytes (0: only when the 4094-entry table is full, as libtiff does).  Every TIFF reader
# accepts either; with the restarts a tile is hundreds of independent pieces for the encoder instead of one serial walk.
LZW_RESTART = 1024

_pinned = {}      # device index -> pinned host buffer the packed code streams land in (grown on demand)

_SAMPLE_FORMAT = {"u": 1, "i": 2, "f": 3}
_NP_OF_TORCH = {torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32,
                torch.float32: np.float32, torch.float64: np.float64}
for _n, _t in (("uint16", np.uint16), ("uint32", np.uint32)):
    if hasattr(torch, _n):
        _NP_OF_TORCH[getattr(torch, _n)] = _t


def _encode_compact(raw, lengths, device=None, restart=None):
    """TIFF-LZW encode consecutive byte ranges of one device buffer (back to back, each padded to 16 bytes).
    Returns (host uint8 array holding all code streams back to back, numpy array of their lengths)."""
    ctx = get_ctx(device)
    restart = LZW_RESTART if restart is None else int(restart)
    n = len(lengths)
    ln = np.asarray(lengths, dtype=np.int64)
    cap = (ln * 3) // 2 + 64
    descs = np.zeros(n, ENC_DESC_DTYPE)
    descs["src_len"], descs["dst_cap"] = ln, cap
    descs["src_off"] = np.concatenate(([0], np.cumsum((ln + 15) & ~15)[:-1]))
    dcap = (cap + 15) & ~15
    descs["dst_off"] = np.concatenate(([0], np.cumsum(dcap)[:-1]))
    out = torch.empty((max(int(dcap.sum()), 16),), dtype=torch.uint8, device=ctx.device)
    out_len = torch.empty((n,), dtype=torch.int32, device=ctx.device)
    if restart:
        check(lib().b2_lzw_encode_restart(ctx.handle, ptr(raw), descs.ctypes.data, n, restart, ptr(out), ptr(out_len), ctx.stream()))
    else:
        d_dev = torch.from_numpy(descs.view(np.uint8).reshape(-1)).to(ctx.device)
        check(lib().b2_lzw_encode(ctx.handle, ptr(raw), ptr(d_dev), n, ptr(out), ptr(out_len), ctx.stream()))
    lens = out_len.cpu().numpy().view(np.uint32).astype(np.int64)          # small read-back; waits for the encoder
    if (lens == 0xFFFFFFFF).any():
        raise B2Error("b2_lzw_encode: output capacity exceeded")
    # compact the code streams on the device (the capacity slots are 1.5x the raw size) with one gather launch, then ONE
    # copy into pinned host memory
    total = int(lens.sum())
    dst_off = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.uint64)
    packed = torch.empty((max(total, 16),), dtype=torch.uint8, device=ctx.device)
    so_d = torch.from_numpy(np.ascontiguousarray(descs["dst_off"])).to(ctx.device)
    do_d = torch.from_numpy(dst_off).to(ctx.device)
    check(lib().b2_gather_ranges(ctx.handle, ptr(out), ptr(so_d), ptr(do_d), ptr(out_len), n, int(lens.max()) if n else 0, ptr(packed),
                                 ctx.stream()))
    host = _pinned.get(ctx.device.index)
    if host is None or host.numel() < total:
        host = _pinned[ctx.device.index] = torch.empty((int(total * 1.25) + 4096,), dtype=torch.uint8).pin_memory()
    host[:total].copy_(packed[:total], non_blocking=True)
    torch.cuda.current_stream(ctx.device).synchronize()
    return host[:total].numpy(), lens


def lzw_encode_tiles(raw, lengths, device=None, restart=None):
    """As _encode_compact, returning one bytes object per stream."""
    host, lens = _encode_compact(raw, lengths, device, restart)
    ends = np.cumsum(lens)
    return [host[int(e - l):int(e)].tobytes() for e, l in zip(ends, lens)]


def _ifd(width, height, bands, np_dtype, tile, block_lens, block_data, nodata, geotransform, epsg):
    """Classic little-endian TIFF around the compressed tiles: header, one IFD (tags ascending), out-of-line values,
    tile data (block_data = the code streams of this file back to back, a buffer)."""
    blocks = [None] * len(block_lens)
    dt = np.dtype(np_dtype)
    ent = {}

    def put(tag, typ, fmt, vals):
        ent[tag] = (typ, len(vals), struct.pack("<%d%s" % (len(vals), fmt), *vals))
    put(256, 3, "H", [width])
    put(257, 3, "H", [height])
    put(258, 3, "H", [dt.itemsize * 8] * bands)
    put(259, 3, "H", [5])                                                  # LZW
    rgb = bands >= 3 and dt == np.uint8                                     # GDAL: RGB only for >= 3 Byte bands
    put(262, 3, "H", [2 if rgb else 1])
    put(277, 3, "H", [bands])
    put(284, 3, "H", [1])                                                  # pixel-interleaved
    extra = bands - (3 if rgb else 1)
    if extra > 0:
        put(338, 3, "H", [0] * extra)
    put(339, 3, "H", [_SAMPLE_FORMAT[dt.kind]] * bands)
    put(322, 3, "H", [tile])
    put(323, 3, "H", [tile])
    if geotransform is not None:
        x0, dx, _, y0, _, dy = [float(v) for v in geotransform]
        put(33550, 12, "d", [dx, -dy, 0.0])
        put(33922, 12, "d", [0.0, 0.0, 0.0, x0, y0, 0.0])
        if epsg:
            put(34735, 3, "H", [1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, int(epsg)])
    if nodata is not None:
        s = str(nodata).encode() + b"\0"
        ent[42113] = (2, len(s), s)
    put(324, 4, "I", [0] * len(blocks))
    put(325, 4, "I", [int(l) for l in block_lens])
    tags = sorted(ent)
    ifd_off = 8
    extra_off = ifd_off + 2 + 12 * len(tags) + 4
    tail = bytearray()
    where = {}
    for t in tags:
        typ, cnt, val = ent[t]
        if len(val) > 4:
            if len(tail) % 2:
                tail += b"\0"
            where[t] = len(tail)
            tail += val
    data_off = extra_off + len(tail)
    offs, cur = [], data_off
    for l in block_lens:
        offs.append(cur)
        cur += int(l)
    ob = struct.pack("<%dI" % len(offs), *offs)
    if 324 in where:
        tail[where[324]:where[324] + len(ob)] = ob
    out = bytearray(b"II" + struct.pack("<HI", 42, ifd_off) + struct.pack("<H", len(tags)))
    for t in tags:
        typ, cnt, val = ent[t]
        if t == 324 and 324 not in where:
            val = ob
        field = struct.pack("<I", extra_off + where[t]) if t in where else val.ljust(4, b"\0")
        out += struct.pack("<HHI", t, typ, cnt) + field
    out += struct.pack("<I", 0) + tail
    if block_data is None:
        return bytes(out)                                                   # the caller writes the tile data after it
    return b"".join((bytes(out), block_data))


def encode_geotiffs(arrays, nodata=None, geotransform=(499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0), epsg=32643, tile=256,
                    device=None, as_parts=False, lzw_restart=None):
    """A batch of (H,W,bands) / (H,W) rasters (CUDA tensors or numpy arrays, any TIFF sample type) -> list of GeoTIFF
    file bytes.  nodata: one value for all, or a list (None entries = no GDAL_NODATA tag).  All tiles of the batch are
    compressed in ONE b2_lzw_encode launch.  as_parts: return (header bytes, tile data view) per file instead of joined
    bytes (the views point into a pinned buffer that the next call reuses): write_geotiffs writes them without a copy."""
    ctx = get_ctx(device)
    n = len(arrays)
    nod = list(nodata) if isinstance(nodata, (list, tuple)) else [nodata] * n
    metas, lengths, parts = [], [], []
    for a in arrays:
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(ctx.device).contiguous()
        if t.dim() == 2:
            t = t[:, :, None]
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        if t.dtype not in _NP_OF_TORCH:
            raise B2Error("encode_geotiffs: unsupported dtype %s" % t.dtype)
        H, W, B = (int(x) for x in t.shape)
        pb = B * t.element_size()
        across, down = (W + tile - 1) // tile, (H + tile - 1) // tile
        tb = tile * tile * pb
        tiles = torch.empty((across * down * tb,), dtype=torch.uint8, device=ctx.device)
        check(lib().b2_tile_split(ctx.handle, ptr(t), H, W, pb, tile, tile, ptr(tiles), ctx.stream()))
        parts.append(tiles)
        lengths += [tb] * (across * down)
        metas.append((W, H, B, _NP_OF_TORCH[t.dtype], across * down))
    raw = torch.cat(parts) if len(parts) > 1 else parts[0]                  # tile sizes are multiples of 16
    host, lens = _encode_compact(raw, lengths, ctx.device, lzw_restart)  # None: LZW_RESTART
    mv = memoryview(host)
    files, k, pos = [], 0, 0
    for (W, H, B, dt, nb), nd in zip(metas, nod):
        size = int(lens[k:k + nb].sum())
        if as_parts:
            files.append((_ifd(W, H, B, dt, tile, lens[k:k + nb], None, nd, geotransform, epsg), mv[pos:pos + size]))
        else:
            files.append(_ifd(W, H, B, dt, tile, lens[k:k + nb], mv[pos:pos + size], nd, geotransform, epsg))
        k += nb
        pos += size
    return files


def write_geotiffs(arrays, paths, nodata=None, geotransform=(499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0), epsg=32643, tile=256,
                   device=None, io_threads=16, lzw_restart=None):
    """Encode a batch of rasters and write each to its path (the save step of create_chips_for_tile for many tiles at
    once): one encode launch, one packed device -> pinned host copy, then header + tile data of every file written by a
    pool of threads straight from the pinned buffer.  Returns the paths."""
    from concurrent.futures import ThreadPoolExecutor
    parts = encode_geotiffs(arrays, nodata=nodata, geotransform=geotransform, epsg=epsg, tile=tile, device=device, as_parts=True,
                            lzw_restart=lzw_restart)

    def write(job):
        path, (head, data) = job
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        try:
            os.write(fd, head)
            os.write(fd, data)
        finally:
            os.close(fd)
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
        list(pool.map(write, zip(paths, parts)))
    return list(paths)


def write_chip_pair(img_arr, lbl_arr, out_base, dltile_key, label_ndv=None, geotransform=(499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0),
                    epsg=32643, device=None):
    """The tail of create_chips_for_tile (reference :735-800): images/<key>.tif and labels/<key>.tif under out_base,
    ':' replaced by '#'; the label gets the nodata value.  Returns (img_file, lbl_file)."""
    out_img_folder, out_lbl_folder = os.path.join(out_base, "images"), os.path.join(out_base, "labels")
    os.makedirs(out_img_folder, exist_ok=True)
    os.makedirs(out_lbl_folder, exist_ok=True)
    fn = dltile_key.replace(":", "#")
    img_file, lbl_file = os.path.join(out_img_folder, fn) + ".tif", os.path.join(out_lbl_folder, fn) + ".tif"
    a, b = encode_geotiffs([img_arr, lbl_arr], nodata=[None, label_ndv], geotransform=geotransform, epsg=epsg, device=device)
    with open(img_file, "wb") as f:
        f.write(a)
    with open(lbl_file, "wb") as f:
        f.write(b)
    return img_file, lbl_file
