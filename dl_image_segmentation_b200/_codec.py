"""Chip decode driver: host header parse (C ABI, no GPU) -> stream / image descriptors -> GPU decode + assembly.

Stands where the reference calls ``rasterio.MemoryFile(bytes).open().read()`` + ``reshape_as_image``
(``_img_to_tf_mp.py:43-75``) or ``tf.image.decode_png`` (``_img_to_tf_threaded.py:59``).  A chip that
cannot be decoded is reported through a per-image status (the reference prints and skips it,
``_img_to_tf_mp.py:133-136``); nothing here decodes on the CPU.
"""
import ctypes
import threading

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
                ("nodata", ctypes.c_double), ("block_bytes", ctypes.c_uint64),
                ("geotransform", ctypes.c_double * 6), ("has_geo", ctypes.c_int32), ("epsg", ctypes.c_int32),
                ("png_bit_depth", ctypes.c_int32), ("png_color_type", ctypes.c_int32)]


STREAM_DESC_DTYPE = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("dst_len", "<u4"),
                              ("codec", "<i4"), ("image", "<i4")])
IMAGE_DESC_DTYPE = np.dtype([("scratch_off", "<u8"), ("out_off", "<u8"), ("block_bytes", "<u8"), ("format", "<i4"),
                             ("width", "<i4"), ("height", "<i4"), ("samples", "<i4"), ("bytes_per_sample", "<i4"),
                             ("predictor", "<i4"), ("planar", "<i4"), ("big_endian", "<i4"), ("block_w", "<i4"),
                             ("block_h", "<i4"), ("blocks_across", "<i4"), ("blocks_down", "<i4"),
                             ("png_bit_depth", "<i4"), ("png_color_type", "<i4"), ("png_flags", "<i4"), ("png_converted", "<i4")])
PLAN_INPLACE = 0x1000   # b2chips.h B2_PLAN_INPLACE: the blobs already lie in the staging buffer (read there), nothing is gathered
PNG_AS_TF = 1      # b2chips.h B2_PNG_AS_TF: present palette / sub-byte / 16-bit PNGs as tf.image.decode_png does (else: as GDAL does)

_vp, _i, _u64, _u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
_lib.register_signatures({
    "b2_image_probe": (_i, [_vp, _u64, _u32, ctypes.POINTER(ImageInfo)]),
    "b2_image_blocks": (_i, [_vp, _u64, ctypes.POINTER(ImageInfo), _vp, _vp, _vp, _i]),
    "b2_decode_streams": (_i, [_vp, _vp, _vp, _i, _u32, _u32, _vp, _vp, _vp]),
    "b2_assemble_images": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
})

IMAGE_INFO_DTYPE = np.dtype(ImageInfo)

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


def probe(blob, png_as_tf=False) -> ImageInfo:
    """Header-only read: height / width / bands / dtype (load_image_rasterio(decode=False), reference :51-53)."""
    a = _host_bytes(blob)
    if is_jpeg(a):
        st, ji = probe_jpeg(a)
        return jpeg_as_image_info(ji, st)
    info = ImageInfo()
    check(lib().b2_image_probe(a.ctypes.data, a.size, PNG_AS_TF if png_as_tf else 0, ctypes.byref(info)))
    return info


def georef_strings(info):
    """(str(src.get_transform()), str(src.read_crs())) as rasterio would print them (reference _img_to_tf_mp.py:49-50)."""
    gt = [float(v) for v in info.geotransform]
    return str(gt), ("EPSG:%d" % info.epsg) if info.epsg else "None"


def _align(x, a=256):
    return (int(x) + a - 1) // a * a


class DecodePlan(ctypes.Structure):
    _fields_ = [("stage_bytes", ctypes.c_uint64), ("scratch_bytes", ctypes.c_uint64), ("out_bytes", ctypes.c_uint64),
                ("compressed_bytes", ctypes.c_uint64), ("n_streams", ctypes.c_int32), ("codec_mask", ctypes.c_uint32),
                ("max_raw_len", ctypes.c_uint32), ("filled", ctypes.c_int32)]


_lib.register_signatures({
    "b2_decode_plan_batch": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _u64, _i, _u32, ctypes.POINTER(DecodePlan)]),
})


class _HostStaging:
    """Pinned host buffers of one device's decode path, grown on demand and reused between batches."""

    def __init__(self):
        self.stage = torch.empty((1 << 20,), dtype=torch.uint8).pin_memory()
        self.streams = torch.empty((4096 * STREAM_DESC_DTYPE.itemsize,), dtype=torch.uint8).pin_memory()
        self.busy = None          # event: the last H2D copies out of these buffers
        self.pending = False      # planned into, not yet uploaded
        self.pool = None

    def wait(self):
        if self.busy is not None:
            self.busy.synchronize()
            self.busy = None

    def ensure_stage(self, nbytes):
        """The pinned byte buffer, at least nbytes long (contents are NOT kept when it grows)."""
        if self.stage.numel() < nbytes:
            self.stage = torch.empty((int(nbytes * 1.25) + 4096,), dtype=torch.uint8).pin_memory()
            pool = getattr(self, "pool", None)
            if pool is not None:
                pool.stage_hwm = max(pool.stage_hwm, self.stage.numel())
        return self.stage


_staging = {}
_shard_staging = {}     # device index -> [two pinned sets, next]: whole shards on their way to the device (parse_encoded_shard)
_pool_lock = threading.Lock()


def shard_staging(device=None):
    """One of two pinned staging sets kept for whole shards (alternating, so that the next shard can be read while the
    previous one is still being uploaded); waits until its last upload has left it."""
    ctx = get_ctx(device)
    with _pool_lock:
        ent = _shard_staging.setdefault(ctx.device.index, [[_HostStaging(), _HostStaging()], 0])
        hs = ent[0][ent[1] & 1]
        ent[1] += 1
    hs.wait()
    return hs


def fill_pinned(hs, src, threads=8, spare=64):
    """A file (path) or host bytes into hs.stage with a few threads (one core copies ~5 GB/s).  Returns the byte count."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    fd = None
    if isinstance(src, (str, os.PathLike)):
        fd = os.open(src, os.O_RDONLY)
        n = os.fstat(fd).st_size
    else:
        src = _host_bytes(src)
        n = int(src.size)
    try:
        dst = hs.ensure_stage(n + spare).numpy()
        step = max(1 << 22, ((n + threads - 1) // max(1, threads) + 4095) & ~4095)
        spans = [(o, min(n, o + step)) for o in range(0, n, step)]

        def one(span):
            o, e = span
            if fd is None:
                np.copyto(dst[o:e], src[o:e])
            else:
                mv = memoryview(dst)
                while o < e:
                    got = os.preadv(fd, [mv[o:e]], o)
                    if got <= 0:
                        raise OSError("short read")
                    o += got
        if len(spans) <= 1:
            for sp in spans:
                one(sp)
        else:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(one, spans))
    finally:
        if fd is not None:
            os.close(fd)
    return n


def _ptr_of(blob):
    """(address, size, keep-alive object) of a host blob without copying it."""
    if isinstance(blob, bytes):
        return ctypes.cast(ctypes.c_char_p(blob), ctypes.c_void_p).value or 0, len(blob), blob
    a = _host_bytes(blob)
    return a.ctypes.data, a.size, a


class _StagingPool:
    """A few pinned staging sets per device, handed out in rotation: a batch can be planned (host work, any thread) while
    the previous one is still being uploaded and decoded."""

    def __init__(self, n=6):
        self.sets = [_HostStaging() for _ in range(n)]
        self.k = 0
        self.lock = threading.Lock()
        self.stage_hwm = 0         # largest buffer any set needed so far: the others grow to it when they are next taken, so
                                   # that a steady stream of equal batches never re-pins memory (cudaHostAlloc stalls every
                                   # CUDA call of the process while it runs)

    def take(self):
        with self.lock:
            for _ in range(len(self.sets)):                    # the next set in rotation that no plan is holding
                hs = self.sets[self.k % len(self.sets)]
                self.k += 1
                if not hs.pending:
                    break
            else:
                raise B2Error("decode staging: more batches planned ahead than there are staging sets")
            hs.pending = True
        hs.wait()                                              # its last uploads have left the pinned buffers
        hs.pool = self
        if hs.stage.numel() < self.stage_hwm:
            hs.stage = torch.empty((self.stage_hwm,), dtype=torch.uint8).pin_memory()
        return hs


class PlannedBatch:
    """Result of plan_blobs: everything decode_planned needs, all of it host-side.  Holds one pinned staging set until
    decode_planned has queued the uploads; a plan that is dropped instead (an exception between planning and decoding,
    a KeyboardInterrupt in a notebook) gives the set back through release() / the destructor."""
    __slots__ = ("n", "hs", "plan", "infos", "status", "images", "png_as_tf", "consumed")

    def release(self):
        hs = getattr(self, "hs", None)
        if hs is not None and not getattr(self, "consumed", True):
            hs.pending = False
        self.consumed = True

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def take_staging(device=None):
    """A free pinned staging set of this device (held until a plan made with it is decoded or released)."""
    ctx = get_ctx(device)
    with _pool_lock:
        pool = _staging.get(ctx.device.index)
        if pool is None:
            pool = _staging[ctx.device.index] = _StagingPool()
    return pool.take()


def reserve_staging(device, nbytes, sets=6):
    """Grow every free staging set of the device to nbytes now (a worker calls this once, before its pipeline starts:
    pinning memory later would stall every CUDA call of the process for the duration of the cudaHostAlloc).  sets: how
    many sets the caller's pipeline can hold at once (the pool only ever grows)."""
    ctx = get_ctx(device)
    with _pool_lock:
        pool = _staging.get(ctx.device.index)
        if pool is None:
            pool = _staging[ctx.device.index] = _StagingPool()
    with pool.lock:
        while len(pool.sets) < sets:
            pool.sets.append(_HostStaging())
        for hs in pool.sets:
            if not hs.pending and hs.stage.numel() < nbytes:
                hs.wait()
                hs.stage = torch.empty((int(nbytes),), dtype=torch.uint8).pin_memory()
        pool.stage_hwm = max(pool.stage_hwm, int(nbytes))


def plan_blobs(blobs, device=None, png_as_tf=False, inplace=None, threads=0):
    """Host half of decode_blobs: header parse, descriptor tables and the gather of all compressed bytes into one pinned
    buffer, in ONE native multi-threaded call.  Touches no GPU state besides pinned host memory, so the translators run
    it on their read-ahead thread while the GPU works on the previous batch.

    inplace: a staging set (take_staging) whose pinned buffer already holds every blob, 16-byte aligned — the translators'
    FileBatchReader reads the files straight into it — so that nothing is gathered and the buffer is uploaded as it is."""
    ctx = get_ctx(device)
    n = len(blobs)
    pb = PlannedBatch()
    pb.n, pb.png_as_tf = n, png_as_tf
    pb.infos = (ImageInfo * max(n, 1))()
    pb.status = np.zeros(n, dtype=np.int32)
    pb.images = np.zeros(n, dtype=IMAGE_DESC_DTYPE)
    pb.plan = DecodePlan()
    pb.hs = None
    pb.consumed = True
    if n == 0:
        return pb
    hs = pb.hs = inplace if inplace is not None else take_staging(ctx.device)
    pb.consumed = False
    ptrs = (ctypes.c_void_p * n)()
    sizes = np.zeros(n, np.uint64)
    keep = []
    for i, b in enumerate(blobs):
        p, sz, k = _ptr_of(b)
        ptrs[i], sizes[i] = p, sz
        keep.append(k)
    ssz = STREAM_DESC_DTYPE.itemsize
    flags = (PNG_AS_TF if png_as_tf else 0) | (PLAN_INPLACE if inplace is not None else 0)
    for _ in range(2):
        check(lib().b2_decode_plan_batch(ptrs, sizes.ctypes.data, n, pb.infos, pb.status.ctypes.data, pb.images.ctypes.data,
                                         hs.streams.data_ptr(), hs.streams.numel() // ssz, hs.stage.data_ptr(),
                                         hs.stage.numel(), int(threads), flags, ctypes.byref(pb.plan)))
        if pb.plan.filled:
            break
        if hs.stage.numel() < pb.plan.stage_bytes:
            if inplace is not None:
                raise B2Error("plan_blobs(inplace): the staging buffer that holds the files is smaller than the plan")
            hs.ensure_stage(int(pb.plan.stage_bytes))
        if hs.streams.numel() // ssz < pb.plan.n_streams:
            hs.streams = torch.empty((int(pb.plan.n_streams * 1.25 + 64) * ssz,), dtype=torch.uint8).pin_memory()
    del keep
    return pb


_side_streams = {}
_status_ring = {}      # device -> [rotating pinned int32 buffers]: pinning per batch would stall the CUDA calls of other threads


def _pinned_status(dev_index, n):
    ring = _status_ring.setdefault(dev_index, {"k": 0, "bufs": [None] * 8})
    ring["k"] = (ring["k"] + 1) % len(ring["bufs"])
    b = ring["bufs"][ring["k"]]
    if b is None or b.numel() < n:
        b = ring["bufs"][ring["k"]] = torch.empty((max(n, 8192),), dtype=torch.int32).pin_memory()
    return b[:n]


def _side_stream(device):
    st = _side_streams.get(device.index)
    if st is None:
        st = _side_streams[device.index] = torch.cuda.Stream(device)
    return st


class DecodeJob:
    """A batch whose uploads and decode kernels are queued on the stream: device buffers + what the host knows."""
    __slots__ = ("n", "out", "scratch", "status_dev", "images", "infos", "plan", "host_status", "blob_dev", "status_pinned",
                 "status_ready")

    def status(self):
        """Per-image status: 0 = decoded.  Waits for THIS batch's decode only (the status words come back on a side
        stream, so work queued behind the batch on the main stream does not delay them)."""
        if self.status_dev is None:
            return self.host_status
        self.status_ready.synchronize()
        return self.status_pinned.numpy()

    def infos_array(self):
        """The ImageInfo records as a NumPy structured array (no copy)."""
        return np.frombuffer(self.infos, dtype=IMAGE_INFO_DTYPE, count=self.n)


def decode_enqueue(pb, device=None, timings=None, blob_dev=None):
    """Device half of decode_blobs without the wait: upload, decode kernels, assembly — all queued, nothing synchronised.
    blob_dev: a device copy of the planned staging buffer that already exists (an in-place plan over an uploaded shard
    whose blobs the planner did not have to compact): the bytes are not uploaded a second time."""
    ctx = get_ctx(device)
    job = DecodeJob()
    job.n, job.plan, job.infos, job.images = pb.n, pb.plan, pb.infos, pb.images
    job.out = job.scratch = job.status_dev = job.blob_dev = None
    job.host_status = pb.status
    n, plan, images = pb.n, pb.plan, pb.images
    if n == 0 or plan.n_streams == 0:
        pb.release()
        return job
    hs = pb.hs
    ssz = STREAM_DESC_DTYPE.itemsize
    if blob_dev is not None and blob_dev.numel() < plan.stage_bytes:
        raise B2Error("decode_enqueue: blob_dev is shorter than the planned staging buffer")
    blob_d = blob_dev if blob_dev is not None else hs.stage[:plan.stage_bytes].to(ctx.device, non_blocking=True)
    sd_d = hs.streams[:plan.n_streams * ssz].to(ctx.device, non_blocking=True)
    im_d = torch.from_numpy(images.view(np.uint8).reshape(-1)).to(ctx.device, non_blocking=True)
    st_d = torch.from_numpy(pb.status).to(ctx.device, non_blocking=True)
    hs.busy = torch.cuda.Event()
    hs.busy.record(torch.cuda.current_stream(ctx.device))
    pb.release()
    scratch = torch.empty((plan.scratch_bytes,), dtype=torch.uint8, device=ctx.device)
    out = torch.empty((plan.out_bytes,), dtype=torch.uint8, device=ctx.device)
    if timings is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
    check(lib().b2_decode_streams(ctx.handle, ptr(blob_d), ptr(sd_d), plan.n_streams, plan.codec_mask, plan.max_raw_len,
                                  ptr(scratch), ptr(st_d), ctx.stream()))
    if timings is not None:
        ev[1].record()
    check(lib().b2_assemble_images(ctx.handle, ptr(scratch), ptr(im_d), images.ctypes.data, n, ptr(out), ptr(st_d), ctx.stream()))
    if timings is not None:
        ev[2].record()
        torch.cuda.synchronize()
        timings.update(decode_ms=ev[0].elapsed_time(ev[1]), assemble_ms=ev[1].elapsed_time(ev[2]),
                       compressed_bytes=int(plan.compressed_bytes), decoded_bytes=int(plan.out_bytes), streams=int(plan.n_streams))
    job.out, job.scratch, job.status_dev, job.blob_dev = out, scratch, st_d, blob_d
    side = _side_stream(ctx.device)
    decoded = torch.cuda.Event()
    decoded.record(torch.cuda.current_stream(ctx.device))
    job.status_pinned = _pinned_status(ctx.device.index, n)
    with torch.cuda.stream(side):
        side.wait_event(decoded)
        job.status_pinned.copy_(st_d, non_blocking=True)
        job.status_ready = torch.cuda.Event()
        job.status_ready.record(side)
    st_d.record_stream(side)
    return job


def job_arrays(job):
    """(arrays, status) of a queued decode: waits for its status words; arrays[i] is a view of the job's output buffer."""
    n, infos, images = job.n, job.infos, job.images
    arrays = [None] * n
    status = job.status()
    out = job.out
    for i in range(n):
        if status[i] != 0 or out is None:
            continue
        info = infos[i]
        bs = _B2_SIZE[info.dtype]
        nbytes = info.width * info.height * info.samples * bs
        o = int(images[i]["out_off"])
        arrays[i] = out[o:o + nbytes].view(_B2_TO_TORCH[info.dtype]).view(info.height, info.width, info.samples)
    return arrays, status


def decode_planned(pb, device=None, timings=None, want_infos=False):
    """Device half of decode_blobs: upload, decode kernels, assembly.  Returns what decode_blobs returns."""
    if pb.n == 0:
        return ([], np.zeros(0, np.int32), []) if want_infos else ([], np.zeros(0, np.int32))
    arrays, status = job_arrays(decode_enqueue(pb, device, timings))
    return (arrays, status, pb.infos) if want_infos else (arrays, status)


def decode_blobs(blobs, device=None, timings=None, want_infos=False, png_as_tf=False):
    """Decode a batch of encoded chips on the GPU.

    png_as_tf: present palette / 1-2-4-bit / 16-bit PNGs the way tf.image.decode_png(dtype=uint8) does (the threaded
    translator and the rgb parser) instead of the way rasterio / GDAL does (the multiprocessing translator).

    blobs: list of bytes / uint8 arrays (host).  Returns (arrays, status): arrays[i] is an (H,W,bands) CUDA
    tensor of the file's dtype (None when status[i] != 0).  All per-file host work (header parse, descriptor
    tables, gathering the compressed bytes into one pinned buffer) happens in ONE native, multi-threaded call
    (plan_blobs); decode_planned then runs the kernels.
    """
    arrays, status, infos = decode_planned(plan_blobs(blobs, device, png_as_tf), device, timings, True)
    arrays, status, infos = merge_jpeg(blobs, arrays, status, infos, device, candidates=np.nonzero(status)[0])
    return (arrays, status, infos) if want_infos else (arrays, status)


def probe_blobs(blobs, png_as_tf=False):
    """Header-only pass over a batch (one native call): list of ImageInfo, status != 0 for unreadable files."""
    n = len(blobs)
    if n == 0:
        return []
    ptrs = (ctypes.c_void_p * n)()
    sizes = np.zeros(n, np.uint64)
    keep = []
    for i, b in enumerate(blobs):
        p, sz, k = _ptr_of(b)
        ptrs[i], sizes[i] = p, sz
        keep.append(k)
    infos = (ImageInfo * n)()
    status = np.zeros(n, dtype=np.int32)
    images = np.zeros(n, dtype=IMAGE_DESC_DTYPE)
    plan = DecodePlan()
    check(lib().b2_decode_plan_batch(ptrs, sizes.ctypes.data, n, infos, status.ctypes.data, images.ctypes.data, None, 0, None, 0,
                                     1, PNG_AS_TF if png_as_tf else 0, ctypes.byref(plan)))
    return infos


# ---------------------------------------------------------------------------------------------- JPEG (.jpg chips)
class JpegInfo(ctypes.Structure):
    """b2chips.h b2_jpeg_info: what the host marker walk learns about one baseline JPEG."""
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("components", ctypes.c_int32),
                ("ycc", ctypes.c_int32), ("h", ctypes.c_int32 * 3), ("v", ctypes.c_int32 * 3),
                ("tq", ctypes.c_int32 * 3), ("td", ctypes.c_int32 * 3), ("ta", ctypes.c_int32 * 3),
                ("restart_interval", ctypes.c_int32), ("mcus_across", ctypes.c_int32), ("mcus_down", ctypes.c_int32),
                ("scan_off", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("qt", (ctypes.c_uint16 * 64) * 4), ("huff_counts", ((ctypes.c_uint8 * 16) * 4) * 2),
                ("huff_syms", ((ctypes.c_uint8 * 256) * 4) * 2)]


JPEG_JOB_DTYPE = np.dtype([("src_off", "<u8"), ("coef_off", "<u8"), ("plane_off", "<u8"), ("out_off", "<u8"),
                           ("src_len", "<u4"), ("image", "<i4")])
FORMAT_JPEG = 3    # ImageInfo.format of a chip that went through the JPEG path (1 TIFF, 2 PNG come from b2_image_probe)

_lib.register_signatures({
    "b2_jpeg_probe": (_i, [_vp, _u64, ctypes.POINTER(JpegInfo)]),
    "b2_jpeg_sizes": (_i, [ctypes.POINTER(JpegInfo), ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "b2_jpeg_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _u64, _vp, _vp, _vp, _vp]),
})


def is_jpeg(blob) -> bool:
    """SOI marker at the start of the file (how libjpeg / GDAL / tf.image.decode_image recognise the format)."""
    return len(blob) >= 3 and blob[0] == 0xFF and blob[1] == 0xD8 and blob[2] == 0xFF


def probe_jpeg(blob):
    """Header-only read of a JPEG: (status, JpegInfo); status 0 ok, 1 corrupt, 3 out-of-scope flavour."""
    b = _host_bytes(blob)
    info = JpegInfo()
    st = lib().b2_jpeg_probe(b.ctypes.data if b.size else None, b.size, ctypes.byref(info))
    return int(st), info


class JpegPlan(ctypes.Structure):
    _fields_ = [("stage_bytes", _u64), ("coef_count", _u64), ("plane_bytes", _u64), ("out_bytes", _u64),
                ("n_jobs", ctypes.c_int32), ("filled", ctypes.c_int32)]


_lib.register_signatures({
    "b2_jpeg_plan_batch": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _u64, _i, ctypes.POINTER(JpegPlan)]),
})

_jpeg_tls = threading.local()      # per host thread: pinned staging for the files' bytes (a worker thread owns one GPU)


def _jpeg_stage(nbytes):
    buf = getattr(_jpeg_tls, "stage", None)
    if buf is None or buf.numel() < nbytes:
        buf = _jpeg_tls.stage = torch.empty((int(nbytes * 1.25) + 4096,), dtype=torch.uint8).pin_memory()
    return buf


def decode_jpeg_blobs(blobs, device=None, timings=None):
    """Decode a batch of baseline JPEG files on the GPU -> (arrays, status, infos): arrays[i] is an (H,W,1) or (H,W,3)
    uint8 CUDA tensor (RGB), None where status[i] != 0; infos[i] the file's JpegInfo (None if the header was refused).
    Replaces tf.image.decode_jpeg behind ImageCoder.decode_jpeg (reference _img_to_tf_threaded.py:36-38,51-56).  All
    per-file host work — the marker walk, the job table, the gather into pinned staging — is ONE native call."""
    ctx = get_ctx(device)
    n = len(blobs)
    status = np.zeros(n, np.int32)
    arrays = [None] * n
    infos = [None] * n
    if n == 0:
        return arrays, status, infos
    ptrs = (ctypes.c_void_p * n)()
    sizes = np.zeros(n, np.uint64)
    keep = []
    for i, b in enumerate(blobs):
        p, sz, k = _ptr_of(b)
        ptrs[i], sizes[i] = p, sz
        keep.append(k)
    jinfos = (JpegInfo * n)()
    jobs = np.zeros(n, JPEG_JOB_DTYPE)
    plan = JpegPlan()
    stage = _jpeg_stage(int(sizes.sum()) + 16 * n)               # an upper bound of stage_bytes: one call suffices
    check(lib().b2_jpeg_plan_batch(ptrs, sizes.ctypes.data, n, jinfos, status.ctypes.data, jobs.ctypes.data,
                                   stage.data_ptr(), stage.numel(), 0, ctypes.byref(plan)))
    del keep
    m = int(plan.n_jobs)
    if m == 0:
        return arrays, status, infos
    if not plan.filled:
        raise B2Error("b2_jpeg_plan_batch: staging buffer smaller than its own bound")
    blob_d = stage[:plan.stage_bytes].to(ctx.device, non_blocking=True)
    isz = ctypes.sizeof(JpegInfo)
    info_d = torch.from_numpy(np.frombuffer(jinfos, dtype=np.uint8, count=m * isz)).to(ctx.device, non_blocking=True)
    jobs_d = torch.from_numpy(jobs[:m].view(np.uint8).reshape(-1)).to(ctx.device, non_blocking=True)
    coef_d = torch.empty((max(int(plan.coef_count), 1),), dtype=torch.int16, device=ctx.device)
    planes_d = torch.empty((max(int(plan.plane_bytes), 1),), dtype=torch.uint8, device=ctx.device)
    out_d = torch.empty((max(int(plan.out_bytes), 1),), dtype=torch.uint8, device=ctx.device)
    st_d = torch.zeros((n,), dtype=torch.int32, device=ctx.device)
    if timings is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
    check(lib().b2_jpeg_decode(ctx.handle, ptr(blob_d), ptr(info_d), ctypes.addressof(jinfos), ptr(jobs_d), jobs.ctypes.data, m,
                               ptr(coef_d), int(plan.coef_count), ptr(planes_d), ptr(out_d), ptr(st_d), ctx.stream()))
    if timings is not None:
        ev[1].record()
        torch.cuda.synchronize()
        timings.update(decode_ms=ev[0].elapsed_time(ev[1]), compressed_bytes=int(jobs["src_len"][:m].sum()),
                       decoded_bytes=int(sum(jinfos[j].width * jinfos[j].height * jinfos[j].components for j in range(m))), files=m)
    st = st_d.cpu().numpy()                                   # also: the uploads out of the pinned buffers have finished
    for j in range(m):
        i = int(jobs[j]["image"])
        fi = infos[i] = jinfos[j]
        status[i] = st[i]
        if st[i] == 0:
            o = int(jobs[j]["out_off"])
            arrays[i] = out_d[o:o + fi.width * fi.height * fi.components].view(fi.height, fi.width, fi.components)
    return arrays, status, infos


JPEG_DIMS_DTYPE = np.dtype([("height", "<i4"), ("width", "<i4"), ("samples", "<i4")])


class JpegBatch:
    """A batch of .jpg chips planned on the host (plan_jpeg_batch) and, after jpeg_decode_enqueue, queued on the device:
    the translators' pipelined form of decode_jpeg_blobs (nothing synchronised until status())."""
    __slots__ = ("n", "m", "hs", "jinfos", "jobs", "plan", "host_status", "out", "status_dev", "keep", "consumed")

    def release(self):
        hs = getattr(self, "hs", None)
        if hs is not None and not getattr(self, "consumed", True):
            hs.pending = False
        self.consumed = True

    def status(self):
        """Per-file status (0 = decoded); waits for the device."""
        if self.status_dev is None:
            return self.host_status
        return np.where(self.host_status != 0, self.host_status, self.status_dev.cpu().numpy())

    def dims(self):
        """(height, width, samples) per file as a structured array (zeros where the header was refused)."""
        d = np.zeros(self.n, JPEG_DIMS_DTYPE)
        for j in range(self.m):
            i = int(self.jobs[j]["image"])
            fi = self.jinfos[j]
            d[i] = (fi.height, fi.width, fi.components)
        return d

    def out_offsets(self):
        """Byte offset of every file's pixels in self.out (valid where status() == 0)."""
        o = np.zeros(self.n, np.int64)
        o[self.jobs["image"][:self.m]] = self.jobs["out_off"][:self.m].astype(np.int64)
        return o


def plan_jpeg_batch(blobs, device=None, threads=0):
    """Host half of a batched JPEG decode: marker walk, job table and the gather of the entropy-coded data into a pinned
    staging set of the device's pool (held until jpeg_decode_enqueue has queued its upload, or release())."""
    ctx = get_ctx(device)
    n = len(blobs)
    jb = JpegBatch()
    jb.n, jb.m, jb.hs, jb.out, jb.status_dev, jb.keep, jb.consumed = n, 0, None, None, None, None, True
    jb.host_status = np.zeros(n, np.int32)
    jb.jinfos = (JpegInfo * max(n, 1))()
    jb.jobs = np.zeros(n, JPEG_JOB_DTYPE)
    jb.plan = JpegPlan()
    if n == 0:
        return jb
    ptrs = (ctypes.c_void_p * n)()
    sizes = np.zeros(n, np.uint64)
    keep = []
    for i, b in enumerate(blobs):
        p, sz, k = _ptr_of(b)
        ptrs[i], sizes[i] = p, sz
        keep.append(k)
    hs = jb.hs = take_staging(ctx.device)
    jb.consumed = False
    stage = hs.ensure_stage(int(sizes.sum()) + 16 * n + 64)        # an upper bound of stage_bytes: one call suffices
    check(lib().b2_jpeg_plan_batch(ptrs, sizes.ctypes.data, n, jb.jinfos, jb.host_status.ctypes.data, jb.jobs.ctypes.data,
                                   stage.data_ptr(), stage.numel(), int(threads), ctypes.byref(jb.plan)))
    del keep
    jb.m = int(jb.plan.n_jobs)
    if jb.m and not jb.plan.filled:
        jb.release()
        raise B2Error("b2_jpeg_plan_batch: staging buffer smaller than its own bound")
    return jb


def jpeg_decode_enqueue(jb, device=None):
    """Upload and queue the decode kernels of a planned JPEG batch; no synchronisation."""
    ctx = get_ctx(device)
    if jb.m == 0:
        jb.release()
        return jb
    hs, plan, m = jb.hs, jb.plan, jb.m
    blob_d = hs.stage[:plan.stage_bytes].to(ctx.device, non_blocking=True)
    hs.busy = torch.cuda.Event()
    hs.busy.record(torch.cuda.current_stream(ctx.device))
    jb.release()
    isz = ctypes.sizeof(JpegInfo)
    info_d = torch.from_numpy(np.frombuffer(jb.jinfos, dtype=np.uint8, count=m * isz).copy()).to(ctx.device, non_blocking=True)
    jobs_d = torch.from_numpy(jb.jobs[:m].view(np.uint8).reshape(-1).copy()).to(ctx.device, non_blocking=True)
    coef_d = torch.empty((max(int(plan.coef_count), 1),), dtype=torch.int16, device=ctx.device)
    planes_d = torch.empty((max(int(plan.plane_bytes), 1),), dtype=torch.uint8, device=ctx.device)
    jb.out = torch.empty((max(int(plan.out_bytes), 1),), dtype=torch.uint8, device=ctx.device)
    jb.status_dev = torch.zeros((jb.n,), dtype=torch.int32, device=ctx.device)
    check(lib().b2_jpeg_decode(ctx.handle, ptr(blob_d), ptr(info_d), ctypes.addressof(jb.jinfos), ptr(jobs_d), jb.jobs.ctypes.data, m,
                               ptr(coef_d), int(plan.coef_count), ptr(planes_d), ptr(jb.out), ptr(jb.status_dev), ctx.stream()))
    jb.keep = (blob_d, info_d, jobs_d, coef_d, planes_d)
    return jb


def merge_jpeg(blobs, arrays, status, infos, device=None, candidates=None):
    """The TIFF / PNG planner reports a .jpg chip as an unknown format; decode those through the JPEG path and put their
    results in place (arrays / status / infos as decode_planned returns them).  candidates: the indices worth a look
    (the ones the planner refused), so that a folder of PNG / TIFF chips pays nothing per file here."""
    cand = range(len(blobs)) if candidates is None else candidates
    idx = [int(k) for k in cand if is_jpeg(_host_bytes(blobs[k]))]
    if not idx:
        return arrays, status, infos
    ja, js, ji = decode_jpeg_blobs([blobs[k] for k in idx], device)
    for j, k in enumerate(idx):
        arrays[k], status[k] = ja[j], js[j]
        infos[k] = jpeg_as_image_info(ji[j], js[j] if js[j] in (1, 3) else 0)   # 2 = header fine, entropy data corrupt
    return arrays, status, infos


def jpeg_as_image_info(jinfo, status) -> ImageInfo:
    """The fields of ImageInfo the translators read, for a chip that went through the JPEG path."""
    info = ImageInfo()
    info.format, info.status, info.dtype = FORMAT_JPEG, int(status), _lib.B2_U8
    if jinfo is not None:
        info.width, info.height, info.samples = jinfo.width, jinfo.height, jinfo.components
    info.geotransform[:] = (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)     # GDAL's default for a file without georeferencing
    return info


JPEG_ENC_JOB_DTYPE = np.dtype([("src_off", "<u8"), ("coef_off", "<u8"), ("out_off", "<u8"), ("out_cap", "<u4"),
                               ("width", "<i4"), ("height", "<i4"), ("components", "<i4")])
TF_JPEG_DENSITY = (1, 300, 300)    # tf.image.encode_jpeg's defaults: density_unit 'in', x_density = y_density = 300

_lib.register_signatures({
    "b2_jpeg_header": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _u64, ctypes.POINTER(_u64)]),
    "b2_jpeg_encode_sizes": (_i, [_i, _i, _i, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "b2_gather_ranges": (_i, [_vp, _vp, _vp, _vp, _vp, _i, ctypes.c_uint32, _vp, _vp]),
    "b2_jpeg_encode_scan": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _u64, _vp, _vp, _vp]),
})


def jpeg_header(height, width, components, quality=100, density=TF_JPEG_DENSITY) -> bytes:
    """SOI .. SOS of the file libjpeg writes for these settings (host only)."""
    buf = (ctypes.c_uint8 * 1024)()
    ln = _u64()
    check(lib().b2_jpeg_header(height, width, components, quality, density[0], density[1], density[2], buf, 1024, ctypes.byref(ln)))
    return bytes(buf[:ln.value])


def encode_jpeg_device(pixels, src_off, height, width, components, quality=100, density=TF_JPEG_DENSITY, device=None,
                       timings=None):
    """Baseline JPEG files of n uint8 images that already lie in ONE device buffer (image j: (height[j], width[j],
    components[j]) pixels at pixels[src_off[j]:], components 1 or 3), assembled on the device: header + scan + EOI of every
    file back to back (16-byte aligned) in one device buffer.  Returns (files buffer, offsets, sizes).  One small
    device -> host read (the scan lengths) is the only synchronisation."""
    ctx = get_ctx(device)
    n = len(src_off)
    shapes = np.stack([np.asarray(height, np.int64), np.asarray(width, np.int64), np.asarray(components, np.int64)], axis=1)
    uniq, inv = np.unique(shapes, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cc_u, cap_u, hdr_off, hdr_len, front = [], [], [], [], bytearray(b"\xff\xd9" + bytes(14))
    cc, cap = _u64(), _u64()
    for h, w, c in uniq:
        if c not in (1, 3):
            raise B2Error("encode_jpeg: (H,W,1) or (H,W,3) uint8 images only (tf.image.encode_jpeg, format='')")
        check(lib().b2_jpeg_encode_sizes(int(h), int(w), int(c), ctypes.byref(cc), ctypes.byref(cap)))
        cc_u.append(cc.value)
        cap_u.append(cap.value)
        hd = jpeg_header(int(h), int(w), int(c), quality, density)
        hdr_off.append(len(front))
        hdr_len.append(len(hd))
        front += hd + bytes((-len(hd)) % 16)
    cc_u, cap_u, hdr_off, hdr_len = (np.asarray(x, np.int64) for x in (cc_u, cap_u, hdr_off, hdr_len))
    caps = (cap_u[inv] + 15) & ~15
    jobs = np.zeros(n, JPEG_ENC_JOB_DTYPE)
    jobs["src_off"] = np.asarray(src_off, np.uint64)
    jobs["coef_off"] = np.concatenate(([0], np.cumsum(cc_u[inv])[:-1]))
    jobs["out_off"] = np.concatenate(([0], np.cumsum(caps)[:-1]))
    jobs["out_cap"] = cap_u[inv]
    jobs["height"], jobs["width"], jobs["components"] = shapes[:, 0], shapes[:, 1], shapes[:, 2]
    coef, nfront = int(cc_u[inv].sum()), len(front)
    buf = torch.empty((nfront + int(caps.sum()) + 16,), dtype=torch.uint8, device=ctx.device)
    buf[:nfront].copy_(torch.from_numpy(np.frombuffer(bytes(front), np.uint8).copy()), non_blocking=True)
    jobs_d = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).to(ctx.device)
    coef_d = torch.empty((max(coef, 1),), dtype=torch.int16, device=ctx.device)
    len_d = torch.zeros((n,), dtype=torch.int32, device=ctx.device)
    if timings is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
    check(lib().b2_jpeg_encode_scan(ctx.handle, ptr(pixels), ptr(jobs_d), jobs.ctypes.data, n, int(quality), ptr(coef_d), coef,
                                    ctypes.c_void_p(buf.data_ptr() + nfront), ptr(len_d), ctx.stream()))
    if timings is not None:
        ev[1].record()
        torch.cuda.synchronize()
        timings.update(encode_ms=ev[0].elapsed_time(ev[1]))
    lens = len_d.cpu().numpy().view(np.uint32).astype(np.int64)
    if (lens == 0xFFFFFFFF).any():
        raise B2Error("b2_jpeg_encode_scan: scan buffer too small (its own bound)")
    # header, scan (far shorter than its bound) and EOI of every file packed by ONE gather launch
    sizes = hdr_len[inv] + lens + 2
    offs = np.concatenate(([0], np.cumsum((sizes + 15) & ~15)[:-1]))
    g_src = np.stack([hdr_off[inv], nfront + jobs["out_off"].astype(np.int64), np.zeros(n, np.int64)], axis=1).reshape(-1)
    g_len = np.stack([hdr_len[inv], lens, np.full(n, 2, np.int64)], axis=1)
    g_dst = (offs[:, None] + np.concatenate((np.zeros((n, 1), np.int64), np.cumsum(g_len, axis=1)[:, :2]), axis=1)).reshape(-1)
    files = torch.empty((int(offs[-1] + ((sizes[-1] + 15) & ~15)) + 16,), dtype=torch.uint8, device=ctx.device)
    so_d = torch.from_numpy(g_src.astype(np.uint64)).to(ctx.device)
    do_d = torch.from_numpy(g_dst.astype(np.uint64)).to(ctx.device)
    ln_d = torch.from_numpy(g_len.reshape(-1).astype(np.uint32)).to(ctx.device)
    check(lib().b2_gather_ranges(ctx.handle, ptr(buf), ptr(so_d), ptr(do_d), ptr(ln_d), 3 * n, int(g_len.max()), ptr(files), ctx.stream()))
    files.record_stream(torch.cuda.current_stream(ctx.device))
    return files, offs.astype(np.uint64), sizes.astype(np.uint64)


def encode_jpeg_arrays(arrays, quality=100, density=TF_JPEG_DENSITY, device=None, timings=None):
    """Encode a batch of (H,W,1) / (H,W,3) uint8 images (CUDA tensors or host arrays) as baseline JPEG files on the GPU
    -> list of bytes.  Replaces tf.image.encode_jpeg(image, format='', quality=100) behind ImageCoder.png_to_jpeg
    (reference _img_to_tf_threaded.py:36-38): grey -> one component, RGB -> YCbCr 4:2:0, standard Huffman tables."""
    ctx = get_ctx(device)
    n = len(arrays)
    if n == 0:
        return []
    flat, src_off, hs, ws, cs, src = [], [], [], [], [], 0
    for a in arrays:
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] not in (1, 3):
            raise B2Error("encode_jpeg_arrays: (H,W,1) or (H,W,3) uint8 images only (tf.image.encode_jpeg, format='')")
        h, w, c = (int(x) for x in t.shape)
        flat.append(t.to(ctx.device, non_blocking=True).contiguous().reshape(-1))
        src_off.append(src)
        hs.append(h)
        ws.append(w)
        cs.append(c)
        src += t.numel()
    pixels = torch.cat(flat)
    files, offs, sizes = encode_jpeg_device(pixels, src_off, hs, ws, cs, quality, density, ctx.device, timings)
    if timings is not None:
        timings.update(pixel_bytes=int(src))
    host = files.cpu().numpy()
    return [host[int(o):int(o + z)].tobytes() for o, z in zip(offs, sizes)]


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
