"""Shared worker loop of the two translators: chip folders -> sharded TFRecords, one GPU per worker.

This is the body of ``_process_image_files_mp_worker`` (``_img_to_tf_mp.py:78-157``) and
``_process_image_files_worker`` (``_img_to_tf_threaded.py:136-219``) with the per-chip native calls batched on
the GPU: file bytes -> [K1 decode] -> [K2 build: protobuf + framing + CRC-32C] -> shard bytes -> file.
Kept from the reference: shard sub-ranges ``np.linspace(lo, hi, S/P + 1).astype(int)`` (``:106-108``), shard name
``'%s-%.5d-of-%.5d'`` (``:115``), records in list order, skip-and-continue on any per-chip failure with the same
messages (``:133-136``), key equality check (``:132``), progress / summary prints (``:145-157``).
"""
import ctypes
import mmap
import os
import sys
from datetime import datetime

import numpy as np
import torch

from . import _codec, ops
from . import _lib as _lib_mod
from ._lib import B2Error, check, get_ctx, lib

_MAPPED_WRITE_MIN = 32 << 20     # pieces this large are written pwrite-head + mapped-rest (see run_worker.write_back)


def tile_key_from_path(path, parse_dltile_filename=True):
    base = os.path.basename(path)
    if parse_dltile_filename:
        return ".".join(base.split(os.extsep)[:-1]).replace("#", ":")       # _img_to_tf_mp.py:61
    return base


def worker_ranges(n_files, num_workers):
    spacing = np.linspace(0, n_files, num_workers + 1).astype(int)          # np.int in the reference (:167)
    return [[int(spacing[i]), int(spacing[i + 1])] for i in range(len(spacing) - 1)]


def _read(path):
    with open(path, "rb") as f:
        return f.read()


class ChipError(Exception):
    pass


_vp, _i, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
_lib_mod.register_signatures({
    "b2_read_files": (_i, [_vp, _i, _vp, _u64, _vp, _vp, _vp, _i, _vp]),
})


_worker_buffers = {}     # device index -> host buffers of run_worker (one worker at a time per device)


class FileBatchReader:
    """Reads whole batches of files with ONE native, multi-threaded call (b2_read_files) into a reusable host buffer:
    the per-file open().read() of the reference's worker loop costs the interpreter ~30 us a file, which is what bounds
    the translators once decode and serialisation run on the GPU.  read() returns one uint8 array view per file (or the
    OSError the reference's except branch would have caught); a view stays valid until `depth` more batches were read."""

    def __init__(self, depth=3, threads=0):
        import threading
        self.bufs = [np.empty(0, np.uint8) for _ in range(depth)]
        self.k = 0
        self.threads = int(threads)
        self.lock = threading.Lock()       # two read-ahead threads share one reader: each call gets a buffer of its own

    def read(self, paths):
        n = len(paths)
        if n == 0:
            return []
        enc = [os.fsencode(p) for p in paths]
        arr = (ctypes.c_char_p * n)(*enc)
        offs, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        status = np.zeros(n, np.int32)
        need = ctypes.c_uint64(0)
        with self.lock:
            slot = self.k % len(self.bufs)
            self.k += 1
        buf = self.bufs[slot]
        for _ in range(2):
            check(lib().b2_read_files(arr, n, buf.ctypes.data if buf.size else None, buf.size, offs.ctypes.data, sizes.ctypes.data,
                                      status.ctypes.data, self.threads, ctypes.byref(need)))
            if buf.size >= need.value and buf.size:
                break
            buf = self.bufs[slot] = np.empty(int(need.value * 1.25) + 4096, np.uint8)
        out = []
        for i in range(n):
            if status[i]:
                out.append(OSError(int(status[i]), os.strerror(int(status[i])), paths[i]))
            else:
                out.append(buf[int(offs[i]):int(offs[i]) + int(sizes[i])])
        return out


def _read_into(self, paths, hs):
    """Read the files straight into the pinned buffer of staging set hs (16-byte aligned each, room for the decoders'
    look-ahead after the last one).  Returns (views, offsets, sizes, clean): one uint8 view per file, where it lies in the
    buffer, and whether every file was read."""
    n = len(paths)
    enc = [os.fsencode(p) for p in paths]
    arr = (ctypes.c_char_p * n)(*enc)
    offs, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    status = np.zeros(n, np.int32)
    need = ctypes.c_uint64(0)
    for _ in range(2):
        stage = hs.stage
        check(lib().b2_read_files(arr, n, stage.data_ptr(), max(0, stage.numel() - 64), offs.ctypes.data, sizes.ctypes.data,
                                  status.ctypes.data, self.threads, ctypes.byref(need)))
        if stage.numel() - 64 >= need.value:
            break
        hs.wait()
        hs.ensure_stage(int(need.value) + 64)
    buf = hs.stage.numpy()
    views = [buf[int(offs[i]):int(offs[i]) + int(sizes[i])] for i in range(n)]
    return views, offs, sizes, not status.any() and bool(sizes.all())


FileBatchReader.read_into = _read_into


def _read_or_error(path):
    try:
        return _read(path)
    except Exception as e:                                                  # unreadable file: the reference skips the chip
        return e


def load_pairs(img_paths, lbl_paths, store_as_array, key_fn, validate=None, device=None, blobs=None, png_as_tf=False,
               planned=None, png_to_jpg=False):
    """Read + (optionally) decode a batch of chip pairs.  Returns one entry per pair: a dict ready for
    ops.build_records, or the Exception that makes the reference skip the chip.  `blobs` = the 2n file contents
    (image, label, image, label, ...) when the caller has already read them (run_worker prefetches on threads);
    `planned` = their _codec.plan_blobs result when the caller has also done the host half of the decode there."""
    ctx = get_ctx(device)
    n = len(img_paths)
    if blobs is None:
        blobs = []
        for i in range(n):
            blobs += [_read_or_error(img_paths[i]), _read_or_error(lbl_paths[i])]
    errs = [None] * n
    for k, b in enumerate(blobs):
        if isinstance(b, Exception):
            if errs[k // 2] is None:
                errs[k // 2] = b
            blobs[k] = b""
    arrays = [None] * (2 * n)
    if png_to_jpg:
        # convert_png_to_jpg (_img_to_tf_threaded.py:92-95): every PNG chip becomes the JPEG that
        # tf.image.encode_jpeg(decode_png(data), format='', quality=100) writes; the record then carries that file (or
        # its decoded pixels), exactly as if the chip had been a .jpg
        blobs = list(blobs)
        pa, pst, pinfos = _codec.decode_blobs(blobs, ctx.device, want_infos=True, png_as_tf=png_as_tf)
        todo = [k for k in range(2 * n) if pinfos[k].format == 2 and pinfos[k].status == 0 and pst[k] == 0
                and pa[k].shape[2] in (1, 3)]
        for k in range(2 * n):                                              # 2 / 4 channels: encode_jpeg refuses them
            if pinfos[k].format == 2 and pinfos[k].status == 0 and pst[k] == 0 and k not in todo and errs[k // 2] is None:
                errs[k // 2] = ChipError("encode_jpeg: image must have 1 or 3 channels (%s)" % (img_paths, lbl_paths)[k % 2][k // 2])
        for k, f in zip(todo, _codec.encode_jpeg_arrays([pa[k] for k in todo], quality=100, device=ctx.device)):
            print("Converting PNG to JPEG for %s" % (img_paths, lbl_paths)[k % 2][k // 2])
            blobs[k] = f
        if planned is not None:
            planned.release()                                               # planned from the PNG bytes: give its staging set back
        planned = None
        del pa
    if store_as_array:                                                      # ONE native planning call + the decode kernels
        if planned is None:
            planned = _codec.plan_blobs(blobs, ctx.device, png_as_tf)
        arrays, st, infos = _codec.decode_planned(planned, ctx.device, want_infos=True)
        arrays, st, infos = _codec.merge_jpeg(blobs, arrays, st, infos, ctx.device, candidates=np.nonzero(st)[0])   # .jpg chips
        for k in range(2 * n):
            if len(blobs[k]) and infos[k].status == 0 and st[k] != 0 and errs[k // 2] is None:
                errs[k // 2] = ChipError("could not decode %s (codec status %d)" % ((img_paths, lbl_paths)[k % 2][k // 2], int(st[k])))
    elif validate is not None:
        # threaded translator: the reference decodes every chip even when it stores the file bytes
        # (_img_to_tf_threaded.py:94-105), so a PNG whose IDAT data does not inflate is skipped, not stored
        if planned is None:
            planned = _codec.plan_blobs(blobs, ctx.device, png_as_tf)
        _, st, infos = _codec.decode_planned(planned, ctx.device, want_infos=True)
        _, st, infos = _codec.merge_jpeg(blobs, [None] * (2 * n), st, infos, ctx.device, candidates=np.nonzero(st)[0])
        for k in range(2 * n):
            if len(blobs[k]) and infos[k].status == 0 and st[k] != 0 and errs[k // 2] is None:
                errs[k // 2] = ChipError("could not decode %s (codec status %d)" % ((img_paths, lbl_paths)[k % 2][k // 2], int(st[k])))
    else:
        infos = _codec.probe_blobs(blobs, png_as_tf=png_as_tf)
        # .jpg chips: the reference decodes them even when it stores the file bytes (_img_to_tf_threaded.py:97-112)
        _, st, infos = _codec.merge_jpeg(blobs, [None] * (2 * n), np.zeros(2 * n, np.int32), infos, ctx.device,
                                         candidates=[k for k in range(2 * n) if infos[k].status != 0])
        for k in np.nonzero(st == 2)[0]:
            if errs[k // 2] is None:
                errs[k // 2] = ChipError("could not decode %s (codec status 2)" % (img_paths, lbl_paths)[k % 2][k // 2])
    out = []
    for i in range(n):
        if errs[i] is not None:
            out.append(errs[i])
            continue
        ii, li = infos[2 * i], infos[2 * i + 1]
        try:
            for info, p, b in ((ii, img_paths[i], blobs[2 * i]), (li, lbl_paths[i], blobs[2 * i + 1])):
                if not len(b) or info.status != 0:
                    raise ChipError("'%s' not recognized as a supported file format." % p)
            if validate is not None:
                validate(ii)
                validate(li)
            ikey, lkey = key_fn(img_paths[i], ii), key_fn(lbl_paths[i], li)
            assert ikey == lkey                                             # _img_to_tf_mp.py:132
        except Exception as e:
            out.append(e)
            continue
        if store_as_array:
            img, lab = arrays[2 * i], arrays[2 * i + 1]
            as_bytes = img.dtype == torch.uint8 and lab.dtype == torch.uint8          # convert_to_example :160-197
            item = dict(img=img.reshape(-1), tgt=lab.reshape(-1), kind=1 if as_bytes else 2)
        else:
            item = dict(img=ops.to_device(blobs[2 * i], ctx.device), tgt=ops.to_device(blobs[2 * i + 1], ctx.device), kind=1)
        item.update(h=ii.height, w=ii.width, c=ii.samples, th=li.height, tw=li.width, identifier=ikey.encode("utf-8"))
        out.append(item)
    return out


_ELEM_SIZE = np.zeros(8, np.int64)
for _code, _sz in ((_lib_mod.B2_U8, 1), (_lib_mod.B2_I8, 1), (_lib_mod.B2_U16, 2), (_lib_mod.B2_I16, 2), (_lib_mod.B2_U32, 4),
                   (_lib_mod.B2_I32, 4), (_lib_mod.B2_F32, 4), (_lib_mod.B2_F64, 8)):
    _ELEM_SIZE[_code] = _sz


class BatchRecords:
    """Every record of one clean decode batch serialised by ONE kernel launch into ONE device buffer (records back to
    back in list order), planned by one native call (b2_example_layout_batch) and a few NumPy expressions: no per-record
    Python.  Shard files then take contiguous byte ranges of that buffer."""

    def __init__(self, ctx, keys, kind, ib, tb, dims, sdt, tdt, img_src, tgt_src, icount, tcount, keep):
        n = len(keys)
        ids = b"".join(keys)
        id_off = np.zeros(n + 1, np.uint64)
        np.cumsum([len(k) for k in keys], out=id_off[1:])
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        ib, tb = np.ascontiguousarray(ib, dtype=np.uint64), np.ascontiguousarray(tb, dtype=np.uint64)
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        descs = np.zeros(n, dtype=np.dtype(_lib_mod.BUILD_DESC_DTYPE))
        cap = 320 * n + len(ids) + 64
        scaf = np.empty(cap, np.uint8)
        sl, tot, mx = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        idb = np.frombuffer(ids, np.uint8) if ids else np.zeros(1, np.uint8)
        check(lib().b2_example_layout_batch(n, kind.ctypes.data, ib.ctypes.data, tb.ctypes.data, dims.ctypes.data, idb.ctypes.data,
                                            id_off.ctypes.data, descs.ctypes.data, scaf.ctypes.data, cap, ctypes.byref(sl),
                                            ctypes.byref(tot), ctypes.byref(mx)))
        descs["src_dtype"], descs["tgt_dtype"] = sdt, tdt
        descs["img_src"], descs["tgt_src"] = img_src, tgt_src
        descs["img_count"], descs["tgt_count"] = icount, tcount
        self.total, self.rec_off = int(tot.value), descs["out_off"].astype(np.int64)
        self.out = torch.empty(((self.total + 15) // 16 * 16 + 16,), dtype=torch.uint8, device=ctx.device)
        descs_d = ops.to_device(descs.view(np.uint8), ctx.device)
        scaf_d = ops.to_device(scaf[:max(1, int(sl.value))], ctx.device)
        isz = descs.dtype.itemsize
        for s0 in range(0, n, 65535):
            m = min(65535, n - s0)
            check(lib().b2_tfrecord_build(ctx.handle, ctypes.c_void_p(descs_d.data_ptr() + s0 * isz), m, int(mx.value), _lib_mod.ptr(scaf_d),
                                          _lib_mod.ptr(self.out), ctx.stream()))
        self.keep = (descs_d, scaf_d, keep)

    @classmethod
    def from_decode(cls, job, keys, ctx):
        """Decoded arrays (store_as_array=True): BytesList iff image and label are both uint8, else FloatList for both
        (convert_to_example, _tfrecord_image_translation.py:160-197)."""
        infos, images = job.infos_array(), job.images
        ii, li = infos[0::2], infos[1::2]
        kind = np.where((ii["dtype"] == _lib_mod.B2_U8) & (li["dtype"] == _lib_mod.B2_U8), 1, 2)
        inum = ii["width"].astype(np.int64) * ii["height"] * ii["samples"]
        lnum = li["width"].astype(np.int64) * li["height"] * li["samples"]
        ib = np.where(kind == 1, inum * _ELEM_SIZE[ii["dtype"]], inum * 4)
        tb = np.where(kind == 1, lnum * _ELEM_SIZE[li["dtype"]], lnum * 4)
        dims = np.stack([ii["height"], ii["width"], ii["samples"], li["height"], li["width"]], axis=1)
        base = job.out.data_ptr()
        return cls(ctx, keys, kind, ib, tb, dims, ii["dtype"], li["dtype"], base + images["out_off"][0::2], base + images["out_off"][1::2],
                   np.where(kind == 1, inum * _ELEM_SIZE[ii["dtype"]], inum), np.where(kind == 1, lnum * _ELEM_SIZE[li["dtype"]], lnum), job)

    @classmethod
    def from_files(cls, dev_buf, offs, sizes, infos, keys, ctx):
        """The files' own bytes (store_as_array=False, _img_to_tf_mp.py:73-75): BytesList of the raw file content."""
        ii, li = infos[0::2], infos[1::2]
        n = len(keys)
        dims = np.stack([ii["height"], ii["width"], ii["samples"], li["height"], li["width"]], axis=1)
        base = dev_buf.data_ptr()
        u8 = np.full(n, _lib_mod.B2_U8, np.int32)
        return cls(ctx, keys, np.ones(n, np.int32), sizes[0::2], sizes[1::2], dims, u8, u8, base + offs[0::2], base + offs[1::2],
                   sizes[0::2], sizes[1::2], dev_buf)

    @classmethod
    def from_jpeg(cls, jb, keys, ctx):
        """Decoded .jpg chips (store_as_array=True): uint8 pixels, BytesList records."""
        d, off = jb.dims(), jb.out_offsets()
        ii, li = d[0::2], d[1::2]
        n = len(keys)
        inum = ii["width"].astype(np.int64) * ii["height"] * ii["samples"]
        lnum = li["width"].astype(np.int64) * li["height"] * li["samples"]
        dims = np.stack([ii["height"], ii["width"], ii["samples"], li["height"], li["width"]], axis=1)
        base = jb.out.data_ptr()
        u8 = np.full(n, _lib_mod.B2_U8, np.int32)
        return cls(ctx, keys, np.ones(n, np.int32), inum, lnum, dims, u8, u8, base + off[0::2], base + off[1::2], inum, lnum, jb)

    def byte_range(self, lo, hi):
        """Bytes of records [lo, hi) of the batch."""
        return int(self.rec_off[lo]), (int(self.rec_off[hi]) if hi < len(self.rec_off) else self.total)


def batch_schedule(shard_ranges, batch_pairs, interleave=8):
    """Decode batches of one worker: a list of batches, each a list of runs (shard, lo, hi) over the file list.  Every file of
    [shard_ranges[0], shard_ranges[-1]) appears once, and each shard's files in order."""
    per = len(shard_ranges) - 1
    lo_all, hi_all = int(shard_ranges[0]), int(shard_ranges[-1])
    # the first batches are small (1/8, 1/4, 1/2 of the full size): the pipeline's start-up latency is the time the first
    # batch takes to go through read -> plan -> decode -> build -> write
    # ... and the last ones taper off again (1/2, 1/4, 1/4): what follows the last read — decode, serialise, copy back,
    # write — is the pipeline's tail, and it is as long as the last batch is large
    sizes, left, size = [], hi_all - lo_all, max(32, batch_pairs // 8)
    while left > 0 and size < batch_pairs:
        sizes.append(min(size, left))
        left -= sizes[-1]
        size *= 2
    tail = []
    if left > batch_pairs:
        for frac in (4, 4, 2):
            t = min(left, max(32, batch_pairs // frac))
            tail.insert(0, t)
            left -= t
    while left > 0:
        sizes.append(min(batch_pairs, left))
        left -= sizes[-1]
    sizes += tail                                                           # e.g. ..., 1024, 512, 256, 256
    # a batch is a list of runs (shard, lo, hi): the same number of pairs from each of up to interleave shards, so that every
    # batch's records go to that many files at once — writes to ONE file serialise in the kernel (see write_back), and with
    # multi-megabyte records a batch of consecutive pairs would touch two or three files only
    batches = []
    for g0 in range(0, per, interleave):
        group = list(range(g0, min(per, g0 + interleave)))
        cursor = {s: int(shard_ranges[s]) for s in group}
        left_g = sum(int(shard_ranges[s + 1]) - cursor[s] for s in group)
        while left_g > 0:
            want = min(sizes.pop(0) if sizes else batch_pairs, left_g)
            if want <= 0:
                continue
            runs, got = [], 0
            while got < want:
                active = [s for s in group if cursor[s] < int(shard_ranges[s + 1])]
                quota = -(-(want - got) // len(active))
                for s in active:
                    k = min(quota, int(shard_ranges[s + 1]) - cursor[s], want - got)
                    if k > 0:
                        if runs and runs[-1][0] == s:
                            runs[-1] = (s, runs[-1][1], cursor[s] + k)
                        else:
                            runs.append((s, cursor[s], cursor[s] + k))
                        cursor[s] += k
                        got += k
            runs.sort()
            merged = []
            for r in runs:                                                  # a second helping from the same shard joins its run
                if merged and merged[-1][0] == r[0] and merged[-1][2] == r[1]:
                    merged[-1] = (r[0], merged[-1][1], r[2])
                else:
                    merged.append(r)
            batches.append(merged)
            left_g -= got
    assert sum(hi - lo for runs in batches for _, lo, hi in runs) == hi_all - lo_all
    return batches


def run_worker(worker_index, ranges, name, img_filenames, lbl_filenames, output_directory, num_shards, key_fn,
               store_as_array, label="process", progress_every=100, validate=None, device=None, batch_pairs=None,
               io_threads=8, png_as_tf=False, png_to_jpg=False, path_key=None, fast_validate=None, info_keys=False):
    """The reference's worker loop as a pipeline over decode batches:

        read-ahead thread : files of batch k+2 -> pinned staging (one native call), header parse + stream tables in place
        main thread       : upload + decode kernels of batch k+1 queued; then batch k: status, records built by one launch,
                            device -> pinned host copy queued
        writer threads    : batch k-1: one pwrite per shard file and batch, different files in parallel

    A batch in which every chip reads, decodes and pairs up cleanly never touches per-chip Python; a batch with any
    irregularity (unreadable / undecodable chip, key mismatch, .jpg chips, convert_png_to_jpg, raw-bytes records, identifiers
    from the georeferencing) goes through load_pairs, which reproduces the reference's skip-and-continue chip by chip.

    path_key(path) -> identifier (when it depends on the file name only); info_keys: the identifier needs the file's header
    as well (georeferencing) — key_fn(path, info) is then called per file on the read-ahead thread with the planner's
    ImageInfo; fast_validate(infos) -> boolean array, False where the reference's shape asserts would fail."""
    from concurrent.futures import ThreadPoolExecutor
    num_workers = len(ranges)
    assert not num_shards % num_workers
    per = int(num_shards / num_workers)
    shard_ranges = np.linspace(ranges[worker_index][0], ranges[worker_index][1], per + 1).astype(int)
    num_files = ranges[worker_index][1] - ranges[worker_index][0]
    ctx = get_ctx(device)
    os.makedirs(output_directory, exist_ok=True)
    # host threads: this worker's share of the box's cores (one worker per GPU runs beside it, in this process or under
    # torchrun); more threads than cores only adds context switches — an 8-GPU box here has 4 cores per GPU
    try:
        cores = len(os.sched_getaffinity(0))                               # the cores this process may run on (taskset, cgroups)
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    share = max(1, cores // max(1, _workers_on_this_host()))
    io_threads = max(2, min(io_threads, share))
    plan_threads = max(2, min(16, share))
    write_threads = max(2, min(16, 2 * share))
    try:
        pair_bytes = os.path.getsize(img_filenames[ranges[worker_index][0]]) + os.path.getsize(lbl_filenames[ranges[worker_index][0]])
    except (OSError, IndexError):
        pair_bytes = 1 << 20
    if batch_pairs is None:
        # a batch should carry thousands of compressed streams for the decoders, but not more than ~1 GB of file bytes
        # (pinned staging) — sized from the first pair
        batch_pairs = int(max(32, min(1024, (512 << 20) // max(1, pair_bytes))))
    # clean batches of decoded-array records, or of raw-file records that need no decode to be validated, skip per-chip Python
    # convert_png_to_jpg: decode, encode and assemble the JPEG files on the device; array records carry the pixels of those
    # files decoded again (the files cross to the host once for the JPEG planner's marker walk)
    to_jpg = bool(png_to_jpg)
    fast = bool((path_key is not None or info_keys) and (store_as_array or validate is None or fast_validate is not None))
    # file-bytes records whose chips must decode before they are accepted (the threaded translator, :94-105): the files go up
    # as they are for the records AND a planned copy of their compressed streams goes through the decoders for the verdict
    check_decode = bool(fast and not store_as_array and validate is not None and not to_jpg)
    kAhead = 3                                                              # batches being read / planned ahead of the GPU
    if fast:
        try:                                                                # .jpg chips are planned into a second set as well
            two_sets = check_decode or str(img_filenames[ranges[worker_index][0]]).lower().endswith((".jpg", ".jpeg"))
        except IndexError:
            two_sets = check_decode
        _codec.reserve_staging(ctx.device, int(pair_bytes * batch_pairs * 1.25) + (1 << 20),
                               sets=(kAhead + 2) * (2 if two_sets else 1))
    n_slots = 4                                                             # rotating pinned write-back buffers, kept per device
    cache = _worker_buffers.setdefault(ctx.device.index, {"pinned": [None] * n_slots, "reader": None})
    pinned = cache["pinned"]
    slot_futs = [[] for _ in range(n_slots)]                                # writes still reading a buffer
    batches = batch_schedule(shard_ranges, batch_pairs)
    files = [open(os.path.join(output_directory, "%s-%.5d-of-%.5d" % (name, worker_index * per + s, num_shards)), "w+b")
             for s in range(per)]
    shard_off = [0] * per
    shard_count = [0] * per
    state = {"counter": 0, "seq": 0}
    copy_stream = torch.cuda.Stream(ctx.device)
    trace = [] if os.environ.get("B2_TRANSLATE_TRACE") else None      # development aid: (seconds, event) per pipeline step
    import time as _time
    t_start = _time.perf_counter()

    def mark(ev):
        if trace is not None:
            trace.append((round(_time.perf_counter() - t_start, 4), ev))

    def count_records(s, k):
        """k more records in shard s: the reference's progress line at every multiple of progress_every."""
        before = state["counter"]
        state["counter"] += k
        shard_count[s] += k
        for mult in range(before // progress_every + 1, state["counter"] // progress_every + 1):
            print("%s [%s %d]: Processed %d of %d images in %s batch." %
                  (datetime.now(), label, worker_index, mult * progress_every, num_files, label))
            sys.stdout.flush()

    def shard_done(s):
        print("%s [%s %d]: Wrote %d images to %s" % (datetime.now(), label, worker_index, shard_count[s], files[s].name))
        sys.stdout.flush()

    with ThreadPoolExecutor(max_workers=kAhead) as pool, ThreadPoolExecutor(max_workers=1) as writer, \
            ThreadPoolExecutor(max_workers=write_threads) as wpool:
        reader = cache["reader"]
        if reader is None:
            reader = cache["reader"] = FileBatchReader(depth=5, threads=io_threads)
        reader.threads = io_threads

        def write_back(buf, total, pieces):
            """Device buffer -> pinned slot -> files.  pieces: [(shard, byte lo, byte hi)] of buf."""
            slot = state["seq"] % n_slots
            state["seq"] += 1
            for x in slot_futs[slot]:                                       # the writes that last used this pinned buffer
                for y in x.result():
                    y.result()
            slot_futs[slot] = []
            if pinned[slot] is None or pinned[slot].numel() < total:
                pinned[slot] = torch.empty((int(total * 1.1) + 4096,), dtype=torch.uint8).pin_memory()
            host = pinned[slot][:total]
            cur = torch.cuda.current_stream(ctx.device)
            copy_stream.wait_stream(cur)                                    # the copy runs beside the next batch's decode
            with torch.cuda.stream(copy_stream):
                host.copy_(buf[:total], non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy_stream)
            buf.record_stream(copy_stream)
            jobs = []
            for s, lo, hi in pieces:
                jobs.append((files[s].fileno(), lo, hi, shard_off[s]))
                shard_off[s] += hi - lo

            def _write(h=host, ev=done, jobs=jobs):
                ev.synchronize()                                             # the records have arrived in pinned memory
                src = h.numpy()
                mv = memoryview(src)
                # write(2) on ONE file serialises on its inode lock (3.3 GB/s on tmpfs whatever the thread count), writes to
                # different files do not (30 GB/s over 8 files, 35-46 GB/s over 16: tools/shm_write_probe.py): one pwrite per
                # shard file and batch.  When a batch touches only a few files (large records), the idle threads copy the
                # rest of each piece into a shared mapping of the file — page faults do not take the inode lock — which
                # nearly doubles the rate (4 files: 12.9 -> 23 GB/s)
                futs = []
                spare = write_threads // max(1, len(jobs)) - 1
                for fd, lo, hi, off in jobs:
                    n = hi - lo
                    k = min(3, spare) if n >= _MAPPED_WRITE_MIN else 0
                    head = n if k <= 0 else (int(n * 0.4) & ~4095)
                    if k > 0:
                        try:
                            os.ftruncate(fd, max(os.fstat(fd).st_size, off + n))
                            a0 = (off + head) & ~(mmap.ALLOCATIONGRANULARITY - 1)
                            mm = mmap.mmap(fd, off + n - a0, offset=a0)
                            dst = np.frombuffer(mm, dtype=np.uint8)[off + head - a0:]
                            step = ((n - head + k - 1) // k + 4095) & ~4095
                            for o in range(0, n - head, step):
                                e = min(o + step, n - head)
                                futs.append(wpool.submit(np.copyto, dst[o:e], src[lo + head + o:lo + head + e]))
                        except (OSError, ValueError):                       # a file system without shared mappings
                            head = n
                    futs.append(wpool.submit(os.pwrite, fd, mv[lo:lo + head], off))
                return futs
            slot_futs[slot].append(writer.submit(_write))

        def read_and_plan(rng):
            mark("thread read start %d" % rng[0][1])
            try:
                return _read_and_plan(rng)
            finally:
                mark("thread read end %d" % rng[0][1])

        def _read_and_plan(rng):
            paths = []
            for _, lo, hi in rng:
                for i in range(lo, hi):
                    paths += [img_filenames[i], lbl_filenames[i]]
            if not fast:
                blobs = reader.read(paths)                                  # one native call; the GIL is free meanwhile
                planned = None
                if (store_as_array or validate is not None) and not png_to_jpg:   # host half of the decode, off the main thread
                    planned = _codec.plan_blobs([b"" if isinstance(b, Exception) else b for b in blobs], ctx.device, png_as_tf,
                                                threads=plan_threads)
                return dict(fast=False, blobs=blobs, planned=planned)
            hs = _codec.take_staging(ctx.device)
            planned = infos = offs = sizes = jpeg = probed = None
            try:
                blobs, offs, sizes, clean = reader.read_into(paths, hs)     # straight into the pinned staging buffer
                if clean and not to_jpg and path_key is not None and sizes.min() >= 3:   # a batch of .jpg chips (SOI marker FF D8 FF)
                    st_np, o64 = hs.stage.numpy(), offs.astype(np.int64)
                    if bool(np.all((st_np[o64] == 0xFF) & (st_np[o64 + 1] == 0xD8) & (st_np[o64 + 2] == 0xFF))):
                        _codec.reserve_staging(ctx.device, 0, sets=(kAhead + 2) * 2)   # a second set per batch in flight
                        jpeg = _codec.plan_jpeg_batch(blobs, ctx.device, threads=plan_threads)
                if jpeg is not None:
                    if jpeg.host_status.any():                              # a file the marker walk refuses: chip by chip
                        jpeg.release()
                        jpeg, clean = None, False
                elif clean and (store_as_array or to_jpg):
                    planned = _codec.plan_blobs(blobs, ctx.device, png_as_tf, inplace=hs, threads=plan_threads)
                elif clean and check_decode:                                # into a staging set of its own: hs keeps the files
                    planned = _codec.plan_blobs(blobs, ctx.device, png_as_tf, threads=plan_threads)
                elif clean:                                                 # raw-bytes records: the header fields only
                    probed = _codec.probe_blobs(blobs, png_as_tf=png_as_tf)
                    infos = np.frombuffer(probed, dtype=_codec.IMAGE_INFO_DTYPE, count=len(paths))
                    clean = not infos["status"].any()
            except Exception:
                planned, clean = None, False
            if not clean or ((store_as_array or check_decode or to_jpg) and planned is None and jpeg is None):
                if planned is not None:
                    planned.release()
                if jpeg is not None:
                    jpeg.release()
                hs.pending = False
                return dict(fast=False, blobs=None, planned=None)           # the chip-by-chip path re-reads the files
            if path_key is not None:
                keys = [path_key(p) for p in paths]
            else:                                                           # identifiers from the header (georeferencing)
                ci = planned.infos if planned is not None else probed
                keys = [key_fn(p, ci[i]) for i, p in enumerate(paths)]
            return dict(fast=True, planned=planned, keys=keys, hs=hs, infos=infos, offs=offs, sizes=sizes, jpeg=jpeg)

        def stage1(bi):
            """Batch bi: wait for its files + plan, queue its upload and decode (nothing synchronised)."""
            mark("wait read %d" % bi)
            b = pending.pop(0).result()
            mark("got read %d" % bi)
            if bi + kAhead < len(batches):
                pending.append(pool.submit(read_and_plan, batches[bi + kAhead]))  # kAhead batches ahead, one thread each
            b["runs"] = batches[bi]
            if b["fast"] and b["planned"] is not None:
                b["job"] = _codec.decode_enqueue(b["planned"], ctx.device)
            if b["fast"] and b.get("jpeg") is not None:                     # the reference decodes a .jpg chip whatever it stores
                _codec.jpeg_decode_enqueue(b["jpeg"], ctx.device)
            if b["fast"] and store_as_array and b.get("jpeg") is not None:
                b["hs"].pending = False                                     # .jpg pixels come from the JPEG plan's own set
            if b["fast"] and not to_jpg and not store_as_array:            # the files as they are: one upload of the staging buffer
                hs = b["hs"]
                used = int(b["offs"][-1] + b["sizes"][-1]) + 16
                b["dev"] = hs.stage[:used].to(ctx.device, non_blocking=True)
                hs.busy = torch.cuda.Event()
                hs.busy.record(torch.cuda.current_stream(ctx.device))
                hs.pending = False
            return b

        def slow_batch(runs, blobs, planned):
            idx = [i for _, lo, hi in runs for i in range(lo, hi)]
            pairs = load_pairs([img_filenames[i] for i in idx], [lbl_filenames[i] for i in idx], store_as_array,
                               key_fn, validate, ctx.device, blobs=blobs, png_as_tf=png_as_tf, planned=planned,
                               png_to_jpg=png_to_jpg)
            pos = 0
            for s, lo, hi in runs:
                items = []
                for i in range(lo, hi):
                    p = pairs[pos + i - lo]
                    if isinstance(p, Exception):
                        print(p)
                        print("SKIPPED: Unexpected eror while decoding %s." % img_filenames[i])
                        continue
                    items.append(p)
                    count_records(s, 1)
                pos += hi - lo
                if items:
                    buf, _, total = ops.build_records(items, ctx.device)
                    write_back(buf, total, [(s, 0, total)])
                if hi == int(shard_ranges[s + 1]):
                    shard_done(s)

        def finish(b):
            runs = b["runs"]
            if b["fast"]:
                keys = b["keys"]
                jpeg = b.get("jpeg")
                if jpeg is not None:
                    mark("wait status")
                    ok = not jpeg.status().any()
                    mark("got status")
                    infos = jpeg.dims()
                elif b["planned"] is not None:
                    job = b["job"]
                    mark("wait status")
                    st = job.status()                                       # waits for this batch's decode only
                    mark("got status")
                    infos = job.infos_array()
                    ok = not st.any() and not infos["status"].any()
                else:
                    infos, ok = b["infos"], True
                ok = ok and keys[0::2] == keys[1::2]
                if ok and fast_validate is not None and jpeg is None:        # (a baseline JPEG has 1 or 3 components: always valid)
                    ok = bool(np.all(fast_validate(infos)))
                if ok and to_jpg:                                           # tf.image.encode_jpeg takes 1 or 3 channels of uint8
                    ok = bool(np.all((infos["samples"] != 2) & (infos["dtype"] == _lib_mod.B2_U8)))
                if ok:
                    ids = [k.encode("utf-8") for k in keys[0::2]]
                    if to_jpg:
                        for _, lo, hi in runs:
                            for i in range(lo, hi):
                                print("Converting PNG to JPEG for %s" % img_filenames[i])
                                print("Converting PNG to JPEG for %s" % lbl_filenames[i])
                        files, foffs, fsizes = _codec.encode_jpeg_device(job.out, job.images["out_off"], infos["height"], infos["width"],
                                                                         infos["samples"], quality=100, device=ctx.device)
                        if store_as_array:
                            used = int(foffs[-1] + fsizes[-1])
                            hbuf = cache.get("jpg_host")
                            if hbuf is None or hbuf.numel() < used:
                                hbuf = cache["jpg_host"] = torch.empty((int(used * 1.25) + 4096,), dtype=torch.uint8).pin_memory()
                            hbuf[:used].copy_(files[:used], non_blocking=True)
                            torch.cuda.current_stream(ctx.device).synchronize()
                            hnp = hbuf.numpy()
                            _codec.reserve_staging(ctx.device, 0, sets=(kAhead + 2) * 2)
                            jb = _codec.plan_jpeg_batch([hnp[int(o):int(o + z)] for o, z in zip(foffs, fsizes)], ctx.device,
                                                        threads=plan_threads)
                            _codec.jpeg_decode_enqueue(jb, ctx.device)
                            if jb.status().any():
                                raise B2Error("convert_png_to_jpg: a JPEG file this package encoded does not decode")
                            rec = BatchRecords.from_jpeg(jb, ids, ctx)
                        else:
                            rec = BatchRecords.from_files(files, foffs, fsizes, infos, ids, ctx)
                    elif jpeg is not None and store_as_array:
                        rec = BatchRecords.from_jpeg(jpeg, ids, ctx)
                    else:
                        rec = BatchRecords.from_decode(job, ids, ctx) if store_as_array else \
                            BatchRecords.from_files(b["dev"], b["offs"], b["sizes"], infos, ids, ctx)
                    pieces, done, pos = [], [], 0
                    for s, lo, hi in runs:
                        pieces.append((s,) + rec.byte_range(pos, pos + hi - lo))
                        pos += hi - lo
                        count_records(s, hi - lo)
                        if hi == int(shard_ranges[s + 1]):
                            done.append(s)
                    mark("built")
                    write_back(rec.out, rec.total, pieces)
                    mark("queued write")
                    for s in done:
                        shard_done(s)
                    return
                b = dict(blobs=None, planned=None)                          # an irregular batch: chip by chip, from the files
            slow_batch(runs, b.get("blobs"), b.get("planned"))

        pending = [pool.submit(read_and_plan, batches[k]) for k in range(min(kAhead, len(batches)))]
        prev = None
        try:
            for bi in range(len(batches)):
                cur = stage1(bi)
                if prev is not None:
                    finish(prev)
                prev = cur
            if prev is not None:
                finish(prev)
                prev = None
        finally:
            for leftover in pending:
                if leftover is not None:                                    # aborted with a read-ahead in flight: hand its
                    try:                                                    # pinned staging set back
                        res = leftover.result()
                        dropped = res.get("planned")
                        if dropped is not None:
                            dropped.release()
                        if res.get("jpeg") is not None:
                            res["jpeg"].release()
                        if res.get("hs") is not None and (not store_as_array or res.get("jpeg") is not None):
                            res["hs"].pending = False                       # the set that holds the files themselves
                    except Exception:
                        pass
            if prev is not None and prev.get("planned") is not None:
                prev["planned"].release()
        for s in range(per):                                                # shards that received no file at all
            if int(shard_ranges[s]) == int(shard_ranges[s + 1]):
                shard_done(s)
        mark("drain writes")
        for fl in slot_futs:
            for x in fl:
                for y in x.result():
                    y.result()
        mark("done")
    if trace is not None:
        sys.stderr.write("translate trace (worker %d): %s\n" % (worker_index, trace))
    for f in files:
        f.close()
    print("%s [%s %d]: Wrote %d images to %d shards." % (datetime.now(), label, worker_index, state["counter"], per))
    sys.stdout.flush()
    return state["counter"]


_concurrent_workers = [1]      # GPU workers running at the same time in THIS process (set by run_workers)


def _workers_on_this_host():
    """How many GPU workers share this host's cores right now: the ranks of a torchrun job on this node, or the device
    threads of a single process."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    return _concurrent_workers[0]


def run_workers(num_workers, fn):
    """Run fn(worker_index, device) for every worker this OS process owns.  Workers on different GPUs run
    concurrently, one host thread per GPU (the C ABI's rule is one context = one device = one thread at a time;
    the decode / serialise work is native or on the device, so the threads do not fight over the GIL); workers
    that share a GPU run one after the other on that GPU's thread.  Results come back in worker order."""
    mine = my_workers(num_workers)
    by_dev = {}
    for p, dev in mine:
        by_dev.setdefault(dev, []).append(p)
    results = {}

    def on_device(dev):
        torch.cuda.set_device(dev)
        for p in by_dev[dev]:
            results[p] = fn(p, dev)
    if len(by_dev) <= 1 or os.environ.get("B2_WORKER_THREADS", "1") == "0":
        _concurrent_workers[0] = 1
        for dev in by_dev:
            on_device(dev)
    else:
        _concurrent_workers[0] = len(by_dev)
        from concurrent.futures import ThreadPoolExecutor
        for dev in by_dev:
            get_ctx(dev)                                                   # contexts are created on the calling thread
        with ThreadPoolExecutor(max_workers=len(by_dev)) as pool:
            for f in [pool.submit(on_device, dev) for dev in by_dev]:
                f.result()
    return [results[p] for p, _ in mine]


def my_workers(num_workers):
    """Which worker indices this OS process runs, and on which device.

    Single process: all of them, worker p on GPU p % visible.  Under torchrun (one rank per GPU): worker p
    belongs to rank p % WORLD_SIZE and runs on LOCAL_RANK — shard ownership is unchanged, so N-rank and
    1-rank runs write byte-identical files.
    """
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    ngpu = max(1, torch.cuda.device_count())
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", rank % ngpu))
        return [(p, local) for p in range(num_workers) if p % world == rank]
    return [(p, p % ngpu) for p in range(num_workers)]
