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
from ._lib import check, get_ctx, lib


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
        self.bufs = [np.empty(0, np.uint8) for _ in range(depth)]
        self.k = 0
        self.threads = int(threads)

    def read(self, paths):
        n = len(paths)
        if n == 0:
            return []
        enc = [os.fsencode(p) for p in paths]
        arr = (ctypes.c_char_p * n)(*enc)
        offs, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        status = np.zeros(n, np.int32)
        need = ctypes.c_uint64(0)
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


def run_worker(worker_index, ranges, name, img_filenames, lbl_filenames, output_directory, num_shards, key_fn,
               store_as_array, label="process", progress_every=100, validate=None, device=None, batch_pairs=None,
               io_threads=8, png_as_tf=False, png_to_jpg=False):
    """The reference's worker loop, restructured as a three-stage pipeline: a thread pool reads the files of batch
    k+1 while the GPU decodes and serialises batch k and a writer thread appends batch k-1 to the shard file."""
    from concurrent.futures import ThreadPoolExecutor
    num_workers = len(ranges)
    assert not num_shards % num_workers
    per = int(num_shards / num_workers)
    shard_ranges = np.linspace(ranges[worker_index][0], ranges[worker_index][1], per + 1).astype(int)
    num_files = ranges[worker_index][1] - ranges[worker_index][0]
    ctx = get_ctx(device)
    counter = 0
    os.makedirs(output_directory, exist_ok=True)
    if batch_pairs is None:
        # the entropy decoders run one warp per compressed stream and are latency-bound: a batch should carry a few
        # thousand streams, but not more than ~1 GB of file bytes (pinned staging) — sized from the first pair
        try:
            pair_bytes = os.path.getsize(img_filenames[ranges[worker_index][0]]) + os.path.getsize(lbl_filenames[ranges[worker_index][0]])
        except (OSError, IndexError):
            pair_bytes = 1 << 20
        batch_pairs = int(max(32, min(2048, (768 << 20) // max(1, pair_bytes))))
    n_slots = 4                                                             # rotating pinned write-back buffers
    # rotating pinned write-back buffers and read-ahead buffers: kept per device between calls (pinning and first-touching
    # a few hundred MB costs ~0.1 s, which is what a worker with a few thousand chips takes altogether)
    cache = _worker_buffers.setdefault(ctx.device.index, {"pinned": [None] * n_slots, "reader": None})
    pinned = cache["pinned"]
    slot_futs = [[] for _ in range(n_slots)]                                # positional writes still reading a buffer
    # decode batches run over the worker's whole file range (the entropy decoders want thousands of streams per launch);
    # records are then serialised and written shard by shard, so a batch may feed several shard files
    lo_all, hi_all = int(shard_ranges[0]), int(shard_ranges[-1])
    # the first batches are small (1/8, 1/4, 1/2 of the full size): the pipeline's start-up latency is the time the first
    # batch takes to go through read -> plan -> decode -> build -> write, and a worker with a few thousand chips is mostly that
    batches, b0, size = [], lo_all, max(32, batch_pairs // 8)
    while b0 < hi_all:
        b1 = min(b0 + size, hi_all)
        batches.append((b0, b1))
        b0, size = b1, min(batch_pairs, size * 2)
    use_mmap = os.environ.get("B2_SHARD_WRITE", "mmap") == "mmap"
    files = [open(os.path.join(output_directory, "%s-%.5d-of-%.5d" % (name, worker_index * per + s, num_shards)), "w+b")
             for s in range(per)]
    shard_off = [0] * per
    shard_count = [0] * per
    seq = 0
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool, ThreadPoolExecutor(max_workers=1) as writer, \
            ThreadPoolExecutor(max_workers=8) as wpool:
        reader = cache["reader"]
        if reader is None:
            reader = cache["reader"] = FileBatchReader(depth=3, threads=io_threads)

        def read_and_plan(paths):
            blobs = reader.read(paths)                                      # one native call; the GIL is free meanwhile
            planned = None
            if (store_as_array or validate is not None) and not png_to_jpg: # host half of the decode, off the main thread
                planned = _codec.plan_blobs([b"" if isinstance(b, Exception) else b for b in blobs], ctx.device, png_as_tf)
            return blobs, planned

        def submit_reads(rng):
            paths = []
            for i in range(*rng):
                paths += [img_filenames[i], lbl_filenames[i]]
            return pool.submit(read_and_plan, paths)
        pending_reads = submit_reads(batches[0]) if batches else None
        try:
            for bi, (b0, b1) in enumerate(batches):
                blobs, planned = pending_reads.result()
                pending_reads = submit_reads(batches[bi + 1]) if bi + 1 < len(batches) else None
                idx = list(range(b0, b1))
                pairs = load_pairs([img_filenames[i] for i in idx], [lbl_filenames[i] for i in idx], store_as_array,
                                   key_fn, validate, ctx.device, blobs=blobs, png_as_tf=png_as_tf, planned=planned,
                                   png_to_jpg=png_to_jpg)
                for s in range(per):
                    lo, hi = max(b0, int(shard_ranges[s])), min(b1, int(shard_ranges[s + 1]))
                    if lo >= hi:
                        continue
                    items = []
                    for i in range(lo, hi):
                        p = pairs[i - b0]
                        if isinstance(p, Exception):
                            print(p)
                            print("SKIPPED: Unexpected eror while decoding %s." % img_filenames[i])
                            continue
                        items.append(p)
                        shard_count[s] += 1
                        counter += 1
                        if not counter % progress_every:
                            print("%s [%s %d]: Processed %d of %d images in %s batch." %
                                  (datetime.now(), label, worker_index, counter, num_files, label))
                            sys.stdout.flush()
                    if items:
                        buf, _, total = ops.build_records(items, ctx.device)
                        slot = seq % n_slots
                        seq += 1
                        for x in slot_futs[slot]:                                # the writes that last used this buffer
                            for y in x.result():
                                y.result()
                        slot_futs[slot] = []
                        if pinned[slot] is None or pinned[slot].numel() < total:
                            pinned[slot] = torch.empty((int(total * 1.1) + 4096,), dtype=torch.uint8).pin_memory()
                        host = pinned[slot][:total]
                        host.copy_(buf[:total], non_blocking=True)
                        done = torch.cuda.Event()
                        done.record(torch.cuda.current_stream(ctx.device))

                        def _write(fd=files[s].fileno(), h=host, ev=done, off=shard_off[s]):
                            ev.synchronize()                                     # the records have arrived in pinned memory
                            src = h.numpy()
                            step = 16 << 20
                            if use_mmap:
                                # write(2) on one file serialises on its inode lock (measured: 8 threads of pwrite = 3.5 GB/s
                                # on tmpfs, the speed of one); page faults on a shared mapping do not, so the chunks are
                                # copied into a mapping of the (grown) file by several threads at once
                                end = off + len(src)
                                try:
                                    os.ftruncate(fd, end)                        # this thread is the only one that grows files
                                    a0 = off & ~(mmap.ALLOCATIONGRANULARITY - 1)
                                    mm = mmap.mmap(fd, end - a0, offset=a0)
                                except (OSError, ValueError):                    # a file system without shared mappings: write(2)
                                    mm = None
                                if mm is not None:
                                    dst = np.frombuffer(mm, dtype=np.uint8)[off - a0:]
                                    return [wpool.submit(np.copyto, dst[o:o + step], src[o:o + step]) for o in range(0, len(src), step)]
                            mv = memoryview(src)                                 # positional writes: order-free, several in flight
                            return [wpool.submit(os.pwrite, fd, mv[o:o + step], off + o) for o in range(0, len(mv), step)]
                        shard_off[s] += total
                        slot_futs[slot].append(writer.submit(_write))
                    if hi == int(shard_ranges[s + 1]):
                        print("%s [%s %d]: Wrote %d images to %s" % (datetime.now(), label, worker_index, shard_count[s], files[s].name))
                        sys.stdout.flush()
        finally:
            if pending_reads is not None:                                   # aborted with a read-ahead in flight: hand its
                try:                                                        # pinned staging set back
                    _, dropped = pending_reads.result()
                    if dropped is not None:
                        dropped.release()
                except Exception:
                    pass
        for fl in slot_futs:
            for x in fl:
                for y in x.result():
                    y.result()
    for f in files:
        f.close()
    print("%s [%s %d]: Wrote %d images to %d shards." % (datetime.now(), label, worker_index, counter, per))
    sys.stdout.flush()
    return counter


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
        for dev in by_dev:
            on_device(dev)
    else:
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
