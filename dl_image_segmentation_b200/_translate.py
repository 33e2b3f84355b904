"""Shared worker loop of the two translators: chip folders -> sharded TFRecords, one GPU per worker.

This is the body of ``_process_image_files_mp_worker`` (``_img_to_tf_mp.py:78-157``) and
``_process_image_files_worker`` (``_img_to_tf_threaded.py:136-219``) with the per-chip native calls batched on
the GPU: file bytes -> [K1 decode] -> [K2 build: protobuf + framing + CRC-32C] -> shard bytes -> file.
Kept from the reference: shard sub-ranges ``np.linspace(lo, hi, S/P + 1).astype(int)`` (``:106-108``), shard name
``'%s-%.5d-of-%.5d'`` (``:115``), records in list order, skip-and-continue on any per-chip failure with the same
messages (``:133-136``), key equality check (``:132``), progress / summary prints (``:145-157``).
"""
import os
import sys
from datetime import datetime

import numpy as np
import torch

from . import _codec, ops
from ._lib import get_ctx


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


def load_pairs(img_paths, lbl_paths, store_as_array, key_fn, validate=None, device=None):
    """Read + (optionally) decode a batch of chip pairs.  Returns one entry per pair: a dict ready for
    ops.build_records, or the Exception that makes the reference skip the chip."""
    ctx = get_ctx(device)
    n = len(img_paths)
    blobs, errs = [], [None] * n
    for i in range(n):
        try:
            blobs += [_read(img_paths[i]), _read(lbl_paths[i])]
        except Exception as e:                                              # unreadable file
            errs[i] = e
            blobs += [b"", b""]
    infos = [_codec.probe(b) if b else None for b in blobs]
    arrays = [None] * (2 * n)
    if store_as_array:
        live = [k for k in range(2 * n) if blobs[k] and infos[k].status == 0]
        dec, st = _codec.decode_blobs([blobs[k] for k in live], device=ctx.device)
        for k, a, s in zip(live, dec, st):
            arrays[k] = a
            if s != 0 and errs[k // 2] is None:
                errs[k // 2] = ChipError("could not decode %s (codec status %d)" % ((img_paths, lbl_paths)[k % 2][k // 2], int(s)))
    out = []
    for i in range(n):
        if errs[i] is not None:
            out.append(errs[i])
            continue
        ii, li = infos[2 * i], infos[2 * i + 1]
        try:
            for info, p in ((ii, img_paths[i]), (li, lbl_paths[i])):
                if info is None or info.status != 0:
                    raise ChipError("'%s' not recognized as a supported file format." % p)
            if validate is not None:
                validate(ii)
                validate(li)
            ikey, lkey = key_fn(img_paths[i]), key_fn(lbl_paths[i])
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
               store_as_array, label="process", progress_every=100, validate=None, device=None, batch_pairs=64):
    num_workers = len(ranges)
    assert not num_shards % num_workers
    per = int(num_shards / num_workers)
    shard_ranges = np.linspace(ranges[worker_index][0], ranges[worker_index][1], per + 1).astype(int)
    num_files = ranges[worker_index][1] - ranges[worker_index][0]
    ctx = get_ctx(device)
    counter = 0
    for s in range(per):
        shard = worker_index * per + s
        output_file = os.path.join(output_directory, "%s-%.5d-of-%.5d" % (name, shard, num_shards))
        os.makedirs(output_directory, exist_ok=True)
        shard_counter = 0
        lo, hi = int(shard_ranges[s]), int(shard_ranges[s + 1])
        with open(output_file, "wb") as f:
            for b0 in range(lo, hi, batch_pairs):
                idx = list(range(b0, min(b0 + batch_pairs, hi)))
                pairs = load_pairs([img_filenames[i] for i in idx], [lbl_filenames[i] for i in idx], store_as_array,
                                   key_fn, validate, ctx.device)
                items = []
                for i, p in zip(idx, pairs):
                    if isinstance(p, Exception):
                        print(p)
                        print("SKIPPED: Unexpected eror while decoding %s." % img_filenames[i])
                        continue
                    items.append(p)
                    shard_counter += 1
                    counter += 1
                    if not counter % progress_every:
                        print("%s [%s %d]: Processed %d of %d images in %s batch." %
                              (datetime.now(), label, worker_index, counter, num_files, label))
                        sys.stdout.flush()
                if items:
                    buf, _, total = ops.build_records(items, ctx.device)
                    f.write(buf[:total].cpu().numpy().tobytes())
        print("%s [%s %d]: Wrote %d images to %s" % (datetime.now(), label, worker_index, shard_counter, output_file))
        sys.stdout.flush()
    print("%s [%s %d]: Wrote %d images to %d shards." % (datetime.now(), label, worker_index, counter, per))
    sys.stdout.flush()
    return counter


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
