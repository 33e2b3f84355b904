"""File discovery, seeded shuffle, range partition, shard naming (oracle; test infrastructure only).

Restates ``_img_to_tf_mp.py``: ``_find_image_files :184-230`` (glob, ``random.seed(12345)``, shuffle of
``list(range(N))`` ``:221-226``), ``_process_image_files_mp :160-181`` (``np.linspace(0,N,P+1).astype(int)``
``:167-170``), worker shard sub-ranges and names ``:102-116``.  Same logic in
``_img_to_tf_threaded.py:163-178,236-239,297-314``.  Two deliberate deviations:
``np.int`` -> ``int`` (removed from NumPy), and both glob results are ``sorted()`` before the shuffle
because ``tf.io.gfile.glob`` order is unspecified and images/labels are paired by position.
"""
import glob
import os
import random

import numpy as np


def find_image_files(data_dir, file_ext="tif", also_jpg=False):
    filenames = sorted(glob.glob("%s/images/*.%s" % (data_dir, file_ext)))
    labels = sorted(glob.glob("%s/labels/*.%s" % (data_dir, file_ext)))
    if also_jpg:                                   # threaded flavour: *.png then *.jpg (:297-304)
        filenames += sorted(glob.glob("%s/images/*.jpg" % data_dir))
        labels += sorted(glob.glob("%s/labels/*.jpg" % data_dir))
    shuffled_index = list(range(len(filenames)))
    random.seed(12345)
    random.shuffle(shuffled_index)
    return [filenames[i] for i in shuffled_index], [labels[i] for i in shuffled_index]


def worker_ranges(n_files, num_proc):
    spacing = np.linspace(0, n_files, num_proc + 1).astype(int)
    return [[int(spacing[i]), int(spacing[i + 1])] for i in range(len(spacing) - 1)]


def shard_plan(n_files, num_shards, num_proc):
    """-> list over shards of (shard_index, lo, hi) in global shard order."""
    ranges = worker_ranges(n_files, num_proc)
    assert not num_shards % len(ranges)
    per = num_shards // len(ranges)
    plan = []
    for p, (lo, hi) in enumerate(ranges):
        sr = np.linspace(lo, hi, per + 1).astype(int)
        for s in range(per):
            plan.append((p * per + s, int(sr[s]), int(sr[s + 1])))
    return plan


def shard_name(name, shard, num_shards):
    return "%s-%.5d-of-%.5d" % (name, shard, num_shards)


def tile_key(path, parse_dltile_filename=True):
    base = os.path.basename(path)
    if parse_dltile_filename:
        return ".".join(base.split(os.extsep)[:-1]).replace("#", ":")      # _img_to_tf_mp.py:61
    return base
