#!/usr/bin/env python3
"""BASELINE.json configs[4]: nearest-to-reference-date mosaic with date / cloud filters, T=32, a 1M-chip dataset
sharded across the GPUs of one box, per-band statistics combined with ONE NCCL allreduce.

    python tools/mosaic_multi.py [--chips 1000000]            (1 GPU)
    torchrun --nproc-per-node N tools/mosaic_multi.py          (N GPUs, strong scaling: the dataset is fixed)

Chips are partitioned exactly as the reference partitions files over workers (linspace ranges).
Stacks come from a resident pool of 256 distinct synthetic stacks (chip % 256), scene metadata from one seeded
device-side table that every rank generates identically, so the statistics are bit-identical for any N.
Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import synthetic as syn  # noqa: E402
from dl_image_segmentation_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chips", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--pool", type=int, default=256)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.get_ctx(dev)
    T, H, W, B = 32, 256, 256, 4
    g = torch.Generator(device=dev)
    g.manual_seed(1005)
    stacks = torch.randint(0, 10001, (args.pool, T, H, W, B), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16)
    coarse = torch.rand((args.pool * T, 1, 16, 16), device=dev, generator=g)
    valids = (torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear") > 0.15).to(torch.uint8).reshape(args.pool, T, H, W)
    # scene metadata of ALL chips (identical on every rank), then this rank's linspace range
    N = args.chips
    day = torch.sort(torch.randint(0, 730, (N, T), dtype=torch.int32, device=dev, generator=g), dim=1).values.contiguous()
    cf = torch.rand((N, T), dtype=torch.float32, device=dev, generator=g)
    spacing = np.linspace(0, N, world + 1).astype(int)
    lo, hi = int(spacing[rank]), int(spacing[rank + 1])
    sp_all = torch.tensor([stacks[c % args.pool].data_ptr() for c in range(args.pool)], dtype=torch.int64, device=dev)
    vp_all = torch.tensor([valids[c % args.pool].data_ptr() for c in range(args.pool)], dtype=torch.int64, device=dev)
    nb = args.batch
    out = torch.empty((nb, H, W, B), dtype=torch.uint16, device=dev)
    mask = torch.empty((nb, H, W), dtype=torch.uint8, device=dev)
    nel = torch.empty((nb,), dtype=torch.int32, device=dev)
    acc = torch.zeros((B, 4), dtype=torch.int64, device=dev)
    none_count = torch.zeros((), dtype=torch.int64, device=dev)
    f = syn.CFG5_FILTER

    def run(c0, c1):
        n = c1 - c0
        idx = torch.arange(c0, c1, device=dev) % args.pool
        sp, vp = sp_all[idx], vp_all[idx]
        _lib.check(_lib.lib().b2_nearest_date_mosaic(ctx.handle, _lib.ptr(sp), _lib.ptr(vp), _lib.ptr(day[c0:c1]), _lib.ptr(cf[c0:c1]),
                                                     f["ref_day"], f["min_day"], f["max_day"], f["max_cf"], n, T, H, W, B, 2,
                                                     _lib.ptr(out), _lib.ptr(mask), None, _lib.ptr(nel), _lib.ptr(acc), ctx.stream()))
        none_count.add_((nel[:n] == 0).sum())                   # statistics are accumulated inside the mosaic kernel

    run(lo, min(hi, lo + nb))                               # warm-up
    acc.zero_()
    none_count.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for c0 in range(lo, hi, nb):
        run(c0, min(hi, c0 + nb))
    if world > 1:
        dist.all_reduce(acc)                                # the single collective of the path: exact integer counters
        dist.all_reduce(none_count)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mean, std = ops.mean_std_from_stats(ops.stats_to_python(acc))
    if rank == 0:
        dense = N * (T * H * W * B * 2 + T * H * W + H * W * B * 2 + H * W)
        print(json.dumps({"workload": "cfg5: nearest-date mosaic + filters, T=32, 256x256x4 u16, %d chips, per-band stats, 1 allreduce" % N,
                          "n_gpus": world, "scaling": "strong", "ms": round(ms, 2), "chips_per_s": round(N / ms * 1e3, 1),
                          "dense_equivalent_GB/s_per_gpu": round(dense / world / ms / 1e6, 1),
                          "chips_without_any_eligible_scene": int(none_count.item()),
                          "band_mean": [float(x) for x in mean], "band_std": [float(x) for x in std]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
