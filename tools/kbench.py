#!/usr/bin/env python3
"""Per-kernel roofline micro-benchmarks (CUDA events, inputs >> L2, >= 3 warm-ups).

    python tools/kbench.py [median] [mosaic] [parse] [build] [stats] [norm] [decode] ...

Prints one JSON line per kernel: algorithmic bytes per launch / event time vs MEASURED_PEAKS.json hbm_gbs.
Used for development and for the ncu captures under profiles/; bench.py stays the headline number.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dl_image_segmentation_b200 import _lib, ops  # noqa: E402

PEAK = 6450.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters, warmup=3):
    if os.environ.get("KB_QUICK"):          # under ncu: one warm-up + one measured launch per variant
        iters, warmup = 1, 1
    for _ in range(warmup):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


QUIET = False          # bench.py imports these benchmarks for its "other_configs" block and prints its own single line


def report(name, ms, algo_bytes, extra=None):
    gbs = algo_bytes / (ms / 1e3) / 1e9
    d = {"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(algo_bytes), "GB/s": round(gbs, 1),
         "frac_of_measured_hbm": round(gbs / PEAK, 4)}
    if extra:
        d.update(extra)
    if not QUIET:
        print(json.dumps(d), flush=True)
    return d


def bench_median(dev, n_tiles=8, iters=16):
    T, H, W, B = 16, 1024, 1024, 8
    g = torch.Generator(device=dev)
    g.manual_seed(1004)
    stacks = [torch.randint(0, 10001, (T, H, W, B), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16)
              for _ in range(n_tiles)]
    valids = [(torch.rand((T, H, W), device=dev, generator=g) > 0.4).to(torch.uint8) for _ in range(n_tiles)]
    ctx = _lib.get_ctx(dev)
    out = torch.empty((H, W, B), dtype=torch.float64, device=dev)
    mask = torch.empty((H, W, B), dtype=torch.uint8, device=dev)

    def fn(i):
        s, v = stacks[i % n_tiles], valids[i % n_tiles]
        _lib.check(_lib.lib().b2_median_composite_u16(ctx.handle, _lib.ptr(s), _lib.ptr(v), None, T, H, W, B,
                                                      _lib.ptr(out), _lib.ptr(mask), ctx.stream()))
    ms = timeit(fn, iters)
    algo = T * H * W * B * 2 + T * H * W + H * W * B * 8 + H * W * B
    return report("median_kernel<16,NV=2> cfg4 (16,1024,1024,8)", ms, algo, {"tiles_per_s": round(1e3 / ms, 1), "Gpix_per_s": round(H * W / ms / 1e6, 3)})


def bench_mosaic(dev, pool=256, chips=4096, iters=5, stats=False):
    import synthetic as syn
    T, H, W, B = 32, 256, 256, 4
    g = torch.Generator(device=dev)
    g.manual_seed(1005)
    stacks = torch.randint(0, 10001, (pool, T, H, W, B), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16)
    # spatially coherent validity (clouds are blobs): coarse noise upsampled
    coarse = torch.rand((pool * T, 1, 16, 16), device=dev, generator=g)
    valids = (torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear") > 0.33).to(torch.uint8).reshape(pool, T, H, W)
    days = np.stack([syn.cfg5_scene_meta(c)[0] for c in range(chips)])
    cfs = np.stack([syn.cfg5_scene_meta(c)[1] for c in range(chips)])
    day_d, cf_d = ops.to_device(days, dev), ops.to_device(cfs, dev)
    sp = torch.tensor([stacks[c % pool].data_ptr() for c in range(chips)], dtype=torch.int64).to(dev)
    vp = torch.tensor([valids[c % pool].data_ptr() for c in range(chips)], dtype=torch.int64).to(dev)
    ctx = _lib.get_ctx(dev)
    out = torch.empty((chips, H, W, B), dtype=torch.uint16, device=dev)
    mask = torch.empty((chips, H, W), dtype=torch.uint8, device=dev)
    src = torch.empty((chips, H, W), dtype=torch.int16, device=dev)
    nel = torch.empty((chips,), dtype=torch.int32, device=dev)
    f = syn.CFG5_FILTER
    acc = torch.zeros((B, 4), dtype=torch.int64, device=dev) if stats else None      # fused per-band statistics

    def fn(i):
        _lib.check(_lib.lib().b2_nearest_date_mosaic(ctx.handle, _lib.ptr(sp), _lib.ptr(vp), _lib.ptr(day_d), _lib.ptr(cf_d),
                                                     f["ref_day"], f["min_day"], f["max_day"], f["max_cf"], chips, T, H, W, B, 2,
                                                     _lib.ptr(out), _lib.ptr(mask), None, _lib.ptr(nel), _lib.ptr(acc), ctx.stream()))
    ms = timeit(fn, iters)
    dense = chips * (T * H * W * B * 2 + T * H * W + H * W * B * 2 + H * W)
    # bytes actually touched: one pixel read + probes of eligible scenes until the first valid one (estimated from src)
    _lib.check(_lib.lib().b2_nearest_date_mosaic(ctx.handle, _lib.ptr(sp), _lib.ptr(vp), _lib.ptr(day_d), _lib.ptr(cf_d),
                                                 f["ref_day"], f["min_day"], f["max_day"], f["max_cf"], chips, T, H, W, B, 2,
                                                 _lib.ptr(out), _lib.ptr(mask), _lib.ptr(src), _lib.ptr(nel), None, ctx.stream()))
    touched_min = chips * (H * W * B * 2 * 2 + H * W + H * W)
    return report("mosaic_kernel<8%s> cfg5 T=32 256x256x4 u16, %d chips" % (", fused band stats" if stats else "", chips), ms, dense,
           {"chips_per_s": round(chips / ms * 1e3, 1), "note": "dense-equivalent bytes; the kernel skips filtered/occluded scenes",
            "GB/s_min_touched": round(touched_min / (ms / 1e3) / 1e9, 1), "mean_eligible": float(nel.float().mean().cpu())})


def bench_decode(dev, kind, n_distinct=16, reps=None):
    reps = reps or int(os.environ.get("KB_DECODE_REPS", "16"))
    import time

    import synthetic as syn
    from dl_image_segmentation_b200 import _codec
    blobs = []
    for i in range(n_distinct):
        if kind == "lzw":
            img, lab, _ = syn.cfg3_chip(i)
            blobs += [syn.tiff_bytes(img, tile=256), syn.tiff_bytes(lab, tile=256, nodata=255)]
        elif kind == "lzw_restart":                  # what the GeoTIFF writer produces: a Clear every 1024 input bytes
            img, lab, _ = syn.cfg3_chip(i)
            blobs += [syn.tiff_bytes(img, tile=256, lzw_restart=1024), syn.tiff_bytes(lab, tile=256, nodata=255, lzw_restart=1024)]
        elif kind == "lzw_strips_pred2":
            img, lab, _ = syn.cfg3_chip(i)
            blobs += [syn.tiff_bytes(img, tile=None, predictor=2, photometric=2), syn.tiff_bytes(lab, tile=None, predictor=2)]
        elif kind == "deflate":
            img, lab, _ = syn.cfg3_chip(i)
            blobs += [syn.tiff_bytes(img, tile=256, compression="deflate"), syn.tiff_bytes(lab, tile=256, compression="deflate")]
        elif kind == "png_images":                   # the two halves of a PNG pair on their own: which one bounds the batch?
            blobs += [syn.png_bytes(syn.cfg1_chip(i)[0]), syn.png_bytes(syn.cfg1_chip(i + 100)[0])]
        elif kind == "png_labels":
            blobs += [syn.png_bytes(syn.cfg1_chip(i)[1]), syn.png_bytes(syn.cfg1_chip(i + 100)[1])]
        else:
            img, lab, _ = syn.cfg1_chip(i)
            blobs += [syn.png_bytes(img), syn.png_bytes(lab)]
    batch = blobs * reps
    best = None
    for it in range(4):
        tm = {}
        t0 = time.time()
        arrays, status = _codec.decode_blobs(batch, device=dev, timings=tm)
        torch.cuda.synchronize()
        tm["wall_ms"] = (time.time() - t0) * 1e3
        assert not status.any()
        if best is None or tm["decode_ms"] + tm["assemble_ms"] < best["decode_ms"] + best["assemble_ms"]:
            best = tm
    pairs = len(batch) // 2
    ms = best["decode_ms"] + best["assemble_ms"]
    algo = best["compressed_bytes"] + best["decoded_bytes"]
    return report("decode %s: %d chip pairs (%d streams)" % (kind, pairs, best["streams"]), ms, algo,
           {"chip_pairs_per_s": round(pairs / ms * 1e3, 1), "decode_ms": round(best["decode_ms"], 3),
            "assemble_ms": round(best["assemble_ms"], 3), "decoded_GB/s": round(best["decoded_bytes"] / ms / 1e6, 1),
            "wall_ms_incl_host_parse_and_h2d": round(best["wall_ms"], 1)})


def bench_jpeg(dev, n_distinct=16, reps=None):
    """cfg1 chips saved as .jpg (quality 100, 4:2:0 — what tf.image.encode_jpeg behind png_to_jpeg writes) + grey labels."""
    reps = reps or int(os.environ.get("KB_DECODE_REPS", "16")) * 4
    import time

    import cv2
    import synthetic as syn
    from dl_image_segmentation_b200 import _codec
    blobs = []
    for i in range(n_distinct):
        img, lab, _ = syn.cfg1_chip(i)
        for arr in (img[..., ::-1], lab.reshape(lab.shape[0], lab.shape[1])):
            ok, buf = cv2.imencode(".jpg", np.ascontiguousarray(arr), [cv2.IMWRITE_JPEG_QUALITY, 100])
            blobs.append(buf.tobytes())
    batch = blobs * reps
    best = None
    for it in range(4):
        tm = {}
        t0 = time.time()
        arrays, status, _ = _codec.decode_jpeg_blobs(batch, device=dev, timings=tm)
        torch.cuda.synchronize()
        tm["wall_ms"] = (time.time() - t0) * 1e3
        assert not status.any()
        if best is None or tm["decode_ms"] < best["decode_ms"]:
            best = tm
    pairs = len(batch) // 2
    ms = best["decode_ms"]
    return report("decode jpeg: %d chip pairs (%d files)" % (pairs, best["files"]), ms,
                  best["compressed_bytes"] + best["decoded_bytes"],
                  {"chip_pairs_per_s": round(pairs / ms * 1e3, 1), "decoded_GB/s": round(best["decoded_bytes"] / ms / 1e6, 1),
                   "compressed_MB": round(best["compressed_bytes"] / 1e6, 1),
                   "wall_ms_incl_host_parse_and_h2d": round(best["wall_ms"], 1)})


def bench_jpeg_encode(dev, n_distinct=16, reps=64):
    """convert_png_to_jpg: cfg1 chips + labels -> quality-100 JPEG files (4:2:0 / grey), 1024 pairs per call."""
    import synthetic as syn
    from dl_image_segmentation_b200 import _codec
    arrays = []
    for i in range(n_distinct):
        img, lab, _ = syn.cfg1_chip(i)
        arrays += [torch.from_numpy(img).to(dev), torch.from_numpy(lab.reshape(lab.shape[0], lab.shape[1], 1)).to(dev)]
    batch = arrays * reps
    best = None
    for it in range(3):
        tm = {}
        files = _codec.encode_jpeg_arrays(batch, quality=100, device=dev, timings=tm)
        if best is None or tm["encode_ms"] < best["encode_ms"]:
            best = tm
    pairs = len(batch) // 2
    out_bytes = sum(len(f) for f in files)
    return report("encode jpeg: %d chip pairs (%d images)" % (pairs, len(batch)), best["encode_ms"], best["pixel_bytes"] + out_bytes,
                  {"chip_pairs_per_s": round(pairs / best["encode_ms"] * 1e3, 1), "file_MB": round(out_bytes / 1e6, 1)})


def bench_parse(dev, n_shards=8):
    """Variants of the fused parse kernel on cfg2 shards + write-only / copy bandwidth references."""
    sys.path.insert(0, ROOT)
    import bench as B
    B.N_SHARDS = n_shards
    if os.environ.get("KB_RECS"):           # longer shards: separates steady-state rate from launch/tail effects
        B.RECS_PER_SHARD = int(os.environ["KB_RECS"])
    shards = B.make_shards_on_device(dev, 7)
    n = B.RECS_PER_SHARD
    tabs = [ops.open_shard_async(s, dev, max_records=n) for s in shards]
    assert all(t.check() == n for t in tabs)
    mean = ops.to_device(np.array([127.0, 128.0, 126.5], np.float32), dev)
    std = ops.to_device(np.array([73.0, 74.0, 72.5], np.float32), dev)
    out = (torch.empty((n, B.H * B.W * B.C), dtype=torch.float32, device=dev),
           torch.empty((n, B.H * B.W * B.K), dtype=torch.float32, device=dev))
    out_raw = (torch.empty((n, B.H * B.W * B.C), dtype=torch.uint8, device=dev),
               torch.empty((n, B.H * B.W), dtype=torch.uint8, device=dev))
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    rec = shards[0].numel() // n
    img_b, hot_b = B.H * B.W * B.C * 4, B.H * B.W * B.K * 4

    def fn_open(i):
        ops.open_shard_async(shards[i % n_shards], dev, max_records=n, table=tabs[i % n_shards].table)
    report("open (scan + index) one shard of %d records" % n, timeit(fn_open, 16), n * 12)
    for name, mode, verify, algo in (
            ("parse crc-only", "none", True, n * rec),
            ("parse norm+onehot no-crc", "norm_onehot", False, n * (rec + img_b + hot_b)),
            ("parse norm+onehot +crc", "norm_onehot", True, n * (rec + img_b + hot_b)),
            ("parse raw +crc", "raw", True, n * (rec + B.H * B.W * (B.C + 1)))):
        def fn(i):
            ops.parse_table(tabs[i % n_shards], mode, B.H * B.W * B.C, B.H * B.W, verify_crc=verify, mean=mean, std=std,
                            num_classes=B.K, out=out if mode == "norm_onehot" else (out_raw if mode == "raw" else None),
                            status=status)
        report(name, timeit(fn, 16), algo)
        assert not status.cpu().numpy().any()
    if os.environ.get("B2_PARSE_PROFILE"):                 # development build only: make -C csrc clean all DEV=1
        import ctypes
        assert hasattr(_lib.lib(), "b2_debug_parse_phases"), "B2_PARSE_PROFILE needs a DEV=1 build of libb2chips.so"
        _lib.lib().b2_debug_parse_phases.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        ph = (ctypes.c_uint64 * 16)()
        _lib.check(_lib.lib().b2_debug_parse_phases(_lib.get_ctx(dev).handle, ph))     # reset
        for i in range(4):
            ops.parse_table(tabs[i], "norm_onehot", B.H * B.W * B.C, B.H * B.W, verify_crc=True, mean=mean, std=std,
                            num_classes=B.K, out=out, status=status)
        _lib.check(_lib.lib().b2_debug_parse_phases(_lib.get_ctx(dev).handle, ph))
        names = ["tile wait", "crc", "image sink", "one-hot sink", "end barrier", "flush", "job fetch"]
        for label, o in (("warp 0", 0), ("warps 1-7", 8)):
            tot = float(sum(ph[o:o + 7]))
            print(json.dumps({"phase share of cycles, norm+onehot +crc, " + label: {k: round(ph[o + i] / tot, 4) for i, k in enumerate(names)},
                              "cycles": tot}))
    if os.environ.get("KB_SPLIT"):          # which sink limits the fused pass?
        for name, wi, wt, algo in (("parse image sink only, no crc", True, False, n * (rec + img_b)),
                                   ("parse one-hot sink only, no crc", False, True, n * (rec + hot_b))):
            def fn2(i):
                ops.parse_table(tabs[i % n_shards], "norm_onehot", B.H * B.W * B.C, B.H * B.W, verify_crc=False, mean=mean,
                                std=std, num_classes=B.K, out=out, status=status, want_img=wi, want_tgt=wt)
            report(name, timeit(fn2, 16), algo)
    big = torch.empty((1 << 30,), dtype=torch.uint8, device=dev)
    big2 = torch.empty((1 << 30,), dtype=torch.uint8, device=dev)
    report("reference: torch zero_ 1 GiB (write-only)", timeit(lambda i: big.zero_(), 10), 1 << 30)
    report("reference: torch copy_ 1 GiB (read+write)", timeit(lambda i: big2.copy_(big), 10), 2 << 30)


def bench_k4(dev):
    """Standalone K4 kernels (cast + normalise, one-hot, exact band statistics)."""
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    N, H, W, C, K = 1024, 256, 256, 3, 10
    imgs = [torch.randint(0, 256, (N, H, W, C), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
    labs = [torch.randint(0, K, (N, H, W), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
    mean = ops.to_device(np.array([127.0, 128.0, 126.5], np.float32), dev)
    std = ops.to_device(np.array([73.0, 74.0, 72.5], np.float32), dev)
    ctx = _lib.get_ctx(dev)
    o_img = torch.empty((N, H, W, C), dtype=torch.float32, device=dev)
    o_hot = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
    L = _lib.lib()

    def f_norm(i):
        _lib.check(L.b2_normalise_onehot(ctx.handle, _lib.ptr(imgs[i & 1]), _lib.B2_U8, None, 0, _lib.ptr(mean), _lib.ptr(std),
                                         N * H * W, C, 1, _lib.ptr(o_img), None, ctx.stream()))
    report("normalise_kernel u8 -> f32, %d chips 256x256x3" % N, timeit(f_norm, 10), N * H * W * C * 5)

    def f_hot(i):
        _lib.check(L.b2_normalise_onehot(ctx.handle, None, 0, _lib.ptr(labs[i & 1]), _lib.B2_U8, None, None, N * H * W, 1, K, None,
                                         _lib.ptr(o_hot), ctx.stream()))
    report("onehot_kernel u8 -> f32 x10, %d chips 256x256" % N, timeit(f_hot, 10), N * H * W * (1 + 4 * K))
    st16 = [torch.randint(0, 10001, (256, 512, 512, 4), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16) for _ in range(2)]
    acc = torch.zeros((4, 4), dtype=torch.int64, device=dev)

    def f_stats(i):
        _lib.check(L.b2_band_stats(ctx.handle, _lib.ptr(st16[i & 1]), _lib.B2_U16, None, 256 * 512 * 512, 4, _lib.ptr(acc), ctx.stream()))
    report("stats_kernel u16 x4 bands, 256 chips 512x512", timeit(f_stats, 10), 256 * 512 * 512 * 4 * 2)


def bench_encode(dev, n_chips=None):
    n_chips = n_chips or int(os.environ.get("KB_ENC_CHIPS", "256"))
    """GeoTIFF writer: tile split + TIFF-LZW encode of cfg3-style chips (512x512x4 u16 + 512x512 u8 labels)."""
    import time

    import synthetic as syn
    from dl_image_segmentation_b200 import _geotiff
    arrs = []
    for i in range(16):
        img, lab, _ = syn.cfg3_chip(i)
        arrs += [torch.from_numpy(img.view(np.int16)).to(dev).view(torch.uint16), torch.from_numpy(lab).to(dev)]
    batch = (arrs * ((2 * n_chips + len(arrs) - 1) // len(arrs)))[:2 * n_chips]
    nd = [None, 255] * n_chips
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        files = _geotiff.encode_geotiffs(batch, nodata=nd, device=dev)
        torch.cuda.synchronize()
        dt = (time.time() - t0) * 1e3
        best = dt if best is None or dt < best else best
    raw = sum(int(a.numel()) * a.element_size() for a in batch)
    report("encode GeoTIFF (tile split + LZW) %d chip pairs, wall incl. D2H + IFD" % n_chips, best, raw + sum(len(f) for f in files),
           {"chip_pairs_per_s": round(n_chips / best * 1e3, 1), "raw_GB/s": round(raw / best / 1e6, 2),
            "compressed_fraction": round(sum(len(f) for f in files) / raw, 3)})
    # the same batch all the way into files (what create_chips_for_tile's save step does, for many tiles at once)
    import shutil
    import tempfile
    root = tempfile.mkdtemp(prefix="b2gt_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        paths = [os.path.join(root, "%s_%d.tif" % ("lbl" if i & 1 else "img", i >> 1)) for i in range(2 * n_chips)]
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            _geotiff.write_geotiffs(batch, paths, nodata=nd, device=dev)
            dt = (time.time() - t0) * 1e3
            best = dt if best is None or dt < best else best
        out_bytes = sum(os.path.getsize(p) for p in paths)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return report("write GeoTIFF files (tile split + LZW + D2H + IFD + file writes) %d chip pairs" % n_chips, best, raw + out_bytes,
                  {"chip_pairs_per_s": round(n_chips / best * 1e3, 1), "raw_GB/s": round(raw / best / 1e6, 2)})


def bench_encode_kernel(dev):
    """lzw_encode_kernel alone (CUDA events): 256x256x4 u16 tiles of cfg3 chips, a few batch sizes."""
    import ctypes

    import synthetic as syn
    from dl_image_segmentation_b200 import _geotiff
    from dl_image_segmentation_b200._lib import check, get_ctx, lib, ptr
    ctx = get_ctx(dev)
    tiles = []
    for i in range(4):
        img, _, _ = syn.cfg3_chip(i)
        for ty in range(2):
            for tx in range(2):
                tiles.append(np.ascontiguousarray(img[ty * 256:(ty + 1) * 256, tx * 256:(tx + 1) * 256]).view(np.uint8).reshape(-1))
    tb = tiles[0].size
    only = os.environ.get("B2_KBENCH_ENC_N")                         # e.g. "64": one batch size, restart encoder only (for ncu)
    for n in ((int(only),) if only else (8, 64, 1024, 4096)):
        raw = torch.from_numpy(np.concatenate([tiles[i % len(tiles)] for i in range(n)])).to(dev)
        cap = tb * 3 // 2 + 64
        descs = np.zeros(n, _geotiff.ENC_DESC_DTYPE)
        descs["src_len"], descs["dst_cap"] = tb, cap
        descs["src_off"] = np.arange(n, dtype=np.uint64) * tb
        descs["dst_off"] = np.arange(n, dtype=np.uint64) * ((cap + 15) & ~15)
        out = torch.empty((n * ((cap + 15) & ~15),), dtype=torch.uint8, device=dev)
        out_len = torch.empty((n,), dtype=torch.int32, device=dev)
        d_dev = torch.from_numpy(descs.view(np.uint8).reshape(-1)).to(dev)

        def fn(i):
            check(lib().b2_lzw_encode(ctx.handle, ptr(raw), ptr(d_dev), n, ptr(out), ptr(out_len), ctx.stream()))
        def fn_r(i):
            check(lib().b2_lzw_encode_restart(ctx.handle, ptr(raw), descs.ctypes.data, n, 1024, ptr(out), ptr(out_len), ctx.stream()))
        ms = timeit(fn_r, 3, warmup=1)
        comp = int(out_len.cpu().numpy().view(np.uint32).astype(np.int64).sum())
        report("lzw restart-1024 encode (segment + concat kernels) %d tiles of 512 KiB" % n, ms, n * tb + comp,
               {"raw_GB/s": round(n * tb / ms / 1e6, 2), "tiles_per_s": round(n / ms * 1e3, 1),
                "cycles_per_byte_per_thread_at_1.9GHz": round(ms * 1e-3 * 1.9e9 / (n * tb / min(n * 512, 148 * 28)), 1),
                "compressed_fraction": round(comp / (n * tb), 3)})
        if only:
            continue
        ms = timeit(fn, 2, warmup=1)
        comp = int(out_len.cpu().numpy().view(np.uint32).astype(np.int64).sum())
        report("lzw_encode_kernel %d tiles of 512 KiB" % n, ms, n * tb + comp,
               {"raw_GB/s": round(n * tb / ms / 1e6, 2), "tiles_per_s": round(n / ms * 1e3, 1),
                "cycles_per_byte_per_stream_at_1.9GHz": round(ms * 1e-3 * 1.9e9 / tb / max(1.0, n / (148 * 9)), 1),
                "compressed_fraction": round(comp / (n * tb), 3)})


def bench_build(dev, n_shards=8):
    """The writer kernel on cfg1 records (uint8 arrays -> framed Examples) and cfg3 records (uint16 -> FloatList)."""
    import bench as B
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    res = []
    for name, (h, c, dt, kind, n) in {"build cfg1 256x256x3 u8 -> BytesList": (256, 3, torch.uint8, 1, 250),
                                      "build cfg3 512x512x4 u16 -> FloatList": (512, 4, torch.int16, 2, 32)}.items():
        plans = []
        for s in range(n_shards):
            hi = 256 if dt == torch.uint8 else 10000
            imgs = torch.randint(0, hi, (n, h, h, c), dtype=torch.int32, device=dev, generator=g).to(dt)
            if dt == torch.int16:
                imgs = imgs.view(torch.uint16)
            labs = torch.randint(0, 10, (n, h, h), dtype=torch.uint8, device=dev, generator=g)
            items = [dict(img=imgs[i].reshape(-1), tgt=labs[i].reshape(-1), kind=kind, h=h, w=h, c=c, th=h, tw=h,
                          identifier=(B.KEY_FMT % (s, i)).encode()) for i in range(n)]
            plans.append(ops.BuildPlan(items, dev))
        ms = timeit(lambda i: plans[i % n_shards].launch(), 16)
        in_b = n * (h * h * c * (1 if dt == torch.uint8 else 2) + h * h)
        res.append(report(name + " (%d records)" % n, ms, in_b + plans[0].total, {"records_per_s": round(n / ms * 1e3, 1)}))
        del plans
    return res


def bench_parse_encoded(dev):
    """parse_encoded_shard: whole shards of encoded-blob records (configs[2] raw-bytes records; configs[0] PNG pairs) ->
    decoded tensors; wall clock from host shard bytes, i.e. what replaces dataset.map(parse_encoded_*_proto)."""
    import time

    import synthetic as syn
    import dl_image_segmentation_b200 as pkg
    from oracle import example_proto as oep
    from oracle import tfrecord as otfr
    for name, parser, n, make in (
            ("cfg3 LZW GeoTIFF blobs, gdal_wrapped (float32)", "gdal_wrapped", 128,
             lambda i: (lambda img, lab, key: (syn.tiff_bytes(img, tile=256), syn.tiff_bytes(lab, tile=256, nodata=255), img.shape, key))(*syn.cfg3_chip(i))),
            ("cfg3 LZW GeoTIFF blobs, gdal_eager (native dtype)", "gdal_eager", 128,
             lambda i: (lambda img, lab, key: (syn.tiff_bytes(img, tile=256), syn.tiff_bytes(lab, tile=256, nodata=255), img.shape, key))(*syn.cfg3_chip(i))),
            ("cfg1 PNG blobs, rgb", "rgb", 1024,
             lambda i: (lambda img, lab, key: (syn.png_bytes(img), syn.png_bytes(lab), img.shape, key))(*syn.cfg1_chip(i)))):
        distinct = [make(i) for i in range(16)]
        recs = []
        for i in range(n):
            ib, lb, (h, w, c), key = distinct[i % 16]
            recs.append(otfr.frame(oep.convert_to_example(ib, lb, h, w, c, h, w, key + "_%d" % i).SerializeToString()))
        shard = np.frombuffer(b"".join(recs), np.uint8)
        best = None
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.time()
            out = pkg.parse_encoded_shard(shard, parser=parser)
            torch.cuda.synchronize()
            dt = (time.time() - t0) * 1e3
            best = dt if best is None or dt < best else best
        decoded = sum(int(a.numel()) * a.element_size() + int(b.numel()) * b.element_size() for a, b, _ in out)
        report("parse_encoded_shard %s: %d records" % (name, n), best, shard.size + decoded,
               {"records_per_s": round(n / best * 1e3, 1), "shard_MB": round(shard.size / 1e6, 1), "decoded_MB": round(decoded / 1e6, 1)})
        del out
        k = 6                                                          # the generator form: the next shard is staged meanwhile
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            for out in pkg.iter_parse_encoded_shards([shard] * k, parser=parser):
                pass
            torch.cuda.synchronize()
            dt = (time.time() - t0) * 1e3
            best = dt if best is None or dt < best else best
        del out
        report("iter_parse_encoded_shards %s: %d shards x %d records" % (name, k, n), best, k * (shard.size + decoded),
               {"records_per_s": round(k * n / best * 1e3, 1)})


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    which = sys.argv[1:] or ["median", "mosaic"]
    if "median" in which:
        bench_median(dev)
    if "mosaic" in which:
        bench_mosaic(dev)
    if "parse" in which:
        bench_parse(dev)
    if "build" in which:
        bench_build(dev)
    if "k4" in which:
        bench_k4(dev)
    if "encode" in which:
        bench_encode(dev)
    if "encode_kernel" in which:
        bench_encode_kernel(dev)
    if "parse_encoded" in which:
        bench_parse_encoded(dev)
    if "jpeg" in which:
        bench_jpeg(dev)
    if "jpeg_encode" in which:
        bench_jpeg_encode(dev)
    for kind in ("lzw", "lzw_restart", "lzw_strips_pred2", "deflate", "png", "png_images", "png_labels"):
        if kind in which or "decode" in which:
            bench_decode(dev, kind)


if __name__ == "__main__":
    main()
