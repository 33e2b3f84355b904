#!/usr/bin/env python3
"""End-to-end translate benchmark (BASELINE.json configs[0] / configs[2], scaled): chip folders on disk -> sharded
TFRecord files, through the drop-in images_to_tfrecords_mp, next to the oracle restatement of the reference's
multiprocessing CPU path (joblib over all host cores) on the same files.  Prints one JSON line per arm.

    python tools/translate_bench.py [png|lzw|jpg|png2jpg] [n_pairs] [gpu_workers]

png2jpg = PNG chips through images_to_tfrecords_mt(convert_png_to_jpg=True): decode, JPEG-encode, store the JPEG files
(GPU arm only: the oracle's JPEG encoder is a pure-Python restatement, not a timing baseline).
"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import synthetic as syn  # noqa: E402


def make_dataset(kind, n, root):
    from joblib import Parallel, delayed
    os.makedirs(os.path.join(root, "images"))
    os.makedirs(os.path.join(root, "labels"))
    ext = {"png": "png", "jpg": "jpg", "png2jpg": "png"}.get(kind, "tif")

    def one(i):
        if kind in ("png", "png2jpg"):
            img, lab, key = syn.cfg1_chip(i)
            a, b = syn.png_bytes(img), syn.png_bytes(lab)
        elif kind == "jpg":                    # quality 100, 4:2:0: what tf.image.encode_jpeg behind png_to_jpeg writes
            import cv2
            img, lab, key = syn.cfg1_chip(i)
            a = cv2.imencode(".jpg", np.ascontiguousarray(img[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 100])[1].tobytes()
            b = cv2.imencode(".jpg", lab, [cv2.IMWRITE_JPEG_QUALITY, 100])[1].tobytes()
        else:
            img, lab, key = syn.cfg3_chip(i)
            a, b = syn.tiff_bytes(img, tile=256), syn.tiff_bytes(lab, tile=256, nodata=255)
        name = key.replace(":", "#") + "." + ext
        open(os.path.join(root, "images", name), "wb").write(a)
        open(os.path.join(root, "labels", name), "wb").write(b)
        return len(a) + len(b)
    distinct = min(n, 64)                      # encode a few distinct chips, then copy them under fresh keys
    sizes = Parallel(n_jobs=os.cpu_count())(delayed(one)(i) for i in range(distinct))
    names = sorted(os.listdir(os.path.join(root, "images")))
    total = sum(sizes)
    for i in range(distinct, n):
        src = names[i % distinct]
        parts = src.rsplit(".", 1)[0].split("#")
        parts[-1] = str(100000 + i)
        dst = "#".join(parts) + "." + ext
        for sub in ("images", "labels"):
            shutil.copyfile(os.path.join(root, sub, src), os.path.join(root, sub, dst))
        total += sizes[i % distinct]
    return ext, total


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "png"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (512 if kind == "lzw" else 1536)
    if kind == "jpg":
        os.environ["B2_ORACLE_JPEG"] = "libjpeg"                # CPU arm decodes with libjpeg-turbo itself, as TF would
    shards = 8
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # GPU workers (num_proc): one host thread per GPU in this process
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    root = tempfile.mkdtemp(prefix="b2tr_", dir=base)
    try:
        t0 = time.time()
        ext, in_bytes = make_dataset(kind, n, root)
        gen_s = time.time() - t0
        import contextlib
        import io

        import torch

        import dl_image_segmentation_b200 as pkg
        out_g = os.path.join(root, "out_gpu")
        if kind == "png2jpg":
            with contextlib.redirect_stdout(io.StringIO()):
                pkg.images_to_tfrecords_mt("warm", root, os.path.join(root, "out_warm"), shards, num_threads=workers, convert_png_to_jpg=True)
                torch.cuda.synchronize()
                t0 = time.time()
                pkg.images_to_tfrecords_mt("bench", root, out_g, shards, num_threads=workers, convert_png_to_jpg=True)
                torch.cuda.synchronize()
                gpu_s = time.time() - t0
            out_bytes = sum(os.path.getsize(os.path.join(out_g, f)) for f in os.listdir(out_g))
            print(json.dumps({"arm": "b200 (%d GPU worker(s), drop-in images_to_tfrecords_mt(convert_png_to_jpg=True), files on %s)" % (workers, base),
                              "kind": kind, "pairs": n, "seconds": round(gpu_s, 3), "pairs_per_s": round(n / gpu_s, 1),
                              "input_MB": round(in_bytes / 1e6, 1), "output_MB": round(out_bytes / 1e6, 1)}), flush=True)
            return
        with contextlib.redirect_stdout(io.StringIO()):
            pkg.images_to_tfrecords_mp("warm", root, os.path.join(root, "out_warm"), shards, num_proc=workers, file_ext=ext)
            torch.cuda.synchronize()
            t0 = time.time()
            pkg.images_to_tfrecords_mp("bench", root, out_g, shards, num_proc=workers, file_ext=ext)
            torch.cuda.synchronize()
            gpu_s = time.time() - t0
        out_bytes = sum(os.path.getsize(os.path.join(out_g, f)) for f in os.listdir(out_g))
        print(json.dumps({"arm": "b200 (%d GPU worker(s) in one process, drop-in images_to_tfrecords_mp, files on %s)" % (workers, base), "kind": kind, "pairs": n,
                          "seconds": round(gpu_s, 3), "pairs_per_s": round(n / gpu_s, 1), "input_MB": round(in_bytes / 1e6, 1),
                          "output_MB": round(out_bytes / 1e6, 1), "dataset_generation_s": round(gen_s, 1)}), flush=True)
        from oracle import translate as otr
        out_c = os.path.join(root, "out_cpu")
        cores = os.cpu_count() or 1
        n_cpu = min(n, 256 if kind == "lzw" else 768)            # bounded sample for the CPU arm
        t0 = time.time()
        otr.images_to_tfrecords("bench", root, out_c, shards, num_proc=shards, file_ext=ext, n_jobs=min(cores, shards), limit=n_cpu)
        cpu_s = time.time() - t0
        print(json.dumps({"arm": "cpu restatement (joblib, %d processes of %d cores)" % (min(cores, shards), cores), "kind": kind,
                          "pairs": n_cpu, "seconds": round(cpu_s, 3), "pairs_per_s": round(n_cpu / cpu_s, 1)}), flush=True)
        # parity: the GPU shards parse back to the same records as the CPU shards of the same (limited) file list
        if n_cpu == n:
            same = all(open(os.path.join(out_g, f), "rb").read() == open(os.path.join(out_c, f), "rb").read()
                       for f in sorted(os.listdir(out_c)))
            print(json.dumps({"shards_byte_identical_gpu_vs_cpu": bool(same)}))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
