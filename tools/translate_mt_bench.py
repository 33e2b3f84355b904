"""Wall time of images_to_tfrecords_mt on a folder of PNG chip pairs in /dev/shm (after one warm-up job): the threaded
translator's modes — file-bytes records (the default), array records, convert_png_to_jpg."""
import contextlib, io, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import translate_bench as tb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
root = tempfile.mkdtemp(prefix="b2mt_", dir="/dev/shm")
try:
    tb.make_dataset("png", n, root)
    import torch
    import dl_image_segmentation_b200 as pkg
    for name, kw in (("file-bytes records (default)", {}), ("array records", dict(store_as_array=True)),
                     ("convert_png_to_jpg, file-bytes records", dict(convert_png_to_jpg=True)),
                     ("convert_png_to_jpg, array records", dict(convert_png_to_jpg=True, store_as_array=True))):
        best = None
        for rep in range(3):
            out = os.path.join(root, "o%d" % rep)
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                pkg.images_to_tfrecords_mt("b", root, out, 24, num_threads=1, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            shutil.rmtree(out, ignore_errors=True)
            if rep and (best is None or dt < best):
                best = dt
        print("images_to_tfrecords_mt %s: %d pairs in %.3f s = %.0f pairs/s" % (name, n, best, n / best))
finally:
    shutil.rmtree(root, ignore_errors=True)
