import sys, os, io, contextlib, time, tempfile, shutil
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import translate_bench as tb
kind=sys.argv[1]; n=int(sys.argv[2])
root=tempfile.mkdtemp(prefix="b2tr_", dir="/dev/shm")
try:
    ext,_=tb.make_dataset(kind,n,root)
    import torch, dl_image_segmentation_b200 as pkg
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("warm", root, os.path.join(root,"w"), 24, num_proc=1, file_ext=ext)
    torch.cuda.synchronize()
    os.environ["B2_TRANSLATE_TRACE"]="1"
    t0=time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("b", root, os.path.join(root,"o"), 24, num_proc=1, file_ext=ext)
    torch.cuda.synchronize()
    print("total", time.perf_counter()-t0)
finally:
    shutil.rmtree(root,ignore_errors=True)
