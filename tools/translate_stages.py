#!/usr/bin/env python3
"""Where the wall time of images_to_tfrecords_mp goes: total seconds spent inside each stage of the worker pipeline
(file reads, decode, serialise, write-back copies, waits), next to the end-to-end rate.  The numbers behind the
translator notes in DESIGN.md (B2_SHARD_WRITE=pwrite shows the inode-lock limit of write(2) on one file).

    python tools/translate_stages.py [png|lzw] [n_pairs]
"""
import sys, os, io, contextlib, time, tempfile, shutil, json, collections
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tools'))
import translate_bench as tb
kind=sys.argv[1]; n=int(sys.argv[2])
root=tempfile.mkdtemp(prefix="b2tr_", dir="/dev/shm")
T=collections.defaultdict(float); C=collections.Counter()
def wrap(obj,name,label):
    f=getattr(obj,name)
    def g(*a,**k):
        t=time.perf_counter()
        try: return f(*a,**k)
        finally:
            T[label]+=time.perf_counter()-t; C[label]+=1
    setattr(obj,name,g)
try:
    ext,_=tb.make_dataset(kind,n,root)
    import torch, dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _translate, ops, _codec
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("warm", root, os.path.join(root,"w"), 8, num_proc=1, file_ext=ext)
    torch.cuda.synchronize()
    wrap(_translate,'load_pairs','load_pairs(main)')
    wrap(ops,'build_records','build_records(main)')
    wrap(_translate.FileBatchReader,'read','read(thread)')
    wrap(_translate.FileBatchReader,'read_into','read_into(thread)')
    wrap(_codec,'decode_enqueue','decode_enqueue(main)'); wrap(_codec.DecodeJob,'status','job.status(main)')
    wrap(_translate.BatchRecords,'__init__','BatchRecords(main)')
    wrap(_codec,'decode_planned','decode_planned(main)'); wrap(_codec,'plan_blobs','plan_blobs(thread)')
    wrap(os,'pwrite','pwrite(threads,sum)'); import numpy as _np; wrap(_np,'copyto','copyto(threads,sum)')
    wrap(torch.cuda.Event,'synchronize','event.sync(any)')
    import concurrent.futures._base as fb
    wrap(fb.Future,'result','future.result(any)')
    t0=time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pkg.images_to_tfrecords_mp("b", root, os.path.join(root,"o"), 8, num_proc=1, file_ext=ext)
        torch.cuda.synchronize()
    dt=time.perf_counter()-t0
    print(json.dumps({"kind":kind,"pairs":n,"s":round(dt,3),"pairs_per_s":round(n/dt,1)}))
    for k,v in sorted(T.items(), key=lambda x:-x[1]): print("%-28s %8.3f s  x%d" % (k,v,C[k]))
finally:
    shutil.rmtree(root,ignore_errors=True)
