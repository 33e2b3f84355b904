#!/usr/bin/env python3
"""How fast can shard bytes leave a (pinned) host buffer for files on /dev/shm?  pwrite to F different files from T threads
versus copies into shared mappings.  Numbers behind the translators' write-back design."""
import json, mmap, os, sys, time, tempfile, shutil
from concurrent.futures import ThreadPoolExecutor
import numpy as np

def main():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    root = tempfile.mkdtemp(prefix="b2w_", dir=base)
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 2 << 30
    src = np.random.default_rng(0).integers(0, 255, total, dtype=np.uint8)
    try:
        for files, threads, chunk in ((1, 8, 16 << 20), (8, 8, 0), (16, 16, 0), (8, 16, 64 << 20), (24, 16, 0)):
            per = total // files
            fds = [os.open(os.path.join(root, "f%d_%d" % (files, i)), os.O_CREAT | os.O_RDWR | os.O_TRUNC) for i in range(files)]
            jobs = []
            for i in range(files):
                step = chunk or per
                for o in range(0, per, step):
                    jobs.append((fds[i], i * per + o, min(step, per - o), o))
            mv = memoryview(src)
            t0 = time.time()
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda j: os.pwrite(j[0], mv[j[1]:j[1] + j[2]], j[3]), jobs))
            dt = time.time() - t0
            print(json.dumps({"mode": "pwrite", "files": files, "threads": threads, "chunk_MB": (chunk or per) >> 20, "GB/s": round(total / dt / 1e9, 2)}), flush=True)
            for fd in fds:
                os.close(fd)
            for i in range(files):
                os.unlink(os.path.join(root, "f%d_%d" % (files, i)))
        # hybrid: few files (what a decode batch of big records touches): one pwrite thread per file for the head of the
        # piece + several threads copying the rest into a shared mapping of the same file (page faults do not take the inode lock)
        for files, mm_threads, head_frac in ((4, 3, 0.4), (4, 3, 0.5), (2, 7, 0.3), (4, 0, 1.0)):
            per = total // files
            fds = [os.open(os.path.join(root, "h%d_%d" % (files, i)), os.O_CREAT | os.O_RDWR | os.O_TRUNC) for i in range(files)]
            jobs = []
            maps = []
            mv = memoryview(src)
            for i in range(files):
                os.ftruncate(fds[i], per)
                head = int(per * head_frac) & ~4095
                jobs.append(("pw", fds[i], i * per, head, 0))
                if mm_threads:
                    mm = mmap.mmap(fds[i], per)
                    maps.append(mm)
                    dst = np.frombuffer(mm, dtype=np.uint8)
                    step = (per - head + mm_threads - 1) // mm_threads
                    for o in range(head, per, step):
                        jobs.append(("mm", dst, i * per + o, min(step, per - o), o))

            def run(j):
                if j[0] == "pw":
                    os.pwrite(j[1], mv[j[2]:j[2] + j[3]], j[4])
                else:
                    np.copyto(j[1][j[4]:j[4] + j[3]], src[j[2]:j[2] + j[3]])
            t0 = time.time()
            with ThreadPoolExecutor(16) as ex:
                list(ex.map(run, jobs))
            dt = time.time() - t0
            print(json.dumps({"mode": "hybrid pwrite head + mmap rest", "files": files, "mmap_threads_per_file": mm_threads,
                              "head_fraction": head_frac, "GB/s": round(total / dt / 1e9, 2)}), flush=True)
            jobs = None
            for mm in maps:
                try:
                    mm.close()
                except BufferError:
                    pass
            for fd in fds:
                os.close(fd)
            for i in range(files):
                os.unlink(os.path.join(root, "h%d_%d" % (files, i)))
        # shared mapping, 8 and 16 threads, one file
        for threads in (8, 16):
            fd = os.open(os.path.join(root, "m"), os.O_CREAT | os.O_RDWR | os.O_TRUNC)
            os.ftruncate(fd, total)
            mm = mmap.mmap(fd, total)
            dst = np.frombuffer(mm, dtype=np.uint8)
            step = 16 << 20
            t0 = time.time()
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda o: np.copyto(dst[o:o + step], src[o:o + step]), range(0, total, step)))
            dt = time.time() - t0
            print(json.dumps({"mode": "mmap copy", "files": 1, "threads": threads, "GB/s": round(total / dt / 1e9, 2)}), flush=True)
            del dst
            mm.close()
            os.close(fd)
            os.unlink(os.path.join(root, "m"))
    finally:
        shutil.rmtree(root, ignore_errors=True)

if __name__ == "__main__":
    main()
