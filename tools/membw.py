#!/usr/bin/env python3
"""Development probe (not part of the product): sustained write-only / read-only HBM bandwidth of this GPU."""
import ctypes, json, os, subprocess, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "_build", "membw.so")
if not os.path.exists(so):
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                           os.path.join(HERE, "membw.cu"), "-o", so])
L = ctypes.CDLL(so)
dev = torch.device("cuda", 0)
N = 2 << 30
buf = torch.empty((N,), dtype=torch.uint8, device=dev)
sink = torch.zeros((4,), dtype=torch.int32, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
p = ctypes.c_void_p(buf.data_ptr())
for grid, block in ((148 * 8, 256), (148 * 16, 256), (148 * 8, 512), (148 * 32, 128)):
    for cs in (0, 1):
        ms = t(lambda: L.mb_write(p, ctypes.c_size_t(N), grid, block, cs, st))
        print(json.dumps({"probe": "write st%s" % (".cs" if cs else ""), "grid": grid, "block": block, "GB/s": round(N / ms / 1e6, 1)}))
    ms = t(lambda: L.mb_read(p, ctypes.c_size_t(N), grid, block, ctypes.c_void_p(sink.data_ptr()), st))
    print(json.dumps({"probe": "read ld.nc", "grid": grid, "block": block, "GB/s": round(N / ms / 1e6, 1)}))
for grid in (148 * 2, 148 * 4, 148 * 8):
    for chunk in (2560, 8192, 32768):
        ms = t(lambda: L.mb_bulk_write(p, ctypes.c_size_t(N), grid, chunk, st))
        print(json.dumps({"probe": "write TMA bulk store", "grid": grid, "chunk": chunk, "GB/s": round(N / ms / 1e6, 1)}))

ctr = torch.zeros((4,), dtype=torch.int32, device=dev)
for grid in (148 * 3, 148 * 4):
    for chunk, per_region in ((2560, 128), (5120, 64)):
        for depth in (1, 2, 10):
            ms = t(lambda: L.mb_bulk_pattern(p, ctypes.c_size_t(N), grid, chunk, per_region, 32, ctypes.c_void_p(ctr.data_ptr()), depth, st))
            print(json.dumps({"probe": "bulk-store pattern (label tiles)", "grid": grid, "chunk": chunk, "chunks_per_region": per_region,
                              "mode": {1: "1 in flight", 2: "2 in flight", 10: "2 in flight + smem writes + proxy fence"}[depth],
                              "GB/s": round(N / ms / 1e6, 1)}))

if os.environ.get("MB_MIX"):
    nrec = 250
    tin = torch.randint(0, 10, (nrec * 32 * 8192,), dtype=torch.uint8, device=dev)
    outs = [(torch.empty((nrec * 24 * 32768,), dtype=torch.uint8, device=dev), torch.empty((nrec * 8 * 327680,), dtype=torch.uint8, device=dev)) for _ in range(4)]
    tins = [tin.clone() for _ in range(4)]
    it = [0]
    def mix(grid):
        k = it[0] % 4; it[0] += 1
        L.mb_mix(ctypes.c_void_p(tins[k].data_ptr()), ctypes.c_void_p(outs[k][0].data_ptr()), ctypes.c_void_p(outs[k][1].data_ptr()), nrec,
                 ctypes.c_void_p(ctr.data_ptr()), grid, st)
    total = nrec * (32 * 8192 + 24 * 32768 + 8 * 327680)
    for grid in (148 * 2, 148 * 3, 148 * 4):
        ms = t(lambda: mix(grid), it=12)
        print(json.dumps({"probe": "fused-parse traffic without arithmetic (TMA tile loads + float4 stores + bulk stores)", "grid": grid,
                          "ms_per_250_records": round(ms, 4), "GB/s": round(total / ms / 1e6, 1)}))
