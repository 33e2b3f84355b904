"""CPU tests (no GPU, no compute calls): the C-ABI library loads and exports every symbol include/b2chips.h
declares; host-side entry points (Example layout, header probe) agree with the oracle; the product fails loudly
without a device; partition helpers of the shim equal the oracle's; 2-rank gloo run of the sharded logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import synthetic as syn
from oracle import example_proto as oep
from oracle import imagecodecs as oic
from oracle import partition as opart

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dl_image_segmentation_b200", "libb2chips.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dl_image_segmentation_b200", "csrc"), "-j8"])
    from dl_image_segmentation_b200 import _codec, _geotiff, _lib  # noqa: F401  (register the codec / writer signatures)
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "b2chips.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    raw = ctypes.CDLL(LIB)
    for name in declared:
        assert hasattr(raw, name), "libb2chips.so does not export %s" % name
    from dl_image_segmentation_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes binding covers the whole header, nothing else
    assert lib.b2_version() == 100


def test_struct_layouts_match_the_header(lib):
    from dl_image_segmentation_b200 import _codec, _lib
    assert ctypes.sizeof(_lib.ExampleIndex) == 80 == np.dtype(_lib.EXAMPLE_INDEX_DTYPE).itemsize
    assert ctypes.sizeof(_lib.BuildDesc) == 80 == np.dtype(_lib.BUILD_DESC_DTYPE).itemsize
    assert ctypes.sizeof(_lib.ParseSink) == 64
    assert ctypes.sizeof(_codec.ImageInfo) == 88 + 6 * 8 + 8 + 8
    assert _codec.STREAM_DESC_DTYPE.itemsize == 32 and _codec.IMAGE_DESC_DTYPE.itemsize == 72 + 16


def test_example_layout_matches_oracle_bytes(lib):
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(0)
    cases = [(1, (24, 24, 3), np.uint8, "256:2:1.0:43:0:0"), (2, (20, 20, 4), np.uint16, "448:32:10.0:43:-3:77"),
             (1, (1, 1, 1), np.uint8, ""), (2, (3, 5, 2), np.int16, "k" * 300), (1, (300, 7, 3), np.uint8, "é:ü")]
    for kind, (h, w, c), dt, key in cases:
        img = rng.integers(0, 200, (h, w, c)).astype(dt)
        lab = rng.integers(0, 10, (h, w)).astype(np.uint8)
        want = oep.convert_to_example(img, lab, h, w, c, h, w, key).SerializeToString()
        ib = img.size * (1 if kind == 1 else 4)
        tb = lab.size * (1 if kind == 1 else 4)
        sc, pl, el = ops.example_layout(kind, ib, tb, h, w, c, h, w, key.encode("utf-8"))
        assert el == len(want)
        ip = img.tobytes() if kind == 1 else img.astype("<f4").tobytes()
        tp = lab.tobytes() if kind == 1 else lab.astype("<f4").tobytes()
        got = sc[:pl[0]] + ip + sc[pl[0]:pl[0] + pl[1]] + tp + sc[pl[0] + pl[1]:]
        assert got == want
    # negative / huge dims are still well-formed varints
    sc, pl, el = ops.example_layout(1, 0, 0, -1, 1 << 40, 0, 0, 0, b"x")
    f = oep.parse_example(sc)
    assert f["image/height"] == ("int64", [-1]) and f["image/width"] == ("int64", [1 << 40])
    assert f["image/image_data"] == ("bytes", [b""])


def test_header_probe_matches_oracle(lib):
    from dl_image_segmentation_b200 import _codec
    img, lab, _ = syn.cfg3_chip(2, size=100)
    files = [syn.tiff_bytes(img, tile=64), syn.tiff_bytes(lab, tile=None, predictor=2, nodata=255),
             syn.tiff_bytes(img, tile=32, planar=2, big_endian=True, compression="deflate"),
             syn.png_bytes(syn.cfg1_chip(0, size=50)[0]), syn.png_bytes(lab)]
    for b in files:
        info = _codec.probe(b)
        assert info.status == 0
        assert (info.height, info.width, info.samples) == oic.image_shape(b)
    t = _codec.probe(files[2])
    assert (t.planar, t.big_endian, t.compression, t.block_w, t.n_blocks) == (2, 1, 8, 32, 4 * 4 * 4)
    assert _codec.probe(files[1]).has_nodata == 1 and _codec.probe(files[1]).nodata == 255.0
    for bad in (b"", b"II*\0", b"\x89PNG\r\n\x1a\n", files[0][:40], os.urandom(64)):
        assert _codec.probe(bad).status != 0
    # BigTIFF and 1-bit images are out of scope, not crashes
    assert _codec.probe(b"II+\0" + bytes(60)).status == 3


def test_no_cpu_fallback(lib):
    import torch

    import dl_image_segmentation_b200 as pkg
    from dl_image_segmentation_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.B2Error):
        pkg.median_composite(np.zeros((2, 4, 4, 2), np.uint16), np.ones((2, 4, 4), np.uint8))
    with pytest.raises(_lib.B2Error):
        pkg.convert_to_example(np.zeros((2, 2, 3), np.uint8), np.zeros((2, 2), np.uint8), 2, 2, 3, 2, 2, "k").SerializeToString()
    src = open(os.path.join(ROOT, "dl_image_segmentation_b200", "ops.py")).read()
    for mod in os.listdir(os.path.join(ROOT, "dl_image_segmentation_b200")):
        if mod.endswith(".py"):
            text = open(os.path.join(ROOT, "dl_image_segmentation_b200", mod)).read()
            assert "import oracle" not in text and "from oracle" not in text, mod     # the product never touches the oracle
    assert "oracle" not in src


def test_shim_partition_equals_oracle(tmp_path):
    from dl_image_segmentation_b200 import _img_to_tf_mp, _img_to_tf_threaded, _translate
    for n, p in ((6000, 12), (1024, 16), (5795, 7), (3, 3), (10, 1)):
        assert _translate.worker_ranges(n, p) == opart.worker_ranges(n, p)
    assert _translate.tile_key_from_path("a/b/60#2#10.0#43#-380#3491.tif") == "60:2:10.0:43:-380:3491"
    assert _translate.tile_key_from_path("a/b/x#y.tar.png", False) == "x#y.tar.png"
    for sub in ("images", "labels"):
        os.makedirs(tmp_path / sub)
        for k in range(9):
            (tmp_path / sub / ("1#2#%d.png" % k)).write_bytes(b"x")
        (tmp_path / sub / "z.jpg").write_bytes(b"x")
    a, b = _img_to_tf_mp._find_image_files(str(tmp_path), "png")
    oa, ob = opart.find_image_files(str(tmp_path), "png")
    assert a == oa and b == ob
    a, b = _img_to_tf_threaded._find_image_files(str(tmp_path))
    oa, ob = opart.find_image_files(str(tmp_path), "png", also_jpg=True)
    assert a == oa and b == ob and len(a) == 10
    os.environ.pop("WORLD_SIZE", None)
    assert [p for p, _ in _translate.my_workers(4)] == [0, 1, 2, 3]


_GLOO_WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from dl_image_segmentation_b200 import _translate, ops
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = [p for p, _ in _translate.my_workers(4)]
# every rank accumulates exact integer band statistics of ITS chips; one allreduce(sum); identical mean/std
rng = np.random.default_rng(0)
data = rng.integers(0, 65536, (8, 16, 16, 4)).astype(np.uint16)          # same on every rank
ranges = _translate.worker_ranges(8, 4)
acc = np.zeros((4, 4), np.int64)
for p in mine:
    for i in range(*ranges[p]):
        x = data[i].reshape(-1, 4).astype(np.uint64)
        acc[:, 0] += x.shape[0]
        acc[:, 1] += x.sum(0).astype(np.int64)
        q = (x * x).sum(0)
        acc[:, 2] += (q & 0xFFFF).astype(np.int64)
        acc[:, 3] += (q >> 16).astype(np.int64)
t = torch.from_numpy(acc)
dist.all_reduce(t)
mean, std = ops.mean_std_from_stats(ops.stats_to_python(t))
json.dump({"mine": mine, "mean": mean.tolist(), "std": std.tolist()}, open(sys.argv[1] + "/r%%d.json" %% rank, "w"))
dist.destroy_process_group()
"""


def test_two_rank_gloo_partition_and_stats(tmp_path):
    """world_size-2 gloo run of the N>1 host logic: worker ownership p %% world == rank, and the single
    statistics allreduce gives bit-identical mean/std to the 1-rank result."""
    import json
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571", CUDA_VISIBLE_DEVICES="")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29571", str(script), str(tmp_path)],
                          env=env, timeout=300, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r0, r1 = (json.load(open(tmp_path / ("r%d.json" % r))) for r in (0, 1))
    assert r0["mine"] == [0, 2] and r1["mine"] == [1, 3]
    assert r0["mean"] == r1["mean"] and r0["std"] == r1["std"]
    from oracle import normalise as onorm
    rng = np.random.default_rng(0)
    data = rng.integers(0, 65536, (8, 16, 16, 4)).astype(np.uint16)
    m, s = onorm.mean_std_from_stats(onorm.band_stats(data))
    assert r0["mean"] == m.tolist() and r0["std"] == s.tolist()


def test_batch_decode_planner_matches_per_file_calls(lib):
    """b2_decode_plan_batch (one native call per batch) == b2_image_probe + b2_image_blocks per file; the compressed bytes
    land in the staging buffer where the stream table says; unreadable files are marked, their streams inert."""
    from dl_image_segmentation_b200 import _codec
    G = os.path.join(ROOT, "tests", "golden")
    names = ["gdalstyle_tiled_lzw_u16x4.tif", "libpng_rgb.png", "libtiff_cv2_lzw_u16x4.tif", "libpng_label.png",
             "libtiff_pil_deflate_u8.tif", "gdalstyle_tiled_lzw_label.tif"]
    blobs = [open(os.path.join(G, f), "rb").read() for f in names] + [b"not an image at all", b""]
    n = len(blobs)
    ptrs = (ctypes.c_void_p * n)()
    sizes = np.zeros(n, np.uint64)
    for i, b in enumerate(blobs):
        ptrs[i] = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value
        sizes[i] = len(b)
    infos = (_codec.ImageInfo * n)()
    status = np.zeros(n, np.int32)
    images = np.zeros(n, _codec.IMAGE_DESC_DTYPE)
    plan = _codec.DecodePlan()
    assert lib.b2_decode_plan_batch(ptrs, sizes.ctypes.data, n, infos, status.ctypes.data, images.ctypes.data, None, 0, None, 0,
                                    2, 0, ctypes.byref(plan)) == 0
    assert plan.filled == 0 and plan.n_streams > 0
    streams = np.zeros(plan.n_streams, _codec.STREAM_DESC_DTYPE)
    stage = np.zeros(plan.stage_bytes, np.uint8)
    assert lib.b2_decode_plan_batch(ptrs, sizes.ctypes.data, n, infos, status.ctypes.data, images.ctypes.data,
                                    streams.ctypes.data, len(streams), stage.ctypes.data, stage.size, 3, 0, ctypes.byref(plan)) == 0
    assert plan.filled == 1
    assert list(status[-2:] != 0) == [True, True] and not status[:-2].any()
    k = 0
    for i, b in enumerate(blobs[:-2]):
        info = _codec.probe(b)
        assert (info.width, info.height, info.samples, info.n_blocks) == (infos[i].width, infos[i].height, infos[i].samples, infos[i].n_blocks)
        nb = info.n_blocks
        offs, cnts, dlen = (np.zeros(nb, np.uint64) for _ in range(3))
        a = np.frombuffer(b, np.uint8)
        assert lib.b2_image_blocks(a.ctypes.data, a.size, ctypes.byref(info), offs.ctypes.data, cnts.ctypes.data, dlen.ctypes.data, nb) == 0
        if info.format == 2:
            sd = streams[k]
            want = b"".join(b[int(o):int(o + c)] for o, c in zip(offs, cnts))
            assert bytes(stage[int(sd["src_off"]):int(sd["src_off"]) + int(sd["src_len"])]) == want
            assert (int(sd["codec"]), int(sd["image"]), int(sd["dst_len"])) == (8, i, int(info.block_bytes))
            k += 1
        else:
            for j in range(nb):
                sd = streams[k + j]
                assert bytes(stage[int(sd["src_off"]):int(sd["src_off"]) + int(sd["src_len"])]) == b[int(offs[j]):int(offs[j] + cnts[j])]
                assert int(sd["dst_len"]) == int(dlen[j]) and int(sd["image"]) == i
                assert int(sd["dst_off"]) == int(images[i]["scratch_off"]) + j * int(info.block_bytes)
            k += nb
    assert k == plan.n_streams


def test_georeferenced_identifier_strings(lib, tmp_path):
    """dltile_from_filename=False: identifier = basename | str(geotransform) | str(crs) (reference _img_to_tf_mp.py:49-50,
    63-67), read from the GeoTIFF tags (ModelPixelScale / ModelTiepoint / GeoKeyDirectory); GDAL defaults for a PNG."""
    from dl_image_segmentation_b200 import _codec
    img = np.arange(16 * 16 * 2, dtype=np.uint16).reshape(16, 16, 2)
    info = _codec.probe(syn.tiff_bytes(img, tile=16))
    assert info.has_geo == 1 and info.epsg == 32643
    assert _codec.georef_strings(info) == ("[499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0]", "EPSG:32643")
    info = _codec.probe(syn.tiff_bytes(img, tile=16, geo=False))
    assert info.has_geo == 0 and _codec.georef_strings(info) == ("[0.0, 1.0, 0.0, 0.0, 0.0, 1.0]", "None")
    info = _codec.probe(syn.png_bytes(np.zeros((4, 4, 3), np.uint8)))
    assert _codec.georef_strings(info) == ("[0.0, 1.0, 0.0, 0.0, 0.0, 1.0]", "None")
    big = _codec.probe(syn.tiff_bytes(img, tile=16, big_endian=True))
    assert _codec.georef_strings(big) == ("[499980.0, 10.0, 0.0, 5300040.0, 0.0, -10.0]", "EPSG:32643")


def test_native_batch_file_reader(tmp_path):
    """b2_read_files (one native multi-threaded call per batch) == open(p,'rb').read() per file; unreadable files come
    back as the OSError the reference's except branch would have caught; buffers are reused between batches."""
    from dl_image_segmentation_b200 import _translate
    rng = np.random.default_rng(3)
    paths, want = [], []
    for i, n in enumerate([0, 1, 15, 16, 17, 4096, 100003, 7]):
        p = tmp_path / ("f%d.bin" % i)
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        p.write_bytes(data)
        paths.append(str(p))
        want.append(data)
    (tmp_path / "adir").mkdir()
    paths += [str(tmp_path / "missing.bin"), str(tmp_path / "adir")]
    rd = _translate.FileBatchReader(depth=2, threads=3)
    for rep in range(3):                                        # the third batch reuses the first buffer
        got = rd.read(paths)
        for g, w in zip(got[:len(want)], want):
            assert isinstance(g, np.ndarray) and g.tobytes() == w
        assert isinstance(got[-2], FileNotFoundError) and isinstance(got[-1], IsADirectoryError)
        assert got[-2].filename == paths[-2]
    assert rd.read([]) == []


def test_batch_schedule_covers_every_file_once_in_shard_order():
    """The translators' decode batches take pairs from up to 8 shards at a time (so that each batch's records go to that
    many files at once): whatever the shard sizes and batch size, every file index appears exactly once, each shard's
    files in order, and no batch is empty."""
    import random
    from dl_image_segmentation_b200._translate import batch_schedule
    rng = random.Random(7)
    for _ in range(200):
        per = rng.choice([1, 2, 3, 5, 8, 9, 24])
        n, lo = rng.randint(0, 3000), rng.randint(0, 50)
        shard_ranges = np.linspace(lo, lo + n, per + 1).astype(int)
        batches = batch_schedule(shard_ranges, rng.choice([1, 8, 32, 100, 227, 1024]))
        seen = {s: int(shard_ranges[s]) for s in range(per)}
        for runs in batches:
            assert runs and [r[0] for r in runs] == sorted({r[0] for r in runs})
            assert len(runs) <= 8
            for s, a, b in runs:
                assert a == seen[s] and a < b <= int(shard_ranges[s + 1])
                seen[s] = b
        assert all(seen[s] == int(shard_ranges[s + 1]) for s in range(per))
