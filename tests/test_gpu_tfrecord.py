"""-m gpu: K2 (CRC-32C, frame scan, Example index, fused parse, record build) and K4 through the C ABI vs the oracle."""
import numpy as np
import pytest

import synthetic as syn
from oracle import example_proto as oep
from oracle import normalise as onorm
from oracle import tfrecord as otfr

pytestmark = pytest.mark.gpu


def _shard(n, size=64, seed0=0, float_mode=False):
    recs, chips = [], []
    for i in range(n):
        if float_mode:
            img, lab, key = syn.cfg3_chip(seed0 + i, size=size)
        else:
            img, lab, key = syn.cfg1_chip(seed0 + i, size=size)
        chips.append((img, lab, key))
        h, w, c = img.shape
        recs.append(oep.convert_to_example(img, lab, h, w, c, h, w, key).SerializeToString())
    return b"".join(otfr.frame(r) for r in recs), recs, chips


def test_crc32c_known_answers_and_random(dev):
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(0)
    msgs = [b"123456789", bytes(32), b"\xff" * 32, bytes(range(32)), bytes(range(31, -1, -1)), b"", b"a", b"ab", b"abc",
            b"abcd", b"abcde"]
    want = [0xE3069283, 0x8A9136AA, 0x62A8AB43, 0x46DD794E, 0x113FDB5C]
    for n in (15, 16, 17, 4095, 4096, 4097, 8191, 8192, 8193, 8208, 16384, 16400, 100000, 262381, 1 << 20):
        msgs.append(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
    # pack at odd offsets so every alignment of start and end is exercised
    buf, offs, lens = bytearray(), [], []
    for i, m in enumerate(msgs):
        buf += bytes((i * 7) % 13 + 1)
        offs.append(len(buf))
        lens.append(len(m))
        buf += m
    got = ops.crc32c(bytes(buf), offs, lens, device=dev)
    for i, m in enumerate(msgs):
        assert int(got[i]) == otfr.crc32c(m), (i, len(m))
    assert [int(x) for x in got[:5]] == want


def test_scan_uniform_ragged_and_corrupt(dev):
    from dl_image_segmentation_b200 import ops
    shard, recs, _ = _shard(9)
    si = ops.open_shard(shard, dev, with_index=False)
    o, l = otfr.scan(shard)
    assert si.n == 9
    np.testing.assert_array_equal(si.rec_off[:9].cpu().numpy().astype(np.uint64), o)
    np.testing.assert_array_equal(si.rec_len[:9].cpu().numpy().astype(np.uint64), l)
    # ragged record lengths force the sequential walk
    rag = b"".join(otfr.frame(bytes(range(k % 251)) * (k % 7 + 1)) for k in range(1, 40)) + otfr.frame(b"")
    si = ops.open_shard(rag, dev, with_index=False)
    o, l = otfr.scan(rag)
    assert si.n == len(o)
    np.testing.assert_array_equal(si.rec_off[:si.n].cpu().numpy().astype(np.uint64), o)
    np.testing.assert_array_equal(si.rec_len[:si.n].cpu().numpy().astype(np.uint64), l)
    assert ops.open_shard(b"\0" * 0 or np.zeros(0, np.uint8), dev, with_index=False).n == 0
    # corrupt a length CRC, and truncate
    bad = bytearray(shard)
    bad[len(otfr.frame(recs[0])) + 9] ^= 1
    with pytest.raises(ops.DataLossError):
        ops.open_shard(bytes(bad), dev, with_index=False)
    with pytest.raises(ops.DataLossError):
        ops.open_shard(shard[:-3], dev, with_index=False)
    with pytest.raises(otfr.DataLossError):
        otfr.scan(shard[:-3])


def test_index_any_key_order_and_errors(dev):
    from dl_image_segmentation_b200 import ops
    img, lab, key = syn.cfg1_chip(1, size=32)
    ex = oep.convert_to_example(img, lab, 32, 32, 3, 32, 32, key)
    orders = [sorted(ex.features), sorted(ex.features, reverse=True), list(oep.KEYS)]
    recs = []
    for order in orders:
        e2 = oep.Example({k: ex.features[k] for k in order})
        recs.append(e2.SerializeToString(deterministic=False))
    # unknown extra feature, duplicate key (last wins), missing key, wrong type, truncated protobuf
    extra = dict(ex.features)
    extra["zzz/other"] = oep.Feature("float", [1.0, 2.0])
    recs.append(oep.Example(extra).SerializeToString())
    missing = {k: v for k, v in ex.features.items() if k != "target/width"}
    recs.append(oep.Example(missing).SerializeToString())
    wrong = dict(ex.features)
    wrong["image/height"] = oep.Feature("float", [32.0])
    recs.append(oep.Example(wrong).SerializeToString())
    recs.append(recs[0][:len(recs[0]) // 2])
    shard = b"".join(otfr.frame(r) for r in recs)
    si = ops.open_shard(shard, dev)
    st = si.index["status"]
    assert list(st) == [0, 0, 0, 0, 2, 2, 1]
    for k in range(4):
        r = si.index[k]
        assert (r["height"], r["width"], r["channels"], r["tgt_height"], r["tgt_width"]) == (32, 32, 3, 32, 32)
        assert r["img_kind"] == 1 and r["img_len"] == 32 * 32 * 3 and r["tgt_len"] == 32 * 32
        o = int(r["img_off"])
        assert shard[o:o + 16] == img.tobytes()[:16]
        assert shard[int(r["id_off"]):int(r["id_off"]) + int(r["id_len"])] == key.encode()


@pytest.mark.parametrize("size,n", [(64, 7), (61, 5), (256, 3)])
def test_parse_raw_8bit_and_crc(dev, size, n):
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from dl_image_segmentation_b200 import ops
    shard, recs, chips = _shard(n, size=size)
    si = ops.open_shard(shard, dev)
    ib, tb, idx = tr.parse_records_raw(si, 1, verify_crc=True)
    for (img, tgt), (wi, wl, _) in zip(tr.rows_to_arrays_8bit(ib, tb, idx), chips):
        np.testing.assert_array_equal(img.cpu().numpy(), wi)
        np.testing.assert_array_equal(tgt.cpu().numpy(), wl)
    assert si.identifiers() == [c[2].encode() for c in chips]
    # flip one payload byte of record 1: CRC must catch it, the others stay fine
    bad = bytearray(shard)
    bad[int(si.index[1]["img_off"]) + 100] ^= 0x40
    si2 = ops.open_shard(bytes(bad), dev)
    _, _, st = ops.parse_shard(si2, "raw", verify_crc=True)
    assert list(st.cpu().numpy()) == [0, 1] + [0] * (n - 2)
    with pytest.raises(ops.DataLossError):
        tr.parse_records_raw(si2, 1, verify_crc=True)


def test_parse_float_records(dev):
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from dl_image_segmentation_b200 import ops
    shard, recs, chips = _shard(3, size=96, float_mode=True)
    si = ops.open_shard(shard, dev)
    assert (si.index["img_kind"] == 2).all()
    ib, tb, idx = tr.parse_records_raw(si, 2, verify_crc=True)
    for (img, tgt), (wi, wl, _), rec in zip(tr.rows_to_arrays_f32(ib, tb, idx), chips, recs):
        oi, ot, _ = oep.parse_higher_dtype_array_proto(rec)
        np.testing.assert_array_equal(img.cpu().numpy(), oi)
        np.testing.assert_array_equal(tgt.cpu().numpy(), ot)
    with pytest.raises(tr.InvalidArgumentError):
        tr.parse_records_raw(si, 1, verify_crc=False)       # bytes template on float records


@pytest.mark.parametrize("size,K", [(64, 10), (61, 10), (32, 3), (40, 7), (256, 10)])
def test_parse_norm_onehot(dev, size, K):
    from dl_image_segmentation_b200 import ops
    n = 4
    shard, recs, chips = _shard(n, size=size, seed0=10)
    si = ops.open_shard(shard, dev)
    mean = np.array([101.5, 99.25, 120.0], np.float32)
    std = np.array([47.0, 51.5, 33.3], np.float32)
    ib, tb, st = ops.parse_shard(si, "norm_onehot", mean=mean, std=std, num_classes=K)
    assert not st.cpu().numpy().any()
    wi, wt = onorm.normalise(np.stack([c[0] for c in chips]), mean, std), onorm.one_hot(np.stack([c[1] for c in chips]), K)
    il, tl = size * size * 3, size * size
    got_i = ib.cpu().numpy()[:, :il].reshape(wi.shape)
    got_t = tb.cpu().numpy()[:, :tl * K].reshape(wt.shape)
    np.testing.assert_allclose(got_i, wi, rtol=1e-6, atol=0)     # north-star tolerance; in practice bit-exact
    np.testing.assert_array_equal(got_i, wi)
    np.testing.assert_array_equal(got_t, wt)


def test_single_example_parsers_and_convert(dev):
    import dl_image_segmentation_b200 as pkg
    img, lab, key = syn.cfg1_chip(5, size=48)
    ex = pkg.convert_to_example(img, lab, 48, 48, 3, 48, 48, key)
    s = ex.SerializeToString()
    assert s == oep.convert_to_example(img, lab, 48, 48, 3, 48, 48, key).SerializeToString()
    gi, gt, ident = pkg.parse_8bit_array_proto(s)
    np.testing.assert_array_equal(gi.cpu().numpy(), img)
    np.testing.assert_array_equal(gt.cpu().numpy(), lab)
    assert ident == key.encode()
    # uint16 image forces FloatList for BOTH payloads (reference :184-197)
    img16, lab16, key16 = syn.cfg3_chip(2, size=40)
    s16 = pkg.convert_to_example(img16, lab16, 40, 40, 4, 40, 40, key16).SerializeToString()
    assert s16 == oep.convert_to_example(img16, lab16, 40, 40, 4, 40, 40, key16).SerializeToString()
    gi, gt, ident = pkg.parse_higher_dtype_array_proto(s16)
    np.testing.assert_array_equal(gi.cpu().numpy(), img16.astype(np.float32))
    np.testing.assert_array_equal(gt.cpu().numpy(), lab16.astype(np.float32))
    # raw bytes payloads are stored verbatim
    blob_i, blob_t = b"\x89PNG fake image bytes" * 37, b"label blob" * 11
    sb = pkg.convert_to_example(blob_i, blob_t, 7, 8, 3, 7, 8, "k").SerializeToString()
    assert sb == oep.convert_to_example(blob_i, blob_t, 7, 8, 3, 7, 8, "k").SerializeToString()
    with pytest.raises(pkg._tfrecord_image_translation.InvalidArgumentError):
        pkg.parse_8bit_array_proto(s16)


@pytest.mark.parametrize("float_mode", [False, True])
def test_build_records_byte_exact(dev, float_mode):
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from dl_image_segmentation_b200 import ops
    sizes = [33, 64, 50, 17, 128]
    items, want = [], b""
    for i, sz in enumerate(sizes):
        img, lab, key = (syn.cfg3_chip if float_mode else syn.cfg1_chip)(i, size=sz)
        h, w, c = img.shape
        items.append(tr.convert_to_example(img, lab, h, w, c, h, w, key + "x" * i).build_item(dev))
        want += otfr.frame(oep.convert_to_example(img, lab, h, w, c, h, w, key + "x" * i).SerializeToString())
    buf, offs, total = ops.build_records(items, dev)
    got = bytes(buf[:total].cpu().numpy())
    assert total == len(want)
    assert got == want
    assert len(otfr.read_records(got, verify=True)) == len(sizes)       # oracle reader accepts the CRCs


def test_build_known_frames(dev):
    """SURVEY.md App. B frame vectors via degenerate Examples is not possible (payload is protobuf), so pin
    the frame maths through round trip: GPU-built shard -> GPU scan + CRC verify -> identical payloads."""
    from dl_image_segmentation_b200 import _tfrecord_image_translation as tr
    from dl_image_segmentation_b200 import ops
    items = []
    for i in range(12):
        img, lab, key = syn.cfg1_chip(100 + i, size=16 + 3 * i)
        h, w, c = img.shape
        items.append(tr.convert_to_example(img, lab, h, w, c, h, w, key).build_item(dev))
    buf, offs, total = ops.build_records(items, dev)
    si = ops.open_shard(buf[:total].clone(), dev)
    assert si.n == 12
    _, _, st = ops.parse_shard(si, "none", verify_crc=True)
    assert not st.cpu().numpy().any()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_normalise_onehot_standalone(dev, dtype):
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(2)
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    img = (rng.integers(info.min, info.max + 1, (3, 37, 41, 5)).astype(dtype) if info is not None
           else rng.normal(0, 1000, (3, 37, 41, 5)).astype(dtype))
    lab = rng.integers(0, 12, (3, 37, 41)).astype(np.uint8)
    lab[0, :5] = 255
    mean = rng.normal(0, 100, 5).astype(np.float32)
    std = (rng.random(5) * 100 + 1).astype(np.float32)
    gi, gh = ops.normalise_onehot(img, lab, mean, std, 10, device=dev)
    np.testing.assert_array_equal(gi.cpu().numpy(), onorm.normalise(img, mean, std))
    np.testing.assert_array_equal(gh.cpu().numpy(), onorm.one_hot(lab, 10))
    _, gh2 = ops.normalise_onehot(None, lab.astype(np.float32), None, None, 10, device=dev)
    np.testing.assert_array_equal(gh2.cpu().numpy(), onorm.one_hot(lab, 10))


@pytest.mark.parametrize("K,C,n", [(1, 1, 1000), (3, 3, 257), (10, 3, 700001), (24, 16, 5000), (25, 17, 5000), (40, 2, 3333)])
def test_standalone_k4_kernel_variants(dev, K, C, n):
    """The shared-memory / bulk-store one-hot (K <= 24) and the tabulated u8 normalise (C <= 16) against their per-element
    fallbacks' definition (the oracle), over ragged sizes and more blocks than resident CTAs (double-buffer reuse)."""
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(K * 100 + C)
    img = rng.integers(0, 256, (n, C), dtype=np.uint8)
    lab = rng.integers(0, K + 3, n).astype(np.uint8)
    lab[::97] = 255
    mean = rng.normal(100, 50, C).astype(np.float32)
    std = (rng.random(C) * 80 + 0.5).astype(np.float32)
    gi, gh = ops.normalise_onehot(img, lab, mean, std, K, device=dev)
    np.testing.assert_array_equal(gi.cpu().numpy(), onorm.normalise(img, mean, std))
    np.testing.assert_array_equal(gh.cpu().numpy(), onorm.one_hot(lab, K))
    _, gh2 = ops.normalise_onehot(None, lab.astype(np.float32), None, None, K, device=dev)   # float labels (array records)
    np.testing.assert_array_equal(gh2.cpu().numpy(), onorm.one_hot(lab, K))


@pytest.mark.parametrize("dtype,B", [(np.uint8, 3), (np.uint16, 4), (np.uint16, 8), (np.uint16, 13)])
def test_band_stats_exact(dev, dtype, B):
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(4)
    img = rng.integers(0, np.iinfo(dtype).max + 1, (5, 64, 50, B)).astype(dtype)
    valid = (rng.random((5, 64, 50)) > 0.2).astype(np.uint8)
    acc = ops.band_stats(img, valid, device=dev)
    acc = ops.band_stats(img[:2], None, acc=acc, device=dev)            # accumulates
    got = ops.stats_to_python(acc)
    a, b = onorm.band_stats(img, valid), onorm.band_stats(img[:2])
    want = [(x[0] + y[0], x[1] + y[1], x[2] + y[2]) for x, y in zip(a, b)]
    assert got == want
    m1, s1 = ops.mean_std_from_stats(got)
    m2, s2 = onorm.mean_std_from_stats(want)
    np.testing.assert_array_equal(m1, m2)
    np.testing.assert_array_equal(s1, s2)


# ------------------------------------------------------------------------------------------------ opened-shard tables
def _small_shard(n, size, seed, id_jitter=True, K=10):
    rng = np.random.default_rng(seed)
    recs, chips = [], []
    for i in range(n):
        img = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
        lab = rng.integers(0, K + 2, (size, size), dtype=np.uint8)
        key = "256:2:1.0:43:%d:%d" % (int(rng.integers(-5000, 5000)) if id_jitter else 7, i if id_jitter else 3)
        chips.append((img, lab, key))
        recs.append(oep.convert_to_example(img, lab, size, size, 3, size, size, key).SerializeToString())
    return b"".join(otfr.frame(r) for r in recs), recs, chips


@pytest.mark.parametrize("n,size", [(300, 8), (70, 40), (1, 16)])
def test_scan_many_records_with_drifting_lengths(dev, n, size):
    """More records than one speculative scan round, identifier lengths drifting: same table as the sequential walk."""
    from dl_image_segmentation_b200 import ops
    shard, recs, chips = _small_shard(n, size, seed=n)
    si = ops.open_shard(shard, dev)
    o, l = otfr.scan(shard)
    assert si.n == n == len(o)
    np.testing.assert_array_equal(si.rec_off[:n].cpu().numpy().astype(np.uint64), o)
    np.testing.assert_array_equal(si.rec_len[:n].cpu().numpy().astype(np.uint64), l)
    assert (si.index["status"] == 0).all()
    assert si.identifiers(shard) == [c[2].encode() for c in chips]
    # a corrupt length CRC deep inside (beyond the first scan round) stops the walk exactly there
    k = min(n - 1, 200)
    bad = bytearray(shard)
    bad[int(o[k]) - 12 + 9] ^= 0x10
    st = ops.open_shard_async(bytes(bad), dev, max_records=512)
    nn, status = st.header()[:2]
    assert (nn, status) == (k, 1)
    # capacity smaller than the shard -> status 2 with the table full
    if n > 4:
        st = ops.open_shard_async(shard, dev, max_records=4)
        assert st.header()[:2] == (4, 2)
        with pytest.raises(ops.B2Error):
            st.check()


def test_scan_wildly_varying_records(dev):
    """Record lengths that defeat the stride prediction (every hop misses its window) still give the sequential result."""
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(5)
    recs = [rng.integers(0, 256, int(rng.integers(0, 9000)), dtype=np.uint8).tobytes() for _ in range(150)]
    shard = b"".join(otfr.frame(r) for r in recs)
    si = ops.open_shard(shard, dev, with_index=False)
    o, l = otfr.scan(shard)
    assert si.n == 150
    np.testing.assert_array_equal(si.rec_off[:150].cpu().numpy().astype(np.uint64), o)
    np.testing.assert_array_equal(si.rec_len[:150].cpu().numpy().astype(np.uint64), l)
    st = ops.parse_shard(si, "none", verify_crc=True)[2]
    assert not st.cpu().numpy().any()
    got = ops.crc32c(shard, o, l, device=dev)
    assert [int(x) for x in got] == [otfr.crc32c(r) for r in recs]


@pytest.mark.parametrize("depth", [1, 3])
def test_shard_pipeline_matches_oracle(dev, depth):
    """open -> fused parse enqueued shard after shard without host syncs == oracle; bad records are reported."""
    import torch

    from dl_image_segmentation_b200 import ops
    size, K = 24, 10
    mean = np.array([101.5, 99.25, 120.0], np.float32)
    std = np.array([47.0, 51.5, 33.3], np.float32)
    counts = [5, 17, 1, 12, 9]
    shards, chips = [], []
    for s, n in enumerate(counts):
        sh, _, ch = _small_shard(n, size, seed=100 + s)
        shards.append(sh)
        chips.append(ch)
    # shard 3: flip a payload byte of record 2 (data CRC must catch it)
    o3, _ = otfr.scan(shards[3])
    bad = bytearray(shards[3])
    bad[int(o3[2]) + 300] ^= 0x01
    shards[3] = bytes(bad)
    host = [torch.from_numpy(np.frombuffer(s, np.uint8).copy()).pin_memory() for s in shards]
    for source in (host, [h.to(dev) for h in host]):
        pipe = ops.ShardPipeline("norm_onehot", size * size * 3, size * size, max_records=20, mean=mean, std=std,
                                 num_classes=K, device=dev, depth=depth)
        seen = 0
        for s, (ib, tb, st, table) in enumerate(pipe.run(source)):
            n, status, tiles, max_len, n_bad = table.header()
            assert (n, status) == (counts[s], 0)
            stat = st[:n].cpu().numpy()
            want_bad = [2] if s == 3 else []
            assert list(np.nonzero(stat)[0]) == want_bad and n_bad == len(want_bad)
            if want_bad:
                with pytest.raises(ops.DataLossError):
                    table.check()
            wi = onorm.normalise(np.stack([c[0] for c in chips[s]]), mean, std)
            wt = onorm.one_hot(np.stack([c[1] for c in chips[s]]), K)
            gi = ib[:n].cpu().numpy().reshape(wi.shape)
            gt = tb[:n].cpu().numpy().reshape(wt.shape)
            for r in range(n):
                if r in want_bad:
                    continue
                np.testing.assert_array_equal(gi[r], wi[r])
                np.testing.assert_array_equal(gt[r], wt[r])
            seen += 1
        assert seen == len(counts)


def test_parse_table_raw_and_reparse(dev):
    """A table can be parsed twice (raw, then CRC only); raw rows equal the stored payload bytes."""
    from dl_image_segmentation_b200 import ops
    shard, recs, chips = _small_shard(11, 30, seed=9)
    st = ops.open_shard_async(shard, dev, max_records=16)
    ib, tb, status = ops.parse_table(st, "raw", 30 * 30 * 3, 30 * 30)
    n = st.check()
    assert n == 11 and not status[:n].cpu().numpy().any()
    for r, (img, lab, _) in enumerate(chips):
        assert bytes(ib[r, :2700].cpu().numpy()) == img.tobytes()
        assert bytes(tb[r, :900].cpu().numpy()) == lab.tobytes()
    _, _, status2 = ops.parse_table(st, "none")
    assert st.check() == 11 and not status2[:n].cpu().numpy().any()
    # rows too small for the payload -> status 3 for every record
    _, _, status3 = ops.parse_table(st, "raw", 100, 900)
    assert list(status3[:n].cpu().numpy()) == [3] * 11


def test_scan_tiny_and_truncated_shards(dev):
    """Shards shorter than a header, ending inside a header, or ending inside the data are DataLoss at the right record."""
    from dl_image_segmentation_b200 import ops
    recs = [bytes(range(40)), b"", bytes(7), bytes(300)]
    shard = b"".join(otfr.frame(r) for r in recs)
    for cut in (5, 11, 12, 20, len(otfr.frame(recs[0])) + 3, len(shard) - 1, len(shard) - 4, len(shard) - 5):
        st = ops.open_shard_async(shard[:cut], dev, max_records=16)
        n, status = st.header()[:2]
        try:
            o, _ = otfr.scan(shard[:cut])
            want = (len(o), 0)
        except otfr.DataLossError:
            # the oracle raises at the first bad frame: count the good ones before it
            good, pos = 0, 0
            while True:
                try:
                    otfr.scan(shard[:cut][pos:pos + 16 + len(recs[good])])
                except Exception:
                    break
                if pos + 16 + len(recs[good]) > cut:
                    break
                pos += 16 + len(recs[good])
                good += 1
            want = (good, 1)
        assert (n, status) == want, cut
    st = ops.open_shard_async(shard, dev, max_records=16)
    assert st.header()[:2] == (4, 0)


def test_captured_pass_graph_replay(dev):
    """A whole pass recorded as one CUDA graph (uploads from pinned memory included) replays to the oracle's values,
    and a replay after the host data changed parses the NEW bytes (the graph holds addresses, not data)."""
    import torch

    from dl_image_segmentation_b200 import ops
    size, K, n = 20, 6, 7
    mean = np.array([90.0, 100.0, 110.0], np.float32)
    std = np.array([40.0, 50.0, 60.0], np.float32)
    shards, chips = [], []
    for s in range(5):
        sh, _, ch = _small_shard(n, size, seed=300 + s, K=K, id_jitter=False)      # equal shard sizes: the graph fixes them
        shards.append(sh)
        chips.append(ch)
    assert len({len(s) for s in shards}) == 1
    host = [torch.from_numpy(np.frombuffer(s, np.uint8).copy()).pin_memory() for s in shards]
    pipe = ops.ShardPipeline("norm_onehot", size * size * 3, size * size, max_records=8, mean=mean, std=std,
                             num_classes=K, device=dev, depth=2, open_ahead=3)
    seen = []

    def consume(img, tgt, status, table):
        seen.append((img[:n].clone(), tgt[:n].clone(), status[:n].clone(), table.hdr_dev.clone()))
        return len(seen) - 1
    cp = ops.CapturedPass(pipe, host, consume)
    del seen[:len(seen) - len(shards)]                     # keep the tensors created during the capture
    for rep in range(2):
        if rep == 1:                                       # new data behind the same pinned addresses
            order = [3, 4, 0, 1, 2]
            new = [bytes(host[k].numpy()) for k in order]
            for h, b in zip(host, new):
                h.copy_(torch.from_numpy(np.frombuffer(b, np.uint8).copy()))
            chips = [chips[k] for k in order]
        cp.replay()
        torch.cuda.synchronize()
        for s, (gi, gt, st, hdr) in enumerate(seen):
            assert int(hdr[0]) == n and int(hdr[1]) == 0 and int(hdr[4]) == 0
            assert not st.cpu().numpy().any()
            wi = onorm.normalise(np.stack([c[0] for c in chips[s]]), mean, std)
            wt = onorm.one_hot(np.stack([c[1] for c in chips[s]]), K)
            np.testing.assert_array_equal(gi.cpu().numpy().reshape(wi.shape), wi)
            np.testing.assert_array_equal(gt.cpu().numpy().reshape(wt.shape), wt)
    assert cp.launches_per_replay == 3 * len(shards)


def test_index_fuzz_against_oracle_parser(dev):
    """Seeded fuzz: random key orders, unknown extra features, duplicate keys, dropped / mistyped required keys, empty and
    multi-kilobyte payloads, float-list payloads, odd identifiers.  The device index must agree with the oracle's
    protobuf walk on every record: same status class, same payload bytes, dims and identifier."""
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(77)
    recs, expect = [], []
    for i in range(160):
        h, w, c = int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(1, 5))
        floats = bool(rng.integers(0, 2))
        if floats:
            img = rng.integers(0, 60000, (h, w, c)).astype(np.uint16)
        else:
            img = rng.integers(0, 256, (h, w, c)).astype(np.uint8)
        lab = rng.integers(0, 256, (h, w)).astype(np.uint8)
        key = "".join(chr(int(x)) for x in rng.integers(33, 127, int(rng.integers(0, 200))))
        ex = oep.convert_to_example(img, lab, h, w, c, h, w, key)
        feats = dict(ex.features)
        order = list(feats)
        rng.shuffle(order)
        status = 0
        mode = int(rng.integers(0, 8))
        if mode == 1:                                        # drop a required key
            order.remove(order[int(rng.integers(0, len(order)))])
            status = 2
        elif mode == 2:                                      # wrong type for a dimension
            feats["image/width"] = oep.Feature("float", [float(w)])
            status = 2
        elif mode == 3:                                      # two values where one is expected
            feats["target/height"] = oep.Feature("int64", [h, h])
            status = 2
        built = {}
        for k in order:
            built[k] = feats[k]
            if rng.random() < 0.3:                           # unknown features sprinkled in between
                j = int(rng.integers(0, 3))
                built["zz/%d/%d" % (i, len(built))] = oep.Feature(["bytes", "float", "int64"][j],
                                                                    [[b"junk" * int(rng.integers(0, 50))], [1.5, -2.0], [7, -9]][j])
        rec = oep.Example(built).SerializeToString(deterministic=False)
        if mode == 4:                                        # duplicate key: the later entry wins (protobuf map semantics)
            dup = oep.Example({"image/height": oep.Feature("int64", [h + 100])}).SerializeToString(deterministic=False)
            # splice the inner Features entries: outer tag 0x0a + varint length
            def inner(b):
                p, ln, sh = 1, 0, 0
                while True:
                    x = b[p]; p += 1
                    ln |= (x & 0x7F) << sh; sh += 7
                    if not x & 0x80:
                        break
                return b[p:p + ln]
            body = inner(rec) + inner(dup)
            ln, var = len(body), b""
            while True:
                var += bytes([(ln & 0x7F) | (0x80 if ln > 0x7F else 0)])
                ln >>= 7
                if not ln:
                    break
            rec = b"\x0a" + var + body
        recs.append(rec)
        expect.append((status, mode, h, w, c, key))
    shard = b"".join(otfr.frame(r) for r in recs)
    si = ops.open_shard(shard, dev)
    assert si.n == len(recs)
    for r, (rec, (status, mode, h, w, c, key)) in enumerate(zip(recs, expect)):
        ix = si.index[r]
        assert int(ix["status"]) == status, (r, mode)
        if status:
            continue
        f = oep.parse_example(rec)
        assert int(ix["height"]) == f["image/height"][1][0] == (h + 100 if mode == 4 else h)
        assert (int(ix["width"]), int(ix["channels"]), int(ix["tgt_height"]), int(ix["tgt_width"])) == (w, c, h, w)
        kind = {"bytes": 1, "float": 2}[f["image/image_data"][0]]
        assert int(ix["img_kind"]) == kind == int(ix["tgt_kind"])
        o, l = int(ix["img_off"]), int(ix["img_len"])
        if kind == 1:
            assert shard[o:o + l] == f["image/image_data"][1][0]
            assert shard[int(ix["tgt_off"]):int(ix["tgt_off"]) + int(ix["tgt_len"])] == f["target/target_data"][1][0]
        else:
            assert np.array_equal(np.frombuffer(shard[o:o + l], "<f4"), np.asarray(f["image/image_data"][1], np.float32))
        assert shard[int(ix["id_off"]):int(ix["id_off"]) + int(ix["id_len"])] == key.encode("utf-8")


@pytest.mark.gpu
def test_context_workspace_is_safe_across_streams_and_host_threads(dev):
    """b2_crc32c / b2_tfrecord_parse / b2_tfrecord_build share one per-context scratch.  Calls on different
    streams, from different host threads, must not corrupt each other: each entry point holds the workspace lock for its
    body and a call waits (on the device) for the previous user's stream."""
    import threading

    import torch

    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(17)
    data = rng.integers(0, 256, 3 << 20, dtype=np.uint8)
    n = 300
    offs = np.sort(rng.integers(0, data.size - 20000, n)).astype(np.uint64)
    lens = rng.integers(1, 20000, n).astype(np.uint64)
    want = np.array([otfr.crc32c(data[int(o):int(o + l)].tobytes()) for o, l in zip(offs, lens)], dtype=np.uint32)
    d_dev = torch.from_numpy(data).to(dev)
    imgs = [rng.integers(0, 256, (40, 40, 3), dtype=np.uint8) for _ in range(40)]
    labs = [rng.integers(0, 10, (40, 40), dtype=np.uint8) for _ in range(40)]
    want_rec = b"".join(otfr.frame(oep.convert_to_example(i, l, 40, 40, 3, 40, 40, "k%d" % k).SerializeToString())
                        for k, (i, l) in enumerate(zip(imgs, labs)))
    errors = []

    def crc_loop():
        try:
            st = torch.cuda.Stream(dev)
            with torch.cuda.stream(st):
                for _ in range(30):
                    assert np.array_equal(ops.crc32c(d_dev, offs, lens, dev), want)
        except Exception as e:            # noqa: BLE001
            errors.append(e)

    def build_loop():
        try:
            st = torch.cuda.Stream(dev)
            with torch.cuda.stream(st):
                items = [dict(img=torch.from_numpy(i).to(dev).reshape(-1), tgt=torch.from_numpy(l).to(dev).reshape(-1), kind=1,
                              h=40, w=40, c=3, th=40, tw=40, identifier=("k%d" % k).encode()) for k, (i, l) in enumerate(zip(imgs, labs))]
                for _ in range(30):
                    buf, _, total = ops.build_records(items, dev)
                    assert bytes(buf[:total].cpu().numpy()) == want_rec
        except Exception as e:            # noqa: BLE001
            errors.append(e)
    ts = [threading.Thread(target=crc_loop), threading.Thread(target=build_loop), threading.Thread(target=crc_loop)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[0]
